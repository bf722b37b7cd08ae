"""BASELINE.json configs[4]: batch sweep of synthetic 4-s 16 kHz clips, N = 30, STRONG scaling -- a fixed global batch sharded over
G ranks (one process per GPU, no data-path collective; NCCL only for the weight broadcast).  One process per rank runs the
whole sweep (engine, weights and NCCL set up once), so a G = 8 sweep costs a minute of box time instead of ten launches.

    python tools/sweep_bench.py --batches 1,2,4,8,16,32,64,128                              # G = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
        tools/sweep_bench.py --batches 1,2,4,8,16,32,64,128,256,512                         # G = 8

Rank 0 prints one JSON line per global batch: utterances/s over the whole job (global batch / max-over-ranks device time of a
step, CUDA events, barrier + synchronize on both sides), the per-rank batches, ms per step (= latency of the batch), clocks.
Every step goes through the host-buffer entry point (pinned wav in, pinned wav out: the end-to-end number).
For global batch 1 it also reports what bounds the latency: launches per step, graph-replay time per network evaluation and
the same evaluation launched eagerly.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ClockSampler                                  # noqa: E402
from sgmse_b200 import Engine, EngineConfig                     # noqa: E402
from sgmse_b200.dist import broadcast_weights, shard_range      # noqa: E402
from sgmse_b200.synth import synthetic_blob, synthetic_speech   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="1,2,4,8,16,32,64,128")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--micro-batch", type=int, default=16)
ap.add_argument("--N", type=int, default=30)
ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="Engine.set_option (e.g. pdl=1 with SGMSE_B200_PDL=1)")
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
eng = Engine(EngineConfig(mode="fp16_tc", max_batch=a.micro_batch, use_graphs=True), device=dev)
eng.load_blob(broadcast_weights(synthetic_blob(eng, 0) if rank == 0 else None, eng.weights_numel(), dev))
for kv in a.opt:
    k, v = kv.split("=")
    eng.set_option(k, int(v))
L = 64000
kw = dict(N=a.N, predictor="reverse_diffusion", corrector="ald", corrector_steps=1, snr=0.5)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for B in [int(b) for b in a.batches.split(",")]:
    lo, hi = shard_range(B, rank, world)
    mine = hi - lo
    wav = synthetic_speech(mine, L, first=lo).pin_memory() if mine else None
    out = torch.empty_like(wav).pin_memory() if mine else None

    def step(i):
        if mine:
            eng.enhance(wav, out=out, seed=1 + i, utt_offset=lo, **kw)

    for i in range(a.warmup):
        step(i)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(i)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clk = clocks.stop() if rank == 0 else None
    ms_step = ms.item() / a.steps
    line = {"metric": "utterances/sec (4 s, 16 kHz, N=30 PC)", "value": round(B / (ms_step * 1e-3), 4), "unit": "utterances/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms_step, 3), "scaling": "strong",
            "higher_is_better": True, "dtype": "f16 (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": "SGMSE+ NCSN++ (VoiceBank-DEMAND config), 16 kHz, 4-s clips, PC reverse_diffusion+ald N=%d snr 0.5" % a.N,
                       "global_batch": B, "per_rank_batch": [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)],
                       "micro_batch": a.micro_batch, **({"options": list(a.opt)} if a.opt else {}), "parallelism": f"dp{world} (fixed global batch sharded, no data-path collective)",
                       "timed": "host-buffer entry point (pinned wav in / out), CUDA events, max over ranks"},
            "latency_ms": round(ms_step, 3), "rtf": round(ms_step * 1e-3 / (B * 4.0), 6), "clocks": clk}
    if B == 1 and rank == 0:
        # what bounds one clip: ~300 dependent launches per network evaluation, most of them single-wave
        l0 = eng.counter("kernel_launches")
        step(0)
        torch.cuda.synchronize()
        launches = eng.counter("kernel_launches") - l0
        Y, _ = eng.analysis(wav.to(dev))
        x = torch.cat([Y, Y], 1)
        t = torch.full((1,), 0.5, device=dev)
        eng.set_option("use_graphs", 0)
        eng.dnn_forward(x, t)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(5):
            eng.dnn_forward(x, t)
        f1.record()
        torch.cuda.synchronize()
        eng.set_option("use_graphs", 1)
        line["b1"] = {"kernel_nodes_per_step": int(launches), "network_evaluations_per_step": 2 * a.N,
                      "launches_per_evaluation": int(eng.counter("launches_last_forward")),
                      "graph_replay_ms_per_evaluation": round(ms_step / (2 * a.N), 4),
                      "eager_ms_per_evaluation": round(f0.elapsed_time(f1) / 5, 4)}
    if rank == 0:
        print(json.dumps(line), flush=True)
eng.close()
if world > 1:
    dist.destroy_process_group()

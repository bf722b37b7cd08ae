#!/usr/bin/env python
"""Benchmark of the reverse-SDE enhancement hot path (BASELINE.json metric: utterances/sec, 4-s 16 kHz
clips, N=30 predictor-corrector steps = 60 score-network evaluations per utterance).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm on the host CPU cores (oracle port)

A "step" is one pass of the hot path over one batch of synthetic noisy speech: STFT -> magnitude
compression -> pad -> N-step PC sampling with the NCSN++ score network -> decompression -> iSTFT.
Workload at every N: configs[1] of BASELINE.json (SGMSE+ NCSN++ VoiceBank config, 16 kHz, batch 16 per GPU,
N=30); utterances are independent, so ranks shard the batch with no data-path collective (weak scaling,
NCCL only for the one-off weight broadcast).

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 16000
CLIP_S = 4
# algorithmic work of one NCSN++ forward on one 4-s 16 kHz utterance ([1,4,256,512])
GFLOP_PER_FORWARD = 1064.7


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="utterances per GPU per step")
    ap.add_argument("--micro-batch", type=int, default=16)
    ap.add_argument("--N", type=int, default=30)
    ap.add_argument("--mode", default="fp16_tc")
    ap.add_argument("--lanes", type=int, default=1, help="concurrent launch sequences inside the sampler graph")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4],
                    help="BASELINE.json configs[i-1]: 2 = the metric's workload (default, the only one the driver runs); "
                         "3 = ncsnpp_48k 48 kHz batch 8; 4 = dereverb settings N=50 snr=0.33 batch 32 (parity-test cases, "
                         "measurable here for the record; no CPU baseline)")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="engine A/B option (Engine.set_option), e.g. --opt pdl=1 with SGMSE_B200_PDL=1; recorded in config")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(state):
    """One bounded sample: 1 utterance, STFT -> 1 of the N=30 PC steps (2 of 60 network evaluations) -> iSTFT.
    Returns seconds; utterances/s is extrapolated by scaling the sampler part to 30 steps."""
    import torch
    from oracle import pipeline as o_pipe, sde as o_sde, spec as o_spec
    sd, ncfg, wav, draws = state
    t0 = time.perf_counter()
    o_pipe.enhance(sd, ncfg, o_spec.SpecConfig(), o_sde.OUVE(), wav, draws, N=1)
    return time.perf_counter() - t0


_CPU_THREADS = None


def pick_cpu_threads(sd, ncfg):
    """Use as many host threads as actually help: a quarter-second probe (one forward on a [1,2,256,64] input)
    per candidate, best wins (oversubscribed MKL-DNN convolutions get slower, not faster)."""
    global _CPU_THREADS
    if _CPU_THREADS is not None:
        return _CPU_THREADS
    import torch
    from oracle import ncsnpp as o_net
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cands = sorted({c for c in (avail, avail // 2, 64, 32, 16, 8) if 1 <= c <= avail}, reverse=True)
    x = torch.complex(torch.randn(1, 2, 256, 64), torch.randn(1, 2, 256, 64))
    t = torch.tensor([0.5])
    best = (float("inf"), avail)
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            o_net.forward(sd, ncfg, x, t)
            t0 = time.perf_counter()
            o_net.forward(sd, ncfg, x, t)
            dt = time.perf_counter() - t0
        if dt < best[0]:
            best = (dt, c)
    _CPU_THREADS = best[1]
    torch.set_num_threads(_CPU_THREADS)
    return _CPU_THREADS


def cpu_state():
    import torch
    from oracle import weights as o_w, sde as o_sde
    from oracle.arch import NetConfig
    ncfg = NetConfig.ncsnpp()
    sd = o_w.make_state_dict(ncfg, seed=0)
    pick_cpu_threads(sd, ncfg)
    from sgmse_b200.synth import synthetic_speech
    wav = synthetic_speech(1, SR * CLIP_S)
    draws = o_sde.make_noise((1, 1, 256, 512), 3, seed=2000)
    return sd, ncfg, wav, draws


def cpu_baseline(n_steps_total=30, reps=1):
    st = cpu_state()
    ts = [cpu_reference_step(st) for _ in range(reps)]
    t = min(ts)
    return {"value": 1.0 / (t * n_steps_total), "unit": "utterances/s", "cores": _CPU_THREADS, "host_cpus": os.cpu_count(), "kind": "port",
            "sample": f"1 utterance (4 s, 16 kHz), STFT + 1 of {n_steps_total} PC steps (2 of {2 * n_steps_total} NCSN++ "
                      f"evaluations) + iSTFT on the fp32 torch-CPU oracle port, {t:.1f} s; utterances/s extrapolated x{n_steps_total}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    st = cpu_state()
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_step(st)
    ts = [cpu_reference_step(st) for _ in range(args.steps)]
    t = sum(ts) / len(ts)
    v = 1.0 / (t * args.N)
    cb = {"value": v, "unit": "utterances/s", "cores": _CPU_THREADS, "host_cpus": os.cpu_count(), "kind": "port",
          "sample": f"per step: 1 utterance, STFT + 1 of {args.N} PC steps + iSTFT; extrapolated x{args.N}"}
    print(json.dumps({
        "impl": "reference", "metric": "utterances/sec (4 s, 16 kHz, N=30 PC)", "value": v, "unit": "utterances/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3 * args.N,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args), "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(args):
    return {"workload": f"{args.workload_name}, 4-s clips, "
                        f"batch {args.batch} per GPU, PC sampler reverse_diffusion+ald N={args.N} snr {args.snr} "
                        f"({2 * args.N} network evaluations), STFT {args.stft}",
            "baseline_config": args.config,
            "global_batch": args.batch * args.gpus, "per_gpu_batch": args.batch, "micro_batch": args.micro_batch, "lanes": args.lanes,
            "parallelism": f"dp{args.gpus} (batch sharded, no data-path collective)",
            **({"options": list(args.opt)} if getattr(args, "opt", None) else {}),
            "l2": "working set per step (>10 GB of activations per micro-batch) exceeds the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from sgmse_b200 import Engine, EngineConfig
    from sgmse_b200.synth import synthetic_blob, synthetic_speech

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    ecfg = (EngineConfig.ncsnpp_48k if args.config == 3 else EngineConfig)(mode=args.mode, max_batch=args.micro_batch, use_graphs=True)
    eng = Engine(ecfg, device=dev)
    eng.set_option("lanes", args.lanes)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    # weights: rank 0 creates them, NCCL broadcast over NVLink, packed per rank
    n = eng.weights_numel()
    if rank == 0:
        blob = synthetic_blob(eng, seed=0).to(dev)
    else:
        blob = torch.empty(n, dtype=torch.float32, device=dev)
    if world > 1:
        dist.broadcast(blob, src=0)
    eng.load_blob(blob)
    del blob

    L = SR * CLIP_S
    wav_host = synthetic_speech(args.batch, L, first=rank * args.batch).pin_memory()
    wav_dev = wav_host.to(dev)
    out_dev = torch.empty_like(wav_dev)
    out_host = torch.empty_like(wav_host).pin_memory()
    kw = dict(N=args.N, predictor="reverse_diffusion", corrector="ald", corrector_steps=1, snr=args.snr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    step_dev = lambda i: eng.enhance(wav_dev, out=out_dev, seed=1 + i, utt_offset=rank * args.batch, **kw)
    step_host = lambda i: eng.enhance(wav_host, out=out_host, seed=1 + i, utt_offset=rank * args.batch, **kw)

    for i in range(args.warmup):
        step_dev(i)
    l0 = eng.counter("kernel_launches")
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = timed(step_dev, args.steps)
    launches = eng.counter("kernel_launches") - l0
    step_host(0)
    ms_e2e = timed(step_host, args.steps)
    clk = clocks.stop() if rank == 0 else None
    assert torch.isfinite(out_dev).all() and torch.isfinite(out_host).all()

    utts = args.batch * world * args.steps
    value = utts / (ms * 1e-3)
    e2e = utts / (ms_e2e * 1e-3)

    roof = None
    if not args.no_roofline and args.mode == "fp16_tc":
        # dominant kernels = the tcgen05 implicit-GEMM convolutions: CUDA events around every launch of one more
        # (eager, un-graphed) step on the launching stream
        eng.set_option("time_convs", 1)
        step_dev(0)
        torch.cuda.synchronize()
        us = eng.counter("timed_conv_tc_us")
        mflop = eng.counter("timed_conv_tc_mflop")
        cnt = eng.counter("timed_conv_tc_count")
        kbytes = eng.counter("timed_conv_tc_kbytes")    # algorithmic bytes of the same launches (inputs once + output once)
        eng.set_option("time_convs", 0)
        pk = peaks()
        ach = mflop / max(us, 1)            # MFLOP/us = TFLOP/s
        # DRAM traffic of the dominant kernel comes from the committed ncu capture (one launch of the dominant shape)
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, "profiles", "r01_conv_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj["traffic_bytes_per_launch"]
            traffic_note = {k: tj[k] for k in ("kernel", "shape", "algorithmic_bytes_per_launch", "source")}
        roof = {"kernel": "tcgen05 implicit-GEMM convolutions (conv_tc6 with fused GroupNorm+SiLU producers; conv_tc4 / conv_tc on the levels below 32 rows)",
                "bound": "tensor", "achieved": round(ach, 1),
                "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": round(ach / pk["tflops_sustained"], 4),
                "traffic": traffic, "traffic_of": traffic_note,
                "launches_timed": cnt, "avg_launch_us": round(us / max(cnt, 1), 1),
                "peak_source": pk["source"] + " (bf16_tflops_sustained: kernel timed inside a long step)",
                "share_of_step": round(us * 1e-3 / (ms / args.steps), 3),
                # the same launches against the other roof: algorithmic HBM bytes / time (KB/us = GB/s)
                "algorithmic_hbm_gbs": round(kbytes / max(us, 1), 1), "hbm_peak_gbs": pk["hbm_gbs"],
                "hbm_frac": round(kbytes / max(us, 1) / pk["hbm_gbs"], 4),
                "algorithmic_gflop_per_step": round(mflop * 1e-3, 1)}
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(args.N)

    if rank == 0:
        line = {
            "metric": "utterances/sec (4 s, 16 kHz, N=30 PC)" if args.config == 2 else
                      f"utterances/sec (4 s, {SR // 1000} kHz, N={args.N} PC)", "value": round(value, 4), "unit": "utterances/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 (fp32 accumulate)" if args.mode != "fp32" else "f32", "data": "synthetic",
            "config": workload_config(args), "rtf": round((ms * 1e-3) / (utts * CLIP_S), 6),
            "e2e": {"value": round(e2e, 4), "unit": "utterances/s", "h2d_bytes_per_step": args.batch * L * 4,
                    "d2h_bytes_per_step": args.batch * L * 4, "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cb,
            "tflops_effective": round(value * 2 * args.N * GFLOP_PER_FORWARD * 1e-3 / world, 1),
        }
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


WORKLOADS = {
    # name, sampling rate, GFLOP per forward per utterance, per-GPU batch, micro-batch, N, snr
    2: dict(name="SGMSE+ NCSN++ (VoiceBank-DEMAND config, 65.6 M params, random init), 16 kHz", sr=16000, gflop=1064.7,
            batch=16, micro=16, N=30, snr=0.5, stft="510/128"),
    3: dict(name="NCSN++ 48 kHz (EARS-WHAM config: backbone ncsnpp_48k, 64.7 M params, random init), 48 kHz", sr=48000,
            gflop=3187.6, batch=8, micro=8, N=30, snr=0.5, stft="1534/384"),
    4: dict(name="SGMSE+ NCSN++ (WSJ0-REVERB dereverberation settings, random init), 16 kHz", sr=16000, gflop=1064.7,
            batch=32, micro=16, N=50, snr=0.33, stft="510/128"),
}


def apply_workload(args):
    """--config 3 / 4 replace the defaults of --batch / --micro-batch / --N (explicit flags still win is NOT attempted:
    a named config means its published settings)."""
    global SR, GFLOP_PER_FORWARD
    w = WORKLOADS[args.config]
    args.snr = w["snr"]
    args.workload_name = w["name"]
    args.stft = w["stft"]
    if args.config != 2:
        args.batch, args.micro_batch, args.N = w["batch"], w["micro"], w["N"]
        args.no_cpu_baseline = True
        SR, GFLOP_PER_FORWARD = w["sr"], w["gflop"]


def main():
    args = parse()
    apply_workload(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

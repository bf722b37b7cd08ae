"""Shard-count invariance on real GPUs (multi-GPU row): the same seeds at G ranks and at 1 rank give
BIT-EQUAL per-utterance outputs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_invariance.py

Every rank: NCCL broadcast of the fp32 weight blob from rank 0 (the only collective of the data path), then
``dist.enhance_sharded(engine.enhance, wav, seed=...)`` (all_gather of the waveforms), then -- on every rank -- the whole batch
once more in a single-process call; the two must be identical bit for bit.  Also checks the documented refusal of the
batch-coupled 'langevin' corrector.  Full-size NCSN++ (65.6 M parameters), product mode, 1-s clips, N = 2.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sgmse_b200 import Engine, EngineConfig                      # noqa: E402
from sgmse_b200.dist import broadcast_weights, enhance_sharded   # noqa: E402
from sgmse_b200.synth import synthetic_blob, synthetic_speech    # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, L, N = int(os.environ.get("INV_BATCH", 5)), 16000, 2
    eng = Engine(EngineConfig(mode="fp16_tc", max_batch=4, use_graphs=True), device=dev)
    blob = synthetic_blob(eng, seed=0) if rank == 0 else None
    eng.load_blob(broadcast_weights(blob, eng.weights_numel(), dev))
    wav = synthetic_speech(B, L).to(dev)
    kw = dict(N=N, predictor="reverse_diffusion", corrector="ald", corrector_steps=1, snr=0.5, seed=1234)
    sharded = enhance_sharded(eng.enhance, wav, gather=True, **kw)
    whole = eng.enhance(wav, utt_offset=0, **kw)                  # the G = 1 result, computed on this rank
    ok = bool(torch.equal(sharded, whole)) and bool(torch.isfinite(whole).all())
    refused = False
    if world > 1:
        try:
            enhance_sharded(eng.enhance, wav, gather=True, **{**kw, "corrector": "langevin"})
        except ValueError as e:
            refused = "langevin" in str(e)
    flags = torch.tensor([int(ok), int(refused or world == 1)], device=dev)
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"nccl_invariance: world {world}, batch {B} x {L} samples, N={N}: sharded == single-process on every rank: "
              f"{bool(flags[0])}; langevin refused when sharded: {bool(flags[1])}; max |x| {whole.abs().max().item():.3f}")
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if bool(flags.min()) else 1)


if __name__ == "__main__":
    main()

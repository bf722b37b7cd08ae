#!/usr/bin/env python
"""Probability-flow ODE sampler at the benchmark shape: wall/device time of one solve, evaluations
per second, and how much of the solve is NOT the score network (RK stage kernels, norm reductions, the one-double D2H
syncs).  Not a BASELINE.json metric -- a measuring tool for round 2.

    python tools/bench_ode.py --batch 16 --rtol 1e-2 --atol 1e-2 [--no-graphs]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgmse_b200 import Engine, EngineConfig          # noqa: E402
from sgmse_b200.synth import synthetic_blob, synthetic_speech   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--micro-batch", type=int, default=16)
    ap.add_argument("--rtol", type=float, default=1e-2)
    ap.add_argument("--atol", type=float, default=1e-2)
    ap.add_argument("--eps", type=float, default=0.03)
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    eng = Engine(EngineConfig(mode="fp16_tc", max_batch=a.micro_batch, use_graphs=not a.no_graphs))
    eng.load_blob(synthetic_blob(eng, seed=0))
    wav = synthetic_speech(a.batch, 64000).cuda()
    Y, _ = eng.analysis(wav)
    # one network evaluation alone, for the split
    x = torch.cat([Y, Y], dim=1)
    t = torch.full((a.batch,), 0.5, device="cuda")
    for _ in range(2):
        eng.dnn_forward(x, t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        eng.dnn_forward(x, t)
    torch.cuda.synchronize()
    fwd_ms = (time.perf_counter() - t0) / 3 * 1e3
    for rep in range(a.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, nfe, st = eng.ode_sample(Y, rtol=a.rtol, atol=a.atol, eps=a.eps, denoise=False, seed=1, return_stats=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"rep {rep}: {dt * 1e3:.1f} ms, nfe {nfe}, {st}, {nfe / dt:.1f} evaluations/s, {a.batch / dt:.2f} utterances/s; "
              f"network alone {fwd_ms:.2f} ms/eval -> {100 * (1 - nfe * fwd_ms * 1e-3 / dt):.1f} % of the solve is outside it; "
              f"the reference's host round trips would add 2 x {out.numel() * 8 / 1e6:.1f} MB of PCIe traffic per evaluation")
    assert torch.isfinite(torch.view_as_real(out)).all()
    eng.close()


if __name__ == "__main__":
    main()

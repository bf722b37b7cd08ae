"""Repeat the benchmark-shaped PC sampler (graph replay, concurrent lanes) and report failures with the barrier-wait code.

    python tools/stress_sampler.py [--reps 20] [--batch 16] [--lanes 2] [--variant 0] [--N 30]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sgmse_b200 import Engine, EngineConfig
from sgmse_b200.synth import synthetic_blob

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--lanes", type=int, default=2)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--N", type=int, default=30)
ap.add_argument("--T", type=int, default=512)
a = ap.parse_args()

cfg = EngineConfig(mode="fp16_tc", max_batch=a.batch)
eng = Engine(cfg)
eng.load_blob(synthetic_blob(eng, 0))
eng.set_option("tc_variant", a.variant)
eng.set_option("lanes", a.lanes)
F = cfg.n_fft // 2 + 1
g = torch.Generator().manual_seed(0)
y = (torch.complex(torch.randn(a.batch, 1, F, a.T, generator=g), torch.randn(a.batch, 1, F, a.T, generator=g)) * 0.3).cuda()
ref = None
for r in range(a.reps):
    t0 = time.time()
    try:
        out, nfe = eng.pc_sample(y, N=a.N, seed=1)
        torch.cuda.synchronize()
    except Exception as ex:  # noqa: BLE001
        print(f"rep {r}: FAILED: {str(ex).splitlines()[0]}  barrier_wait_code={eng.counter('barrier_wait_code')}", flush=True)
        sys.exit(1)
    dt = time.time() - t0
    fin = bool(torch.isfinite(torch.view_as_real(out)).all())
    same = True if ref is None else bool((out == ref).all())
    if ref is None:
        ref = out.clone()
    print(f"rep {r}: {dt*1e3:8.1f} ms  finite {fin}  bitwise-equal-to-rep0 {same}", flush=True)
    if not (fin and same):
        sys.exit(2)
print("ok")

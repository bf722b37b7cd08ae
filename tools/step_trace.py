"""Per-step trace of the benchmark workload: every step timed with its own CUDA events, nvidia-smi sampled every 100 ms (SM clock,
power, temperature, throttle reasons).  Shows whether the step time drifts with the thermal / power state of the chip.

    python tools/step_trace.py --steps 40 [--opt key=value ...]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sgmse_b200 import Engine, EngineConfig
from sgmse_b200.synth import synthetic_blob, synthetic_speech

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--lanes", type=int, default=0, help="0 = engine default (2)")
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
eng = Engine(EngineConfig(mode="fp16_tc", max_batch=16, use_graphs=True))
eng.load_blob(synthetic_blob(eng, 0))
if a.lanes:
    eng.set_option("lanes", a.lanes)
for kv in a.opt:
    k, v = kv.split("=")
    eng.set_option(k, int(v))
wav = synthetic_speech(a.batch, 64000).cuda()
out = torch.empty_like(wav)
kw = dict(N=30, predictor="reverse_diffusion", corrector="ald", corrector_steps=1, snr=0.5)
eng.enhance(wav, out=out, seed=1, **kw)            # capture
torch.cuda.synchronize()
rows = []
q = "timestamp,clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_slowdown"
proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
t_start = time.time()


def reader():
    for line in proc.stdout:
        rows.append((time.time() - t_start, [c.strip() for c in line.split(",")]))


threading.Thread(target=reader, daemon=True).start()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
evs[0].record()
for i in range(a.steps):
    eng.enhance(wav, out=out, seed=2 + i, **kw)
    evs[i + 1].record()
torch.cuda.synchronize()
proc.terminate()
ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(a.steps)]
print("options", a.opt, "lanes", a.lanes or "default")
print("ms per step:", " ".join(f"{m:.1f}" for m in ms))
cum = 0.0
for i, m in enumerate(ms):
    t0, t1 = cum * 1e-3, (cum + m) * 1e-3
    cum += m
    sel = [r for t, r in rows if t0 <= t < t1 and len(r) >= 7]
    if not sel:
        continue
    clk = sorted(float(r[1]) for r in sel)
    pw = [float(r[2]) for r in sel]
    tp = [float(r[3]) for r in sel]
    print(f"step {i:3d}: {m:8.1f} ms  clock median {clk[len(clk) // 2]:6.0f} MHz (min {clk[0]:.0f} max {clk[-1]:.0f})  power {sum(pw) / len(pw):6.1f} W  temp {max(tp):.0f} C  "
          f"power_cap {sum(r[4].startswith('Active') for r in sel)}/{len(sel)} sw_thermal {sum(r[5].startswith('Active') for r in sel)} hw_slowdown {sum(r[6].startswith('Active') for r in sel)}")
eng.close()

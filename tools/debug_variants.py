"""Compare per-module activations of two tc_variant settings on the full-size network (debug helper)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgmse_b200 import Engine, EngineConfig
from sgmse_b200.synth import synthetic_blob

B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 2, int(sys.argv[2]) if len(sys.argv) > 2 else 128
va, vb = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (4, 0)
eng = Engine(EngineConfig(mode="fp16_tc", max_batch=B, use_graphs=False))
eng.load_blob(synthetic_blob(eng, 0))
eng.set_option("record_taps", 1)
if len(sys.argv) > 5:
    eng.set_option("tc_mask", int(sys.argv[5], 0))
g = torch.Generator().manual_seed(0)
x = (torch.complex(torch.randn(B, 2, 256, T, generator=g), torch.randn(B, 2, 256, T, generator=g)) * 0.3).cuda()
t = torch.linspace(0.9, 0.2, B).cuda()
names = ["in_conv"] + [f"m{i}" for i in range(4, 77)] + [f"pyr{i}" for i in range(7)]
res = {}
for v in (va, vb):
    eng.set_option("tc_variant", v)
    eng.dnn_forward(x, t); torch.cuda.synchronize()
    t0 = time.perf_counter(); out = eng.dnn_forward(x, t); torch.cuda.synchronize()
    print(f"variant {v}: forward {1e3 * (time.perf_counter() - t0):.2f} ms, tc {eng.counter('tc_convs_last_forward')}, launches {eng.counter('launches_last_forward')}")
    taps = {}
    for nme in names:
        try:
            taps[nme] = eng.tap(nme)
        except RuntimeError:
            pass
    res[v] = (out.cpu(), taps)
a, b = res[va], res[vb]
for nme in names:
    if nme in a[1] and nme in b[1]:
        ta, tb = a[1][nme], b[1][nme]
        errs = [((ta[n] - tb[n]).norm() / ta[n].norm()).item() for n in range(B)]
        flag = " <<<" if max(errs) > 1e-2 else ""
        print(f"{nme:8s} " + " ".join(f"{e:.2e}" for e in errs) + flag)
print("out", [((a[0][n] - b[0][n]).norm() / a[0][n].norm()).item() for n in range(B)])

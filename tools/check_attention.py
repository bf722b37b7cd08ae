"""Gate of the tcgen05 attention kernel (csrc/attn_umma.cu, attn_variant 0 = default): the full-size network at the benchmark
shape against the mma.sync kernel (attn_variant 3, the round-1 default) and the fp32 CUDA-core kernel (attn_variant 1) --
outputs of the three 512-token attention blocks (oracle tap names m21, m23, m50) and of the whole network.

    python tools/check_attention.py [--batch 2]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sgmse_b200 import Engine, EngineConfig
from sgmse_b200.synth import synthetic_blob

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2)
a = ap.parse_args()
eng = Engine(EngineConfig(mode="fp16_tc", max_batch=a.batch, use_graphs=False))
eng.load_blob(synthetic_blob(eng, 0))
g = torch.Generator().manual_seed(0)
x = (torch.complex(torch.randn(a.batch, 2, 256, 512, generator=g), torch.randn(a.batch, 2, 256, 512, generator=g)) * 0.3).cuda()
t = torch.linspace(0.9, 0.1, a.batch).cuda()
eng.set_option("record_taps", 1)


def run(variant):
    eng.set_option("attn_variant", variant)
    out = eng.dnn_forward(x, t)
    torch.cuda.synchronize()
    return out.clone(), {k: eng.tap(k).clone() for k in ("m21", "m23", "m50")}


def rel(p, q):
    p, q = torch.view_as_real(p) if p.is_complex() else p, torch.view_as_real(q) if q.is_complex() else q
    return (torch.linalg.vector_norm((p - q).float()) / torch.linalg.vector_norm(q.float())).item()


ref_out, ref_taps = run(3)
ok = False
for v in (0, 1):
    try:
        out, taps = run(v)
        errs = {k: rel(taps[k], ref_taps[k]) for k in taps}
        e = rel(out, ref_out)
        good = all(torch.isfinite(taps[k]).all().item() for k in taps) and max(errs.values()) < 5e-3 and e < 5e-3
        ok = ok or (good and v == 0)
        print(f"attn_variant={v}: " + ", ".join(f"{k} rel-L2 {errs[k]:.3e}" for k in errs) + f"; network output rel-L2 {e:.3e} -> {'OK' if good else 'MISMATCH'}", flush=True)
    except RuntimeError as ex:
        print(f"attn_variant={v}: FAILED {ex}", flush=True)
        break
eng.set_option("attn_variant", 0)
eng.close()
sys.exit(0 if ok else 1)

"""Gate between the round-2 default kernels (option value 0) and the round-1 kernels they replaced (which keep a number of their own).

    python tools/check_candidates.py            # on the B200 box

For each option it evaluates the score network with the default and with the round-1 kernel -- the reduced config of the golden
fixtures and the full-size network at [2, 256, 128] and [2, 256, 512] -- and checks the promise: bit-identical output for
`outconv_variant`, `inconv_variant`, `combine_variant`, `tc1_narrow`, `gn_self`, `gnfin_variant`, `tc6_lean` (same arithmetic, different
memory pipelining / tiling / thread mapping); rel-L2 <= 2e-3 for `fir_variant` (half2 FIR-up arithmetic) and for `attn_variant`
(tcgen05 vs mma.sync).  First run in round 2 with the old numbering (profiles/r02_candidates_gate.txt); the same gate is
tests/test_gpu_zz_next_rows.py::test_round1_kernels_agree_with_the_defaults.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sgmse_b200 import Engine, EngineConfig
from sgmse_b200.synth import synthetic_blob

# (option, number of the ROUND-1 kernel, tolerance against the current default = 0)
CANDIDATES = [("outconv_variant", 4, 0.0), ("inconv_variant", 3, 0.0), ("attn_variant", 3, 2e-3), ("combine_variant", 2, 0.0),
              ("tc1_narrow", 2, 0.0), ("gn_self", 2, 0.0), ("gnfin_variant", 2, 0.0), ("fir_variant", 3, 2e-3), ("tc6_lean", 4, 0.0),
              ("tc6_lean", 1, 0.0)]
SMALL = dict(nf=32, ch_mult=(1, 2, 2), image_size=64, num_res_blocks=1, attn_resolutions=(16,), n_fft=126, hop_length=32)
CASES = [("small nf=32 [2,64,64]", EngineConfig(mode="fp16_tc", max_batch=2, use_graphs=False, **SMALL), (2, 2, 64, 64)),
         ("full size [2,256,128]", EngineConfig(mode="fp16_tc", max_batch=2, use_graphs=False), (2, 2, 256, 128)),
         ("full size [2,256,512]", EngineConfig(mode="fp16_tc", max_batch=2, use_graphs=False), (2, 2, 256, 512))]

failed = []
for label, cfg, shape in CASES:
    eng = Engine(cfg)
    eng.load_blob(synthetic_blob(eng, 0))
    g = torch.Generator().manual_seed(3)
    x = (torch.complex(torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)) * 0.3).cuda()
    t = torch.tensor([0.7, 0.2]).cuda()
    ref = eng.dnn_forward(x, t)
    torch.cuda.synchronize()
    for key, val, tol in CANDIDATES:
        eng.set_option(key, val)
        try:
            got = eng.dnn_forward(x, t)
            torch.cuda.synchronize()
            err = (torch.linalg.vector_norm(torch.view_as_real(got - ref)) / torch.linalg.vector_norm(torch.view_as_real(ref))).item()
            ok = torch.equal(got, ref) if tol == 0.0 else err <= tol
            verdict = "ok" if ok else "FAIL"
        except RuntimeError as ex:                     # a trap poisons the context: report and stop this case
            err, ok, verdict = float("nan"), False, f"ERROR {ex}"
        print(f"{label:24s} {key}={val}: rel-L2 vs default {err:.3e} ({'bitwise' if tol == 0.0 else f'<= {tol:g}'}) {verdict}", flush=True)
        if not ok:
            failed.append((label, key))
            if verdict.startswith("ERROR"):
                sys.exit(f"stopping: {failed}")
        eng.set_option(key, 0)
    eng.close()
print("all candidates keep their promise" if not failed else f"FAILED: {failed}")
sys.exit(1 if failed else 0)

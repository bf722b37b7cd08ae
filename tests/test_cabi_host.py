"""Host-only checks of the C-ABI library: it loads, exports every symbol include/sgmse_b200.h declares,
and its architecture walk (weight manifest) equals the reference's state_dict layout.  No GPU needed."""
import os
import re

import numpy as np
import pytest
import torch

from oracle.arch import NetConfig, state_dict_manifest
from sgmse_b200 import _lib, Engine, EngineConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "sgmse_b200.h")).read()
    declared = set(re.findall(r"\b(sgmse_b200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.sgmse_b200_version()


CASES = [
    (NetConfig.ncsnpp(), EngineConfig.ncsnpp_16k()),
    (NetConfig.ncsnpp_48k(), EngineConfig.ncsnpp_48k()),
    (NetConfig.ncsnpp(nf=16, ch_mult=(1, 2, 2), image_size=64, attn_resolutions=(16,)),
     EngineConfig(nf=16, ch_mult=(1, 2, 2), image_size=64, attn_resolutions=(16,), n_fft=126, hop_length=32)),
    (NetConfig.ncsnpp_48k(nf=16, ch_mult=(1, 2, 2), image_size=64),
     EngineConfig.ncsnpp_48k(nf=16, ch_mult=(1, 2, 2), image_size=64, n_fft=126, hop_length=32)),
]


@pytest.mark.parametrize("ncfg,ecfg", CASES)
def test_manifest_equals_reference_state_dict_layout(ncfg, ecfg):
    eng = Engine(ecfg)
    man = eng.manifest()
    ref = [(k, int(np.prod(s))) for k, s in state_dict_manifest(ncfg)]
    assert man == ref
    assert eng.weights_numel() == sum(n for _, n in ref)
    eng.close()


def test_golden_state_dict_is_accepted(golden_dir):
    z = np.load(os.path.join(golden_dir, "ncsnpp_small.npz"))
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w/")}
    eng = Engine(CASES[2][1])
    blob = eng.flatten_state_dict(sd)
    assert blob.numel() == eng.weights_numel()
    with pytest.raises(KeyError):
        eng.flatten_state_dict({k: v for k, v in list(sd.items())[1:]})
    eng.close()


def test_errors_are_reported_not_thrown():
    lib = _lib.load()
    import ctypes as C
    cfg = EngineConfig(nf=12).to_c()          # nf must be a multiple of 8
    h = C.c_void_p()
    assert lib.sgmse_b200_create(C.byref(cfg), C.byref(h)) != 0
    assert b"nf=12" in lib.sgmse_b200_last_error()
    with pytest.raises(ValueError, match="Predictor with name 'bogus' unknown."):
        Engine(CASES[2][1]).sampler_struct(predictor="bogus")


def test_no_cpu_fallback():
    eng = Engine(CASES[2][1])
    with pytest.raises(RuntimeError, match="no CPU path"):
        eng.score(torch.zeros(1, 1, 64, 64, dtype=torch.complex64), torch.zeros(1, 1, 64, 64, dtype=torch.complex64),
                  torch.ones(1))
    eng.close()

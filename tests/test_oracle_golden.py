"""Pins the oracle (oracle/) to outputs of the unmodified reference (tests/golden/*.npz,
produced by oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import ncsnpp, sde as sde_mod, spec as spec_mod, pipeline
from oracle.arch import NetConfig, state_dict_manifest

SMALL = dict(nf=16, ch_mult=(1, 2, 2), image_size=64, num_res_blocks=2)
CASES = {
    "ncsnpp_small": NetConfig.ncsnpp(attn_resolutions=(16,), **SMALL),
    "ncsnpp48k_small": NetConfig.ncsnpp_48k(**SMALL),
}


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w/")}
    return z, sd


def _rel(a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("name", list(CASES))
def test_manifest_matches_reference_state_dict(golden_dir, name):
    z, sd = _load(golden_dir, name)
    man = state_dict_manifest(CASES[name])
    assert [k for k, _ in man] == list(sd.keys())
    for k, s in man:
        assert tuple(sd[k].shape) == s


@pytest.mark.parametrize("name", list(CASES))
def test_forward_and_score(golden_dir, name):
    z, sd = _load(golden_dir, name)
    cfg = CASES[name]
    x, y, t = (torch.from_numpy(z[k]) for k in ("x", "y", "t"))
    out = ncsnpp.forward(sd, cfg, torch.cat([x, y], 1), t)
    assert _rel(out, z["dnn_out"]) < 1e-5
    assert _rel(ncsnpp.score(sd, cfg, x, y, t), z["score"]) < 1e-5


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("pred,corr", [("reverse_diffusion", "ald"), ("reverse_diffusion", "langevin"),
                                       ("none", "ald"), ("reverse_diffusion", "none")])
def test_pc_sampler(golden_dir, name, pred, corr):
    z, sd = _load(golden_dir, name)
    cfg = CASES[name]
    y = torch.from_numpy(z["y"])
    N = 3
    draws = sde_mod.make_noise(tuple(y.shape), sde_mod.n_noise_draws(N, pred, corr, 1), seed=7)
    with torch.no_grad():
        smp, nfe = sde_mod.pc_sample(lambda a, b, c: ncsnpp.score(sd, cfg, a, b, c), y, sde_mod.OUVE(), N=N,
                                     predictor=pred, corrector=corr, corrector_steps=1, snr=0.5, noise=draws)
    assert nfe == int(z[f"nfe_{pred}_{corr}"])
    assert _rel(smp, z[f"pc_{pred}_{corr}"]) < 2e-4


@pytest.mark.parametrize("name", list(CASES))
def test_enhance_chain(golden_dir, name):
    z, sd = _load(golden_dir, name)
    cfg = CASES[name]
    scfg = spec_mod.SpecConfig(n_fft=126, hop_length=32)
    wav = torch.from_numpy(z["wav"])
    B, N = wav.shape[0], 3
    draws = sde_mod.make_noise((B, 1, 64, 64), sde_mod.n_noise_draws(N, "reverse_diffusion", "ald", 1), seed=11)
    xh, X, Y = pipeline.enhance(sd, cfg, scfg, sde_mod.OUVE(), wav, draws, N=N, return_spec=True)
    assert _rel(Y, z["Y"]) < 1e-4
    assert _rel(xh, z["enh"]) < 1e-3


def test_fir_and_stft_ops(golden_dir):
    z = np.load(os.path.join(golden_dir, "ops.npz"))
    x = torch.from_numpy(z["fir_x"])
    assert _rel(ncsnpp.fir_down2(x), z["fir_down"]) < 1e-6
    assert _rel(ncsnpp.fir_up2(x), z["fir_up"]) < 1e-6
    scfg = spec_mod.SpecConfig(n_fft=126, hop_length=32)
    wav = torch.from_numpy(z["wav"])
    S = spec_mod.stft(wav, scfg)
    assert _rel(S, z["stft"]) < 1e-5
    assert _rel(spec_mod.spec_fwd(S, scfg), z["spec_fwd"]) < 1e-5
    assert _rel(spec_mod.spec_back(spec_mod.spec_fwd(S, scfg), scfg), z["spec_back"]) < 1e-4
    assert _rel(spec_mod.istft(torch.from_numpy(z["stft"]), scfg, 2000), z["istft"]) < 1e-5
    s48 = spec_mod.SpecConfig.cfg_48k()
    w48 = torch.from_numpy(z["wav48"])
    S48 = spec_mod.stft(w48, s48)
    assert _rel(S48, z["stft48"]) < 1e-5
    assert _rel(spec_mod.spec_fwd(S48, s48), z["spec_fwd48"]) < 1e-5
    assert _rel(spec_mod.istft(torch.from_numpy(z["stft48"]), s48, 6000), z["istft48"]) < 1e-5

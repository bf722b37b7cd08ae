"""Host logic of the drop-in layer (sgmse_b200/api.py) with stand-in model and engine objects: which engine call each
rebound ScoreModel method makes, with which arguments, and which reference error it reproduces.  No GPU, no reference
checkout (the live-reference variants are in tests/test_oracle_vs_reference.py)."""
import types

import numpy as np
import pytest
import torch

import sgmse_b200
from sgmse_b200 import api


class OUVESDE:                       # the class NAME is what the drop-in layer dispatches on (model.py:441,450)
    def __init__(self, sampler_type="pc", N=30):
        self.sampler_type, self.N = sampler_type, N


class SBVESDE:
    def __init__(self, sampler_type="ode", N=50):
        self.sampler_type, self.N = sampler_type, N


class FakeEngine:
    def __init__(self, backbone="ncsnpp"):
        self.cfg = types.SimpleNamespace(backbone=backbone, t_eps=0.03, sr=16000)
        self.device = "cpu"
        self.calls = []

    def enhance(self, wav, **kw):
        self.calls.append(("enhance", tuple(wav.shape), kw))
        return wav * 2

    def analysis(self, wav, pad_mode="zero_pad"):
        self.calls.append(("analysis", tuple(wav.shape), pad_mode))
        return torch.zeros(wav.shape[0], 1, 4, 64, dtype=torch.complex64), wav.abs().amax(dim=1)

    def synthesis(self, X, norm, length):
        self.calls.append(("synthesis", tuple(X.shape), length))
        return torch.ones(X.shape[0], length) * norm[:, None]

    def enhance_ode(self, wav, denoise=True, **kw):
        self.calls.append(("enhance_ode", tuple(wav.shape), denoise, kw))
        if denoise:
            raise TypeError("ReverseDiffusionPredictor.update_fn() missing 1 required positional argument: 'stepsize'")
        return wav * 2, [38] * wav.shape[0]

    def ode_sample(self, y, prior_noise=None, denoise=True, **kw):
        self.calls.append(("ode_sample", tuple(y.shape), denoise, prior_noise is not None, kw))
        if denoise:
            raise TypeError("ReverseDiffusionPredictor.update_fn() missing 1 required positional argument: 'stepsize'")
        return y + 1, 38

    def pc_sample(self, y, noise=None, **kw):
        self.calls.append(("pc_sample", tuple(y.shape), kw))
        return y, kw["N"] * 2


def make_model(sde):
    m = types.SimpleNamespace(sde=sde, sr=16000)
    m.__dict__["_orig"] = True
    return m


def test_install_binds_the_ode_sampler_only_where_the_engine_has_it():
    m = make_model(OUVESDE())
    eng = FakeEngine()
    sgmse_b200.install(m, engine=eng, rebind_forward=False)
    assert {"get_pc_sampler", "get_ode_sampler", "get_sb_sampler", "enhance"} <= set(m.__dict__)
    sgmse_b200.uninstall(m)
    assert "get_ode_sampler" not in m.__dict__ and "enhance" not in m.__dict__
    m2 = make_model(SBVESDE())
    sgmse_b200.install(m2, engine=FakeEngine("ncsnpp_v2"), rebind_forward=False)
    assert "get_ode_sampler" not in m2.__dict__          # the reference's own torch path stays in place
    sgmse_b200.uninstall(m2)


def test_get_ode_sampler_arguments_and_minibatch_loop():
    m = make_model(OUVESDE())
    eng = FakeEngine()
    sgmse_b200.install(m, engine=eng, rebind_forward=False)
    y = torch.zeros(5, 1, 4, 64, dtype=torch.complex64)
    with pytest.raises(TypeError, match="stepsize"):     # the reference's default (denoise=True) fails the same way
        m.get_ode_sampler(y)()
    x, nfe = m.get_ode_sampler(y, denoise=False, rtol=1e-3, seed=11)()
    name, shape, denoise, has_noise, kw = eng.calls[-1]
    assert (name, shape, denoise, has_noise) == ("ode_sample", (5, 1, 4, 64), False, False)
    assert kw["rtol"] == 1e-3 and kw["atol"] == 1e-5 and kw["eps"] == 0.03 and kw["method"] == "RK45" and kw["seed"] == 11
    assert nfe == 38 and torch.equal(x, y + 1)
    # minibatch: one ODE system per slice, noise ids continue across slices, nfe is a list (model.py:380-390)
    noise = torch.zeros_like(y)
    eng.calls.clear()
    x, ns = m.get_ode_sampler(y, minibatch=2, denoise=False, noise=noise, seed=3)()
    assert ns == [38, 38, 38] and x.shape == y.shape
    assert [c[1][0] for c in eng.calls] == [2, 2, 1] and [c[4]["utt_offset"] for c in eng.calls] == [0, 2, 4]
    assert all(c[3] for c in eng.calls)
    sgmse_b200.uninstall(m)


def test_enhance_dispatch_follows_model_py():
    """model.py:441-454: OUVESDE + 'pc' -> PC sampler, OUVESDE + 'ode' -> ODE sampler, anything else -> ValueError."""
    wav = 0.5 * torch.ones(1, 1000)
    eng = FakeEngine()
    m = make_model(OUVESDE("pc"))
    sgmse_b200.install(m, engine=eng, rebind_forward=False)
    out = m.enhance(wav, N=7, corrector_steps=2, snr=0.33, seed=1)
    assert isinstance(out, np.ndarray) and out.shape == (1000,)
    name, shape, kw = eng.calls[-1]
    assert name == "enhance" and kw["N"] == 7 and kw["corrector_steps"] == 2 and kw["snr"] == 0.33
    xh, nfe, rtf = m.enhance(wav, N=7, corrector_steps=2, timeit=True)
    assert nfe == 7 * 3 and rtf > 0

    m.sde.sampler_type = "ode"
    eng.calls.clear()
    with pytest.raises(TypeError, match="stepsize"):     # enhance() forwards kwargs only: denoise stays True (model.py:447)
        m.enhance(wav)
    xh, nfe, rtf = m.enhance(wav, denoise=False, rtol=1e-2, timeit=True, seed=4)
    name, shape, denoise, kw = eng.calls[-1]
    assert (name, shape, denoise) == ("enhance_ode", (1, 1000), False) and nfe == 38 and xh.shape == (1000,)
    assert kw["rtol"] == 1e-2 and kw["seed"] == 4 and kw["pad_mode"] == "zero_pad"

    m.sde.sampler_type = "bogus"
    with pytest.raises(ValueError, match="Invalid sampler type"):
        m.enhance(wav)
    sgmse_b200.uninstall(m)


def test_forward_under_no_grad_goes_to_the_engine():
    """install(rebind_forward="no_grad"): training_step keeps autograd, the `_step` of validation_step (model.py:189-198,
    257-258, run by Lightning under torch.no_grad()) evaluates the network on the engine."""
    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(()))
            self.sde = OUVESDE()

        def forward(self, x_t, y, t):
            return self.w * x_t

    eng = FakeEngine()
    eng.model_forward = lambda x_t, y, t: (eng.calls.append(("model_forward", tuple(x_t.shape))), x_t * 3)[1]
    m = M()
    sgmse_b200.install(m, engine=eng, rebind_forward="no_grad")
    x = torch.ones(2, 1, 4, 4)
    out = m(x, x, torch.ones(2))
    assert out.requires_grad and torch.equal(out, x) and not eng.calls          # autograd path untouched
    with torch.no_grad():
        out = m(x, x, torch.ones(2))
    assert torch.equal(out, 3 * x) and eng.calls == [("model_forward", (2, 1, 4, 4))]
    sgmse_b200.uninstall(m)
    with torch.no_grad():
        assert torch.equal(m(x, x, torch.ones(2)), x)


def test_installing_twice_starts_from_the_models_own_methods():
    m = make_model(OUVESDE())
    e1, e2 = FakeEngine(), FakeEngine()
    sgmse_b200.install(m, engine=e1, rebind_forward=False)
    sgmse_b200.install(m, engine=e2, rebind_forward=False)
    assert m._sgmse_b200_engine is e2
    m.get_ode_sampler(torch.zeros(1, 1, 4, 64, dtype=torch.complex64), denoise=False)()
    assert e2.calls and not e1.calls
    sgmse_b200.uninstall(m)
    assert not any(k in m.__dict__ for k in ("get_pc_sampler", "get_ode_sampler", "get_sb_sampler", "enhance", "_sgmse_b200_engine"))
    sgmse_b200.uninstall(m)                              # idempotent


def test_refresh_hooks_the_ema_swap_in_train_not_only_eval():
    """ScoreModel.train(mode, no_ema) is where the reference swaps EMA weights (model.py:111-122).  Lightning + DDP never
    call ScoreModel.eval(): on_validation_model_eval() runs nn.Module.eval() on the wrapper, which recurses as
    child.train(False).  The refresh must fire on that route, on model.eval() (which delegates to train, model.py:124-125)
    and on model.train(False) -- and not on model.train(True)."""
    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.dnn = torch.nn.Linear(2, 2)
            self.sde = OUVESDE()
            self.swaps = []

        def train(self, mode=True, no_ema=False):
            res = super().train(mode)
            self.swaps.append(bool(mode))
            return res

        def eval(self, no_ema=False):
            return self.train(False, no_ema=no_ema)

    eng = FakeEngine()
    loads = []
    eng.load_state_dict = lambda sd, on_device=False: loads.append(sorted(sd))
    m = M()
    sgmse_b200.install(m, engine=eng, rebind_forward=False, refresh_on_eval=True)
    wrapper = torch.nn.Sequential(m)                 # stands for DistributedDataParallel(model)
    torch.nn.Module.eval(wrapper)                    # what Lightning's on_validation_model_eval does
    assert len(loads) == 1 and m.swaps[-1] is False
    wrapper.train()                                  # back to training: no refresh
    assert len(loads) == 1 and m.swaps[-1] is True
    m.eval()
    assert len(loads) == 2
    m.train(False)
    assert len(loads) == 3
    sgmse_b200.uninstall(m)
    m.eval()
    assert len(loads) == 3 and "train" not in m.__dict__


def test_enhance_forwards_sampler_kwargs_and_refuses_what_it_cannot_honour():
    """The reference forwards enhance(**kwargs) into get_pc_sampler (model.py:443-445): denoise / probability_flow take
    effect there, so they must here; eps other than t_eps and unknown keywords raise instead of being dropped."""
    wav = 0.5 * torch.ones(1, 1000)
    eng = FakeEngine()
    m = make_model(OUVESDE("pc"))
    sgmse_b200.install(m, engine=eng, rebind_forward=False)
    m.enhance(wav, denoise=False, probability_flow=True, seed=2)
    kw = eng.calls[-1][2]
    assert kw["denoise"] is False and kw["probability_flow"] is True
    m.enhance(wav, seed=2)
    kw = eng.calls[-1][2]
    assert kw["denoise"] is True and kw["probability_flow"] is False
    m.enhance(wav, eps=0.03, intermediate=False)     # what the reference itself passes (model.py:444)
    with pytest.raises(NotImplementedError, match="eps"):
        m.enhance(wav, eps=0.05)
    with pytest.raises(NotImplementedError, match="bogus_kw"):
        m.enhance(wav, bogus_kw=1)
    with pytest.raises(NotImplementedError, match="intermediate"):
        m.enhance(wav, intermediate=True)
    sgmse_b200.uninstall(m)

"""Import shim for the *live* reference (build container only).

``/root/reference`` is mounted read-only in the build container and does not
exist on the GPU box.  ``sgmse.model`` / ``sgmse.data_module`` / ``sgmse.util.other``
import six packages that are not installed (pytorch_lightning, torch_ema, librosa,
pesq, pystoi, torch_pesq); none carries hot-path arithmetic, so
inert stand-ins are injected into ``sys.modules`` before import.

Used by oracle/make_golden.py and tests/test_oracle_vs_reference.py.
TEST INFRASTRUCTURE – see oracle/__init__.py.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
from typing import List

import torch

REFERENCE_ROOT = os.environ.get("SGMSE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sgmse"))


def _install_stubs():
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(torch.nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

        class LightningDataModule:
            def __init__(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        pl.LightningDataModule = LightningDataModule
        sys.modules["pytorch_lightning"] = pl
    if "torch_ema" not in sys.modules:
        m = types.ModuleType("torch_ema")

        class ExponentialMovingAverage:
            def __init__(self, params, decay):
                self.collected_params = None

            def update(self, *a):
                pass

            def store(self, *a):
                pass

            def copy_to(self, *a):
                pass

            def restore(self, *a):
                pass

            def to(self, *a, **k):
                pass

            def state_dict(self):
                return {}

            def load_state_dict(self, *a):
                pass

        m.ExponentialMovingAverage = ExponentialMovingAverage
        sys.modules["torch_ema"] = m
    for name, attrs in (("librosa", ["resample"]), ("pesq", ["pesq"]), ("pystoi", ["stoi"]),
                        ("torch_pesq", ["PesqLoss"])):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, lambda *args, **kw: (_ for _ in ()).throw(RuntimeError(f"{name} stub")))
            sys.modules[name] = m


def import_reference():
    """Returns the reference's ``sgmse`` package (imported from REFERENCE_ROOT)."""
    if not reference_available():
        raise RuntimeError("reference not available at " + REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import sgmse  # noqa
    import sgmse.model, sgmse.data_module, sgmse.sdes, sgmse.sampling, sgmse.backbones  # noqa
    import sgmse.util.other  # noqa
    return sgmse


def make_score_model(backbone="ncsnpp", seed=0, **kw):
    """ScoreModel with ``init_scale=1.0`` by default, ``eval()``-ed (EMA swap is a no-op stub)."""
    sg = import_reference()
    from sgmse.model import ScoreModel
    from sgmse.data_module import SpecsDataModule
    args = dict(backbone=backbone, sde="ouve", data_module_cls=SpecsDataModule, base_dir="/nonexistent",
                theta=1.5, sigma_min=0.05, sigma_max=0.5, init_scale=1.0)
    args.update(kw)
    torch.manual_seed(seed)
    model = ScoreModel(**args)
    model.eval()
    return model


@contextlib.contextmanager
def injected_noise(draws: List[torch.Tensor]):
    """Monkeypatch ``torch.randn_like`` so the reference sampler consumes ``draws`` in call order."""
    it = iter(draws)
    orig = torch.randn_like

    def fake(x, *a, **k):
        z = next(it)
        assert z.shape == x.shape and z.dtype == x.dtype, (z.shape, x.shape, z.dtype, x.dtype)
        return z.to(x.device)

    torch.randn_like = fake
    try:
        yield
    finally:
        torch.randn_like = orig

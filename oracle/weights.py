"""Deterministic random-init weights in the reference's state_dict layout (oracle side).

The GPU box has neither /root/reference nor any checkpoint, so tests and the
benchmark need weights they can generate themselves.  The distributions follow
the reference initialisers (variance-scaling fan_avg uniform,
ncsnpp_utils/layers.py:54-91; NIN init_scale 0.1, layers.py:547-549;
GaussianFourierProjection W ~ N(0,1)*scale, layerspp.py:35-37) with
``init_scale=1.0`` for the zero-initialised convs (the default 0.
gives a constant-output net).  Biases and GroupNorm affine parameters are
perturbed away from their 0/1 defaults so that parity tests exercise them.

TEST INFRASTRUCTURE – see oracle/__init__.py.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .arch import NetConfig, state_dict_manifest


def _var_scale_uniform(shape, scale, g):
    # layers.py:62-83 with in_axis=1, out_axis=0, mode fan_avg, uniform
    rf = 1
    for d in shape[2:]:
        rf *= d
    fan_in, fan_out = shape[1] * rf, shape[0] * rf
    var = scale / ((fan_in + fan_out) / 2.0)
    return (torch.rand(shape, generator=g) * 2.0 - 1.0) * math.sqrt(3.0 * var)


def make_state_dict(cfg: NetConfig, seed: int = 0, perturb: float = 0.1) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape in state_dict_manifest(cfg):
        leaf = key.split(".")[-1]
        parent = key.split(".")[-2]
        if leaf == "W" and len(shape) == 1:                       # GaussianFourierProjection
            t = torch.randn(shape, generator=g) * cfg.fourier_scale
        elif leaf == "W":                                         # NIN
            scale = 0.1 if parent != "NIN_3" else (cfg.init_scale or 1e-10)
            t = _var_scale_uniform(shape, scale, g)
        elif leaf == "b":
            t = perturb * torch.randn(shape, generator=g)
        elif leaf == "weight" and len(shape) == 1:                # GroupNorm gamma
            t = 1.0 + perturb * torch.randn(shape, generator=g)
        elif leaf == "weight":
            t = _var_scale_uniform(shape, 1.0, g)
        elif leaf == "bias":
            t = perturb * torch.randn(shape, generator=g)
        else:
            raise KeyError(key)
        sd[key] = t.float().contiguous()
    return sd


def flatten_state_dict(sd: Dict[str, torch.Tensor], cfg: NetConfig) -> torch.Tensor:
    """fp32 blob in ``state_dict()`` order – the layout ``sgmse_b200_load_weights`` takes."""
    return torch.cat([sd[k].reshape(-1).float() for k, _ in state_dict_manifest(cfg)])

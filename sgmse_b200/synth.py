"""Synthetic (random-init) weights and inputs for benchmarking and smoke runs.

There is no network in the build/bench environment, hence no checkpoint: the benchmark uses
random-init weights of the named architecture and synthetic noisy speech (BASELINE.md §3,
SURVEY.md §8d).  Built from the engine's own weight manifest, independent of oracle/.
"""
from __future__ import annotations

import math

import torch


def synthetic_blob(engine, seed: int = 0) -> torch.Tensor:
    """fp32 CPU blob in manifest (= state_dict()) order: fan-in scaled uniform weights, small biases,
    GroupNorm affine ~ (1, 0) with a perturbation so that it is exercised."""
    g = torch.Generator().manual_seed(seed)
    man = engine.manifest()
    sizes = dict(man)
    parts = []
    for name, n in man:
        stem, leaf = name.rsplit(".", 1)
        if leaf == "W" and n == engine.cfg.nf and stem == "all_modules.0":     # Gaussian Fourier frequencies
            t = torch.randn(n, generator=g) * 16.0
        elif leaf in ("weight", "W"):
            n_bias = sizes[stem + (".b" if leaf == "W" else ".bias")]
            if n_bias == n:                                                  # GroupNorm gamma (1-D, like its bias)
                t = 1.0 + 0.1 * torch.randn(n, generator=g)
            else:                                                            # conv / linear / NIN
                fan_in = max(1, n // n_bias)
                t = (torch.rand(n, generator=g) * 2 - 1) * math.sqrt(3.0 / fan_in)
        else:                                                                # biases, GroupNorm beta
            t = 0.05 * torch.randn(n, generator=g)
        parts.append(t.float())
    return torch.cat(parts).contiguous()


def synthetic_speech(batch: int, length: int, seed: int = 1000, first: int = 0) -> torch.Tensor:
    """0.1 * randn waveforms; one generator per *global* utterance index so that any
    sharding of the batch sees the same utterances."""
    out = torch.empty(batch, length)
    for b in range(batch):
        g = torch.Generator().manual_seed(seed + first + b)
        out[b] = 0.1 * torch.randn(length, generator=g)
    return out

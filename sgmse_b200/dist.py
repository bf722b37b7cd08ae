"""Multi-GPU driver: one process per GPU (torch.distributed), utterances sharded across ranks.

The sampling loop needs no cross-GPU traffic (utterances are independent: GroupNorm is per sample,
ALD / reverse-diffusion are per sample, noise is keyed by the *global* utterance id), so the only
collectives are the one-off broadcast of the fp32 weight blob from rank 0 (NCCL over NVLink on GPUs)
and an optional gather of the enhanced waveforms.  SURVEY.md §8(e).

The reference has no inference parallelism of its own (only the sequential ``minibatch`` loop of
/root/reference/sgmse/model.py:354-368); this module is the replacement for running that loop on 8 GPUs.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous ceil(total/world) slices; trailing ranks may be empty."""
    per = (total + world - 1) // world
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def broadcast_weights(blob: Optional[torch.Tensor], numel: int, device: torch.device, src: int = 0) -> torch.Tensor:
    """Rank ``src`` passes the fp32 blob (any device); every rank returns it resident on ``device``."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_rank() == src:
            buf = blob.to(device=device, dtype=torch.float32).contiguous()
            assert buf.numel() == numel
        else:
            buf = torch.empty(numel, dtype=torch.float32, device=device)
        dist.broadcast(buf, src=src)
        return buf
    return blob.to(device=device, dtype=torch.float32).contiguous()


def enhance_sharded(enhance_fn: Callable[..., torch.Tensor], wav: torch.Tensor, gather: bool = True, **kw):
    """``wav`` [B, L] (identical on every rank).  Each rank enhances its slice with
    ``enhance_fn(wav_slice, utt_offset=lo, **kw)``; with ``gather`` every rank receives the full [B, L].

    Per-utterance results do not depend on the number of ranks (noise is keyed by the global utterance id; bit-equal to
    the single-process call, tools/nccl_invariance.py).  The one sampler for which that cannot hold is
    ``corrector='langevin'``: its step size is a mean over the batch (/root/reference/sgmse/sampling/correctors.py:50-52),
    so the utterances of a batch are coupled and sharding the batch would change every result -- refused for world > 1."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world > 1 and kw.get("corrector", "ald") == "langevin":
        raise ValueError("corrector='langevin' couples the utterances of a batch (batch-mean step size, correctors.py:50-52): "
                         "sharding the batch over ranks would change the results; run it on one rank or use 'ald'")
    B, L = wav.shape
    lo, hi = shard_range(B, rank, world)
    mine = enhance_fn(wav[lo:hi], utt_offset=lo, **kw) if hi > lo else wav.new_zeros((0, L))
    if world == 1 or not gather:
        return mine
    per = (B + world - 1) // world
    pad = mine.new_zeros((per, L))
    pad[: hi - lo] = mine
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat(parts, dim=0)[:B]

"""N>1 host logic on CPU: two gloo ranks shard a batch, broadcast a weight blob, gather results.
The per-utterance function stands in for the engine call (noise keyed by global utterance id)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sgmse_b200.dist import shard_range, broadcast_weights, enhance_sharded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def fake_enhance(wav, utt_offset=0, scale=2.0):
    # depends on the *global* utterance index exactly like the Philox keying of the engine
    idx = torch.arange(wav.shape[0], dtype=wav.dtype)[:, None] + utt_offset
    return wav * scale + idx


def _worker(rank, world, port, B, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        wav = torch.randn(B, 50, generator=g)
        blob = torch.arange(1000, dtype=torch.float32) if rank == 0 else None
        got_blob = broadcast_weights(blob, 1000, torch.device("cpu"))
        full = enhance_sharded(fake_enhance, wav, gather=True, scale=3.0)
        try:                                   # the batch-coupled corrector is refused when the batch is sharded (correctors.py:50-52)
            enhance_sharded(fake_enhance, wav, gather=True, corrector="langevin")
            refused = False
        except ValueError as e:
            refused = "langevin" in str(e)
        out_q.put((rank, got_blob.sum().item(), full, refused))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 4, 1])
def test_two_rank_sharding_matches_single_process(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    wav = torch.randn(B, 50, generator=g)
    want = fake_enhance(wav, 0, scale=3.0)
    for rank, blob_sum, full, refused in res:
        assert blob_sum == float(sum(range(1000)))
        assert torch.equal(full, want) and refused


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 16, 17):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d

"""Host logic of the batched file service (sgmse_b200/service.py) with a stand-in engine: bucketing by
padded frame count, batch cutting, noise-id assignment, per-clip front/back end, resampling hook.  No GPU."""
import os
import types

import pytest
import torch

from sgmse_b200.service import BatchedEnhancer, default_pad_mode, plan_batches


def padded_frames(L, hop=128):
    nT = 1 + L // hop                       # engine.cu: frames_of / padded_frames (util/other.py:76-90: multiple of 64)
    return (nT + 63) // 64 * 64


def test_plan_buckets_by_padded_frames_and_keeps_ids_contiguous():
    lengths = [64000, 8000, 64100, 8100, 16000, 7000, 63000]
    plan = plan_batches(lengths, padded_frames, max_batch=2)
    # 8000/8100/7000 -> 64 frames; 16000 -> 128; 63000/64000/64100 -> 512
    assert [tp for tp, _ in plan.batches] == [64, 64, 128, 512, 512]
    assert [idx for _, idx in plan.batches] == [[1, 3], [5], [4], [0, 2], [6]]
    assert sorted(plan.utt_id) == list(range(len(lengths)))
    for _, idx in plan.batches:             # a batch is a contiguous id range starting at its first clip
        assert [plan.utt_id[i] for i in idx] == list(range(plan.utt_id[idx[0]], plan.utt_id[idx[0]] + len(idx)))
    with pytest.raises(ValueError):
        plan_batches([10, 0], padded_frames, 4)
    with pytest.raises(ValueError):
        plan_batches([10], padded_frames, 0)


class FakeEngine:
    """Identity 'enhancement' that records what it was asked to do."""

    def __init__(self, backbone="ncsnpp", sr=16000, max_batch=3):
        self.cfg = types.SimpleNamespace(backbone=backbone, sr=sr, max_batch=max_batch)
        self.calls = []

    def padded_frames(self, L):
        return padded_frames(L)

    def analysis(self, wav, pad_mode="zero_pad"):
        self.calls.append(("analysis", tuple(wav.shape), pad_mode))
        B, L = wav.shape
        Y = torch.zeros(B, 1, 4, self.padded_frames(L), dtype=torch.complex64)
        Y[:, 0, 0, 0] = wav.sum(dim=1)      # a tag that survives the round trip
        return Y, wav.abs().amax(dim=1)

    def pc_sample(self, Y, seed=0, utt_offset=0, **kw):
        self.calls.append(("pc_sample", tuple(Y.shape), seed, utt_offset, tuple(sorted(kw.items()))))
        return Y + utt_offset, 60

    def synthesis(self, X, norm, length):
        self.calls.append(("synthesis", tuple(X.shape), length))
        return (X[:, 0, 0, 0].real[:, None] * torch.ones(1, length)) * norm[:, None]


def test_batched_enhancer_routes_every_clip_through_its_own_front_and_back_end():
    eng = FakeEngine(max_batch=2)
    waves = [torch.full((8000,), 0.5), torch.full((64000,), -0.25), torch.full((8100,), 0.1)]
    outs, ids = BatchedEnhancer(eng, device="cpu")(waves, seed=7, N=30, corrector="ald")
    assert ids == [0, 2, 1]                                  # 64-frame bucket first (clips 0 and 2), then 512
    assert [c[1] for c in eng.calls if c[0] == "pc_sample"] == [(2, 1, 4, 64), (1, 1, 4, 512)]
    assert [c[3] for c in eng.calls if c[0] == "pc_sample"] == [0, 2]
    assert all(c[2] == 7 and ("N", 30) in c[4] for c in eng.calls if c[0] == "pc_sample")
    assert [c[2] for c in eng.calls if c[0] == "analysis"] == ["zero_pad"] * 3
    for k, (w, o) in enumerate(zip(waves, outs)):
        assert o.shape == w.shape
        batch_first = {0: 0, 2: 0, 1: 2}[k]                  # utt_offset of the batch the clip travelled in
        assert torch.allclose(o, (w.sum() + batch_first) * w.abs().max() * torch.ones_like(w))


def test_pad_mode_and_resampling_follow_enhancement_py():
    assert default_pad_mode("ncsnpp") == "zero_pad" and default_pad_mode("ncsnpp_48k") == "reflection"
    assert default_pad_mode("ncsnpp_v2") == "reflection"
    eng = FakeEngine(backbone="ncsnpp_48k", sr=48000)
    enh = BatchedEnhancer(eng, device="cpu")
    with pytest.raises(ValueError):
        enh([torch.ones(16000)], sr=16000)
    outs, _ = enh([torch.ones(16000)], sr=16000, resample=lambda w, a, b: w.repeat_interleave(b // a))
    assert outs[0].numel() == 48000 and eng.calls[0] == ("analysis", (1, 48000), "reflection")


def test_langevin_corrector_is_sampled_clip_by_clip():
    """correctors.py:50-52: the Langevin step size is a batch mean, so clips sharing a batch are coupled; the reference's file
    loop (enhancement.py:58-103) never batches.  The service keeps that: one clip per sampler call, ids still contiguous."""
    eng = FakeEngine(max_batch=4)
    waves = [torch.full((8000,), 0.5), torch.full((8100,), 0.1), torch.full((7000,), 0.2)]
    outs, ids = BatchedEnhancer(eng, device="cpu")(waves, seed=1, corrector="langevin")
    pcs = [c for c in eng.calls if c[0] == "pc_sample"]
    assert [c[1][0] for c in pcs] == [1, 1, 1] and [c[3] for c in pcs] == [0, 1, 2] and ids == [0, 1, 2]
    eng.calls.clear()
    BatchedEnhancer(eng, device="cpu")(waves, seed=1, corrector="ald")
    assert [c[1][0] for c in eng.calls if c[0] == "pc_sample"] == [3]


def test_directory_enhancer_mirrors_the_enhancement_py_file_loop(tmp_path):
    """sgmse_b200/files.py: glob order and relative output paths of enhancement.py:38-60,100-103, resampling to the model's rate
    (:65-66), reader / GPU / writer windows, noise ids continuing across windows; WAV I/O through scipy (no soundfile here)."""
    import numpy as np
    from scipy.io import wavfile
    from sgmse_b200.files import DirectoryEnhancer, list_audio_files, read_audio
    src, dst = tmp_path / "noisy", tmp_path / "enhanced"
    (src / "spk1").mkdir(parents=True)
    rng = np.random.default_rng(0)
    specs = {"b.wav": (16000, 8000, np.float32), "a.wav": (16000, 8100, np.int16), os.path.join("spk1", "c.wav"): (8000, 4000, np.float32)}
    for name, (sr, n, dt) in specs.items():
        x = 0.1 * rng.standard_normal(n)
        wavfile.write(str(src / name), sr, (x * 32767).astype(np.int16) if dt == np.int16 else x.astype(np.float32))
    files = list_audio_files(str(src))
    assert [os.path.relpath(f, str(src)) for f in files] == ["a.wav", "b.wav", os.path.join("spk1", "c.wav")]
    y, sr = read_audio(files[0])
    assert y.shape == (1, 8100) and sr == 16000 and y.dtype == torch.float32 and y.abs().max() < 1.0
    eng = FakeEngine(max_batch=2)
    outs, ids = DirectoryEnhancer(eng, window=2, io_workers=2, device="cpu")(str(src), str(dst), seed=5, N=30)
    assert [os.path.relpath(o, str(dst)) for o in outs] == ["a.wav", "b.wav", os.path.join("spk1", "c.wav")]
    assert sorted(ids) == [0, 1, 2] and ids[2] == 2                     # window 1 = {a, b}, window 2 = {c}: ids continue
    lens = []
    for o in outs:
        sr_o, x = wavfile.read(o)
        assert sr_o == 16000 and x.dtype == np.float32
        lens.append(len(x))
    assert lens == [8100, 8000, 8000]                                   # c.wav: 4000 samples at 8 kHz -> 8000 at the model's 16 kHz
    assert [c[1] for c in eng.calls if c[0] == "analysis"] == [(1, 8100), (1, 8000), (1, 8000)] or \
           sorted(c[1] for c in eng.calls if c[0] == "analysis") == [(1, 8000), (1, 8000), (1, 8100)]
    with pytest.raises(ValueError, match="mono"):
        wavfile.write(str(src / "st.wav"), 16000, np.zeros((100, 2), np.float32))
        DirectoryEnhancer(eng, device="cpu")(str(src), str(dst))

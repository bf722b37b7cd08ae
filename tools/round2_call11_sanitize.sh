#!/bin/bash
# compute-sanitizer memcheck over one full-size evaluation (conv_tc6 strip producers, tcgen05 attention, every small kernel) and a small
# sampler run (graphs off), device-side weight packing included (profile_forward loads a CUDA blob).
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/profile_forward.py --batch 1 --evals 1 > gpurun_out/sanitize_forward.log 2>&1
tail -6 gpurun_out/sanitize_forward.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python - > gpurun_out/sanitize_sampler.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, ".")
from sgmse_b200 import Engine, EngineConfig
from sgmse_b200.synth import synthetic_blob, synthetic_speech
eng = Engine(EngineConfig(nf=64, ch_mult=(1, 2, 2), image_size=64, attn_resolutions=(16,), num_res_blocks=1, n_fft=126, hop_length=32,
                          mode="fp16_tc", max_batch=2, use_graphs=False))
eng.load_blob(synthetic_blob(eng, 0).cuda())
wav = synthetic_speech(2, 4000).cuda()
out = eng.enhance(wav, N=2, seed=3)
torch.cuda.synchronize()
print("sampler ok", bool(torch.isfinite(out).all()), eng.counter("tc_convs_last_forward"))
PY
tail -6 gpurun_out/sanitize_sampler.log

#!/bin/bash
# First GPU call of round 2: everything round 1 could not run, in order of importance.  Build both libraries
# in the build container first (python -m sgmse_b200.build && python -m sgmse_b200.build --pdl), then
#   gpurun --timeout 2400 -- 'bash tools/round2_first_call.sh'
# Every step writes its own log under gpurun_out/; a failing step does not stop the next one.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -s --maxfail=20 > gpurun_out/gpu_tests_full.log 2>&1; tail -120 gpurun_out/gpu_tests_full.log > gpurun_out/gpu_tests.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
SGMSE_B200_CANDIDATES=1 timeout 900 python -m pytest tests/test_gpu_zz_next_rows.py -q -m gpu -s -k round2_candidate 2>&1 | tail -30 > gpurun_out/gpu_tests_candidates.log
timeout 600 python tools/check_candidates.py > gpurun_out/candidates.log 2>&1
timeout 900 python tools/ab_forward.py fir_variant=2 outconv_variant=3 inconv_variant=2 attn_variant=2 combine_variant=1 tc1_narrow=1 gn_self=1 gnfin_variant=1 \
    fir_variant=2,outconv_variant=3,inconv_variant=2,attn_variant=2,combine_variant=1,tc1_narrow=1,gn_self=1,gnfin_variant=1 > gpurun_out/ab_small.log 2>&1
# ablations of conv_tc6 (twin library; outputs are wrong on purpose, only the conv time matters): 1 = weights loaded once per CTA,
# 2 = producers copy instead of GroupNorm+SiLU, 3 = both; tc_variant=6 = TMA-fed operands
SGMSE_B200_PDL=1 timeout 900 python tools/ab_forward.py tc6_ablate=1 tc6_ablate=2 tc6_ablate=3 tc_variant=6 tc_variant=6,tc6_ablate=1 > gpurun_out/ab_ablate.log 2>&1
SGMSE_B200_PDL=1 timeout 900 python tools/check_pdl.py > gpurun_out/pdl.log 2>&1
SGMSE_B200_PDL=1 timeout 600 python bench.py --steps 3 --warmup 3 --opt pdl=1 --no-cpu-baseline > gpurun_out/bench_c2_pdl.json 2> gpurun_out/bench_c2_pdl.err
timeout 600 python bench.py --steps 3 --warmup 3 --lanes 2 --no-cpu-baseline --no-roofline > gpurun_out/bench_c2_lanes2.json 2> gpurun_out/bench_c2_lanes2.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-roofline --opt fir_variant=2 --opt outconv_variant=3 --opt inconv_variant=2 --opt attn_variant=2 --opt combine_variant=1 --opt tc1_narrow=1 --opt gn_self=1 --opt gnfin_variant=1 > gpurun_out/bench_c2_candidates.json 2> gpurun_out/bench_c2_candidates.err
(nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_bench tools/mufu_bench.cu && timeout 120 /tmp/mufu_bench) > gpurun_out/mufu.log 2>&1
timeout 600 python bench.py --config 3 --steps 2 --warmup 3 --no-roofline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 900 python bench.py --config 4 --steps 2 --warmup 3 --no-roofline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 900 python tools/bench_ode.py --batch 16 --rtol 1e-2 --atol 1e-2 > gpurun_out/ode_b16.log 2>&1
timeout 500 bash tools/ncu_l2_ridge.sh > gpurun_out/l2_ridge.log 2>&1
tail -3 gpurun_out/gpu_tests.log; cat gpurun_out/bench_c2.json; tail -8 gpurun_out/candidates.log; tail -9 gpurun_out/ab_small.log; tail -6 gpurun_out/pdl.log; cat gpurun_out/mufu.log

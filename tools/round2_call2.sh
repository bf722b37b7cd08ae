#!/bin/bash
# Second GPU call of round 2: new parity tests, error decomposition, warp-role priority experiment, reference CPU arm on the box.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -s --maxfail=20 > gpurun_out/gpu_tests_full2.log 2>&1; tail -40 gpurun_out/gpu_tests_full2.log > gpurun_out/gpu_tests2.log
timeout 600 python tools/parity_decompose.py > gpurun_out/parity_decompose.log 2>&1
timeout 900 python tools/ab_forward.py tc6_roles=1 tc6_roles=1,tc6_rings=1 tc_variant=6 tc_variant=9,tc6_roles=1 tc_variant=10,tc6_roles=1 \
   tc6_roles=1,fir_variant=2,outconv_variant=3,inconv_variant=2,attn_variant=2,combine_variant=1,tc1_narrow=1,gn_self=1,gnfin_variant=1 > gpurun_out/ab_roles.log 2>&1
M=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__inst_executed_pipe_xu.sum,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed
timeout 400 ncu --clock-control none -k regex:conv_tc6 -s 54 -c 8 --metrics $M --csv --log-file gpurun_out/roles0.csv python tools/profile_forward.py --batch 16 --evals 2 > gpurun_out/roles0.log 2>&1
timeout 400 ncu --clock-control none -k regex:conv_tc6 -s 54 -c 8 --metrics $M --csv --log-file gpurun_out/roles1.csv python tools/profile_forward.py --batch 16 --evals 2 --opt tc6_roles=1 > gpurun_out/roles1.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt tc6_roles=1 > gpurun_out/bench_c2_roles.json 2> gpurun_out/bench_c2_roles.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -5 gpurun_out/gpu_tests2.log; cat gpurun_out/parity_decompose.log; cat gpurun_out/ab_roles.log; cat gpurun_out/bench_c2_roles.json; cat gpurun_out/bench_ref.json

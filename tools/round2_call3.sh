#!/bin/bash
# Third GPU call of round 2: tightened tolerances, tcgen05 attention bring-up, lean conv_tc6 producers, ncu source-level capture of
# the fused convolution, batch sweep on one GPU, launch list of a batch-1 evaluation, DRAM traffic of the dominant launch.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -s --maxfail=20 > gpurun_out/gpu_tests_full3.log 2>&1; tail -40 gpurun_out/gpu_tests_full3.log > gpurun_out/gpu_tests3.log
timeout 300 python tools/check_attention.py > gpurun_out/check_attention.log 2>&1
timeout 900 python tools/ab_forward.py tc6_lean=1 tc6_lean=1,tc6_roles=1 attn_variant=5 tc6_lean=1,attn_variant=5,fir_variant=2,outconv_variant=3,inconv_variant=2,combine_variant=1,tc1_narrow=1,gn_self=1,gnfin_variant=1 > gpurun_out/ab_lean.log 2>&1
M=gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --clock-control none -k regex:conv_tc6 -s 54 -c 4 --metrics $M --csv --log-file gpurun_out/lean1.csv python tools/profile_forward.py --batch 16 --evals 2 --opt tc6_lean=1 > gpurun_out/lean1.log 2>&1
timeout 300 ncu --clock-control none -k regex:conv_tc6 -s 54 -c 1 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/conv_traffic.csv python tools/profile_forward.py --batch 16 --evals 2 > gpurun_out/conv_traffic.log 2>&1
python tools/make_conv_traffic.py gpurun_out/conv_traffic.csv > gpurun_out/r02_conv_traffic.json 2> gpurun_out/conv_traffic.err
timeout 400 ncu --set full --import-source on --clock-control none -k regex:conv_tc6 -s 54 -c 1 -f -o gpurun_out/tc6_fused python tools/profile_forward.py --batch 16 --evals 2 > gpurun_out/tc6_fused.log 2>&1
timeout 400 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_b1.csv python tools/profile_forward.py --batch 1 --evals 2 > gpurun_out/launches_b1.log 2>&1
timeout 400 ncu --clock-control none -k regex:attention --metrics gpu__time_duration.sum --csv --log-file gpurun_out/attn_time.csv python tools/profile_forward.py --batch 16 --evals 1 --opt attn_variant=5 > gpurun_out/attn_time5.log 2>&1
timeout 400 ncu --clock-control none -k regex:attention --metrics gpu__time_duration.sum --csv --log-file gpurun_out/attn_time0.csv python tools/profile_forward.py --batch 16 --evals 1 > gpurun_out/attn_time0.log 2>&1
timeout 900 python tools/sweep_bench.py --batches 1,2,4,8,16,32,64,128 --steps 2 --warmup 2 > gpurun_out/sweep_g1.jsonl 2> gpurun_out/sweep_g1.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt tc6_lean=1 > gpurun_out/bench_c2_lean.json 2> gpurun_out/bench_c2_lean.err
tail -6 gpurun_out/gpu_tests3.log; cat gpurun_out/check_attention.log; cat gpurun_out/ab_lean.log; cat gpurun_out/sweep_g1.jsonl | cut -c1-300; cut -c1-400 gpurun_out/bench_c2_lean.json; ls -la gpurun_out/tc6_fused.ncu-rep

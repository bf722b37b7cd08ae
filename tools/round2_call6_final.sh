#!/bin/bash
# Verification call with the round-2 defaults: whole GPU suite, smoke, benches of BASELINE configs 2 / 3 / 4, launch list and DRAM
# traffic of the current build, ncu --set full of the dominant launch, batch sweep on one GPU.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -s --maxfail=30 > gpurun_out/gpu_tests_full6.log 2>&1; tail -40 gpurun_out/gpu_tests_full6.log > gpurun_out/gpu_tests6.log
python __graft_entry__.py smoke > gpurun_out/smoke6.log 2>&1
timeout 300 ncu --clock-control none -k regex:conv_tc6 -s 54 -c 1 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/conv_traffic.csv python tools/profile_forward.py --batch 16 --evals 2 > gpurun_out/conv_traffic.log 2>&1
python tools/make_conv_traffic.py gpurun_out/conv_traffic.csv > gpurun_out/r02_conv_traffic.json 2> gpurun_out/conv_traffic.err
cp gpurun_out/r02_conv_traffic.json profiles/r02_conv_traffic.json
timeout 900 python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/bench6_c2.json 2> gpurun_out/bench6_c2.err
timeout 600 python bench.py --config 3 --steps 4 --warmup 3 > gpurun_out/bench6_c3.json 2> gpurun_out/bench6_c3.err
timeout 900 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/bench6_c4.json 2> gpurun_out/bench6_c4.err
timeout 400 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_b16.csv python tools/profile_forward.py --batch 16 --evals 2 > gpurun_out/launches_b16.log 2>&1
M=gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --clock-control none -k regex:conv_tc6 -s 54 -c 54 --metrics $M --csv --log-file gpurun_out/convs6.csv python tools/profile_forward.py --batch 16 --evals 2 > gpurun_out/convs6.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:conv_tc6 -s 54 -c 1 -f -o gpurun_out/tc6_final python tools/profile_forward.py --batch 16 --evals 2 > gpurun_out/tc6_final.log 2>&1
timeout 900 python tools/sweep_bench.py --batches 1,2,4,8,16,32,64,128 --steps 2 --warmup 2 > gpurun_out/sweep6_g1.jsonl 2> gpurun_out/sweep6_g1.err
tail -8 gpurun_out/gpu_tests6.log; cat gpurun_out/smoke6.log | tail -2; for f in gpurun_out/bench6_c*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], (d.get('roofline') or {}).get('frac'))"; done; cut -c1-200 gpurun_out/sweep6_g1.jsonl

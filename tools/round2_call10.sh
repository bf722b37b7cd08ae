#!/bin/bash
# Product library without the superseded producer forms: whole GPU suite + smoke + bench; the lab twin once through the variant tests.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -s --maxfail=30 > gpurun_out/gpu_tests_full10.log 2>&1; tail -30 gpurun_out/gpu_tests_full10.log > gpurun_out/gpu_tests10.log
python __graft_entry__.py smoke > gpurun_out/smoke10.log 2>&1
SGMSE_B200_PDL=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_next_rows.py -q -m gpu -s -k "variants_agree or round1_kernels" > gpurun_out/gpu_tests_lab10.log 2>&1
timeout 300 ncu --clock-control none -k regex:conv_tc6 -s 54 -c 1 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/conv_traffic.csv python tools/profile_forward.py --batch 16 --evals 2 > gpurun_out/conv_traffic.log 2>&1
python tools/make_conv_traffic.py gpurun_out/conv_traffic.csv > gpurun_out/r02_conv_traffic.json 2> gpurun_out/conv_traffic.err
cp gpurun_out/r02_conv_traffic.json profiles/r02_conv_traffic.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench10_c2.json 2> gpurun_out/bench10_c2.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench10_ref.json 2> gpurun_out/bench10_ref.err
tail -4 gpurun_out/gpu_tests10.log; tail -1 gpurun_out/smoke10.log; tail -3 gpurun_out/gpu_tests_lab10.log; cut -c1-400 gpurun_out/bench10_c2.json; cut -c1-300 gpurun_out/bench10_ref.json

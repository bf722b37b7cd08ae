#!/bin/bash
# Last call: comment-only source edits since call 10 -> rebuild verified by the whole GPU suite once more, traffic capture with the new
# digest, the driver's bench line.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/gpu_tests16.log 2>&1; tail -5 gpurun_out/gpu_tests16.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 ncu --clock-control none -k regex:conv_tc6 -s 54 -c 1 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/conv_traffic.csv python tools/profile_forward.py --batch 16 --evals 2 > gpurun_out/conv_traffic.log 2>&1
python tools/make_conv_traffic.py gpurun_out/conv_traffic.csv > gpurun_out/r02_conv_traffic.json 2> gpurun_out/conv_traffic.err
cp gpurun_out/r02_conv_traffic.json profiles/r02_conv_traffic.json
timeout 900 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench16_c2.json 2> gpurun_out/bench16_c2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench16_c2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['roofline']['traffic'])"

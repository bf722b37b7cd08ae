#!/bin/bash
# What the driver's bench line will see: bench.py --steps 20 --warmup 5 with 1 and 2 lanes, and per-step traces of both.
mkdir -p gpurun_out
timeout 600 python tools/step_trace.py --steps 25 --lanes 1 > gpurun_out/trace_lanes1.log 2>&1
sleep 15
timeout 600 python tools/step_trace.py --steps 25 --lanes 2 > gpurun_out/trace_lanes2.log 2>&1
sleep 15
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench8_l1.json 2> gpurun_out/bench8_l1.err
sleep 15
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --lanes 2 > gpurun_out/bench8_l2.json 2> gpurun_out/bench8_l2.err
head -3 gpurun_out/trace_lanes1.log | cut -c1-400; head -3 gpurun_out/trace_lanes2.log | cut -c1-400
for f in gpurun_out/bench8_l1.json gpurun_out/bench8_l2.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'], (d.get('roofline') or {}).get('frac'))"; done

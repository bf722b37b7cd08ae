#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/step_trace.py --steps 25 > gpurun_out/trace13_default.log 2>&1
sleep 10
timeout 600 python tools/step_trace.py --steps 25 --opt outconv_variant=2 > gpurun_out/trace13_oc2.log 2>&1
sleep 10
timeout 600 python tools/step_trace.py --steps 25 > gpurun_out/trace13_default_b.log 2>&1
sleep 10
timeout 600 python tools/step_trace.py --steps 25 --opt outconv_variant=2 > gpurun_out/trace13_oc2_b.log 2>&1
for f in gpurun_out/trace13_default.log gpurun_out/trace13_oc2.log gpurun_out/trace13_default_b.log gpurun_out/trace13_oc2_b.log; do head -2 $f | cut -c1-330; done

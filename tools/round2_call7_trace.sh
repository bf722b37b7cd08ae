#!/bin/bash
# Step-time drift: 40 steps each, round-2 defaults vs the round-1 kernel selection, with clocks / power / temperature per step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=power.limit,power.default_limit,power.max_limit,temperature.gpu,clocks.max.sm --format=csv > gpurun_out/power_limits.txt 2>&1
timeout 600 python tools/step_trace.py --steps 40 > gpurun_out/trace_default.log 2>&1
sleep 20
timeout 600 python tools/step_trace.py --steps 40 --opt tc6_lean=4 --opt fir_variant=3 --opt outconv_variant=4 --opt inconv_variant=3 --opt combine_variant=2 --opt tc1_narrow=2 --opt gn_self=2 --opt gnfin_variant=2 --opt attn_variant=3 > gpurun_out/trace_round1.log 2>&1
sleep 20
timeout 600 python tools/step_trace.py --steps 40 --opt tc6_lean=3 > gpurun_out/trace_lean3.log 2>&1
cat gpurun_out/power_limits.txt; head -3 gpurun_out/trace_default.log | cut -c1-600; head -3 gpurun_out/trace_round1.log | cut -c1-600; head -3 gpurun_out/trace_lean3.log | cut -c1-600

#!/bin/bash
# Final checks of the round-2 defaults (one lane, new config-4 fixture test) + the remaining conv_tc6 ring / issue switches on top of them.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -s --maxfail=30 > gpurun_out/gpu_tests_full9.log 2>&1; tail -30 gpurun_out/gpu_tests_full9.log > gpurun_out/gpu_tests9.log
python __graft_entry__.py smoke > gpurun_out/smoke9.log 2>&1
timeout 900 python tools/ab_forward.py --rounds 7 tc6_rings=1 tc6_mma=1 tc6_tma_poll=1 tc6_lean=3 tc6_rings=1,tc6_lean=3 > gpurun_out/ab_rings9.log 2>&1
timeout 600 python tools/step_trace.py --steps 20 --opt tc6_rings=1 > gpurun_out/trace_rings1.log 2>&1
sleep 10
timeout 600 python tools/step_trace.py --steps 20 > gpurun_out/trace_default9.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 5 > gpurun_out/bench9_c2.json 2> gpurun_out/bench9_c2.err
tail -5 gpurun_out/gpu_tests9.log; tail -1 gpurun_out/smoke9.log; cat gpurun_out/ab_rings9.log; head -2 gpurun_out/trace_rings1.log | cut -c1-300; head -2 gpurun_out/trace_default9.log | cut -c1-300; cut -c1-600 gpurun_out/bench9_c2.json

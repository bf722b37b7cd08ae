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
# programmatic dependent launch where the step is latency-bound (small batches, no power cap): lab twin, pdl off / on
SGMSE_B200_PDL=1 timeout 600 python tools/sweep_bench.py --batches 1,2,4,8 --steps 3 --warmup 2 > gpurun_out/sweep9_pdl0.jsonl 2> gpurun_out/sweep9_pdl0.err
SGMSE_B200_PDL=1 timeout 600 python tools/sweep_bench.py --batches 1,2,4,8 --steps 3 --warmup 2 --opt pdl=1 > gpurun_out/sweep9_pdl1.jsonl 2> gpurun_out/sweep9_pdl1.err
for f in gpurun_out/sweep9_pdl0.jsonl gpurun_out/sweep9_pdl1.jsonl; do echo $f; python -c "
import json
for l in open('$f'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['global_batch'], d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"; done
tail -5 gpurun_out/gpu_tests9.log; tail -1 gpurun_out/smoke9.log; cat gpurun_out/ab_rings9.log; head -2 gpurun_out/trace_rings1.log | cut -c1-300; head -2 gpurun_out/trace_default9.log | cut -c1-300; cut -c1-600 gpurun_out/bench9_c2.json

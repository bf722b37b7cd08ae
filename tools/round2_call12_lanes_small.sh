#!/bin/bash
# Small batches are latency-bound, not power-bound: do two concurrent lanes help there?
mkdir -p gpurun_out
timeout 600 python tools/sweep_bench.py --batches 1,2,4,8,16 --steps 3 --warmup 2 > gpurun_out/sweep12_l1.jsonl 2> gpurun_out/sweep12_l1.err
timeout 600 python tools/sweep_bench.py --batches 2,4,8,16 --steps 3 --warmup 2 --opt lanes=2 > gpurun_out/sweep12_l2.jsonl 2> gpurun_out/sweep12_l2.err
timeout 600 python tools/sweep_bench.py --batches 4,8,16 --steps 3 --warmup 2 --opt lanes=4 > gpurun_out/sweep12_l4.jsonl 2> gpurun_out/sweep12_l4.err
for f in gpurun_out/sweep12_l1.jsonl gpurun_out/sweep12_l2.jsonl gpurun_out/sweep12_l4.jsonl; do echo $f; python -c "
import json
for l in open('$f'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['global_batch'], d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; done

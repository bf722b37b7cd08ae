#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --clock-control none --metrics gpu__time_duration.sum,sm__cycles_active.avg,launch__grid_size,launch__block_size --csv --log-file gpurun_out/launches_b1_final.csv python tools/profile_forward.py --batch 1 --evals 2 > gpurun_out/launches_b1_final.log 2>&1
tail -2 gpurun_out/launches_b1_final.log

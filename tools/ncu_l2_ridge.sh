#!/bin/bash
# Round-2 measurement for DESIGN.md §3 "conv_tc6 at the L2 -> SM ridge": bytes through L2 and into the SMs, DRAM bytes,
# tensor-pipe activity and duration of the tcgen05 convolutions of ONE forward at the benchmark shape -- fused producers
# (default) and TMA-fed operands (tc_variant 6).  Metrics pass only (a few replays per kernel); run under gpurun on ONE GPU:
#   gpurun --timeout 900 -- 'bash tools/ncu_l2_ridge.sh'
# Read with: python tools/ncu_raw.py gpurun_out/l2_ridge_fused.csv   (or any csv reader): per launch
#   L2->SM bytes = lts__t_sectors_srcunit_tex_op_read.sum * 32, compare with 375 KB x tiles of the launch.
M=gpu__time_duration.sum,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.sum,smsp__inst_executed.sum
mkdir -p gpurun_out
ncu --clock-control none -k regex:conv_tc6 -s 54 -c 54 --metrics $M --csv --log-file gpurun_out/l2_ridge_fused.csv \
    python tools/profile_forward.py --batch 16 --evals 2 > gpurun_out/l2_ridge_fused.log 2>&1
ncu --clock-control none -k regex:conv_tc6 -s 54 -c 54 --metrics $M --csv --log-file gpurun_out/l2_ridge_tma.csv \
    python tools/profile_forward.py --batch 16 --evals 2 --opt tc_variant=6 > gpurun_out/l2_ridge_tma.log 2>&1
tail -2 gpurun_out/l2_ridge_fused.log gpurun_out/l2_ridge_tma.log

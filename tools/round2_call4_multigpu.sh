#!/bin/bash
# Multi-GPU call of round 2 (gpurun --gpus 8): shard-count invariance over NCCL on 2 ranks, strong-scaling batch sweep on 8 ranks.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_invariance.py > gpurun_out/nccl_invariance.log 2>&1
INV_BATCH=11 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 tools/nccl_invariance.py >> gpurun_out/nccl_invariance.log 2>&1
timeout 600 python -m pytest tests/test_gpu_zz_next_rows.py -q -m gpu -s -k two_rank_nccl > gpurun_out/gpu_test_nccl.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/sweep_bench.py --batches 1,2,4,8,16,32,64,128,256,512 --steps 2 --warmup 2 > gpurun_out/sweep_g8.jsonl 2> gpurun_out/sweep_g8.err
grep -h "nccl_invariance" gpurun_out/nccl_invariance.log; tail -3 gpurun_out/gpu_test_nccl.log; cut -c1-260 gpurun_out/sweep_g8.jsonl; tail -3 gpurun_out/sweep_g8.err

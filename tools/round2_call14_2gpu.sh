#!/bin/bash
# The driver's multi-GPU launch of bench.py on 2 ranks with the final library (NCCL broadcast -> device-side weight packing), + invariance.
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench14_n2.json 2> gpurun_out/bench14_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench14_ref_n2.json 2> gpurun_out/bench14_ref_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/nccl_invariance.py > gpurun_out/nccl14.log 2>&1
cut -c1-500 gpurun_out/bench14_n2.json; tail -2 gpurun_out/bench14_n2.err; cut -c1-200 gpurun_out/bench14_ref_n2.json; grep nccl_invariance gpurun_out/nccl14.log

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_zz_next_rows.py -q -m gpu -s -k "directory_enhancer" > gpurun_out/gpu_tests17.log 2>&1; tail -8 gpurun_out/gpu_tests17.log

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vposer.py tests/test_gpu_dropin.py -x -q -m gpu > gpurun_out/k_vposer.txt 2>&1; tail -25 gpurun_out/k_vposer.txt

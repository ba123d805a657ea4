#!/bin/bash
# first GPU bring-up: parity tests + closure timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/a_pytest.txt
timeout 600 python scripts/time_closure.py > gpurun_out/a_time.txt 2>&1
cat gpurun_out/a_pytest.txt gpurun_out/a_time.txt

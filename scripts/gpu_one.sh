#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/tail_latency.py 1 > gpurun_out/k_tail1.txt 2>&1; tail -14 gpurun_out/k_tail1.txt
timeout 300 python scripts/tail_latency.py 4 > gpurun_out/k_tail4.txt 2>&1; tail -12 gpurun_out/k_tail4.txt

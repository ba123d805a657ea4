#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/phase_times.py 1 > gpurun_out/r2t_phase_b1.txt 2>&1; cat gpurun_out/r2t_phase_b1.txt | tail -40
timeout 300 python scripts/phase_times.py 256 > gpurun_out/r2t_phase_b256.txt 2>&1; tail -36 gpurun_out/r2t_phase_b256.txt

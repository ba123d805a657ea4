#!/bin/bash
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|Socket|Core|Thread" | head -5
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/m_ref.txt 2> gpurun_out/m_ref.err; tail -c 700 gpurun_out/m_ref.txt; tail -2 gpurun_out/m_ref.err

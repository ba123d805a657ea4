#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/m_ref.txt 2> gpurun_out/m_ref.err
echo "elapsed $(( $(date +%s) - t0 )) s"; tail -c 500 gpurun_out/m_ref.txt

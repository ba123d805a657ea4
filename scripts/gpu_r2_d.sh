#!/bin/bash
mkdir -p gpurun_out
for B in 5 256; do
  timeout 300 python scripts/dense_ab.py gpurun_out/ab_p_$B.npz $B 2>&1 | tail -1
done
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -15 > gpurun_out/r2d_tests.log
tail -6 gpurun_out/r2d_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
head -c 500 gpurun_out/r2d_bench.json; tail -3 gpurun_out/r2d_bench.err

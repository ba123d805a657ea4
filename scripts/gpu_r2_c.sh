#!/bin/bash
# round 2, third GPU call: first run of the persistent dense-round kernel -- A/B against the multi-kernel rounds, suite, bench
mkdir -p gpurun_out
for B in 5 40 256; do
  timeout 300 python scripts/dense_ab.py gpurun_out/ab_p_$B.npz $B > gpurun_out/r2c_ab_p_$B.log 2>&1; echo "persistent B=$B rc=$?" >> gpurun_out/r2c_ab.log
  MVS_DENSE_MULTIKERNEL=1 timeout 300 python scripts/dense_ab.py gpurun_out/ab_m_$B.npz $B > gpurun_out/r2c_ab_m_$B.log 2>&1; echo "multikernel B=$B rc=$?" >> gpurun_out/r2c_ab.log
  python scripts/dense_ab.py --compare gpurun_out/ab_p_$B.npz gpurun_out/ab_m_$B.npz >> gpurun_out/r2c_ab.log 2>&1
  tail -2 gpurun_out/r2c_ab_p_$B.log >> gpurun_out/r2c_ab.log; tail -1 gpurun_out/r2c_ab_m_$B.log >> gpurun_out/r2c_ab.log
done
cat gpurun_out/r2c_ab.log
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v Warning | tail -25 > gpurun_out/r2c_tests.log
tail -8 gpurun_out/r2c_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
head -c 700 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err

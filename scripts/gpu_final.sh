#!/bin/bash
# end-of-round: tests, bench (default flags), the VPoser workload, smoke, full ncu captures of the round kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/f_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.txt; tail -3 gpurun_out/f_pytest.txt
timeout 600 python bench.py > gpurun_out/f_bench.txt 2> gpurun_out/f_bench.err; tail -c 300 gpurun_out/f_bench.txt; tail -2 gpurun_out/f_bench.err
timeout 300 python bench.py --frames 1 --views 4 --vposer 1 --steps 5 --warmup 3 > gpurun_out/f_cfg1.txt 2> gpurun_out/f_cfg1.err; tail -c 300 gpurun_out/f_cfg1.txt; tail -2 gpurun_out/f_cfg1.err
timeout 300 python bench.py --frames 256 --views 4 --vposer 1 --steps 3 --warmup 2 > gpurun_out/f_vp256.txt 2> gpurun_out/f_vp256.err; tail -c 300 gpurun_out/f_vp256.txt; tail -2 gpurun_out/f_vp256.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.txt 2>&1; tail -2 gpurun_out/f_smoke.txt
for k in sdf_fused_kernel frame_step_kernel skin_kernel posedirs_gemm_tc_kernel; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:$k -s 10 -c 1 -o gpurun_out/q_$k python scripts/prof_closure.py lbfgs > gpurun_out/q_ncu_$k.log 2>&1
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lbfgs_resident_kernel -c 1 -o gpurun_out/q_lbfgs_resident_kernel python scripts/prof_closure.py resident > gpurun_out/q_ncu_res.log 2>&1
ls gpurun_out | grep "^q_.*rep"

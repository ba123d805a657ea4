#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
python bench.py --steps 3 --warmup 2 > gpurun_out/e_bench.txt 2> gpurun_out/e_bench.err; tail -c 1600 gpurun_out/e_bench.txt; tail -3 gpurun_out/e_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vertex_fwd_tc_kernel -s 3 -c 1 -o gpurun_out/e_vertex_fwd_tc python scripts/prof_closure.py lbfgs > gpurun_out/e_ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/e_launches_lbfgs.csv python scripts/prof_closure.py lbfgs > gpurun_out/e_ncu3.log 2>&1
ls -la gpurun_out | grep e_

#!/bin/bash
# one development iteration on the GPU box: tests, per-phase timing of frame_step, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/i_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.txt
tail -3 gpurun_out/i_pytest.txt
timeout 120 python scripts/phase_times.py 256 > gpurun_out/i_phase.txt 2>&1; tail -3 gpurun_out/i_phase.txt
timeout 120 python scripts/phase_times.py 256 resident > gpurun_out/i_phase_res.txt 2>&1; tail -3 gpurun_out/i_phase_res.txt
timeout 600 python bench.py --steps 3 --warmup 2 > gpurun_out/i_bench.txt 2> gpurun_out/i_bench.err; tail -c 600 gpurun_out/i_bench.txt; tail -3 gpurun_out/i_bench.err
timeout 300 python scripts/tail_latency.py 1 > gpurun_out/i_tail1.txt 2>&1; tail -12 gpurun_out/i_tail1.txt

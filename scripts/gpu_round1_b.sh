#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/b_pytest.txt
cat gpurun_out/b_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b_smoke.txt 2>&1; tail -5 gpurun_out/b_smoke.txt
timeout 900 python bench.py --steps 2 --warmup 2 > gpurun_out/b_bench.txt 2> gpurun_out/b_bench.err; tail -3 gpurun_out/b_bench.txt; tail -5 gpurun_out/b_bench.err

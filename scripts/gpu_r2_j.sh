#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err
tail -c 1500 gpurun_out/r2j_bench_n2.json; tail -5 gpurun_out/r2j_bench_n2.err
timeout 600 python scripts/cfg5_convergence.py 2>&1 | tail -4
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -12 > gpurun_out/r2j_tests.log; tail -5 gpurun_out/r2j_tests.log

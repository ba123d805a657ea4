#!/bin/bash
# 2-GPU validation of the bench contract (torchrun, NCCL) + the reference arm
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/m_bench2.txt 2> gpurun_out/m_bench2.err; tail -c 700 gpurun_out/m_bench2.txt; tail -3 gpurun_out/m_bench2.err
timeout 600 python -m pytest tests/test_gpu_sequence.py -x -q -m gpu > gpurun_out/m_seq.txt 2>&1; tail -2 gpurun_out/m_seq.txt
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/m_ref.txt 2> gpurun_out/m_ref.err; tail -c 900 gpurun_out/m_ref.txt; tail -3 gpurun_out/m_ref.err

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2s_ref_cfg4.json 2> gpurun_out/r2s_ref_cfg4.err; tail -c 700 gpurun_out/r2s_ref_cfg4.json; tail -3 gpurun_out/r2s_ref_cfg4.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 --frames 64 --views 4 --sdf 0 --ref-seconds 60 > gpurun_out/r2s_ref_cfg3.json 2> gpurun_out/r2s_ref_cfg3.err; tail -c 500 gpurun_out/r2s_ref_cfg3.json

#!/bin/bash
mkdir -p gpurun_out
/usr/bin/time -v timeout 900 python bench.py --impl reference > gpurun_out/m_ref.txt 2> gpurun_out/m_ref.err; grep -E "Elapsed|Maximum resident" gpurun_out/m_ref.err; tail -c 500 gpurun_out/m_ref.txt

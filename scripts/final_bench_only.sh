#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/final_bench2.json 2> gpurun_out/final_bench2.err
cut -c1-200 gpurun_out/final_bench2.json

#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multical_b200/csrc -diag-suppress 550 -o /tmp/chol_bench scripts/chol_bench.cu && /tmp/chol_bench | tee gpurun_out/chol_bench.txt

#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multical_b200/csrc -diag-suppress 550 -o /tmp/chol_bench scripts/chol_bench.cu && /tmp/chol_bench | grep -E "n= 32" | tee gpurun_out/chol_bench.txt
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for wl in cfg2 cfg3 cfg4; do
  timeout 300 python scripts/profile_one.py $wl time 2>&1 | tail -1
  timeout 300 python scripts/profile_one.py $wl solve 2>&1 | tail -1
  MCBA_PROF=1 timeout 300 python scripts/profile_one.py $wl solve 2>&1 | grep "k_lm phases" | tail -3 | head -1
done

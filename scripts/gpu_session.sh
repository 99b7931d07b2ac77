#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -x > gpurun_out/pytest_gpu_parity.log 2>&1; tail -3 gpurun_out/pytest_gpu_parity.log
for wl in cfg2 cfg3 cfg4 cfg5; do
  timeout 300 python scripts/profile_one.py $wl time 2>&1 | tail -1
  timeout 300 python scripts/profile_one.py $wl solve 2>&1 | tail -1
done 2>&1 | tee gpurun_out/profile_all.txt

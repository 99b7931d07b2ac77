#!/bin/bash
# A/B after: warp-wide view search + cross-view prefetch in k_linearize; 64-column panels in the cooperative Cholesky
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for wl in cfg2 cfg3 cfg4 cfg5; do
  timeout 300 python scripts/profile_one.py $wl time 2>&1 | tail -1
  timeout 300 python scripts/profile_one.py $wl solve 2>&1 | tail -2
done > gpurun_out/profile_all.txt 2>&1
cat gpurun_out/profile_all.txt
for wl in cfg4 cfg5; do
  MCBA_PROF=1 timeout 300 python scripts/profile_one.py $wl solve > gpurun_out/phases_$wl.txt 2>&1; grep phases gpurun_out/phases_$wl.txt | tail -2
done

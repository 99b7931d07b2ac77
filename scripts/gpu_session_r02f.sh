#!/bin/bash
# full GPU suite + bench line + small-grid A/B (cfg2 / cfg3) + k_lm phases of every workload
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for g in 148 0; do
  for wl in cfg2 cfg3; do echo "MCBA_LM_GRID=$g $wl"; MCBA_LM_GRID=$g timeout 300 python scripts/profile_one.py $wl solve 2>&1 | tail -1; done
done 2>&1 | tee gpurun_out/lm_grid_ab.txt
for wl in cfg2 cfg3 cfg4 cfg5; do
  MCBA_PROF=1 timeout 300 python scripts/profile_one.py $wl solve > gpurun_out/phases_$wl.txt 2>&1; grep phases gpurun_out/phases_$wl.txt | tail -2 | head -1
done
( time timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2>&1 | grep real; python scripts/show_bench.py gpurun_out/bench_n1.json

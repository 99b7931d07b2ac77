#!/bin/bash
# One GPU call: smoke, the GPU suite (graph loop and host-driven loop), timings at cfg2/cfg3/cfg4, launch lists, ncu captures.
mkdir -p gpurun_out
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
MCBA_GRAPH=0 timeout 1500 python -m pytest tests -q -m gpu --tb=line -x > gpurun_out/pytest_gpu_nograph.log 2>&1; tail -3 gpurun_out/pytest_gpu_nograph.log
for wl in cfg2 cfg3 cfg4; do
  timeout 300 python scripts/profile_one.py $wl time 2>&1 | tail -1
  timeout 300 python scripts/profile_one.py $wl solve 2>&1 | tail -1
  MCBA_GRAPH=0 timeout 300 python scripts/profile_one.py $wl solve 2>&1 | tail -1
done
for wl in cfg2 cfg4; do
  MCBA_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$wl.csv python scripts/profile_one.py $wl solve > gpurun_out/ncu_$wl.log 2>&1
  python scripts/summarize_launches.py gpurun_out/launches_$wl.csv 2>/dev/null | head -16
done
MCBA_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_linearize -s 1 -c 1 -o gpurun_out/ncu_linearize_cfg4 python scripts/profile_one.py cfg4 solve > gpurun_out/ncu_full_cfg4.log 2>&1; tail -2 gpurun_out/ncu_full_cfg4.log
MCBA_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_lm -s 1 -c 1 -o gpurun_out/ncu_lm_cfg2 python scripts/profile_one.py cfg2 solve > gpurun_out/ncu_full_cfg2.log 2>&1; tail -2 gpurun_out/ncu_full_cfg2.log
ls -la gpurun_out/*.ncu-rep

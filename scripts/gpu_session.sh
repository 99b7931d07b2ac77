#!/bin/bash
# One GPU call: smoke, the GPU suite, bench, timings of the linearisation kernel and of whole solves at cfg2/cfg3/cfg4.
mkdir -p gpurun_out
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
for wl in cfg2 cfg3 cfg4; do
  timeout 300 python scripts/profile_one.py $wl time 2>&1 | tail -1
  timeout 300 python scripts/profile_one.py $wl solve 2>&1 | tail -1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg4.csv python scripts/profile_one.py cfg4 solve > gpurun_out/ncu_cfg4.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_cfg4.csv 2>/dev/null | head -24
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv python scripts/profile_one.py cfg2 solve > gpurun_out/ncu_cfg2.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_cfg2.csv 2>/dev/null | head -24
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/cond scripts/probe_cond_graph.cu && /tmp/cond

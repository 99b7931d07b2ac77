#!/bin/bash
# Hardware session of the final round-2 code: GPU suite, bench line, cfg5, phase times, launch list, ncu --set full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 180 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 1200 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python scripts/show_bench.py gpurun_out/bench_n1.json
for wl in cfg4 cfg5; do
  timeout 300 python scripts/profile_one.py $wl time 2>&1 | tail -1
  timeout 300 python scripts/profile_one.py $wl solve 2>&1 | tail -3
done > gpurun_out/profile_all.txt 2>&1
cat gpurun_out/profile_all.txt
for wl in cfg2 cfg4 cfg5; do
  MCBA_PROF=1 timeout 300 python scripts/profile_one.py $wl solve > gpurun_out/phases_$wl.txt 2>&1
done
tail -8 gpurun_out/phases_cfg4.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_cfg4.csv \
  python bench.py --steps 2 --warmup 1 --no-secondary > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_bench_cfg4.csv | head -20
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_linearize -s 2 -c 1 -f -o gpurun_out/ncu_k_linearize_cfg4 \
  python scripts/profile_one.py cfg4 kernels > gpurun_out/ncu_lin.log 2>&1; tail -2 gpurun_out/ncu_lin.log
MCBA_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_lm -s 1 -c 1 -f -o gpurun_out/ncu_k_lm_cfg4 \
  python scripts/profile_one.py cfg4 solve > gpurun_out/ncu_lm.log 2>&1; tail -2 gpurun_out/ncu_lm.log
ls -la gpurun_out

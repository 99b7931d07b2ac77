#!/bin/bash
# final evidence of the round: GPU suite, bench line (cfg4 + cfg2/cfg3/cfg5 blocks), reference arm on the box, phase stamps, launch list with the loop kernels
# visible (MCBA_GRAPH=0: ncu does not list the body of a CUDA-graph WHILE node), ncu --set full of k_linearize and k_lm
mkdir -p gpurun_out
timeout 180 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
timeout 1200 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2>&1 | grep real; python scripts/show_bench.py gpurun_out/bench_n1.json
for wl in cfg2 cfg3 cfg4 cfg5; do
  MCBA_PROF=1 timeout 300 python scripts/profile_one.py $wl solve > gpurun_out/phases_$wl.txt 2>&1; grep phases gpurun_out/phases_$wl.txt | tail -2 | head -1
done
MCBA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_solve_cfg4.csv \
  python scripts/profile_one.py cfg4 solve > gpurun_out/ncu_solve.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_solve_cfg4.csv | head -14
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_linearize -s 2 -c 1 -f -o gpurun_out/ncu_k_linearize_cfg4 \
  python scripts/profile_one.py cfg4 kernels > gpurun_out/ncu_lin.log 2>&1; tail -1 gpurun_out/ncu_lin.log
MCBA_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_lm -s 1 -c 1 -f -o gpurun_out/ncu_k_lm_cfg4 \
  python scripts/profile_one.py cfg4 solve > gpurun_out/ncu_lm.log 2>&1; tail -1 gpurun_out/ncu_lm.log
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2>&1 | grep real; tail -c 600 gpurun_out/bench_reference.json

#!/bin/bash
# one GPU call that produces everything judged: test log, bench line, reference arm, launch list, ncu of the hot kernel
mkdir -p gpurun_out
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 900 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest.log 2>&1; tail -4 gpurun_out/pytest.log
timeout 300 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -2 gpurun_out/bench_cfg2.err; cat gpurun_out/bench_cfg2.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench_cfg2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_l.log 2>&1
for w in cfg2 cfg3 cfg4; do echo "== $w"; timeout 200 python scripts/profile_one.py $w solve 2>&1 | tail -1; done
timeout 120 python scripts/e2e_breakdown.py cfg2 2>&1 | tail -14

#!/bin/bash
# First GPU call of the next session: everything that was written after the round-1 GPU budget was spent and is so far verified
# only on the SIMT interpreter (tests/test_simt_kernels.py).  One call, outputs under gpurun_out/.
#   1. the GPU suite, new files last (tests/conftest.py)          -> gpurun_out/pytest_gpu.log
#   2. A/B of the opt-in candidates (MCBA_CHOL=blocked, MCBA_FUSE=1) -> gpurun_out/bench_chol_{small,blocked}.json, bench_fuse*.json
#   3. launch lists of both                                         -> gpurun_out/launches_chol_*.csv
#   4. timing of the two motion models and of the batched pose initialisation -> gpurun_out/motion_pnp_timing.txt
mkdir -p gpurun_out
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 1200 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_chol_small.json 2> gpurun_out/bench_chol_small.err; python scripts/show_bench.py gpurun_out/bench_chol_small.json
MCBA_CHOL=blocked timeout 300 python bench.py > gpurun_out/bench_chol_blocked.json 2> gpurun_out/bench_chol_blocked.err; python scripts/show_bench.py gpurun_out/bench_chol_blocked.json
MCBA_FUSE=1 timeout 300 python bench.py > gpurun_out/bench_fuse.json 2> gpurun_out/bench_fuse.err; python scripts/show_bench.py gpurun_out/bench_fuse.json
MCBA_FUSE=1 MCBA_CHOL=blocked timeout 300 python bench.py > gpurun_out/bench_fuse_blocked.json 2> gpurun_out/bench_fuse_blocked.err; python scripts/show_bench.py gpurun_out/bench_fuse_blocked.json
for v in small blocked; do
  MCBA_CHOL=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_chol_$v.csv \
    python scripts/profile_one.py cfg2 solve > gpurun_out/ncu_chol_$v.log 2>&1
  python scripts/summarize_launches.py gpurun_out/launches_chol_$v.csv 2>/dev/null | head -12
done
for e in serial parallel; do echo "== expand maps $e (cfg4)"; MCBA_EXPAND=$e timeout 300 python scripts/profile_one.py cfg4 solve 2>&1 | tail -1; done
for m in mma f32; do echo "== moments $m (cfg4: 5.5 M corners)"; MCBA_MOMENTS=$m timeout 300 python scripts/profile_one.py cfg4 time 2>&1 | tail -1; MCBA_MOMENTS=$m timeout 300 python scripts/profile_one.py cfg4 solve 2>&1 | tail -1; done
timeout 300 python scripts/motion_pnp_timing.py > gpurun_out/motion_pnp_timing.txt 2>&1; cat gpurun_out/motion_pnp_timing.txt
# 5. (needs --gpus 2 or more) the same A/B on several GPUs: scripts/run_ngpu_ab.sh with MCBA_FUSE=1 (five exchanges per iteration instead of six)

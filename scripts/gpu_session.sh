#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$? $(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/bench_n1.err
python scripts/show_bench.py gpurun_out/bench_n1.json
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$? $(( $(date +%s) - t0 ))s"; tail -1 gpurun_out/bench_ref.json | cut -c1-400

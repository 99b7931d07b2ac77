#!/bin/bash
# developer check on a 2-GPU box: correctness of the sharded solve, then bench.py with and without peer-memory exchanges
mkdir -p gpurun_out
t0=$(date +%s)
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/mgc.log 2>&1
echo "multi_gpu_check rc=$? $(( $(date +%s) - t0 ))s"; tail -6 gpurun_out/mgc.log
port=29600
for p in 1 0; do
  t1=$(date +%s); port=$((port+7))
  MCBA_PEER=$p timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 5 --warmup 2 > gpurun_out/bench_n2_peer$p.json 2> gpurun_out/bench_n2_peer$p.err
  echo "bench peer=$p rc=$? $(( $(date +%s) - t1 ))s"
  tail -3 gpurun_out/bench_n2_peer$p.err
  python - <<PY
import json
try:
  d = json.load(open("gpurun_out/bench_n2_peer$p.json"))
  print("peer=$p value %.4g ms_per_step %.3f evals %.1f e2e_ms %.3f launches %d" % (d["value"], d["ms_per_step"], d["nfev_plus_njev_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
except Exception as e:
  print("no json:", e)
PY
done

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python scripts/show_bench.py gpurun_out/bench_n1.json
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n1.json') if l.startswith('{')][-1])
print('affinity', d['config'].get('host_affinity')); print('e2e f64', d['e2e']['ms_per_step'], 'e2e f32 table', d['e2e_float32_table']['ms_per_step'])
for k,v in d['other_workloads'].items(): print(k, round(v['ms_per_step'],3), round(v['e2e']['ms_per_step'],3))
PY
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_table.py -q -m gpu -x 2>&1 | tail -2

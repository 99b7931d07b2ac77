#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "device_packing or partial_step" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python scripts/show_bench.py gpurun_out/bench_n1.json
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n1.json') if l.startswith('{')][-1])
print('e2e f64', d['e2e']['ms_per_step'], 'e2e f32 table', d.get('e2e_float32_table'))
PY

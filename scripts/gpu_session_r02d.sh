#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/e2e_breakdown.py cfg4 2>&1 | tee gpurun_out/e2e_breakdown_cfg4.txt
timeout 300 python scripts/e2e_breakdown.py cfg2 2>&1 | tee gpurun_out/e2e_breakdown_cfg2.txt

#!/bin/bash
mkdir -p gpurun_out
for wl in cfg4 cfg5; do
  timeout 300 python scripts/profile_one.py $wl solve 2>&1 | tail -1
  MCBA_PROF=1 timeout 300 python scripts/profile_one.py $wl solve > gpurun_out/phases_$wl.txt 2>&1; grep phases gpurun_out/phases_$wl.txt | tail -2 | head -1
done

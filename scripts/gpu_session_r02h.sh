#!/bin/bash
for wl in cfg2 cfg4 cfg5; do timeout 300 python scripts/profile_one.py $wl time 2>&1 | tail -1; timeout 300 python scripts/profile_one.py $wl solve 2>&1 | tail -1; done

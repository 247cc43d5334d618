#!/bin/bash
BDS_TRK_TIMING=1 timeout 90 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | cut -c1-330

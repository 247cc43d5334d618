#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
BDS_LIB_NAME=libtest16.so timeout 600 python -m pytest tests/test_gpu_tracking.py -m gpu -x -q > gpurun_out/pytest16.log 2>&1; echo "pytest16 rc=$?"; tail -3 gpurun_out/pytest16.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
run() { echo "== $*"; env "$@" $B 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('x_rt',round(d['config']['x_realtime'],1),'kernel_ms',round(d['roofline']['kernel_ms_per_launch'],2),'e2e',d['e2e'])
    else: print(l.rstrip()[:300])
"; }
run BDS_TRK_PASSES=3
run BDS_TRK_PASSES=2
run BDS_LIB_NAME=libtest16.so BDS_TRK_PASSES=3
run BDS_LIB_NAME=libtest16.so BDS_TRK_PASSES=2
run BDS_LIB_NAME=libtest16.so BDS_TRK_PASSES=1

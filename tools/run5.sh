#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tracking.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
BDS_LIB_NAME=libtest16.so timeout 900 python -m pytest tests/test_gpu_tracking.py -m gpu -x -q > gpurun_out/pytest16.log 2>&1; echo "pytest16 rc=$?"; tail -4 gpurun_out/pytest16.log
B="timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { echo "== $*"; env "$@" $B 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('x_rt',round(d['config']['x_realtime'],1),'kernel_ms',round(d['roofline']['kernel_ms_per_launch'],2))
    else: print(l.rstrip()[:400])
"; }
run BDS_TRK_PASSES=3 BDS_TRK_STAGES=3 BDS_TRK_TIMING=1
for lib in libbdsgpu.so libtest16.so; do
 for st in 3 4; do
  for ah in 1 2 3; do
   for pa in 2 3; do
     run BDS_LIB_NAME=$lib BDS_TRK_STAGES=$st BDS_TRK_AHEAD=$ah BDS_TRK_PASSES=$pa
   done
  done
 done
done

#!/bin/bash
mkdir -p gpurun_out/r2
LOG=gpurun_out/r2/call6.log
: > $LOG
timeout 400 python -m pytest tests/test_gpu_tracking.py -m gpu -x -q > gpurun_out/r2/pytest6.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2/pytest6.log
run() {
  echo "== $*" >> $LOG
  env "$@" timeout 30 python tools/variant_check.py closed gpurun_out/r2/closed6 5 2>&1 | grep -E "^\{|bds timing" | cut -c1-400 >> $LOG || echo "failed rc=$?" >> $LOG
}
run BDS_TRK_TIMING=1
for pa in 1 2; do for ah in 2 3; do run BDS_TRK_PASSES=$pa BDS_TRK_AHEAD=$ah; done; done
echo "== open" >> $LOG
timeout 30 python tools/variant_check.py open gpurun_out/r2/closed6.npz gpurun_out/r2/open6 2>&1 | grep -E "^\{" | cut -c1-300 >> $LOG
for n in 8 30; do run BDS_NCH=$n; done
cat $LOG


#!/bin/bash
# Round 2, GPU call 4: the rebuilt chip-synchronous kernel (table builder warp in the consumer CTA, prefix-mask body).
mkdir -p gpurun_out/r2
LOG=gpurun_out/r2/call4.log
: > $LOG
timeout 400 python -m pytest tests/test_gpu_tracking.py -m gpu -x -q > gpurun_out/r2/pytest4.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2/pytest4.log
run() {
  echo "== $*" >> $LOG
  env "$@" timeout 30 python tools/variant_check.py closed gpurun_out/r2/closed4 5 2>&1 | grep -E "^\{|bds timing" | cut -c1-400 >> $LOG || echo "failed rc=$?" >> $LOG
}
run BDS_NOP=1
run BDS_TRK_TIMING=1
for pa in 1 2 3; do for ah in 1 2 3; do run BDS_TRK_PASSES=$pa BDS_TRK_AHEAD=$ah; done; done
echo "== open" >> $LOG
timeout 30 python tools/variant_check.py open gpurun_out/r2/closed4.npz gpurun_out/r2/open4 2>&1 | grep -E "^\{" | cut -c1-300 >> $LOG
for n in 8 15 30; do run BDS_NCH=$n; done
cat $LOG

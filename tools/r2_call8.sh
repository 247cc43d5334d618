#!/bin/bash
mkdir -p gpurun_out/r2
LOG=gpurun_out/r2/call8.log
: > $LOG
timeout 400 python -m pytest tests/test_gpu_tracking.py -m gpu -x -q > gpurun_out/r2/pytest8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2/pytest8.log
run() {
  lib=$1; shift
  echo "== $lib $*" >> $LOG
  env BDS_LIB_NAME=$lib "$@" timeout 30 python tools/variant_check.py closed gpurun_out/r2/closed8 5 2>&1 | grep -E "^\{|bds timing" | cut -c1-400 >> $LOG || echo "failed rc=$?" >> $LOG
}
for lib in libbdsgpu.so libbds_svclo.so; do
run $lib BDS_TRK_TIMING=1
run $lib BDS_TRK_PASSES=1 BDS_TRK_AHEAD=2
run $lib BDS_TRK_PASSES=2 BDS_TRK_AHEAD=3
run $lib BDS_NCH=8
run $lib BDS_NCH=30
done
echo "== open" >> $LOG
timeout 30 python tools/variant_check.py open gpurun_out/r2/closed8.npz gpurun_out/r2/open8 2>&1 | grep -E "^\{" | cut -c1-300 >> $LOG
cat $LOG

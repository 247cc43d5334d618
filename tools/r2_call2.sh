#!/bin/bash
# Round 2, GPU call 2: service-warps-high variant, pass/prefetch knob sweep on the f2rec build, timing breakdown.
mkdir -p gpurun_out/r2
LOG=gpurun_out/r2/call2.log
: > $LOG
run() {  # lib, env...
  lib=$1; shift
  echo "== $lib $*" >> $LOG
  env BDS_LIB_NAME=$lib "$@" timeout 20 python tools/variant_check.py closed gpurun_out/r2/tmp 5 2>&1 | grep -E "^\{|bds timing" | cut -c1-400 >> $LOG || echo "failed rc=$?" >> $LOG
}
for lib in libbds_f2rec.so libbds_svchi.so; do
  run $lib BDS_NOP=1
  for pa in 1 2 3; do for ah in 0 1 2 3; do run $lib BDS_TRK_PASSES=$pa BDS_TRK_AHEAD=$ah; done; done
done
run libbds_f2rec.so BDS_TRK_TIMING=1
run libbds_svchi.so BDS_TRK_TIMING=1
cat $LOG

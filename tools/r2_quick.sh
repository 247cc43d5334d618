#!/bin/bash
# quick closed-loop timing of the default build at 60 / 30 / 8 channels (+ timing breakdown, + open loop)
mkdir -p gpurun_out/r2
LOG=gpurun_out/r2/quick.log
: > $LOG
run() {
  echo "== $*" >> $LOG
  env "$@" timeout 30 python tools/variant_check.py closed gpurun_out/r2/closedq 5 2>&1 | grep -E "^\{|bds timing" | cut -c1-400 >> $LOG || echo "failed rc=$?" >> $LOG
}
run BDS_TRK_TIMING=1
run BDS_NOP=1
echo "== open" >> $LOG
timeout 30 python tools/variant_check.py open gpurun_out/r2/closedq.npz gpurun_out/r2/openq 2>&1 | grep -E "^\{" | cut -c1-300 >> $LOG
run BDS_NCH=30
run BDS_NCH=8
for x in "$@"; do run $x; done
cat $LOG

#!/bin/bash
mkdir -p gpurun_out/r2
LOG=gpurun_out/r2/quick.log
: > $LOG
run() {
  lib=$1; shift
  echo "== $lib $*" >> $LOG
  env BDS_LIB_NAME=$lib "$@" timeout 30 python tools/variant_check.py closed gpurun_out/r2/closedq_$lib 5 2>&1 | grep -E "^\{|bds timing" | cut -c1-400 >> $LOG || echo "failed rc=$?" >> $LOG
}
run libbds_dev.so BDS_TRK_TIMING=1
run libbds_dev.so BDS_TRK_TIMING=1 BDS_NCH=8
run libbdsgpu.so BDS_NOP=1
echo "== open abl4" >> $LOG
BDS_LIB_NAME=libbds_abl4.so timeout 30 python tools/variant_check.py open gpurun_out/r2/closedq_libbdsgpu.so.npz gpurun_out/r2/openq 2>&1 | grep -E "^\{" | cut -c1-300 >> $LOG
cat $LOG

#!/bin/bash
# B2a cluster kernel: GPU tests of the B2a paths, then closed-loop timing at several cluster sizes / channel counts
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_b2a_unit.py tests/test_gpu_tracking.py -m gpu -x -q -k "b2a or B2a" > gpurun_out/r2/pytest_b2a.log 2>&1; echo "pytest b2a rc=$?"; tail -12 gpurun_out/r2/pytest_b2a.log
LOG=gpurun_out/r2/b2a_cs.log; : > $LOG
run() { echo "== $*" >> $LOG; env "$@" timeout 120 python tools/variant_check.py closed_b2a gpurun_out/r2/b2a_cs 2 2>&1 | grep -E "^\{|bds timing|rror" | grep -v "producer detail" | cut -c1-330 >> $LOG || echo "failed" >> $LOG; }
run BDS_B2A_CS=1
run BDS_B2A_CS=2
run BDS_NCH=30 BDS_B2A_CS=4
run BDS_NCH=8 BDS_B2A_CS=8
run BDS_LIB_NAME=libbds_dev.so BDS_TRK_TIMING=1 BDS_B2A_CS=1
run BDS_LIB_NAME=libbds_dev.so BDS_TRK_TIMING=1 BDS_B2A_CS=2
run BDS_LIB_NAME=libbds_dev.so BDS_TRK_TIMING=1 BDS_NCH=8 BDS_B2A_CS=8
cat $LOG

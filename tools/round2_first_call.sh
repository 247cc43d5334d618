#!/bin/bash
# First GPU call of the next round (one B200, under gpurun): everything prepared offline at the end of round 1, each
# step with its own timeout.  Build the variants first:  bash tools/variant_build.sh
mkdir -p gpurun_out/r2
# 1. the per-channel B2a kernel: hardware gate + speed against the general kernel
BDS_TEST_B2A_UNIT=1 timeout 300 python -m pytest tests/test_gpu_b2a_unit.py -m gpu -x -q > gpurun_out/r2/pytest_b2a_unit.log 2>&1
echo "b2a unit tests rc=$?"; tail -3 gpurun_out/r2/pytest_b2a_unit.log
timeout 60 python tools/variant_check.py closed_b2a gpurun_out/r2/b2a_general 1 2>&1 | tail -1
BDS_TRK_B2A_UNIT=1 timeout 60 python tools/variant_check.py closed_b2a gpurun_out/r2/b2a_unit 1 2>&1 | tail -1
# 2. the B1C variants, closed and open loop
timeout 200 bash tools/variant_run.sh 5 "libbdsgpu.so libbds_f2.so libbds_rec.so libbds_f2rec.so libbds_w20.so libbds_w20f2rec.so" ""
# 3. the GPU suite on the most promising build (becomes the default if green)
BDS_LIB_NAME=libbds_f2rec.so timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_f2rec.log 2>&1
echo "f2rec suite rc=$?"; tail -3 gpurun_out/r2/pytest_f2rec.log

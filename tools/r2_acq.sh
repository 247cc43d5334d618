#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_acquisition.py -m gpu -x -q > gpurun_out/r2/pytest_acq.log 2>&1; echo "pytest acq rc=$?"; tail -25 gpurun_out/r2/pytest_acq.log

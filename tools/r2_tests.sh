#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2/pytest_all.log

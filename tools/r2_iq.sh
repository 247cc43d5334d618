#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q -k "iq" > gpurun_out/r2/pytest_iq.log 2>&1; echo "pytest iq rc=$?"; tail -25 gpurun_out/r2/pytest_iq.log
timeout 900 python -m pytest tests -m gpu -x -q -k "not iq" > gpurun_out/r2/pytest_rest.log 2>&1; echo "pytest rest rc=$?"; tail -8 gpurun_out/r2/pytest_rest.log

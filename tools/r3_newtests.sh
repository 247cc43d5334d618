#!/bin/bash
O=gpurun_out/r3x
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_acquisition.py tests/test_gpu_tracking.py -m gpu -x -q -k "53_mhz or sixty_channels" > $O/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_new.log

#!/bin/bash
mkdir -p gpurun_out/r2
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest3.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2/pytest3.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2/bench3.json 2> gpurun_out/r2/bench3.err; echo "bench rc=$?"; cat gpurun_out/r2/bench3.json; tail -5 gpurun_out/r2/bench3.err

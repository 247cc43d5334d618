#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_acquisition.py -m gpu -x -q > gpurun_out/pytest_acq.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_acq.log
timeout 120 python bench.py --workload acq_b2a --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_acq.json 2> gpurun_out/bench_acq.err; echo "acq rc=$?"; cut -c1-900 gpurun_out/bench_acq.json; tail -3 gpurun_out/bench_acq.err

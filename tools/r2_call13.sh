#!/bin/bash
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_b2a_unit.py tests/test_gpu_tracking.py -m gpu -x -q > gpurun_out/r2/pytest13.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2/pytest13.log
timeout 60 python tools/variant_check.py closed_b2a gpurun_out/r2/b2a_unit 1 2>&1 | tail -1
BDS_NCH=8 timeout 60 python tools/variant_check.py closed_b2a gpurun_out/r2/b2a_unit8 1 2>&1 | tail -1
timeout 400 python bench.py --workload track_b2a --steps 3 --warmup 3 > gpurun_out/r2/bench_b2a.json 2> gpurun_out/r2/bench_b2a.err; echo "bench rc=$?"; cut -c1-1800 gpurun_out/r2/bench_b2a.json; tail -5 gpurun_out/r2/bench_b2a.err

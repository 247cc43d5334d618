#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest.log
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
BDS_TRK_TIMING=1 timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | cut -c1-600 | tail -5
timeout 200 python bench.py --workload acq_b2a --steps 2 --warmup 1 > gpurun_out/bench_acq.json 2> gpurun_out/bench_acq.err; echo "acq rc=$?"; cat gpurun_out/bench_acq.json; tail -3 gpurun_out/bench_acq.err

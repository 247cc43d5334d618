#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_acquisition.py -m gpu -x -q 2>&1 | tail -2
timeout 120 python bench.py --workload acq_b2a --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_acq.json 2> gpurun_out/bench_acq.err; cut -c1-330 gpurun_out/bench_acq.json; tail -2 gpurun_out/bench_acq.err
timeout 200 python bench.py --workload acq_b1c --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_acq_b1c.json 2> gpurun_out/bench_acq_b1c.err; cut -c1-330 gpurun_out/bench_acq_b1c.json; tail -2 gpurun_out/bench_acq_b1c.err

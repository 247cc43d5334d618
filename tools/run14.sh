#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tracking.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('x_rt',round(d['config']['x_realtime'],1),'e2e',round(d['e2e']['value']),d['e2e']['ms_breakdown'],'clocks',d['clocks'])"
timeout 200 python bench.py --workload acq_b1c --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_acq_b1c.json 2> gpurun_out/bench_acq_b1c.err; cut -c1-1000 gpurun_out/bench_acq_b1c.json; tail -2 gpurun_out/bench_acq_b1c.err
timeout 100 python bench.py --workload acq_b2a --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_acq.json 2> gpurun_out/bench_acq.err; cut -c1-400 gpurun_out/bench_acq.json; tail -2 gpurun_out/bench_acq.err

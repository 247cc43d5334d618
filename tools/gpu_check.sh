#!/bin/bash
# Round-end style validation on one B200 (run under gpurun): parity tests, smoke, bench arms, acquisition lines.
# Every step has its own short timeout so that a hang cannot burn the GPU budget.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json
timeout 120 python bench.py --workload acq_b2a --steps 5 --warmup 3 > gpurun_out/bench_acq.json 2> gpurun_out/bench_acq.err; cut -c1-260 gpurun_out/bench_acq.json; tail -2 gpurun_out/bench_acq.err

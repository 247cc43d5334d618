#!/bin/bash
# developer sweep of the tracking kernel's tuning knobs on one B200 (run under gpurun)
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { echo "== $*" >> gpurun_out/sweep.log; env "$@" $B 2>> gpurun_out/sweep.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['config']['x_realtime'],1), round(d['roofline']['kernel_ms_per_launch'],2))" >> gpurun_out/sweep.log; }
run BDS_TRK_TIMING=1
run BDS_TRK_PASSES=2
run BDS_TRK_PASSES=4
run BDS_TRK_TUNE=2
run BDS_TRK_TUNE=4
run BDS_TRK_AHEAD=0
run BDS_TRK_AHEAD=2
run BDS_TRK_STAGES=2
cat gpurun_out/sweep.log

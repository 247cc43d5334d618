#!/bin/bash
O=gpurun_out/r3b
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_b2a_unit.py tests/test_gpu_tracking.py -m gpu -x -q -k "b2a or B2a" > $O/pytest_b2a.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_b2a.log
run() { name=$1; shift; echo "== $name: $*"; timeout 900 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?"; tail -c 300 $O/$name.err; python tools/bench_show.py $O/$name.json; echo; }
run bench_track_b2a python bench.py --workload track_b2a --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
run bench_track_b2a_8ch python bench.py --workload track_b2a --channels 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e

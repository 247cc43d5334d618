#!/bin/bash
O=gpurun_out/r2f
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_all.log
run() { name=$1; shift; echo "== $name: $*"; timeout 900 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?"; tail -c 300 $O/$name.err; python tools/bench_show.py $O/$name.json; echo; }
run bench_track_b2a python bench.py --workload track_b2a --steps 3 --warmup 3
run bench_dual python bench.py --workload dual --steps 3 --warmup 3
run bench_dual_cs2 python bench.py --workload dual --steps 3 --warmup 3 --b2a-cluster 2 --no-e2e --no-cpu-baseline
run bench_dual_ref python bench.py --workload dual --impl reference --steps 2 --warmup 1

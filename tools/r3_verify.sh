#!/bin/bash
# last verification of the round: whole GPU suite, smoke, default bench + reference arm
O=gpurun_out/r3v
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
run() { name=$1; shift; echo "== $name: $*"; timeout 900 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?"; tail -c 300 $O/$name.err; python tools/bench_show.py $O/$name.json; echo; }
run bench_track python bench.py
run bench_reference python bench.py --impl reference --steps 2 --warmup 1
run bench_acq_b2a python bench.py --workload acq_b2a --steps 5 --warmup 3 --no-cpu-baseline

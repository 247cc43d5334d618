#!/bin/bash
# final state of round 2, last session, on one B200: GPU suite, smoke, every bench workload
O=gpurun_out/r3final
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
run() { name=$1; shift; echo "== $name: $*"; timeout 900 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?"; tail -c 300 $O/$name.err; python tools/bench_show.py $O/$name.json; echo; }
run bench_track python bench.py
run bench_reference python bench.py --impl reference --steps 2 --warmup 1
run bench_track_12ch python bench.py --channels 12 --no-cpu-baseline
run bench_track_b2a python bench.py --workload track_b2a --steps 3 --warmup 3
run bench_dual python bench.py --workload dual --steps 3 --warmup 3
run bench_acq_b2a python bench.py --workload acq_b2a --steps 5 --warmup 3
run bench_acq_b1c python bench.py --workload acq_b1c --steps 3 --warmup 3
BDS_BENCH_ACQ_PRNS=63 run bench_acq_b1c_63prn python bench.py --workload acq_b1c --steps 2 --warmup 1 --no-cpu-baseline
run bench_track_53mhz python bench.py --fs 53e6 --no-cpu-baseline --no-e2e-file
run bench_track_53mhz_10ch python bench.py --fs 53e6 --channels 10 --no-cpu-baseline --no-e2e-file
run bench_pipeline python bench.py --workload pipeline --steps 2 --warmup 1

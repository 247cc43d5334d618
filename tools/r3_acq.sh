#!/bin/bash
# acquisition: GPU parity tests, bench of both grids, optional tile variant / 63-PRN grid
O=gpurun_out/r3a
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_acquisition.py -m gpu -x -q > $O/pytest_acq.log 2>&1; echo "pytest acq rc=$?"; tail -6 $O/pytest_acq.log
run() { name=$1; shift; echo "== $name: $*"; timeout 900 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?"; tail -c 200 $O/$name.err; python tools/bench_show.py $O/$name.json; echo; }
run bench_acq_b2a python bench.py --workload acq_b2a --steps 5 --warmup 3 --no-cpu-baseline
run bench_acq_b1c python bench.py --workload acq_b1c --steps 3 --warmup 3 --no-cpu-baseline
BDS_BENCH_ACQ_TUNE=8 run bench_acq_b2a_t8 python bench.py --workload acq_b2a --steps 5 --warmup 3 --no-cpu-baseline
BDS_BENCH_ACQ_TUNE=8 run bench_acq_b1c_t8 python bench.py --workload acq_b1c --steps 3 --warmup 3 --no-cpu-baseline
if [ "$1" = full ]; then BDS_BENCH_ACQ_PRNS=63 run bench_acq_b1c_63 python bench.py --workload acq_b1c --steps 2 --warmup 1 --no-cpu-baseline; fi

#!/bin/bash
# chip-synchronous B1C kernel at the reference's shipped 53 MHz: GPU tests of the tracking suite + bench lines
O=gpurun_out/r3t
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_tracking.py -m gpu -x -q > $O/pytest_trk.log 2>&1; echo "pytest trk rc=$?"; tail -6 $O/pytest_trk.log
run() { name=$1; shift; echo "== $name: $*"; timeout 900 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?"; tail -c 300 $O/$name.err; python tools/bench_show.py $O/$name.json; echo; }
run bench_track_53 python bench.py --fs 53e6 --no-cpu-baseline --no-e2e-file
run bench_track_53_general python bench.py --fs 53e6 --kernel general --seconds 3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-e2e-file
run bench_track python bench.py --no-cpu-baseline --no-e2e-file

#!/bin/bash
# narrow-band chip bodies: GPU tracking suite + bench lines (NB on its own body, WB unchanged)
O=gpurun_out/r3nb
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_tracking.py -m gpu -x -q -k "not one_second_trajectory and not sixty_channels" > $O/pytest_trk.log 2>&1; echo "pytest trk rc=$?"; tail -6 $O/pytest_trk.log
run() { name=$1; shift; echo "== $name: $*"; timeout 900 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?"; tail -c 300 $O/$name.err; python tools/bench_show.py $O/$name.json; echo; }
run bench_track_nb python bench.py --workload track_nb --no-cpu-baseline --no-e2e-file
run bench_track_nb_53 python bench.py --workload track_nb --fs 53e6 --no-cpu-baseline --no-e2e-file --no-e2e
run bench_track python bench.py --no-cpu-baseline --no-e2e-file --no-e2e

#!/bin/bash
# acquisition A/B over cfg.tune values (bdsgpu.h: 1 generic kernels, 2 other row tiling, 4 one stream, 8 four streams): bash tools/r3_acq_ab.sh 0 4 8
O=gpurun_out/r3ab
mkdir -p $O
run() { name=$1; shift; echo "== $name: $*"; timeout 900 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?"; tail -c 160 $O/$name.err; python tools/bench_show.py $O/$name.json; echo; }
for t in "$@"; do
BDS_BENCH_ACQ_TUNE=$t run bench_acq_b2a_t$t python bench.py --workload acq_b2a --steps 5 --warmup 3 --no-cpu-baseline --no-e2e
[ -n "$NO_B1C" ] || BDS_BENCH_ACQ_TUNE=$t run bench_acq_b1c_t$t python bench.py --workload acq_b1c --steps 3 --warmup 2 --no-cpu-baseline --no-e2e
done

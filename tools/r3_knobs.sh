#!/bin/bash
# knob A/B of the chip-synchronous kernel on the final build: passes per task, prefetch depth
O=gpurun_out/r3kn
mkdir -p $O
run() { name=$1; shift; echo "== $name: $*"; timeout 900 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?"; tail -c 200 $O/$name.err; python tools/bench_show.py $O/$name.json; echo; }
for t in "fwPassesPerTask=1" "fwPassesPerTask=3" "fwPassesPerTask=3,fwPrefetch=2" "fwPassesPerTask=4" "fwPassesPerTask=2,fwPrefetch=4"; do
BDS_BENCH_TRK_TUNING=$t run "bench_$(echo $t | tr '=,' '__')" python bench.py --no-cpu-baseline --no-e2e-file --no-e2e --steps 3 --warmup 3
done

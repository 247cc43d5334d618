#!/bin/bash
# 8-GPU lines of the last session: PRN-sharded full B1C grid and BASELINE config 5 as a joint acquisition -> tracking run
N=${1:-8}
O=gpurun_out/r3n$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() { name=$1; shift; echo "== $name: $*"; timeout 600 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" $O/$name.err | tail -c 300; python tools/bench_show.py $O/$name.json; echo; }
BDS_BENCH_ACQ_PRNS=63 run acq_b1c_63 $TR bench.py --gpus $N --workload acq_b1c --steps 3 --warmup 2 --no-cpu-baseline
run pipeline $TR bench.py --gpus $N --workload pipeline --steps 3 --warmup 2 --no-cpu-baseline

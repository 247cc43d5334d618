#!/bin/bash
# PRN-sharded acquisition on N GPUs (bds_acquire prn_lo / prn_hi): the full 63 PRN grids
N=${1:-8}
O=gpurun_out/r2n$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() { name=$1; shift; echo "== $name: $*"; timeout 600 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" $O/$name.err | tail -c 300; python tools/bench_show.py $O/$name.json; echo; }
BDS_BENCH_ACQ_PRNS=63 run acq_b1c_63 $TR bench.py --gpus $N --workload acq_b1c --steps 3 --warmup 2 --no-cpu-baseline
run acq_b2a_63 $TR bench.py --gpus $N --workload acq_b2a --steps 5 --warmup 3 --no-cpu-baseline

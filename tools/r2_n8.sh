#!/bin/bash
# 8-GPU run of the round: strong / weak scaling of B1C tracking, e2e with SMs left to NCCL, BASELINE config 5 (dual band), gather cadence
N=${1:-8}
O=gpurun_out/r2n$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() { name=$1; shift; echo "== $name: $*"; timeout 600 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" $O/$name.err | tail -c 400; python tools/bench_show.py $O/$name.json; echo; }
run track_strong $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline
BDS_BENCH_E2E_FW_CTAS=136 run track_strong_e2e136 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline
run dual $TR bench.py --gpus $N --workload dual --steps 3 --warmup 3 --no-cpu-baseline
run track_weak7 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --scaling weak --channels 7 --no-e2e
run b2a_strong $TR bench.py --gpus $N --workload track_b2a --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
run gather_1 $TR bench.py --gpus $N --steps 2 --warmup 1 --seconds 5 --no-cpu-baseline --no-e2e --gather-every 1
run gather_step $TR bench.py --gpus $N --steps 2 --warmup 1 --seconds 5 --no-cpu-baseline --no-e2e
for f in track_strong track_strong_e2e136; do python -c "import json; j=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]); print('$f', j['e2e'])"; done

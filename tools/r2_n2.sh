#!/bin/bash
# N-GPU developer run: e2e exchange paths, gather cadence, dual band
N=${1:-2}
O=gpurun_out/r2n$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() { name=$1; shift; echo "== $name: $*"; timeout 900 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" $O/$name.err | tail -c 600; python tools/bench_show.py $O/$name.json; echo; }
run track_peer $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline
BDS_BENCH_E2E_EXCHANGE=nccl run track_nccl $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline
run gather_1 $TR bench.py --gpus $N --steps 2 --warmup 1 --seconds 5 --no-cpu-baseline --no-e2e --gather-every 1
run gather_10 $TR bench.py --gpus $N --steps 2 --warmup 1 --seconds 5 --no-cpu-baseline --no-e2e --gather-every 10
run gather_100 $TR bench.py --gpus $N --steps 2 --warmup 1 --seconds 5 --no-cpu-baseline --no-e2e --gather-every 100
run gather_step $TR bench.py --gpus $N --steps 2 --warmup 1 --seconds 5 --no-cpu-baseline --no-e2e
run dual $TR bench.py --gpus $N --workload dual --steps 3 --warmup 3 --no-cpu-baseline
grep -h "e2e\b" $O/track_peer.json | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['e2e'])"
python -c "import json; j=json.loads(open('$O/track_nccl.json').read().strip().splitlines()[-1]); print(j['e2e'])"

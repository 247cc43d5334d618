#!/bin/bash
# One GPU call that records the state of every bench workload plus the ncu evidence of the two tracking kernels
# (launch list of the bench command, one --set full capture each).  Outputs under gpurun_out/r2s/.
O=gpurun_out/r2s
mkdir -p $O
run() { name=$1; shift; echo "== $name: $*"; timeout 600 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?"; tail -c 600 $O/$name.json; echo; }
run bench_track        python bench.py --steps 5 --warmup 3
run bench_reference    python bench.py --impl reference --steps 2 --warmup 1
run bench_track_b2a    python bench.py --workload track_b2a --steps 3 --warmup 3
run bench_track_12ch   python bench.py --channels 12 --steps 5 --warmup 3 --no-cpu-baseline
run bench_acq_b2a      python bench.py --workload acq_b2a --steps 5 --warmup 3
run bench_acq_b1c      python bench.py --workload acq_b1c --steps 3 --warmup 3
# launch list of the bench command (serialised, cold-cache: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_steps2_warmup1.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/launches_bench.log 2>&1
echo "launch list rc=$?"
# --set full of the dominant kernel at the bench configuration (60 ch, 30 s); second launch = after the sizing run
timeout 900 ncu --set full --import-source on --clock-control none -k regex:trk_fw_kernel -s 1 -c 1 -f -o $O/trk_fw_bench \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_fw.log 2>&1
echo "ncu fw rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:trk_b2a_unit_kernel -s 1 -c 1 -f -o $O/trk_b2a_bench \
    python bench.py --workload track_b2a --seconds 5 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_b2a.log 2>&1
echo "ncu b2a rc=$?"
ls -la $O

#!/bin/bash
# launch lists (per-kernel durations) of one acquisition call per band
O=gpurun_out/r3l
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_acq_b2a.csv \
    python bench.py --workload acq_b2a --steps 1 --warmup 0 --no-cpu-baseline > $O/b2a.log 2>&1; echo "rc=$?"
BDS_BENCH_ACQ_PRNS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_acq_b1c_2prn.csv \
    python bench.py --workload acq_b1c --steps 1 --warmup 0 --no-cpu-baseline > $O/b1c.log 2>&1; echo "rc=$?"
ls -la $O

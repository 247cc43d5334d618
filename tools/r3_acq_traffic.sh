#!/bin/bash
# DRAM traffic of the inverse passes with the caches left as the previous launch left them (ncu flushes them by default,
# which hides the L2 reuse between the row pass and the column pass)
O=gpurun_out/r3tr
mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
BDS_BENCH_ACQ_PRNS=2 timeout 600 ncu --metrics $M --cache-control none --clock-control none -k regex:acq_inv_ -s 40 -c 12 --csv --log-file $O/traffic_b1c.csv \
    python bench.py --workload acq_b1c --steps 1 --warmup 0 --no-cpu-baseline > $O/b1c.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics $M --cache-control none --clock-control none -k regex:acq_inv_ -s 40 -c 12 --csv --log-file $O/traffic_b2a.csv \
    python bench.py --workload acq_b2a --steps 1 --warmup 0 --no-cpu-baseline > $O/b2a.log 2>&1; echo "rc=$?"

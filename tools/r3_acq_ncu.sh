#!/bin/bash
O=gpurun_out/r3n2
mkdir -p $O
# one --set full capture of each inverse kernel, B2a (P = 2^19) and B1C (P = 2^22), taken from the bench command itself
timeout 600 ncu --set full --import-source on --clock-control none -k regex:acq_inv_ -s 40 -c 2 -f -o $O/acq_inv_b2a \
    python bench.py --workload acq_b2a --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_b2a.log 2>&1; echo "rc=$?"
BDS_BENCH_ACQ_PRNS=2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:acq_inv_ -s 40 -c 2 -f -o $O/acq_inv_b1c \
    python bench.py --workload acq_b1c --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_b1c.log 2>&1; echo "rc=$?"
ls -la $O

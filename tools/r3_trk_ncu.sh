#!/bin/bash
# launch list of the default bench command and one --set full capture of trk_fw_kernel per geometry (short records)
O=gpurun_out/r3k
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_steps2_warmup1.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/launches.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trk_fw_kernel -c 1 -f -o $O/trk_fw_g53 \
    python bench.py --fs 53e6 --steps 1 --warmup 0 --seconds 5 --no-e2e --no-cpu-baseline > $O/ncu_g53.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trk_fw_kernel -c 1 -f -o $O/trk_fw_g99 \
    python bench.py --steps 1 --warmup 0 --seconds 5 --no-e2e --no-cpu-baseline > $O/ncu_g99.log 2>&1; echo "rc=$?"
ls -la $O

#!/bin/bash
# ncu --set full of one trk_fw_kernel launch (open loop: same correlator work as the closed loop, deterministic)
mkdir -p gpurun_out/r2
MODE=${1:-open}
if [ "$MODE" = open ]; then
  timeout 60 python tools/variant_check.py closed gpurun_out/r2/closed_ncu 5 > /dev/null 2>&1
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:trk_fw_kernel -s 1 -c 1 -f -o gpurun_out/r2/fw_open \
     python tools/variant_check.py open gpurun_out/r2/closed_ncu.npz gpurun_out/r2/open_ncu 2>&1 | tail -3
else
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:trk_fw_kernel -s 1 -c 1 -f -o gpurun_out/r2/fw_closed \
     python tools/variant_check.py closed gpurun_out/r2/closed_ncu 5 2>&1 | tail -3
fi
ls -la gpurun_out/r2/*.ncu-rep

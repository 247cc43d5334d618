#!/bin/bash
# Builds the opt-in variants of the tracking kernel next to libbdsgpu.so (developer A/B runs, see tools/variant_run.sh):
#   f2      chip tail on packed fp32 pairs (FFMA2)                      -DBDS_FAST_F32X2=1
#   rec     rank search through per-bin threshold records               -DBDS_FAST_BINREC=1
#   f2rec   both
#   w20     20 compute warps at 96 registers, 3 stages (5 warps per SM sub-partition instead of 4)
#   w20f2rec  all of the above
# usage: bash tools/variant_build.sh [names...]      (default: all)
set -e
cd "$(dirname "$0")/.."
PKG=bds-3-b1c-b2a-sdr-receiver_b200
declare -A FLAGS=(
  [f2]="-DBDS_FAST_F32X2=1"
  [rec]="-DBDS_FAST_BINREC=1"
  [f2rec]="-DBDS_FAST_F32X2=1 -DBDS_FAST_BINREC=1"
  [w20]="-DBDS_FW_COMPUTE_WARPS=20 -DBDS_FW_SETMAXNREG=96 -DBDS_FW_STAGES=3"
  [w20f2rec]="-DBDS_FW_COMPUTE_WARPS=20 -DBDS_FW_SETMAXNREG=96 -DBDS_FW_STAGES=3 -DBDS_FAST_F32X2=1 -DBDS_FAST_BINREC=1"
)
names=("$@")
[ ${#names[@]} -eq 0 ] && names=(f2 rec f2rec w20 w20f2rec)
for n in "${names[@]}"; do
  BDS_LIB_NAME=libbds_$n.so BDS_OBJ_SUFFIX=_$n BDS_EXTRA_FLAGS="${FLAGS[$n]}" python $PKG/build.py > /dev/null
  echo "built $PKG/libbds_$n.so  (${FLAGS[$n]})"
  grep -A3 "13trk_fw_kernel" $PKG/build_$n/bds_track.cu.ptxas.log | grep "Used" || true
done

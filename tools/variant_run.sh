#!/bin/bash
# Developer A/B run of alternative tracking-kernel builds on one B200 (under gpurun).  Every step has a short timeout.
# usage: bash tools/variant_run.sh SECONDS "closed-loop libs" "open-loop-only (ablation) libs"
mkdir -p gpurun_out/variants
SECS=${1:-5}
CLOSED=${2:-"libbdsgpu.so"}
OPEN=${3:-""}
LOG=gpurun_out/variants/variants.log
: > $LOG
first=""
for lib in $CLOSED; do
  BDS_LIB_NAME=$lib timeout 15 python tools/variant_check.py closed gpurun_out/variants/closed_$lib $SECS >> $LOG 2>&1 || echo "{\"lib\": \"$lib\", \"failed\": \"closed rc=$?\"}" >> $LOG
  [ -z "$first" ] && first=gpurun_out/variants/closed_$lib.npz
done
for lib in $CLOSED $OPEN; do
  [ -f "$first" ] || break
  BDS_LIB_NAME=$lib timeout 15 python tools/variant_check.py open $first gpurun_out/variants/open_$lib >> $LOG 2>&1 || echo "{\"lib\": \"$lib\", \"failed\": \"open rc=$?\"}" >> $LOG
done
cat $LOG

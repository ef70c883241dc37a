#!/bin/bash
# compute-sanitizer over small workloads (the reference has no sanitizer runs; the in-kernel epoch handshake and the
# out-of-place buffer rotation read by a neighbour are what racecheck / memcheck exist for).  On the GPU box:
#   bash tools/sanitize.sh [tag]   ->  gpurun_out/<tag>_sanitize_{memcheck,racecheck,initcheck}.log + a one-line summary each
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; T=${1:-r2}; mkdir -p $O
for tool in memcheck racecheck initcheck; do
  for what in single slabs; do
    log=$O/${T}_sanitize_${tool}_${what}.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_scene.py $what > $log 2>&1
    echo "$tool $what: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
  done
done

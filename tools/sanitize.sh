#!/bin/bash
# compute-sanitizer over every kernel family on small shapes (run on a GPU box: gpurun -- bash tools/sanitize.sh).
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards (transpose tiles, reduction tiles);
# synccheck: barrier misuse.  Summaries land in gpurun_out/sanitize_*.log; copy the tails into profiles/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
rc=0
for tool in memcheck racecheck synccheck; do
  for set in small big multi; do
    log=gpurun_out/sanitize_${tool}_${set}.log
    timeout 900 $CS --tool $tool --error-exitcode 7 python tools/sanitize_cases.py $set > $log 2>&1
    code=$?
    echo "== $tool $set: exit $code: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
    [ $code -ne 0 ] && rc=1
  done
done
exit $rc

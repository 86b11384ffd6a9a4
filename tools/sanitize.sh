#!/bin/bash
# compute-sanitizer evidence for profiles/ (run under gpurun, 1 GPU): memcheck, racecheck and synccheck over one small invocation of
# every kernel family (tools/sanitizer_cases.py).  Usage: bash tools/sanitize.sh <tag> [per-tool timeout seconds]
TAG=${1:-r2}
LIMIT=${2:-600}
mkdir -p gpurun_out
for TOOL in memcheck racecheck synccheck; do
  timeout $LIMIT compute-sanitizer --tool $TOOL --print-limit 30 --error-exitcode 1 \
      python tools/sanitizer_cases.py > gpurun_out/sanitizer_${TOOL}_${TAG}.log 2>&1
  echo "== $TOOL exit code $?" >> gpurun_out/sanitizer_${TOOL}_${TAG}.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|exit code|SANITIZER_CASES_DONE" gpurun_out/sanitizer_${TOOL}_${TAG}.log | tail -4
done

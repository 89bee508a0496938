#!/bin/bash
# compute-sanitizer over tools/sanitize_run.py: memcheck, racecheck, initcheck, synccheck.  Summary -> gpurun_out/sanitizer.txt
set -u
mkdir -p gpurun_out
out=gpurun_out/sanitizer.txt
: > $out
echo "== plain run" | tee -a $out
python tools/sanitize_run.py 2>&1 | tee gpurun_out/sanitize_plain.log | tee -a $out
for tool in memcheck racecheck initcheck synccheck; do
  echo "== compute-sanitizer --tool $tool" | tee -a $out
  extra=""
  [ $tool = racecheck ] && extra="--racecheck-report all"
  [ $tool = initcheck ] && extra="--track-unused-memory no"
  timeout 420 compute-sanitizer --tool $tool $extra --error-exitcode 9 --log-file gpurun_out/sanitize_$tool.log python tools/sanitize_run.py > gpurun_out/sanitize_$tool.out 2>&1
  rc=$?
  echo "rc=$rc" | tee -a $out
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -20 | tee -a $out
  if diff -q <(grep checksum gpurun_out/sanitize_plain.log) <(grep checksum gpurun_out/sanitize_$tool.out) > /dev/null; then echo "checksums equal the plain run" | tee -a $out; else echo "CHECKSUMS DIFFER (or run incomplete)" | tee -a $out; tail -5 gpurun_out/sanitize_$tool.out | tee -a $out; fi
done

#!/bin/bash
# compute-sanitizer memcheck + racecheck on a small invocation of every kernel (smoke path).
OUT=gpurun_out/${1:-san}; mkdir -p $OUT
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" $OUT/$tool.log | tail -3
done

#!/bin/bash
# Dev helper (run under gpurun): every GPU test in its own process with a timeout, so that one
# faulting kernel cannot poison the rest; logs land in gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/gpu_tests.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv >> $LOG 2>&1
FILTER=${1:-}
TESTS=$(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::" | grep -E "${FILTER}")
PASS=0; FAIL=0
for t in $TESTS; do
  echo "=== $t" >> $LOG
  timeout 600 python -m pytest "$t" -x -q -s -m gpu --tb=short -p no:cacheprovider > gpurun_out/_one.log 2>&1
  rc=$?
  tail -n 60 gpurun_out/_one.log >> $LOG
  if [ $rc -eq 0 ]; then PASS=$((PASS+1)); else FAIL=$((FAIL+1)); echo "FAILED($rc): $t" >> $LOG; fi
done
echo "SUMMARY pass=$PASS fail=$FAIL" | tee -a $LOG
grep -E "^FAILED" $LOG
exit 0

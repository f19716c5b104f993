#!/bin/bash
# compute-sanitizer over the library's kernels at tiny batch.  Runs on the GPU box; one summary line
# per (tool, target) is appended to gpurun_out/sanitizer_summary.txt, full logs next to it.
# usage: [TARGETS="bf16 ..."] tools/sanitize.sh [tool ...]   (default: memcheck racecheck synccheck over fp32 generic aux bf16)
set -u
TOOLS=${@:-memcheck racecheck synccheck}
mkdir -p gpurun_out
: > gpurun_out/sanitizer_summary.txt
for tool in $TOOLS; do
  for target in ${TARGETS:-fp32 generic aux bf16}; do
    # racecheck does not model the async proxy (bulk copies, tcgen05) of the bf16 kernels: memcheck / synccheck only there
    if [ "$tool" = racecheck ] && { [ "$target" = bf16 ] || [ "$target" = generic ]; }; then continue; fi
    log=gpurun_out/sanitizer_${tool}_${target}.log
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python tools/sanitize_target.py $target > $log 2>&1
    rc=$?
    echo "$tool $target rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize target' $log | tr '\n' ' ')" | tee -a gpurun_out/sanitizer_summary.txt
  done
done

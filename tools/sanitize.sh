#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the GPU parity tests that reach every kernel variant; the
# full-size / exhaustive tests are left out (they only repeat the same kernels on more data).  Run under gpurun:
#   bash tools/sanitize.sh <tag>      -> gpurun_out/<tag>_sanitizer_{memcheck,racecheck,synccheck}.txt
tag=${1:-r02}
SEL='not (exhaustive or full_size or 1m or standard_normal or baseline_sizes or many_keys or config2 or mt19937_stream or coscheduled or host_pipe or dropin or between_processes or local_two_devices or counter_word3)'
for tool in memcheck racecheck synccheck; do
  ( time timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$SEL" ) \
      > gpurun_out/${tag}_sanitizer_${tool}.txt 2>&1
  echo "$tool rc=$? $(grep -E 'passed|failed' gpurun_out/${tag}_sanitizer_${tool}.txt | tail -1) $(grep 'ERROR SUMMARY' gpurun_out/${tag}_sanitizer_${tool}.txt | tail -1)"
done

#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/sanitizer_workload.py > gpurun_out/sanitizer_plain.log 2>&1; echo "plain exit $?"; tail -2 gpurun_out/sanitizer_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitizer_workload.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitizer_workload.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 gpurun_out/sanitizer_racecheck.log

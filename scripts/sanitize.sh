#!/bin/bash
# compute-sanitizer over scripts/sanitize_case.py, ONE tool per gpurun call (B200_PROFILING.md):
#   gpurun --timeout 900 -- 'bash scripts/sanitize.sh racecheck'
# The plain run comes first (the sanitizer only runs on a program that exited 0 without it);
# the log lands in gpurun_out/sanitizer_<tool>.log and its summary is copied to profiles/.
set -u
tool=${1:-memcheck}
mkdir -p gpurun_out
python scripts/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
timeout ${SANITIZE_TIMEOUT:-1500} compute-sanitizer --tool "$tool" --print-limit 50 \
    --log-file gpurun_out/sanitizer_${tool}.log python scripts/sanitize_case.py \
    > gpurun_out/sanitize_${tool}_stdout.log 2>&1
rc=$?
echo "compute-sanitizer --tool $tool exit code $rc"
tail -5 gpurun_out/sanitizer_${tool}.log
tail -3 gpurun_out/sanitize_${tool}_stdout.log
exit $rc

#!/bin/bash
# GPU test suite in separate processes (a CUDA fault in one file does not poison the next)
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.jsonl
for f in tests/test_gpu_footprints.py "tests --ignore=tests/test_gpu_footprints.py"; do
    name=$(echo $f | tr -c 'a-zA-Z0-9' '_' | cut -c1-40)
    timeout 1500 python -m pytest $f -m gpu -q > gpurun_out/pytest_$name.log 2>&1
    echo "pytest $f exit $?"
    grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_$name.log | tail -15
done

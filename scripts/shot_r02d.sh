#!/bin/bash
# GPU tests in separate processes, default bench line, combine-kernel benches.
set -u
o=gpurun_out; mkdir -p $o
tag=${1:-r02d}
bash scripts/shot_tests.sh > $o/${tag}_tests.txt 2>&1
tail -12 $o/${tag}_tests.txt
timeout 600 python bench.py > $o/${tag}_bench_default.json 2> $o/${tag}_bench_default.err
echo "bench exit $?"; tail -3 $o/${tag}_bench_default.err; cut -c1-200 $o/${tag}_bench_default.json
[ -f scripts/bench_combine.py ] && timeout 300 python scripts/bench_combine.py > $o/${tag}_combine.jsonl 2> $o/${tag}_combine.err
tail -3 $o/${tag}_combine.err
du -sh $o

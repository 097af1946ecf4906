#!/bin/bash
# One GPU call that refreshes every measured artefact under profiles/ (run through gpurun):
#   tests, default bench line, reduced config sweep, new-kernel benches, ncu launch list and
#   one full ncu capture of the hot kernel (each ncu pass after the same command exited 0).
o=gpurun_out
mkdir -p $o
(time timeout 400 python -m pytest tests -m gpu -x -q) > $o/r01f_pytest.log 2>&1
timeout 240 python bench.py > $o/r01f_bench_default.json 2> $o/r01f_bench_default.err
run() { timeout 200 python bench.py --cpu-frames 0 --e2e-steps 0 --warmup 5 "$@" | tail -1 >> $o/r01f_sweep.jsonl; }
: > $o/r01f_sweep.jsonl
run --config 1 --frames 1 --steps 400
run --config 2 --frames 64 --steps 100 --resampling 1
run --config 2 --frames 64 --steps 100 --resampling 3
run --config 2 --frames 64 --steps 50 --dtype float64
run --config 3 --frames 256 --steps 30
run --config 4 --frames 512 --steps 15 --dtype float64
run --config 4 --frames 512 --steps 15
run --config 2 --frames 64 --steps 100 --combine abba
run --config 2 --frames 64 --steps 100 --combine abba_median
timeout 120 python scripts/bench_stack_median.py > $o/r01f_stack_median.jsonl 2> $o/r01f_stack_median.err
timeout 120 python scripts/bench_prereduce.py > $o/r01f_prereduce.jsonl 2> $o/r01f_prereduce.err
timeout 120 python scripts/bench_median.py > $o/r01f_median.jsonl 2> $o/r01f_median.err
# ncu passes (numbers printed under ncu are never bench values)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $o/r01f_launches.csv python bench.py --steps 3 --warmup 3 --cpu-frames 0 --e2e-steps 0 \
    > $o/r01f_ncu_launch.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_rectwv -s 3 -c 1 \
    -o $o/r01f_rectwv -f python bench.py --steps 2 --warmup 3 --frames 64 --cpu-frames 0 --e2e-steps 0 \
    > $o/r01f_ncu_full.log 2>&1
tail -n 3 $o/r01f_pytest.log
cut -c1-200 $o/r01f_bench_default.json
wc -l $o/r01f_sweep.jsonl $o/r01f_launches.csv
ls -la $o/r01f_rectwv.ncu-rep

#!/bin/bash
# Round-2 measurement pass: the GPU test suite (separate processes), the default bench line,
# the ncu launch list of the same command and full ncu captures of the hot kernel for
# float32 / float64 / LR / nearest (each ncu pass after the same command exited 0 without it).
set -u
o=gpurun_out; mkdir -p $o
tag=${1:-r02c}
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $o/${tag}_gpu.txt 2>&1
if [ "${2:-tests}" = "tests" ]; then
    bash scripts/shot_tests.sh > $o/${tag}_tests.txt 2>&1
    tail -8 $o/${tag}_tests.txt
fi
timeout 400 python bench.py > $o/${tag}_bench_default.json 2> $o/${tag}_bench_default.err
echo "bench exit $?"; cut -c1-300 $o/${tag}_bench_default.json
run() { timeout 200 python bench.py --cpu-frames 0 --e2e-steps 0 --warmup 5 "$@" 2>> $o/${tag}_sweep.err | tail -1 >> $o/${tag}_sweep.jsonl; }
: > $o/${tag}_sweep.jsonl
run --config 1 --frames 1 --steps 400
run --config 2 --frames 64 --steps 100 --resampling 1
run --config 2 --frames 64 --steps 100 --resampling 3
run --config 2 --frames 64 --steps 50 --dtype float64
run --config 3 --frames 256 --steps 30
run --config 4 --frames 512 --steps 15 --dtype float64
run --config 4 --frames 512 --steps 15
run --config 4b --frames 512 --steps 15
wc -l $o/${tag}_sweep.jsonl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $o/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --cpu-frames 0 --e2e-steps 0 \
    > $o/${tag}_ncu_launch.log 2>&1
echo "launch list exit $?"
cap() {  # name, bench flags
    name=$1; shift
    cmd="python bench.py --steps 2 --warmup 3 --cpu-frames 0 --e2e-steps 0 $*"
    $cmd > $o/${tag}_${name}_plain.log 2>&1 &&
    timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_rectwv -s 3 -c 1 \
        -o $o/${tag}_${name} -f $cmd > $o/${tag}_${name}_ncu.log 2>&1
    echo "ncu $name exit $?"
    # gpurun_out travels back only below 64 MiB: keep the text pages, drop the report
    # (the float32 one stays: its source page is read here)
    ncu -i $o/${tag}_${name}.ncu-rep --page raw --csv > $o/${tag}_${name}_raw.csv 2>/dev/null
    ncu -i $o/${tag}_${name}.ncu-rep --page details > $o/${tag}_${name}_details.txt 2>/dev/null
    python scripts/ncu_summary.py $o/${tag}_${name}.ncu-rep 30 > $o/${tag}_${name}_summary.txt 2>&1
    [ "$name" = "f32" ] || rm -f $o/${tag}_${name}.ncu-rep
}
cap f32 --frames 64
cap f64 --frames 64 --dtype float64
cap lr --config 4 --frames 64
cap r1 --frames 64 --resampling 1
ls -la $o/*.ncu-rep; du -sh $o

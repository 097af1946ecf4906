#!/bin/bash
# full ncu captures of k_rectwv for the configurations bench.py can be asked for beyond the four of
# scripts/shot_measure.sh, so that roofline.traffic is never null (text pages only travel back)
set -u
o=gpurun_out; mkdir -p $o
tag=${1:-r02x}
cap() {  # name, bench flags
    name=$1; shift
    cmd="python bench.py --steps 2 --warmup 3 --cpu-frames 0 --e2e-steps 0 $*"
    $cmd > $o/${tag}_${name}_plain.log 2>&1 &&
    timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_rectwv -s 3 -c 1 \
        -o $o/${tag}_${name} -f $cmd > $o/${tag}_${name}_ncu.log 2>&1
    echo "ncu $name exit $?"
    python scripts/ncu_summary.py $o/${tag}_${name}.ncu-rep 12 > $o/${tag}_${name}_summary.txt 2>&1
    rm -f $o/${tag}_${name}.ncu-rep
}
cap c3 --config 3 --frames 64
cap c4f64 --config 4 --frames 64 --dtype float64
cap c4b --config 4b --frames 64
cap r3 --frames 64 --resampling 3
cap c1 --config 1 --frames 1
timeout 600 python -m pytest tests/test_gpu_abba_combine.py -m gpu -q -k "sigmaclip or combine_abba" 2>&1 | tail -2
timeout 300 python scripts/bench_combine.py 2>/dev/null | grep sigmaclip | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['kernel'], d['frames'], '%.3f ms frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))"

#!/bin/bash
# One full ncu capture of the hot kernel (after the same command exited 0 without ncu).
#   gpurun -- 'bash scripts/shot_ncu.sh r02b "--dtype float32"'
set -u
tag=${1:-r02}; extra=${2:-}
o=gpurun_out; mkdir -p $o
cmd="python bench.py --steps 2 --warmup 3 --frames 64 --cpu-frames 0 --e2e-steps 0 $extra"
$cmd > $o/${tag}_plain.log 2>&1 &&
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_rectwv -s 3 -c 1 \
    -o $o/${tag}_rectwv -f $cmd > $o/${tag}_ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 $o/${tag}_ncu_full.log; ls -la $o/${tag}_rectwv.ncu-rep

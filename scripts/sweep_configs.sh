#!/bin/bash
# Device-resident throughput of every BASELINE config on one GPU.
# usage: scripts/sweep_configs.sh out.jsonl
out=${1:-gpurun_out/sweep.jsonl}
: > "$out"
run() { python bench.py --cpu-frames 0 --e2e-steps 0 --warmup 5 "$@" | tail -1 >> "$out"; }
run --config 1 --frames 1 --steps 400
run --config 1 --frames 1 --steps 400 --resampling 1
run --config 2 --frames 64 --steps 100 --resampling 1
run --config 2 --frames 64 --steps 100 --resampling 3
run --config 2 --frames 64 --steps 50 --dtype float64
run --config 3 --frames 256 --steps 30
run --config 3 --frames 256 --steps 30 --resampling 1
run --config 4 --frames 512 --steps 15 --dtype float64
run --config 4b --frames 512 --steps 15 --dtype float64
run --config 4 --frames 512 --steps 15
run --config 5 --frames 1024 --steps 8

#!/bin/bash
# A/B of experiment variants of the library on one box: every variant twice, alternating.
#   gpurun -- 'bash scripts/ab.sh slim slim:EMIRWV_FILL=kernel'   (names from
#   scripts/build_variants.py, "default" is the product library; name:VAR=value:... sets
#   environment switches of the library for that run).  Extra bench flags: AB_FLAGS="--dtype float64".
set -u
mkdir -p gpurun_out
out=gpurun_out/ab_$(date +%H%M%S).jsonl
for rep in 1 2; do
    for spec in "$@"; do
        v=${spec%%:*}; envs=""; [ "$spec" != "$v" ] && envs=$(echo "${spec#*:}" | tr ':' ' ')
        lib=libemirwv_$v.so; [ "$v" = default ] && lib=libemirwv.so
        env $envs EMIRWV_LIB=$lib AB_LABEL="$spec" python bench.py --steps 40 --warmup 5 --cpu-frames 0 --e2e-steps 0 ${AB_FLAGS:-} >> $out 2>> gpurun_out/ab.err || echo "$v failed"
    done
done
python - $out <<'PY'
import json, sys
for line in open(sys.argv[1]):
    try: d = json.loads(line)
    except Exception: continue
    t = d.get('tuning', {})
    print('%-34s %s st %d ctas %d smem %6d win %2dx%3d tk %3d | %7.0f f/s frac %.3f | own %.3f ms frac %.3f | clk %s %s' % (
        t.get('label') or t.get('library'), d['dtype'], t.get('stages'), t.get('ctas_per_sm'), t.get('smem_bytes'),
        t.get('window_rows'), t.get('window_cols'), t.get('tile_k'), d['value'], d['roofline']['frac'],
        d['roofline']['k_rectwv_alone']['launch_ms'], d['roofline']['k_rectwv_alone']['frac'],
        d['clocks']['sm_mhz'] if d.get('clocks') else None, d['clocks']['reasons'] if d.get('clocks') else None))
PY
tail -3 gpurun_out/ab.err 2>/dev/null

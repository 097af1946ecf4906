#!/bin/bash
# Round-2 first GPU call: the whole GPU test suite, the default bench line, the experiment
# variants of the hot kernel (short runs), float64, then one compute-sanitizer tool.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
python bench.py > gpurun_out/r02a_bench_default.json 2> gpurun_out/r02a_bench_default.err
echo "bench exit $?"; tail -c 600 gpurun_out/r02a_bench_default.json
short="--steps 40 --warmup 5 --cpu-frames 0 --e2e-steps 0"
for run in "default:" "stages2:--stages 2" ; do
    name=${run%%:*}; extra=${run#*:}
    python bench.py $short $extra >> gpurun_out/r02a_variants.jsonl 2>> gpurun_out/r02a_variants.err
done
for lib in x3; do
    EMIRWV_LIB=libemirwv_$lib.so python bench.py $short >> gpurun_out/r02a_variants.jsonl 2>> gpurun_out/r02a_variants.err
done
python bench.py $short --dtype float64 >> gpurun_out/r02a_variants.jsonl 2>> gpurun_out/r02a_variants.err
EMIRWV_LIB=libemirwv_x3.so python bench.py $short --dtype float64 >> gpurun_out/r02a_variants.jsonl 2>> gpurun_out/r02a_variants.err
python bench.py $short --config 4 >> gpurun_out/r02a_variants.jsonl 2>> gpurun_out/r02a_variants.err
python bench.py $short --resampling 1 >> gpurun_out/r02a_variants.jsonl 2>> gpurun_out/r02a_variants.err
python - <<'PY'
import json
for line in open('gpurun_out/r02a_variants.jsonl'):
    try: d = json.loads(line)
    except Exception: continue
    t = d.get('tuning', {})
    print(t.get('library'), d['dtype'], d['config']['workload'][:18], 'stages', t.get('stages'), 'ctas', t.get('ctas_per_sm'),
          'smem', t.get('smem_bytes'), 'win', t.get('window_rows'), 'tile_k', t.get('tile_k'),
          '| %.0f f/s frac %.3f | own %.3f ms frac %.3f' % (d['value'], d['roofline']['frac'],
          d['roofline']['k_rectwv_alone']['launch_ms'], d['roofline']['k_rectwv_alone']['frac']))
PY
SANITIZE_TIMEOUT=600 bash scripts/sanitize.sh ${1:-racecheck}

#!/bin/bash
# host zero-fill + cooperative bad-pixel kernel: tests, default bench, side-kernel benches
set -u
o=gpurun_out; mkdir -p $o
tag=${1:-r02f}
timeout 900 python -m pytest tests/test_gpu_abba_combine.py tests/test_prereduce.py tests/test_gpu_parity.py -m gpu -q > $o/${tag}_tests.txt 2>&1
tail -4 $o/${tag}_tests.txt
timeout 600 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err
echo "bench exit $?"; tail -3 $o/${tag}_bench.err
python - <<PY
import json
d=json.load(open('$o/${tag}_bench.json'))
e=d['e2e']
print('value %.0f frac %.3f | e2e %.0f all-dma %.0f sparse %.0f raw %.0f abba %.0f (f32 in %.0f) | dropin %.2f ms' % (d['value'], d['roofline']['frac'], e['value'], e['all_bytes_over_pcie']['value'], e['zeroed_buffer']['value'], e['raw_u16']['value'], e['abba_product']['value'], e['abba_product']['float32_reduced_frames_in'], d['dropin_single_frame']['ms_per_call']))
PY
for t in 1 2 4 6; do EMIRWV_HOST_ZERO_THREADS=$t timeout 300 python bench.py --steps 20 --cpu-frames 0 2>>$o/${tag}_threads.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('zero threads $t: e2e %.0f' % d['e2e']['value'])"; done
timeout 300 python scripts/bench_combine.py > $o/${tag}_combine.jsonl 2> $o/${tag}_combine.err
tail -2 $o/${tag}_combine.jsonl | cut -c1-300

#!/bin/bash
set -u
o=gpurun_out; mkdir -p $o
tag=${1:-r02g}
timeout 900 python -m pytest tests/test_gpu_guards.py tests/test_gpu_abba_combine.py -m gpu -q > $o/${tag}_tests.txt 2>&1
tail -6 $o/${tag}_tests.txt
for t in 3 1 2 0 3; do EMIRWV_HOST_ZERO_THREADS=$t timeout 300 python bench.py --steps 20 --cpu-frames 0 2>>$o/${tag}_threads.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('zero threads $t: e2e %.0f all-dma %.0f sparse %.0f raw %.0f abba %.0f' % (e['value'], e['all_bytes_over_pcie']['value'], e['zeroed_buffer']['value'], e['raw_u16']['value'], e['abba_product']['value']))"; done

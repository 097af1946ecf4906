#!/bin/bash
# the reworked pre-reduction (tests), A/B of the record prefetch, side-kernel benches, default bench
set -u
o=gpurun_out; mkdir -p $o
tag=${1:-r02e}
timeout 900 python -m pytest tests/test_gpu_abba_combine.py tests/test_prereduce.py -m gpu -q > $o/${tag}_tests.txt 2>&1
tail -4 $o/${tag}_tests.txt
bash scripts/ab.sh slim recpf > $o/${tag}_ab.txt 2>&1
cat $o/${tag}_ab.txt
AB_FLAGS="--config 3 --frames 64" bash scripts/ab.sh slim recpf > $o/${tag}_ab3.txt 2>&1
cat $o/${tag}_ab3.txt
timeout 300 python scripts/bench_combine.py > $o/${tag}_combine.jsonl 2> $o/${tag}_combine.err
tail -3 $o/${tag}_combine.err
timeout 600 python bench.py --cpu-frames 0 --steps 50 > $o/${tag}_bench.json 2> $o/${tag}_bench.err
python - <<PY
import json
d=json.load(open('$o/${tag}_bench.json'))
print(d['value'], {k:(v['value'] if isinstance(v,dict) and 'value' in v else v) for k,v in d['e2e'].items() if k in ('value','zeroed_buffer','raw_u16','abba_product')}, d['e2e']['abba_product'].get('float32_reduced_frames_in'), d['dropin_single_frame'])
PY

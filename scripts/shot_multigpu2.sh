#!/bin/bash
# N-GPU call: bench with the collective block only (no end-to-end legs) + the multi-GPU tests
set -u
n=${1:-8}; tag=${2:-r02p}
o=gpurun_out; mkdir -p $o
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --steps 30 --warmup 3 --e2e-steps 0 > $o/${tag}_bench_${n}.json 2> $o/${tag}_bench_${n}.err
echo "bench exit $?"; tail -3 $o/${tag}_bench_${n}.err
python - <<PY
import json
d = json.loads(open("$o/${tag}_bench_${n}.json").read().strip().splitlines()[-1])
print("value %.0f" % d["value"])
for k, v in d["collective"].items():
    print("  ", k, json.dumps({kk: vv for kk, vv in v.items() if kk != "how"} if isinstance(v, dict) else v)[:500])
PY
timeout 240 python -m pytest tests/test_gpu_multigpu.py -m gpu -q 2>&1 | tail -3

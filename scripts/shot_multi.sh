#!/bin/bash
# N-GPU call: copy-bandwidth probe, bench with the collective block (NUMA placement on / off),
# the multi-GPU ABBA test.   gpurun --gpus N -- 'bash scripts/shot_multi.sh N tag'
set -u
n=${1:-2}; tag=${2:-r02m}
o=gpurun_out; mkdir -p $o
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
nvidia-smi topo -m > $o/${tag}_topo_${n}.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" > $o/${tag}_cpu_${n}.txt 2>&1
timeout 300 bash -c "$(declare -f run); n=$n; run 29512 scripts/pcie_probe.py" > $o/${tag}_pcie_${n}.json 2> $o/${tag}_pcie_${n}.err
echo "probe exit $?"; cut -c1-1500 $o/${tag}_pcie_${n}.json
timeout 600 bash -c "$(declare -f run); n=$n; run 29513 bench.py --gpus $n --steps 30 --warmup 3 --e2e-steps 3" > $o/${tag}_bench_${n}.json 2> $o/${tag}_bench_${n}.err
echo "bench exit $?"; tail -3 $o/${tag}_bench_${n}.err
EMIRWV_NUMA_BIND=0 timeout 600 bash -c "$(declare -f run); n=$n; run 29514 bench.py --gpus $n --steps 30 --warmup 3 --e2e-steps 3 --no-collective" > $o/${tag}_bench_${n}_nobind.json 2> $o/${tag}_bench_${n}_nobind.err
echo "bench (no bind) exit $?"; tail -3 $o/${tag}_bench_${n}_nobind.err
timeout 900 python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_abba_combine.py tests/test_prereduce.py -m gpu -q > $o/${tag}_tests_${n}.txt 2>&1
tail -3 $o/${tag}_tests_${n}.txt
python - <<PY
import json
for name in ("$o/${tag}_bench_${n}.json", "$o/${tag}_bench_${n}_nobind.json"):
    try:
        d = json.loads(open(name).read().strip().splitlines()[-1])
    except Exception as exc:
        print(name, "unreadable", exc); continue
    e = d["e2e"]
    print(name, "value %.0f" % d["value"], "e2e %.0f" % e["value"], "sparse %.0f" % e["zeroed_buffer"]["value"],
          "raw %.0f" % e.get("raw_u16", {}).get("value", 0), "abba %.0f" % e.get("abba_product", {}).get("value", 0), d.get("host_placement"))
    if d.get("collective"):
        for k, v in d["collective"].items():
            print("  ", k, json.dumps(v)[:400])
PY

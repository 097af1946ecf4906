#!/bin/bash
# N-GPU call, second pass: copy probe (with the device -> host only leg), bench with the
# collective block, and the end-to-end figures with one host zero-fill thread per rank.
set -u
n=${1:-8}; tag=${2:-r02n}
o=gpurun_out; mkdir -p $o
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 300 bash -c "$(declare -f run); n=$n; run 29512 scripts/pcie_probe.py" > $o/${tag}_pcie_${n}.json 2> $o/${tag}_pcie_${n}.err
echo "probe exit $?"; cut -c1-2500 $o/${tag}_pcie_${n}.json
timeout 600 bash -c "$(declare -f run); n=$n; run 29513 bench.py --gpus $n --steps 30 --warmup 3 --e2e-steps 3" > $o/${tag}_bench_${n}.json 2> $o/${tag}_bench_${n}.err
echo "bench exit $?"; tail -3 $o/${tag}_bench_${n}.err
EMIRWV_HOST_ZERO_THREADS=1 timeout 600 bash -c "$(declare -f run); n=$n; run 29514 bench.py --gpus $n --steps 30 --warmup 3 --e2e-steps 3 --no-collective" > $o/${tag}_bench_${n}_t1.json 2> $o/${tag}_bench_${n}_t1.err
echo "bench (1 zero thread) exit $?"; tail -3 $o/${tag}_bench_${n}_t1.err
python - <<PY
import json
for name in ("$o/${tag}_bench_${n}.json", "$o/${tag}_bench_${n}_t1.json"):
    try:
        d = json.loads(open(name).read().strip().splitlines()[-1])
    except Exception as exc:
        print(name, "unreadable", exc); continue
    e = d["e2e"]
    print(name, "value %.0f" % d["value"], "e2e %.0f" % e["value"], "all-dma %.0f" % e["all_bytes_over_pcie"]["value"], "sparse %.0f" % e["zeroed_buffer"]["value"],
          "raw %.0f" % e.get("raw_u16", {}).get("value", 0), "abba %.0f" % e.get("abba_product", {}).get("value", 0))
PY

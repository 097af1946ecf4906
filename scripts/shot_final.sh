#!/bin/bash
# last check of the tree as the driver will see it: smoke, the GPU test suite, the driver's bench line
set -u
o=gpurun_out; mkdir -p $o
python -c "import __graft_entry__ as g; g.smoke()" > $o/final_smoke.txt 2>&1; echo "smoke exit $?"; tail -2 $o/final_smoke.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $o/final_pytest.txt 2>&1; echo "pytest exit $?"; tail -3 $o/final_pytest.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $o/final_bench.json 2> $o/final_bench.err; echo "bench exit $?"; cut -c1-250 $o/final_bench.json

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/x9; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu (PDL) rc=$?"; tail -3 $O/pytest_gpu.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_generators.py tests/test_gpu_inference.py -x -q > $O/pytest_rep$i.log 2>&1; echo "repeat $i rc=$?"; done
V2V_PDL=0 timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_pdl0.json 2> $O/bench_pdl0.err
timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_pdl1.json 2> $O/bench_pdl1.err
timeout 300 python bench.py --no-cpu-baseline --steps 20 --workload cfg2 > $O/bench_cfg2.json 2> $O/bench_cfg2.err
grep -o '"ms_per_step": [0-9.]*' $O/bench*.json

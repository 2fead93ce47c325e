#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/x4; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $O/pytest_gpu.log
V2V_APPLY=0 timeout 600 python -m pytest tests/test_gpu_conv.py -x -q > $O/pytest_conv_a0.log 2>&1; echo "pytest conv (old apply) rc=$?"
timeout 300 python tools/time_conv.py > $O/time_conv.log 2>&1; grep conv_ms $O/time_conv.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench.json 2> $O/bench.err
V2V_APPLY=0 timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_a0.json 2> $O/bench_a0.err
for sk in 16 2; do V2V_SKIP=$sk timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_skip$sk.json 2> $O/bench_skip$sk.err; done
grep -o '"ms_per_step": [0-9.]*' $O/bench*.json
V2V_DBG=4 timeout 120 python tools/time_conv.py c64_512x1024 > $O/trace_c64.log 2>&1; grep "trace it" $O/trace_c64.log | tail -30 | head -4
timeout 300 python tools/profile_frame.py cfg4 $O/profile_cfg4 > /dev/null 2>&1

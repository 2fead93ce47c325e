#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/x6; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $O/pytest_gpu.log
S="c1024 c128_256x512 stem108_48 stem108_96 stem108_192 stem108_32"
V2V_MG=0 timeout 300 python tools/time_conv.py $S > $O/tc_mg1.log 2>&1; echo "== MG off"; grep conv_ms $O/tc_mg1.log
V2V_MG=2 timeout 300 python tools/time_conv.py $S > $O/tc_mg2.log 2>&1; echo "== MG<=2"; grep conv_ms $O/tc_mg2.log
timeout 300 python tools/time_conv.py $S > $O/tc_mg4.log 2>&1; echo "== MG<=4"; grep conv_ms $O/tc_mg4.log
V2V_MG=0 timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_mg1.json 2> $O/bench_mg1.err
timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench.json 2> $O/bench.err
grep -o '"ms_per_step": [0-9.]*' $O/bench*.json
timeout 300 python tools/profile_frame.py cfg4 $O/profile_cfg4 > /dev/null 2>&1; head -30 $O/profile_cfg4.txt

#!/bin/bash
# Experiment session: 2-D patch mode, warp-wide MMA issue, in-graph cost per kernel kind.
cd "$(dirname "$0")/.."
O=gpurun_out/x2; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv.py -x -q > $O/pytest_conv.log 2>&1; echo "pytest conv (patch2d) rc=$?"; tail -3 $O/pytest_conv.log
V2V_PATCH2D=0 timeout 900 python -m pytest tests/test_gpu_conv.py -x -q > $O/pytest_conv_p0.log 2>&1; echo "pytest conv (patch2d off) rc=$?"; tail -2 $O/pytest_conv_p0.log
V2V_DBG=8 timeout 900 python -m pytest tests/test_gpu_conv.py -x -q > $O/pytest_conv_ww.log 2>&1; echo "pytest conv (warp-wide issue) rc=$?"; tail -2 $O/pytest_conv_ww.log
V2V_CONV_IMPL=simt timeout 900 python -m pytest tests/test_gpu_conv.py -x -q > $O/pytest_conv_simt.log 2>&1; echo "pytest conv (simt) rc=$?"; tail -2 $O/pytest_conv_simt.log
V2V_PATCH2D=0 timeout 300 python tools/time_conv.py > $O/time_conv_p0.log 2>&1
timeout 300 python tools/time_conv.py > $O/time_conv_p1.log 2>&1
V2V_DBG=8 timeout 300 python tools/time_conv.py > $O/time_conv_p1_ww.log 2>&1
for f in p0 p1 p1_ww; do echo "== $f"; grep conv_ms $O/time_conv_$f.log; done
V2V_PATCH2D=0 timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_p0.json 2> $O/bench_p0.err
timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_p1.json 2> $O/bench_p1.err
V2V_DBG=8 timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_p1_ww.json 2> $O/bench_p1_ww.err
for sk in 8 16 2 33 64; do V2V_SKIP=$sk timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_skip$sk.json 2> $O/bench_skip$sk.err; done
grep -o '"ms_per_step": [0-9.]*' $O/bench_*.json
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $O/pytest_gpu.log

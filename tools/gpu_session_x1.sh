#!/bin/bash
# Experiment session: epilogue groups, MMA rate by swizzle layout, in-graph cost per kernel kind, ncu of norm-apply.
cd "$(dirname "$0")/.."
O=gpurun_out/x1; mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vid2vid_b200/csrc -o /tmp/umma_rate tools/micro/umma_rate.cu && timeout 120 /tmp/umma_rate > $O/umma_rate.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q > $O/pytest_conv.log 2>&1; echo "pytest conv rc=$?"
for eg in 1 2; do V2V_EG=$eg timeout 300 python tools/time_conv.py > $O/time_conv_eg$eg.log 2>&1; done
cat $O/time_conv_eg*.log | grep -v Warn
for eg in 1 2; do V2V_EG=$eg timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_eg$eg.json 2> $O/bench_eg$eg.err; done
for sk in 8 16 2 33 64; do V2V_SKIP=$sk timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_skip$sk.json 2> $O/bench_skip$sk.err; done
grep -o '"ms_per_step": [0-9.]*' $O/bench_*.json
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $O/pytest_gpu.log
for s in stem108_32 c64_512x1024; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:norm_apply -s 2 -c 1 -o $O/apply_$s python tools/time_conv.py $s > $O/ncu_apply_$s.log 2>&1; echo "ncu apply $s rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -o $O/conv_head32 python tools/time_conv.py head32_2048 > $O/ncu_head.log 2>&1
ls $O

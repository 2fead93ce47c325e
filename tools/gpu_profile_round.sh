#!/bin/bash
# Evidence run for profiles/: tests, smoke, both bench arms, ncu launch lists with DRAM bytes per arithmetic mode, full ncu
# captures of the top conv launch and the normalise pass, stand-alone op timings.  Usage: tools/gpu_profile_round.sh [notests]
cd "$(dirname "$0")/.."
O=gpurun_out/prof; mkdir -p $O
if [ "$1" != "notests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -2 $O/pytest_gpu.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_cfg4_reference.json 2> $O/bench_cfg4_reference.err; echo "reference arm rc=$?"
for m in precise fast; do
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --print-units base --clock-control none -c 4000 --csv --log-file $O/launches_cfg4_$m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --modes $m > $O/ncu_bench_$m.log 2>&1; echo "launch list $m rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -o $O/conv_c1024_precise python tools/time_conv.py c1024 > $O/ncu_c1024.log 2>&1; echo "ncu c1024 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:norm_apply -s 2 -c 1 -o $O/apply_stem48_precise python tools/time_conv.py stem108_48 > $O/ncu_apply.log 2>&1; echo "ncu apply rc=$?"
timeout 300 python tools/time_conv.py > $O/time_conv.log 2>&1
timeout 300 python tools/time_ops.py > $O/time_ops.log 2>&1
timeout 600 python tools/profile_frame.py cfg4 $O/profile_cfg4 > /dev/null 2>&1
ls $O; tail -c 600 $O/bench_cfg4.json

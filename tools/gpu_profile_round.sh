#!/bin/bash
# Evidence run for profiles/: tests, smoke, bench lines, ncu launch list with DRAM bytes, full ncu captures.
cd "$(dirname "$0")/.."
O=gpurun_out/prof; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -2 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --print-units base --clock-control none -c 6000 --csv --log-file $O/launches_cfg4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "launch list rc=$?"
for s in c1024; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -o $O/conv_$s python tools/time_conv.py $s > $O/ncu_$s.log 2>&1; echo "ncu $s rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:norm_apply -s 2 -c 1 -o $O/apply_stem48 python tools/time_conv.py stem108_48 > $O/ncu_apply.log 2>&1; echo "ncu apply rc=$?"
timeout 300 python tools/time_conv.py > $O/time_conv.log 2>&1
timeout 600 python tools/profile_frame.py cfg4 $O/profile_cfg4 > /dev/null 2>&1
timeout 600 python tools/profile_frame.py cfg2 $O/profile_cfg2 > /dev/null 2>&1
timeout 600 python bench.py --workload cfg2 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err
ls $O; tail -c 900 $O/bench_cfg4.json

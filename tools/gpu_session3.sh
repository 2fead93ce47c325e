#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/s3
mkdir -p $O
PT="python -m pytest -q -s -p no:cacheprovider --timeout=300"
timeout 1500 $PT tests -m gpu > $O/full.log 2>&1; echo "full rc=$?" | tee $O/summary.txt
timeout 600 python tools/profile_frame.py cfg4 $O/profile_cfg4 > $O/profile_cfg4.log 2>&1; echo "profile rc=$?" | tee -a $O/summary.txt
timeout 600 python tools/profile_frame.py cfg2 $O/profile_cfg2 > $O/profile_cfg2.log 2>&1
# launch list (cold-cache, serialised) of a short bench run
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches_cfg4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
# full capture of the roofline kernel: 1024->1024 3x3 at 32x64 (cfg2 shape: launch index inside G0)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 40 -c 3 -o $O/prof_conv python bench.py --workload cfg2 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $O/summary.txt
grep -h -E "passed|failed" $O/full.log | tail -3; grep -E "^FAILED" $O/full.log
cat $O/summary.txt; head -40 $O/profile_cfg4.txt

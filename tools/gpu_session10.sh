#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/s10; mkdir -p $O
PT="python -m pytest -q -s -p no:cacheprovider --timeout=300"
timeout 1500 $PT tests -m gpu > $O/full.log 2>&1; echo "full rc=$?" | tee $O/summary.txt
timeout 600 python tools/profile_frame.py cfg4 $O/profile_cfg4 > $O/profile_cfg4.log 2>&1; echo "profile rc=$?" | tee -a $O/summary.txt
timeout 900 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 rc=$?" | tee -a $O/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench cfg4 rc=$?" | tee -a $O/summary.txt

grep -h -E "passed|failed" $O/full.log | tail -2; grep -h -E "^FAILED" $O/full.log | head
cat $O/summary.txt; head -3 $O/profile_cfg4.txt; python -c "
import json
for f in ['$O/bench_cfg2.json','$O/bench_cfg4.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f,'fps %.2f e2e %.2f ms %.2f conv_ms %.2f other %.2f TF %.1f'%(d['value'],d['e2e']['value'],d['ms_per_step'],r['conv_kernel_ms_per_frame'],r['other_kernels_ms_per_frame'],r['achieved']), r['ms_by_kernel_kind'])
    except Exception as e: print(f,e)
"

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/s9; mkdir -p $O
PT="python -m pytest -q -s -p no:cacheprovider --timeout=300"
timeout 900 $PT tests/test_gpu_conv.py > $O/conv.log 2>&1; echo "conv rc=$?" | tee $O/summary.txt
timeout 300 python tools/time_conv.py > $O/time.log 2>&1
V2V_DBG=4 timeout 300 python tools/time_conv.py c64_512x1024 2>&1 | grep -E "trace it=1[0-3]|conv_ms" | cut -c1-300 >> $O/time.log
timeout 1200 $PT tests -m gpu --deselect tests/test_gpu_conv.py > $O/rest.log 2>&1; echo "rest rc=$?" | tee -a $O/summary.txt
timeout 600 python tools/profile_frame.py cfg4 $O/profile_cfg4 > $O/profile_cfg4.log 2>&1; echo "profile rc=$?" | tee -a $O/summary.txt
timeout 900 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 rc=$?" | tee -a $O/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench cfg4 rc=$?" | tee -a $O/summary.txt
grep -h -E "passed|failed" $O/conv.log $O/rest.log | tail -3; grep -h -E "^FAILED" $O/*.log | head -20
cat $O/summary.txt $O/time.log; head -24 $O/profile_cfg4.txt | cut -c1-110; python -c "
import json
for f in ['$O/bench_cfg2.json','$O/bench_cfg4.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f,'fps %.2f e2e %.2f ms %.2f conv_ms %.2f other %.2f TF %.1f'%(d['value'],d['e2e']['value'],d['ms_per_step'],r['conv_kernel_ms_per_frame'],r['other_kernels_ms_per_frame'],r['achieved']), r['ms_by_kernel_kind'])
    except Exception as e: print(f,e)
"

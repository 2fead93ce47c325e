#!/bin/bash
# Final evidence run of a round: full GPU test suite, smoke, both bench arms, training and FlowNet2 lines, training profile.
cd "$(dirname "$0")/.."
O=gpurun_out/final; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_cfg4_reference.json 2> $O/bench_cfg4_reference.err; echo "reference arm rc=$?"
timeout 900 python bench.py > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench rc=$?"
timeout 400 python bench.py --workload cfg3 --steps 6 > $O/bench_cfg3_1gpu.json 2> $O/bench_cfg3_1gpu.err; echo "cfg3 rc=$?"
timeout 300 python bench.py --workload flownet2 --steps 10 > $O/bench_flownet2.json 2> $O/bench_flownet2.err; echo "flownet2 rc=$?"
timeout 300 python tools/profile_train.py cfg3 $O/profile_train_cfg3.txt > /dev/null 2>&1
python - <<'PY'
import json
for f in ('bench_cfg4', 'bench_cfg4_reference', 'bench_cfg3_1gpu', 'bench_flownet2'):
    try:
        d = json.load(open('gpurun_out/final/%s.json' % f))
        print(f, round(d['value'], 3), d['unit'], 'ms/step', round(d.get('ms_per_step', 0), 2), 'e2e', round(d.get('e2e', {}).get('value', 0), 3),
              'fast', round(d.get('fast', {}).get('value', 0), 2), 'roofline', round(d.get('roofline', {}).get('frac', 0), 4))
    except Exception as e:
        print(f, 'unreadable', e)
PY
head -8 $O/profile_train_cfg3.txt

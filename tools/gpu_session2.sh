#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/s2
mkdir -p $O
PT="python -m pytest -q -s -p no:cacheprovider --timeout=300"
echo "=== full suite (default: no tap reuse)" | tee $O/summary.txt
timeout 1500 $PT tests -m gpu > $O/full.log 2>&1; echo "full rc=$?" | tee -a $O/summary.txt
echo "=== tap-reuse variants" | tee -a $O/summary.txt
for c in c3s1_rowtile_R3_8x160 c7_stem_R7_108_32; do
  V2V_TAP_REUSE=1 timeout 180 $PT "tests/test_gpu_conv.py::test_conv_unit[$c]" > $O/reuse_doc_$c.log 2>&1; echo "reuse(desc per doc) $c rc=$?" | tee -a $O/summary.txt
  V2V_TAP_REUSE=1 V2V_DESC_MODE=1 timeout 180 $PT "tests/test_gpu_conv.py::test_conv_unit[$c]" > $O/reuse_d1_$c.log 2>&1; echo "reuse(base_offset=0) $c rc=$?" | tee -a $O/summary.txt
done
V2V_TAP_REUSE=1 V2V_DESC_MODE=1 timeout 900 $PT tests/test_gpu_generators.py > $O/gen_reuse_d1.log 2>&1; echo "generators reuse(base_offset=0) rc=$?" | tee -a $O/summary.txt
echo "=== SIMT cross-check" | tee -a $O/summary.txt
V2V_CONV_IMPL=simt timeout 900 $PT tests/test_gpu_conv.py -k "not 1024 and not 512_512" > $O/conv_simt.log 2>&1; echo "conv_simt rc=$?" | tee -a $O/summary.txt
echo "=== smoke + bench" | tee -a $O/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 900 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 rc=$?" | tee -a $O/summary.txt
timeout 1500 python bench.py --steps 5 --warmup 3 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench cfg4 rc=$?" | tee -a $O/summary.txt
V2V_TAP_REUSE=1 V2V_DESC_MODE=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_cfg4_reuse.json 2> $O/bench_cfg4_reuse.err; echo "bench cfg4 reuse rc=$?" | tee -a $O/summary.txt
grep -h -E "passed|failed" $O/*.log | tail -20
cat $O/summary.txt
tail -c 3000 $O/bench_cfg4.json; tail -5 $O/bench_cfg4.err

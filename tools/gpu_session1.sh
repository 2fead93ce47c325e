#!/bin/bash
# First GPU session: validate everything with the SIMT cross-check kernel first (isolates plan/layout/norm bugs from
# tcgen05 bugs), then the tcgen05 kernel case by case in separate processes (a trap poisons a CUDA context),
# then the full suite, smoke, and a short bench.  Everything is logged under gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out/s1
mkdir -p $O
nvidia-smi > $O/nvidia-smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" > $O/host.txt 2>&1
PT="python -m pytest -q -s -p no:cacheprovider --timeout=300"

echo "=== ops" | tee $O/summary.txt
timeout 600 $PT tests/test_gpu_ops.py > $O/ops.log 2>&1; echo "ops rc=$?" | tee -a $O/summary.txt

echo "=== SIMT conv units" | tee -a $O/summary.txt
V2V_CONV_IMPL=simt timeout 900 $PT tests/test_gpu_conv.py -k "not 1024 and not 512_512" > $O/conv_simt.log 2>&1; echo "conv_simt rc=$?" | tee -a $O/summary.txt
V2V_CONV_IMPL=simt timeout 900 $PT tests/test_gpu_generators.py -k "not cfg1 and not simt_vs_umma" > $O/gen_simt.log 2>&1; echo "gen_simt rc=$?" | tee -a $O/summary.txt

echo "=== UMMA conv units, one process each" | tee -a $O/summary.txt
for c in c3s1_reflect_64_128_32x64 c3s2_zero_64_128 deconv_128_64 c7_stem_small_map resblock_128 d_first_layer_lrelu c3_512_512_16x32 c3_1024_1024_32x64; do
  timeout 180 $PT "tests/test_gpu_conv.py::test_conv_unit[$c]" > $O/umma_$c.log 2>&1; echo "umma $c rc=$?" | tee -a $O/summary.txt
done
echo "=== UMMA tap-reuse cases (shifted smem descriptors)" | tee -a $O/summary.txt
for c in c3s1_rowtile_R3_8x160 c7_stem_R7_108_32; do
  timeout 180 $PT "tests/test_gpu_conv.py::test_conv_unit[$c]" > $O/umma_reuse_$c.log 2>&1; echo "umma reuse(desc per doc) $c rc=$?" | tee -a $O/summary.txt
  V2V_DESC_MODE=1 timeout 180 $PT "tests/test_gpu_conv.py::test_conv_unit[$c]" > $O/umma_reuse_d1_$c.log 2>&1; echo "umma reuse(base_offset=0) $c rc=$?" | tee -a $O/summary.txt
  V2V_TAP_REUSE=0 timeout 180 $PT "tests/test_gpu_conv.py::test_conv_unit[$c]" > $O/umma_noreuse_$c.log 2>&1; echo "umma no-reuse $c rc=$?" | tee -a $O/summary.txt
done

echo "=== full suite (default path)" | tee -a $O/summary.txt
timeout 1500 $PT tests -m gpu > $O/full.log 2>&1; echo "full rc=$?" | tee -a $O/summary.txt
echo "=== full suite with V2V_TAP_REUSE=0" | tee -a $O/summary.txt
V2V_TAP_REUSE=0 timeout 1500 $PT tests -m gpu > $O/full_noreuse.log 2>&1; echo "full_noreuse rc=$?" | tee -a $O/summary.txt

echo "=== smoke + bench" | tee -a $O/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 900 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 rc=$?" | tee -a $O/summary.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench cfg4 rc=$?" | tee -a $O/summary.txt
V2V_TAP_REUSE=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_cfg4_noreuse.json 2> $O/bench_cfg4_noreuse.err; echo "bench cfg4 noreuse rc=$?" | tee -a $O/summary.txt
grep -h -E "passed|failed|error" $O/*.log | tail -40
cat $O/summary.txt
tail -c 1500 $O/bench_cfg4.json

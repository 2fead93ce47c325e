#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/x5; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_generators.py -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for i in 1 2; do
V2V_LIB=$PWD/vid2vid_b200/ab/libv2v_prev.so timeout 300 python tools/time_conv.py c1024 c512 c128_256x512 c64_512x1024 stem108_32 head32_2048 > $O/tc_prev$i.log 2>&1; echo "== prev $i"; grep conv_ms $O/tc_prev$i.log
timeout 300 python tools/time_conv.py c1024 c512 c128_256x512 c64_512x1024 stem108_32 head32_2048 > $O/tc_new$i.log 2>&1; echo "== new $i"; grep conv_ms $O/tc_new$i.log
done
V2V_PATCH2D=0 timeout 300 python tools/time_conv.py c1024 c512 c128_256x512 c64_512x1024 stem108_32 head32_2048 > $O/tc_new_p0.log 2>&1; echo "== new patch2d=0"; grep conv_ms $O/tc_new_p0.log
V2V_LIB=$PWD/vid2vid_b200/ab/libv2v_prev.so timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_prev.json 2> $O/bench_prev.err
timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_new.json 2> $O/bench_new.err
V2V_APPLY=0 timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_new_a0.json 2> $O/bench_new_a0.err
V2V_SKIP=16 timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_skip16.json 2> $O/bench_skip16.err
grep -o '"ms_per_step": [0-9.]*' $O/bench*.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:norm_apply -s 2 -c 1 -o $O/apply_stem python tools/time_conv.py stem108_32 > $O/ncu_apply.log 2>&1; echo "ncu rc=$?"

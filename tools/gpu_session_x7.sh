#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/x7; mkdir -p $O
timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_first.json 2> $O/bench_first.err
V2V_MG=2 V2V_DBG=4 timeout 120 python tools/time_conv.py c128_256x512 > $O/trace_c128_mg2.log 2>&1; grep "trace it" $O/trace_c128_mg2.log | tail -24 | head -8
V2V_MG=0 V2V_DBG=4 timeout 120 python tools/time_conv.py c128_256x512 > $O/trace_c128_mg0.log 2>&1; grep "trace it" $O/trace_c128_mg0.log | tail -24 | head -5
V2V_DBG=4 timeout 120 python tools/time_conv.py stem108_48 > $O/trace_stem_mg4.log 2>&1; grep "trace it" $O/trace_stem_mg4.log | tail -24 | head -5
V2V_MG=2 V2V_DBG=4 timeout 120 python tools/time_conv.py stem108_48 > $O/trace_stem_mg2.log 2>&1; grep "trace it" $O/trace_stem_mg2.log | tail -24 | head -5
for eg in 1 2; do V2V_EG=$eg timeout 200 python tools/time_conv.py c128_256x512 stem108_48 > $O/tc_eg$eg.log 2>&1; grep conv_ms $O/tc_eg$eg.log; done
timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench_second.json 2> $O/bench_second.err
grep -o '"ms_per_step": [0-9.]*' $O/bench*.json

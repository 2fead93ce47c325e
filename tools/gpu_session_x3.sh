#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/x3; mkdir -p $O
for d in 0 1 2 3; do V2V_DBG=$d timeout 300 python tools/time_conv.py c64_512x1024 c32_512x1024 up64_32 c128_256x512 head32_2048 > $O/tc_dbg$d.log 2>&1; echo "== dbg $d"; grep conv_ms $O/tc_dbg$d.log; done
for eg in 1 2; do V2V_EG=$eg V2V_DBG=3 timeout 300 python tools/time_conv.py c64_512x1024 c32_512x1024 > $O/tc_dbg3_eg$eg.log 2>&1; echo "== dbg 3 eg $eg"; grep conv_ms $O/tc_dbg3_eg$eg.log; done
V2V_DBG=4 timeout 120 python tools/time_conv.py c64_512x1024 > $O/trace_c64.log 2>&1; grep "trace it" $O/trace_c64.log | tail -30 | head -12

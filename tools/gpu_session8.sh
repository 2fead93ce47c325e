#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/s8; mkdir -p $O
V2V_DBG=4 timeout 300 python tools/time_conv.py c64_512x1024 2>&1 | grep -E "trace|conv_ms" | head -30 > $O/trace_c64.log
V2V_DBG=4 timeout 300 python tools/time_conv.py c1024 2>&1 | grep -E "trace|conv_ms" | head -30 > $O/trace_c1024.log
cat $O/trace_c64.log

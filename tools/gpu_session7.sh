#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/s7; mkdir -p $O; rm -f $O/time.log
for dbg in 0 1 2 3; do V2V_DBG=$dbg timeout 300 python tools/time_conv.py c64_512x1024 c32_512x1024 c128_256x512 stem108_32 >> $O/time.log 2>&1; done
V2V_MG=1 timeout 300 python tools/time_conv.py c1024 c128_256x512 stem108_32 >> $O/time.log 2>&1
cat $O/time.log

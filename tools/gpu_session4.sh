#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/s4; mkdir -p $O
for dbg in 0 1 2; do V2V_DBG=$dbg timeout 300 python tools/time_conv.py c1024 c512 c64_512x1024 >> $O/time.log 2>&1; done
for st in 2 3 4; do V2V_STAGES=$st timeout 300 python tools/time_conv.py c1024 >> $O/time.log 2>&1; done
cat $O/time.log

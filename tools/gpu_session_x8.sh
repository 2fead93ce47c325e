#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/x8; mkdir -p $O
run() { V2V_FORCE=$1 timeout 120 python tools/time_conv.py $2 2>&1 | grep conv_ms | sed "s/^/force=$1 /"; }
for f in 64,128,1 64,128,2 32,128,2 64,64,2 64,64,4; do run $f c128_256x512; done
for f in 64,64,1 64,64,2 32,64,2 32,64,4 64,32,2 64,32,4; do run $f stem108_48; done
for f in 64,64,1 64,64,2 32,96,2 64,96,1 32,128,2 32,96,1; do run $f stem108_96; done
for f in 64,64,2 32,96,2 64,96,1 32,128,2; do run $f stem108_192; done
timeout 300 python bench.py --no-cpu-baseline --steps 20 > $O/bench.json 2> $O/bench.err
grep -o '"ms_per_step": [0-9.]*' $O/bench*.json

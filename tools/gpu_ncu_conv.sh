#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/ncu1; mkdir -p $O
for s in c1024 c64_512x1024 stem108_32; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -o $O/$s python tools/time_conv.py $s > $O/$s.log 2>&1
  echo "$s rc=$?"
done
ls -la $O

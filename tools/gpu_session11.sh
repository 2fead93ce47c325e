#!/bin/bash
# 2-GPU session: default bench (with cpu_baseline), reference arm, torchrun N=2
cd "$(dirname "$0")/.."
O=gpurun_out/s11; mkdir -p $O
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?" | tee $O/summary.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?" | tee -a $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" | tee -a $O/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "bench ref n2 rc=$?" | tee -a $O/summary.txt
cat $O/summary.txt; for f in bench_default bench_reference bench_n2 bench_ref_n2; do echo "== $f"; tail -1 $O/$f.json | cut -c1-1500; tail -2 $O/$f.err | cut -c1-300; done

#!/bin/bash
# Multi-GPU evidence run (gpurun --gpus 8): both drop-in routes checked against the single-GPU search and
# the full-size golden, then the scaling bench the driver runs, at N = 1, 2, 4, 8, both routes.
set -u
out=gpurun_out/r02_multi_gpu.txt
: > $out
run() { echo "\$ $*" >> $out; "$@" >> $out 2>&1; echo "rc=$?" >> $out; echo >> $out; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run $TR --nproc-per-node 8 --master-port 29541 tests/multi_gpu_check.py
run python tests/single_process_multi_gpu_check.py 8
run python bench.py --gpus 1 --no-other --no-cpu
for n in 2 4 8; do
  run $TR --nproc-per-node $n --master-port 2955$n bench.py --gpus $n --steps 5 --warmup 3
  run python bench.py --gpus $n --single-process --steps 5 --warmup 3
done
grep -c '"parity"' $out

#!/usr/bin/env bash
# Round-2 GPU job (4 GPUs): the 4-rank bench line (2x2x1 partition, parity block).
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 --cg-iters 20 > $O/r2_bench_n4.json 2> $O/r2_bench_n4.err
echo "bench rc=$?"
tail -c 600 $O/r2_bench_n4.json
tail -3 $O/r2_bench_n4.err

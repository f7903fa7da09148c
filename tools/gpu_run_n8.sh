#!/usr/bin/env bash
# Round-2 GPU job (8 GPUs): the 8-rank bench line (parity block included).
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2_bench_n8.json 2> $O/r2_bench_n8.err
echo "bench rc=$?"
tail -c 1500 $O/r2_bench_n8.json
echo done

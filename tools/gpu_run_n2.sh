#!/usr/bin/env bash
# Round-2 GPU job (2 GPUs): multi-process parity over CUDA IPC + NCCL, and the
# 2-rank bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/check_multi_gpu.py > $O/r2_check_multi_gpu_2ranks.log 2>&1
echo "check_multi_gpu rc=$?"
timeout 600 $TR --master-port 29517 tools/check_comm_nccl.py > $O/r2_check_comm_nccl_2ranks.log 2>&1
echo "check_comm_nccl rc=$?"
timeout 900 $TR --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2_bench_n2.json 2> $O/r2_bench_n2.err
echo "bench rc=$?"
tail -3 $O/r2_check_multi_gpu_2ranks.log $O/r2_check_comm_nccl_2ranks.log
echo done

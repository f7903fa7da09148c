#!/usr/bin/env bash
# Round-2 GPU job (8 GPUs): the two single-process CG tests, then the 8-rank
# bench line (parity block included) with exchange mode 3 as the default.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --tb=short -k "fused_distributed_cg_single_process" > $O/r2_n8b_pytest.log 2>&1
tail -3 $O/r2_n8b_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2_bench_n8_mode3.json 2> $O/r2_bench_n8_mode3.err
echo "bench rc=$?"
tail -c 1800 $O/r2_bench_n8_mode3.json
echo done

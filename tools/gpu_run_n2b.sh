#!/usr/bin/env bash
# 2 GPUs at the per-rank size of the 8-GPU run (ne=42: 37 k elements per rank):
# where do the 36 us between the bare kernel and the step with the exchange go?
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -k "peer_memory_halo or fused_distributed" > gpurun_out/r2_n2b_pytest.log 2>&1; tail -3 gpurun_out/r2_n2b_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --ne 42 --steps 40 --warmup 5 --no-e2e --no-parity --cg-iters 40"
i=0
for env in "SFEM_HALO_FUSE_UNPACK=2" "SFEM_HALO_FUSE_UNPACK=3" "SFEM_HALO_FUSE_UNPACK=3 SFEM_WAIT_CTAS=148" "SFEM_HALO_FUSE_UNPACK=3 SFEM_WAIT_CTAS=1184"; do
  i=$((i+1))
  env $env timeout 600 $TR --master-port $((29530+i)) $B > $O/r2_n2_ne42_$i.json 2> $O/r2_n2_ne42_$i.err
  python - "$O/r2_n2_ne42_$i.json" "$env" <<'PY'
import json,sys
try:
  d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print(sys.argv[2], '| step %.4f ms, bare kernel %.4f ms, cg %.4f ms/it'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['cg']['ms_per_iteration']), d['details']['halo_timeline'])
except Exception as e: print(sys.argv[2],'ERR',e)
PY
done
echo done

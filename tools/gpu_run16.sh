#!/usr/bin/env bash
# Round-2 GPU job 16 (1 GPU): Stokes pieces -- specialised D / D^T kernels,
# CG graph heuristic; tests, then the Stokes step at ne = 64 and 256.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_navier_stokes_gpu.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --tb=short -k "navier or stokes or fused_div or kolmogorov or graph_replay or device_state" > $O/r2_run16_pytest.log 2>&1
tail -5 $O/r2_run16_pytest.log
for ne in 64 256; do
timeout 400 python tools/bench_ns.py --ne $ne --order 7 --reps 10 > $O/r2_bench_ns_ne${ne}_detail.json 2> $O/r2_bench_ns_ne${ne}_detail.err
tail -3 $O/r2_bench_ns_ne${ne}_detail.err
python - $ne <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r2_bench_ns_ne{sys.argv[1]}_detail.json').read().strip().splitlines()[-1])
print({k: round(v, 1) for k, v in d['us'].items()})
print({k: round(v, 1) for k, v in d['gbs'].items()})
for k, v in d.items():
  if k not in ('us', 'gbs'):
    print(k, v)
PY
done
echo done

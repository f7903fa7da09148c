#!/usr/bin/env bash
# Round-2 GPU job 9 (1 GPU): fused Stokes D / D^T, CG step with the deferred x
# update.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_navier_stokes_gpu.py -m gpu -q -p no:cacheprovider --tb=short > $O/r2_run9_pytest_ns.log 2>&1
tail -3 $O/r2_run9_pytest_ns.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > $O/r2_run9_pytest.log 2>&1
tail -3 $O/r2_run9_pytest.log
timeout 600 python tools/bench_ns.py --ne 64 --order 7 > $O/r2_bench_ns_fused.json 2> $O/r2_bench_ns_fused.err
timeout 600 python tools/bench_ns.py --ne 256 --order 7 > $O/r2_bench_ns_fused_ne256.json 2> $O/r2_bench_ns_fused_ne256.err
cat $O/r2_bench_ns_fused.json $O/r2_bench_ns_fused_ne256.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extra --no-parity --cg-iters 50 > $O/r2_bench_cg_xdefer.json 2> $O/r2_bench_cg_xdefer.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_cg_xdefer.json').read().strip().splitlines()[-1])
print('apply %.2f frac %.3f | cg %.4f ms/it frac %.3f'%(d['value'],d['roofline']['frac'],d['cg']['ms_per_iteration'],d['cg']['roofline_frac']))
PY
echo done

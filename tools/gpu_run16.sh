#!/usr/bin/env bash
# Round-2 GPU job 16 (1 GPU): line-per-lane 2-D Stokes D / D^T kernels --
# tests (fused vs composed, goldens, Stokes step, Kolmogorov), then timings.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_navier_stokes_gpu.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --tb=short -k "navier or stokes or fused_div or kolmogorov or graph_replay or device_state" > $O/r2_run16_pytest.log 2>&1
tail -3 $O/r2_run16_pytest.log
for lines in 1; do
SFEM_STOKES_LINES=$lines timeout 400 python tools/bench_ns.py --ne 256 --order 7 --reps 20 > $O/r2_bench_ns_ne256_lines${lines}.json 2> $O/r2_bench_ns_ne256_lines${lines}.err
tail -3 $O/r2_bench_ns_ne256_lines${lines}.err
python - $lines <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r2_bench_ns_ne256_lines{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('lines', sys.argv[1], {k: round(v, 1) for k, v in d['us'].items()}, 'step', round(d['stokes_one_step_ms']), round(d['stokes_one_step_second_call_ms']), d['dp_iterations'])
PY
done
timeout 400 python tools/bench_ns.py --ne 64 --order 7 --reps 20 > $O/r2_bench_ns_ne64_detail.json 2> $O/r2_bench_ns_ne64_detail.err
python - <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r2_bench_ns_ne64_detail.json').read().strip().splitlines()[-1])
print('ne64', {k: round(v, 1) for k, v in d['us'].items()}, 'step', round(d['stokes_one_step_ms']), round(d['stokes_one_step_second_call_ms']), d['dp_iterations'])
PY
echo done

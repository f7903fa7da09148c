#!/usr/bin/env bash
# Round-2 GPU job 12 (1 GPU): lazy zero fill by a companion kernel -- test,
# then the bench line with the eager fill and with several pacings.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --tb=short -k "lazy_zero" > $O/r2_run12_pytest.log 2>&1
tail -4 $O/r2_run12_pytest.log
grep -h "timeout in\|rejected" $O/r2_run12_pytest.log | head -5
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extra --no-parity --cg-iters 30"
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 200 $B > $O/r2_lazy_$name.json 2> $O/r2_lazy_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
  d = json.loads(open(f'gpurun_out/r2_lazy_{name}.json').read().strip().splitlines()[-1])
  print('%-14s apply %.2f GDOF/s %.4f ms frac %.3f | cg %.4f ms/it frac %.3f | clocks %s %s | %s' % (
      name, d['value'], d['ms_per_step'], d['roofline']['frac'], d['cg']['ms_per_iteration'],
      d['cg']['roofline_frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['details']['zero_fill'][:150]))
  sys.exit(0 if d['ms_per_step'] < 50 and 'rejected' not in d['details']['zero_fill'] else 3)
except Exception as e:
  print(name, 'FAILED', e)
  print(open(f'gpurun_out/r2_lazy_{name}.err').read()[-1500:])
  sys.exit(3)
PY
}
run eager SFEM_LAZY_ZERO=0
run a6_r4 SFEM_LAZY_ZERO=1 SFEM_LAZY_AHEAD=6 SFEM_LAZY_REPORT=4 || { echo "lazy path broken: stopping"; cat $O/r2_lazy_a6_r4.err | tail -5; exit 0; }
run a4_r4 SFEM_LAZY_ZERO=1 SFEM_LAZY_AHEAD=4 SFEM_LAZY_REPORT=4
run a3_r2 SFEM_LAZY_ZERO=1 SFEM_LAZY_AHEAD=3 SFEM_LAZY_REPORT=2
run a2_r2 SFEM_LAZY_ZERO=1 SFEM_LAZY_AHEAD=2 SFEM_LAZY_REPORT=2
run a3_r1 SFEM_LAZY_ZERO=1 SFEM_LAZY_AHEAD=3 SFEM_LAZY_REPORT=1 || { echo "r1 broken: stopping"; cat $O/r2_lazy_a3_r1.err | tail -5; exit 0; }
run a2_r1 SFEM_LAZY_ZERO=1 SFEM_LAZY_AHEAD=2 SFEM_LAZY_REPORT=1
run a1_r1 SFEM_LAZY_ZERO=1 SFEM_LAZY_AHEAD=1 SFEM_LAZY_REPORT=1
run a3_r1_c74 SFEM_LAZY_ZERO=1 SFEM_LAZY_AHEAD=3 SFEM_LAZY_REPORT=1 SFEM_LAZY_CTAS=74
echo done

#!/usr/bin/env bash
# Round-2 GPU job 12 (1 GPU): lazy zero fill by a companion kernel + claimed
# CTA steps -- test, then the bench line with the eager fill and with several
# chunk sizes / leads.
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
run c128_a0 SFEM_LAZY_ZERO=1 SFEM_LAZY_CHUNK=128 SFEM_LAZY_AHEAD=0 || { echo "lazy path broken: stopping"; tail -5 $O/r2_lazy_c128_a0.err; exit 0; }
run c128_a740 SFEM_LAZY_ZERO=1 SFEM_LAZY_CHUNK=128 SFEM_LAZY_AHEAD=740
run c512_a0 SFEM_LAZY_ZERO=1 SFEM_LAZY_CHUNK=512 SFEM_LAZY_AHEAD=0
run c32_a0 SFEM_LAZY_ZERO=1 SFEM_LAZY_CHUNK=32 SFEM_LAZY_AHEAD=0
run c128_a2200 SFEM_LAZY_ZERO=1 SFEM_LAZY_CHUNK=128 SFEM_LAZY_AHEAD=2200
run c128_a0_c74 SFEM_LAZY_ZERO=1 SFEM_LAZY_CHUNK=128 SFEM_LAZY_AHEAD=0 SFEM_LAZY_CTAS=74
echo done

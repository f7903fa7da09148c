#!/usr/bin/env bash
# Round-2 GPU job 14 (1 GPU): final state -- full GPU suite, smoke, the
# driver's bench line (both arms), launch list of a short bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > $O/r2_final_pytest.log 2>&1
tail -4 $O/r2_final_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py > $O/r2_bench_n1_final.json 2> $O/r2_bench_n1_final.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/r2_bench_ref_final.json 2> $O/r2_bench_ref_final.err
echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-e2e \
  --no-cpu-baseline --no-extra --no-parity --cg-iters 5 > $O/r2_ncu_launches_final.log 2>&1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1_final.json').read().strip().splitlines()[-1])
print('apply %.2f GDOF/s %.4f ms frac %.3f | e2e %.2f ms | cg %.4f ms/it frac %.3f | launches %s | clocks %s' % (
    d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'],
    d['cg']['ms_per_iteration'], d['cg']['roofline_frac'], d['gpu_launches'], d['clocks']))
print('cpu_baseline', d['cpu_baseline'])
r = json.loads(open('gpurun_out/r2_bench_ref_final.json').read().strip().splitlines()[-1])
print('reference arm', r.get('value'), r.get('unit'), r.get('cpu_baseline'))
PY
du -sh $O
echo done

#!/usr/bin/env bash
# Round-2 GPU job 4 (1 GPU): lazy zero fill -- tests, on/off and chunking at
# ne=68, DRAM traffic of the kernel (ncu --set full), launch list.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -k "lazy" > $O/r2_run4_pytest_lazy.log 2>&1
tail -3 $O/r2_run4_pytest_lazy.log
B="python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-extra --no-parity --cg-iters 30"
SFEM_LAZY_ZERO=0 timeout 600 $B > $O/r2_lazy_off.json 2> $O/r2_lazy_off.err
rm -f $O/r2_lazy_c*.json
for cfg in "512 2 4 32" "512 2 4 8" "512 2 8 32" "1024 2 4 64" "256 2 6 16"; do
  set -- $cfg
  SFEM_LAZY_CHUNK=$1 SFEM_LAZY_AHEAD=$2 SFEM_LAZY_MAX_AHEAD=$3 SFEM_LAZY_DUTY=$4 timeout 600 $B > $O/r2_lazy_c$1_a$2_m$3_d$4.json 2> $O/r2_lazy_c$1_a$2_m$3_d$4.err
done
for f in $O/r2_lazy_*.json; do python - "$f" <<'PY'
import json,sys
try:
  d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print(sys.argv[1], 'apply %.2f GDOF/s %.4f ms frac %.3f | cg %.4f ms/it frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['cg']['ms_per_iteration'],d['cg']['roofline_frac']))
except Exception as e: print(sys.argv[1],'ERR',e)
PY
done
# DRAM traffic of one launch of the (lazy) kernel
timeout 900 ncu --set full --clock-control none -k regex:apply3d_v2 -s 8 -c 1 -o $O/prof_lazy -f \
  python bench.py --steps 4 --warmup 5 --no-e2e --no-cpu-baseline --no-extra --no-parity --cg-iters 0 > $O/r2_ncu_lazy.log 2>&1
if [ -f $O/prof_lazy.ncu-rep ]; then
  ncu -i $O/prof_lazy.ncu-rep --page raw --csv > $O/prof_lazy_raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/prof_lazy_raw.csv > $O/r2_ncu_apply3d_ne68_lazy.txt
  ncu -i $O/prof_lazy.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r2_source_apply3d_ne68_lazy.csv.gz
  rm -f $O/prof_lazy_raw.csv $O/prof_lazy.ncu-rep
fi
du -sh $O
echo done

#!/usr/bin/env bash
# Round-2 GPU job 17 (1 GPU): ncu --set full of the fused Stokes D / D^T
# kernels at 256^2 elements, order 7.
set -u
mkdir -p gpurun_out
O=gpurun_out
for k in stokes_div stokes_grad_t; do
  rep=$O/prof_$k
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 3 -c 1 -o $rep -f \
    python tools/bench_ns.py --ne 256 --order 7 --reps 3 --ops-only > $O/ncu_$k.log 2>&1
  if [ -f $rep.ncu-rep ]; then
    ncu -i $rep.ncu-rep --page raw --csv > ${rep}_raw.csv 2>/dev/null
    python tools/ncu_summary.py ${rep}_raw.csv > $O/r2_ncu_${k}_ne256.txt
    ncu -i $rep.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r2_ncu_${k}_ne256_source.csv.gz
    rm -f ${rep}_raw.csv $rep.ncu-rep
    echo "== $k"; cat $O/r2_ncu_${k}_ne256.txt | head -60
  else
    echo "== $k: no report"; tail -5 $O/ncu_$k.log
  fi
done
du -sh $O
echo done

#!/usr/bin/env bash
# Round-2 GPU job 1 (1 GPU): full -m gpu suite (Kolmogorov un-gated, fp32 at
# 1e-5), tuning variants 9/10/11 of the headline kernel, first timing of the
# Stokes operators, ncu captures across orders.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -150 > $O/r2_run1_pytest.log
EXP=$PWD/swirl_fem_b200/lib_exp/libswirl_b200.so
for v in 0 9 10 11 0; do
  SFEM_LIB=$EXP SFEM_VARIANT=$v python bench.py --steps 20 --warmup 5 --no-e2e \
    --no-cpu-baseline --cg-iters 0 > $O/r2_variant_${v}.json 2> $O/r2_variant_${v}.err
done
SFEM_LIB=$EXP python tools/bench_apply.py --dim 3 --orders 4,5,6,8 --target-dofs 16e6 \
  --variants 0,9,10,11 --dtypes f64,f32 --check 0 > $O/r2_variants_orders.log 2>&1
SFEM_LIB=$EXP python tools/bench_apply.py --dim 3 --orders 7 --target-dofs 16e6 \
  --variants 0,3,4,5,6,7,9,10,11 --dtypes f64,f32 --check 0 >> $O/r2_variants_orders.log 2>&1
python tools/bench_ns.py --ne 64 --order 7 > $O/r2_bench_ns.json 2> $O/r2_bench_ns.err
bash tools/profile_orders.sh 3 "3 7 11 13 15" "f64 f32" > $O/r2_profile_orders_3d.log 2>&1
bash tools/profile_orders.sh 2 "4 8 11 15" "f64 f32" > $O/r2_profile_orders_2d.log 2>&1
# source-level stall pages of two captures (the .ncu-rep files are too big to
# bring back: gpurun merges at most 64 MiB)
for k in d2_p8_f64 d3_p11_f64; do
  if [ -f $O/prof_${k}.ncu-rep ]; then
    ncu -i $O/prof_${k}.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r2_source_${k}.csv.gz
  fi
done
rm -f $O/prof_d*_p*.ncu-rep
du -sh $O
echo done

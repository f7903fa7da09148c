#!/usr/bin/env bash
# Round-2 GPU job 8 (1 GPU): full suite and the driver's bench line on the
# final tree, EPB / occupancy variants across orders (both precisions), launch
# list and ncu --set full of the headline kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > $O/r2_run8_pytest.log 2>&1
tail -3 $O/r2_run8_pytest.log
timeout 900 python bench.py > $O/r2_bench_n1_final.json 2> $O/r2_bench_n1_final.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/r2_bench_ref_final.json 2> $O/r2_bench_ref_final.err
EXP=$PWD/swirl_fem_b200/lib_exp/libswirl_b200.so
SFEM_LIB=$EXP timeout 900 python tools/bench_apply.py --dim 3 --orders 4,5,6,7,8 --target-dofs 16e6 \
  --variants 0,3,4,5,6,7 --dtypes f64,f32 --check 0 > $O/r2_variants_epb_orders.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/r2_launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-e2e \
  --no-cpu-baseline --no-extra --no-parity --cg-iters 5 > $O/r2_ncu_launches_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:apply3d_v2 -s 8 -c 1 -o $O/prof_final -f \
  python bench.py --steps 4 --warmup 5 --no-e2e --no-cpu-baseline --no-extra --no-parity --cg-iters 0 > $O/r2_ncu_final.log 2>&1
if [ -f $O/prof_final.ncu-rep ]; then
  ncu -i $O/prof_final.ncu-rep --page raw --csv > $O/prof_final_raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/prof_final_raw.csv > $O/r2_ncu_apply3d_ne68_final.txt
  ncu -i $O/prof_final.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r2_source_apply3d_ne68_final.csv.gz
  rm -f $O/prof_final_raw.csv $O/prof_final.ncu-rep
fi
timeout 900 ncu --set full --clock-control none -k regex:cg_step -s 3 -c 1 -o $O/prof_step -f \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra --no-parity --cg-iters 8 > $O/r2_ncu_step.log 2>&1
if [ -f $O/prof_step.ncu-rep ]; then
  ncu -i $O/prof_step.ncu-rep --page raw --csv > $O/prof_step_raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/prof_step_raw.csv > $O/r2_ncu_cg_step_ne68.txt
  rm -f $O/prof_step_raw.csv $O/prof_step.ncu-rep
fi
du -sh $O
echo done

#!/usr/bin/env bash
# ncu captures for the order sweep (BASELINE config 5): one `--set full` launch
# of the apply kernel per (dim, order, dtype), condensed with
# tools/ncu_summary.py (DRAM traffic, FP64 / FMA pipe utilisation, stall
# breakdown) into gpurun_out/ncu_orders_<tag>.txt.  Run under gpurun on ONE GPU,
# only after the same bench_apply command has exited 0 without ncu:
#   gpurun --timeout 900 -- 'bash tools/profile_orders.sh 3 "3 7 11 13 15" "f64 f32"'
set -u
DIM=${1:-3}
ORDERS=${2:-"3 7 11 13 15"}
DTYPES=${3:-"f64 f32"}
mkdir -p gpurun_out
for p in $ORDERS; do
  for t in $DTYPES; do
    tag="d${DIM}_p${p}_${t}"
    rep="gpurun_out/prof_${tag}"
    timeout 300 ncu --set full --clock-control none --import-source on \
      -k "regex:apply${DIM}d_v2" -s 5 -c 1 -o "$rep" -f \
      python tools/bench_apply.py --dim "$DIM" --orders "$p" --dtypes "$t" \
        --target-dofs 16e6 --check 0 --reps 5 > "gpurun_out/ncu_${tag}.log" 2>&1
    if [ -f "${rep}.ncu-rep" ]; then
      ncu -i "${rep}.ncu-rep" --page raw --csv > "${rep}_raw.csv" 2>/dev/null
      python tools/ncu_summary.py "${rep}_raw.csv" > "gpurun_out/ncu_orders_${tag}.txt"
      rm -f "${rep}_raw.csv"
      echo "== ${tag}"; grep -E "duration|dram__bytes|pipe_fp64|pipe_fma|warps_active" \
        "gpurun_out/ncu_orders_${tag}.txt"
    else
      echo "== ${tag}: no report (see gpurun_out/ncu_${tag}.log)"
    fi
  done
done

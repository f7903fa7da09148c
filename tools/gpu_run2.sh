#!/usr/bin/env bash
# Round-2 GPU job 2 (1 GPU): failing tests in full, the driver's bench line
# with all its new parts, launch list of a short bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short \
  -k "apply_matches_oracle or single_process or kolmogorov or fespace_matches" \
  > $O/r2_run2_pytest_failing.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=line 2>&1 | tail -40 > $O/r2_run2_pytest_all.log
timeout 900 python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
timeout 600 python bench.py --impl reference > $O/r2_bench_ref.json 2> $O/r2_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e \
  --no-cpu-baseline --no-extra --no-parity --cg-iters 5 > $O/r2_ncu_launches.log 2>&1
du -sh $O
echo done

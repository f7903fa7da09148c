#!/usr/bin/env bash
# Round-2 GPU job 3 (1 GPU): full suite, PDL on/off at two sizes, Stokes step
# with the device-state (graph-captured) CG.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > $O/r2_run3_pytest.log 2>&1
tail -5 $O/r2_run3_pytest.log
for pdl in 1 0; do for ne in 68 34; do
  SFEM_PDL=$pdl timeout 600 python bench.py --ne $ne --steps 30 --warmup 5 --no-e2e \
    --no-cpu-baseline --no-extra --no-parity --cg-iters 30 \
    > $O/r2_pdl${pdl}_ne${ne}.json 2> $O/r2_pdl${pdl}_ne${ne}.err
done; done
timeout 900 python tools/bench_apply.py --dim 2 --orders 1,2,3,4,5,6,7,8,9,11,15 --target-dofs 16e6 --variants 0,3 --dtypes f64,f32 --check 1 > $O/r2_2d_warp_vs_block.log 2>&1
timeout 600 python tools/bench_ns.py --ne 64 --order 7 > $O/r2_bench_ns_graph.json 2> $O/r2_bench_ns_graph.err
SFEM_CG_GRAPH=0 timeout 600 python tools/bench_ns.py --ne 64 --order 7 > $O/r2_bench_ns_eager.json 2> $O/r2_bench_ns_eager.err
timeout 600 python tools/bench_ns.py --ne 256 --order 7 > $O/r2_bench_ns_graph_ne256.json 2> $O/r2_bench_ns_graph_ne256.err
du -sh $O
echo done

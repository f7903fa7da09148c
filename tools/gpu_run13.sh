#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( echo "== default"; timeout 120 python tools/debug_lazy.py ) > $O/r2_debug_lazy.log 2>&1
cat $O/r2_debug_lazy.log | cut -c1-220 | head -30
bash tools/gpu_run12.sh

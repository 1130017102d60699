#!/usr/bin/env bash
# A/B of contraction-kernel builds (k-loop unroll factor of strip_steps) on one box
set -u
mkdir -p gpurun_out
for rep in 1 2; do
for v in "" _u1 _u4 _u8; do
  lib=$PWD/qcdmet_b200/libqcdmet_b200$v.so
  [ -f "$lib" ] || continue
  for wl in sn2 polyyne; do
    QCDMET_B200_LIB=$lib timeout 100 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rep$rep lib=[$v] $wl kernel_ms', round(d['roofline']['kernel_ms'],4))"
  done
done
done

#!/usr/bin/env bash
# A/B of fused-kernel builds on one box: $QCDMET_B200_LIB selects the library
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2l}
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print(sys.argv[2], "no line", e)
PY
}
for rep in 1 2; do
for v in "" _v1 _v3 ${VARIANTS:-}; do
  lib=$PWD/qcdmet_b200/libqcdmet_b200$v.so
  [ -f "$lib" ] || continue
  for n in 576 1000 2000; do
    QCDMET_B200_LEAD=0 QCDMET_B200_LIB=$lib timeout 200 python bench.py --norb $n --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_ab${v}_N$n.json 2> gpurun_out/${TAG}_ab${v}_N$n.err
    line gpurun_out/${TAG}_ab${v}_N$n.json "rep$rep lib=[$v] N=$n"
  done
  QCDMET_B200_LEAD=0 QCDMET_B200_LIB=$lib timeout 300 python bench.py --norb 4000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_ab${v}_N4000.json 2> gpurun_out/${TAG}_ab${v}_N4000.err
  line gpurun_out/${TAG}_ab${v}_N4000.json "rep$rep lib=[$v] N=4000"
done
done

#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2u}
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bounds.py -m gpu -q --tb=short -k "resident or factorised or bounds" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"], 4), "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print(sys.argv[2], "no line", e)
PY
}
for wl in sn2 polyyne; do
  for var in 0 1 2; do
    QCDMET_B200_CONTRACT_VARIANT=$var timeout 100 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_v${var}_$wl.json 2> gpurun_out/${TAG}_v${var}_$wl.err
    line gpurun_out/${TAG}_v${var}_$wl.json "variant $var $wl"
  done
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:contract_resident_ws -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_contract_ws_sn2 \
  python bench.py --workload sn2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_ncu_contract.log 2>&1
echo "ncu rc=$?"

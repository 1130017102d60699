#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2j}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_misc.py -m gpu -q --tb=short -k "resident or misc or factorised or auto_reports or norotation or hessian" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log
timeout 600 python -m pytest tests/test_dmet_loop.py -m gpu -q --tb=short -k "lm" > gpurun_out/${TAG}_pytest_lm.log 2>&1; echo "pytest lm rc=$?"; tail -8 gpurun_out/${TAG}_pytest_lm.log
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"], 4), "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "frac", round(d["roofline"]["frac"], 4), "value", round(d["value"]))
except Exception as e:
    print(sys.argv[2], "no line", e)
PY
}
for wl in sn2 polyyne; do
  QCDMET_B200_CONTRACT_RESIDENT=0 timeout 200 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_chunked_$wl.json 2> gpurun_out/${TAG}_chunked_$wl.err
  line gpurun_out/${TAG}_chunked_$wl.json "chunked $wl"
  timeout 200 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
  line gpurun_out/${TAG}_bench_$wl.json "resident $wl"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:contract_resident -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_contract_sn2 \
  python bench.py --workload sn2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_ncu_contract.log 2>&1
echo "ncu contract rc=$?"

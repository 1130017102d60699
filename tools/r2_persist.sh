#!/usr/bin/env bash
# round-2 check of the persistent fused kernel: parity tests (DMET loops excluded), then the size sweep
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2c}
timeout 600 python -m pytest tests -m gpu -q --tb=short -x --deselect tests/test_dmet_loop.py > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
for n in 200 400 576 1000 2000; do
  timeout 200 python bench.py --norb $n --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_bench_N$n.json 2> gpurun_out/${TAG}_bench_N$n.err
  echo "bench N=$n rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_N$n.json").read().strip().splitlines()[-1])
    print($n, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
except Exception as e:
    print("no line", e)
PY
done
for wl in hubbard24 hubbard24_bath; do
  timeout 200 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
  echo "bench $wl rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_$wl.json").read().strip().splitlines()[-1])
    print("$wl", d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
except Exception as e:
    print("no line", e)
PY
done
timeout 400 python bench.py --norb 4000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_N4000.json 2> gpurun_out/${TAG}_bench_N4000.err
echo "bench N=4000 rc=$?"; cat gpurun_out/${TAG}_bench_N4000.json | cut -c1-900; tail -3 gpurun_out/${TAG}_bench_N4000.err

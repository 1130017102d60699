#!/usr/bin/env bash
# sweep of the consumer-warp stagger ($QCDMET_B200_LEAD) of the persistent fused kernel
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2d}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -m gpu -q --tb=short -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"], 4), "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print(sys.argv[2], "no line", e)
PY
}
for lead in 0 1 2 3 4 5 7; do
  for n in 200 400 576 1000; do
    QCDMET_B200_LEAD=$lead timeout 200 python bench.py --norb $n --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_lead${lead}_N$n.json 2> gpurun_out/${TAG}_lead${lead}_N$n.err
    line gpurun_out/${TAG}_lead${lead}_N$n.json "lead=$lead N=$n"
  done
done
for lead in 0 2 3 4; do
  QCDMET_B200_LEAD=$lead timeout 200 python bench.py --norb 2000 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_lead${lead}_N2000.json 2> gpurun_out/${TAG}_lead${lead}_N2000.err
  line gpurun_out/${TAG}_lead${lead}_N2000.json "lead=$lead N=2000"
  QCDMET_B200_LEAD=$lead timeout 200 python bench.py --workload hubbard24_bath --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_lead${lead}_h24bath.json 2> gpurun_out/${TAG}_lead${lead}_h24bath.err
  line gpurun_out/${TAG}_lead${lead}_h24bath.json "lead=$lead hubbard24_bath"
done
for lead in 0 3 5; do
  QCDMET_B200_LEAD=$lead timeout 300 python bench.py --norb 4000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_lead${lead}_N4000.json 2> gpurun_out/${TAG}_lead${lead}_N4000.err
  line gpurun_out/${TAG}_lead${lead}_N4000.json "lead=$lead N=4000"
done

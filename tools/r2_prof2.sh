#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2f}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py tests/test_gpu_bath.py -m gpu -q --tb=short -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
rm -f gpurun_out/${TAG}_phases.txt
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"], 4), "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print(sys.argv[2], "no line", e)
PY
}
for mode in 0 1; do
  for lead in 0 4; do
    echo "== N=576 lead=$lead mode=$mode" >> gpurun_out/${TAG}_phases.txt
    QCDMET_B200_LEAD_MODE=$mode QCDMET_B200_LEAD=$lead QCDMET_B200_PROFILE=gpurun_out/${TAG}_phases.txt timeout 200 python bench.py --norb 576 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > /dev/null 2> gpurun_out/${TAG}_prof.err
    for n in 400 576 1000; do
      QCDMET_B200_LEAD_MODE=$mode QCDMET_B200_LEAD=$lead timeout 200 python bench.py --norb $n --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_m${mode}l${lead}_N$n.json 2> gpurun_out/${TAG}_m${mode}l${lead}_N$n.err
      line gpurun_out/${TAG}_m${mode}l${lead}_N$n.json "mode=$mode lead=$lead N=$n"
    done
  done
done
cat gpurun_out/${TAG}_phases.txt

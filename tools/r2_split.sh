#!/usr/bin/env bash
# tile-granular balancing of the fused kernel (SPLIT instantiation): parity, then timing next to the unsplit launch
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2v}
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "tile_split" > gpurun_out/${TAG}_pytest_split.log 2>&1
echo "pytest split rc=$?"; tail -6 gpurun_out/${TAG}_pytest_split.log
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py tests/test_gpu_bounds.py -m gpu -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"], 4), "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print(sys.argv[2], "no line", e)
PY
}
for n in 400 576 1000 2000; do
  for sp in 0 auto; do
    if [ $sp = auto ]; then unset QCDMET_B200_SPLIT; else export QCDMET_B200_SPLIT=$sp; fi
    timeout 120 python bench.py --norb $n --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_split${sp}_N$n.json 2> gpurun_out/${TAG}_split${sp}_N$n.err
    line gpurun_out/${TAG}_split${sp}_N$n.json "split=$sp N=$n"
  done
done
unset QCDMET_B200_SPLIT
for wl in hubbard24 hubbard24_bath; do
  for sp in 0 auto; do
    if [ $sp = auto ]; then unset QCDMET_B200_SPLIT; else export QCDMET_B200_SPLIT=$sp; fi
    timeout 120 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_split${sp}_$wl.json 2> gpurun_out/${TAG}_split${sp}_$wl.err
    line gpurun_out/${TAG}_split${sp}_$wl.json "split=$sp $wl"
  done
done

#!/usr/bin/env bash
# Round-trip used under gpurun: smoke -> GPU tests -> bench at increasing sizes.  Every stage has its
# own timeout so a hung kernel cannot hold the box.
set -u
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 600 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for n in "$@"; do
  timeout 900 python bench.py --norb $n --steps 3 --warmup 3 > gpurun_out/bench_N$n.json 2> gpurun_out/bench_N$n.err
  echo "bench N=$n rc=$?"; cat gpurun_out/bench_N$n.json; tail -3 gpurun_out/bench_N$n.err
done

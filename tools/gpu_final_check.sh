#!/usr/bin/env bash
# Last gpurun round-trip of a round: smoke, GPU tests, and quick benches of the mid-size shapes.
set -u
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -6 gpurun_out/smoke.log
timeout 400 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 100 python bench.py --workload hubbard24 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_v4_hubbard24.json 2> gpurun_out/bench_v4_hubbard24.err
echo "bench hubbard24 rc=$?"; cat gpurun_out/bench_v4_hubbard24.json
for n in 1000 2000; do
  timeout 150 python bench.py --norb $n --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_v4_N$n.json 2> gpurun_out/bench_v4_N$n.err
  echo "bench N=$n rc=$?"; cat gpurun_out/bench_v4_N$n.json
done

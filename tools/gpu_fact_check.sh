#!/usr/bin/env bash
# gpurun round-trip for the factorised path: GPU tests, bench of the large-fragment shapes, ncu of the
# contraction and panel kernels.  Every stage has its own timeout.
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu_fact.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu_fact.log
for wl in sn2 polyyne; do
  timeout 200 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fact_$wl.json 2> gpurun_out/bench_fact_$wl.err
  echo "bench $wl rc=$?"; cat gpurun_out/bench_fact_$wl.json; tail -3 gpurun_out/bench_fact_$wl.err
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_fact_sn2.csv \
  python bench.py --workload sn2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_fact.log 2>&1
echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:contract_panels -s 3 -c 1 -f -o gpurun_out/prof_contract_sn2 \
  python bench.py --workload sn2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_fact.log 2>&1
echo "ncu full rc=$?"

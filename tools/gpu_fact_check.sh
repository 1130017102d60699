#!/usr/bin/env bash
# gpurun round-trip: GPU tests, bench of the large-fragment shapes and of the north-star shape, ncu of the
# contraction kernel and (optionally, NCU4000=1) of the main DMMA kernel at N = 4000.  Every stage has its
# own timeout.
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu_fact.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu_fact.log
for wl in sn2 polyyne; do
  timeout 200 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fact_$wl.json 2> gpurun_out/bench_fact_$wl.err
  echo "bench $wl rc=$?"; cat gpurun_out/bench_fact_$wl.json; tail -3 gpurun_out/bench_fact_$wl.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:contract_panels -s 3 -c 1 -f -o gpurun_out/prof_contract_sn2 \
  python bench.py --workload sn2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_fact.log 2>&1
echo "ncu contract rc=$?"
if [ "${NCU4000:-0}" = "1" ]; then
  timeout 400 python bench.py --norb 4000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l2order_N4000.json 2> gpurun_out/bench_l2order_N4000.err
  echo "bench N=4000 rc=$?"; cat gpurun_out/bench_l2order_N4000.json; tail -3 gpurun_out/bench_l2order_N4000.err
  timeout 500 ncu --set full --clock-control none -k regex:jac_dmma -c 1 -f -o gpurun_out/prof_dmma_v4_N4000 \
    python bench.py --norb 4000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --check-cols 0 > gpurun_out/ncu_full_N4000.log 2>&1
  echo "ncu N=4000 rc=$?"; tail -3 gpurun_out/ncu_full_N4000.log
fi
if [ "${WS:-0}" = "1" ]; then   # the opt-in warp-specialised contraction variant
  QCDMET_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests -m gpu -q --tb=short -k experimental > gpurun_out/pytest_ws.log 2>&1
  echo "pytest ws rc=$?"; tail -5 gpurun_out/pytest_ws.log
  for wl in sn2 polyyne; do
    QCDMET_B200_CONTRACT_WS=1 timeout 100 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ws_$wl.json 2> gpurun_out/bench_ws_$wl.err
    echo "bench ws $wl rc=$?"; cat gpurun_out/bench_ws_$wl.json
  done
fi

#!/usr/bin/env bash
# round-2 full check on one GPU: smoke, whole GPU suite, bench lines of every shape, ncu launch list + captures
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2x}
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short --durations=12 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/${TAG}_pytest.log
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print(sys.argv[2], "ms/step", round(d["ms_per_step"], 4), "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "frac", round(d["roofline"]["frac"], 4), "value", round(d["value"]), "e2e", e.get("value"))
except Exception as e:
    print(sys.argv[2], "no line", e)
PY
}
for wl in sn2 polyyne hubbard24 hubbard24_bath hubbard24_ti hubbard24_ti_bath; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
  line gpurun_out/${TAG}_bench_$wl.json "$wl"
done
for n in 200 400 576 1000 2000; do
  timeout 300 python bench.py --norb $n --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_N$n.json 2> gpurun_out/${TAG}_bench_N$n.err
  line gpurun_out/${TAG}_bench_N$n.json "N=$n"
done
timeout 600 python bench.py --lm > gpurun_out/${TAG}_bench_N4000.json 2> gpurun_out/${TAG}_bench_N4000.err
line gpurun_out/${TAG}_bench_N4000.json "N=4000"; tail -4 gpurun_out/${TAG}_bench_N4000.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_N4000.json 2> gpurun_out/${TAG}_bench_reference_N4000.err
echo "reference arm rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_reference_N4000.json
# ncu: launch list of the default bench command, then full captures of the dominant kernels
# (cuSOLVER's syevd alone is > 400 launches and belongs to the set-up: list this library's kernels)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:jac_dmma|build_|split_transpose|gather_rows|dblocks|checksum|contract_|sum_subcolumns|zero_columns|bath_|embed_|rdm_diag' --csv --log-file gpurun_out/${TAG}_launches_N4000.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
if [ "${NCU_ALL:-0}" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jac_dmma -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_dmma_N4000 \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_ncu_N4000.log 2>&1
echo "ncu N4000 rc=$?"
fi
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jac_dmma -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_dmma_hubbard24 \
  python bench.py --workload hubbard24 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_ncu_h24.log 2>&1
echo "ncu hubbard24 rc=$?"
[ "${NCU_ALL:-0}" = "1" ] && for wl in sn2 polyyne; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:contract_resident -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_contract_$wl \
    python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_ncu_contract_$wl.log 2>&1
  echo "ncu contract $wl rc=$?"
done

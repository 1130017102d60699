#!/usr/bin/env bash
# last check of the round: smoke, whole GPU suite, the default bench line and the large-fragment shapes
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2final}
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log
for wl in sn2 polyyne; do
  timeout 200 python bench.py --workload $wl --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
  python -c "import json; d=json.loads(open('gpurun_out/${TAG}_bench_$wl.json').read().strip().splitlines()[-1]); print('$wl', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity'])"
done
timeout 600 python bench.py > gpurun_out/${TAG}_bench_N4000.json 2> gpurun_out/${TAG}_bench_N4000.err
python -c "import json; d=json.loads(open('gpurun_out/${TAG}_bench_N4000.json').read().strip().splitlines()[-1]); print('N4000', d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['parity']['max_abs_err'], d['clocks'])"

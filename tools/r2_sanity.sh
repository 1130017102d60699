#!/usr/bin/env bash
# short post-cleanup check: smoke, the parity file, the default bench line
set -u
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r2s_smoke.log
timeout 140 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short > gpurun_out/r2s_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2s_pytest.log
timeout 110 python bench.py --steps 3 --warmup 3 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python -c "import json; d=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1]); print('N4000', d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['parity']['max_abs_err'])"

#!/usr/bin/env bash
# round-2 check: smoke, whole GPU suite (durations listed), one bench line at the target size
set -u
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/r2b_smoke.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short --durations=25 > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?"; tail -45 gpurun_out/r2b_pytest.log

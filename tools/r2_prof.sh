#!/usr/bin/env bash
# phase-cycle profile of the fused kernel (instrumented instantiation) + the batched-bath tests
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2e}
rm -f gpurun_out/${TAG}_phases.txt
for lead in 0 4; do
  for n in 576 1000 4000; do
    echo "== N=$n lead=$lead" >> gpurun_out/${TAG}_phases.txt
    QCDMET_B200_LEAD=$lead QCDMET_B200_PROFILE=gpurun_out/${TAG}_phases.txt timeout 200 python bench.py --norb $n --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity > /dev/null 2> gpurun_out/${TAG}_prof.err
  done
done
cat gpurun_out/${TAG}_phases.txt
timeout 900 python -m pytest tests/test_gpu_bath.py -m gpu -q --tb=short -s > gpurun_out/${TAG}_bath.log 2>&1
echo "bath pytest rc=$?"; tail -40 gpurun_out/${TAG}_bath.log

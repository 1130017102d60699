#!/usr/bin/env bash
# multi-GPU check on NG GPUs of one box: group tests on distinct devices, the driver's torchrun launch and the
# single-process (in-library) path behind the C ABI, at the north-star size and (SWEEP=1) at smaller sizes
set -u
NG=${NG:-2}
ONLY_SWEEP=${ONLY_SWEEP:-0}
mkdir -p gpurun_out
TAG=${TAG:-r2m}_g$NG
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
if [ "$ONLY_SWEEP" != "1" ]; then
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -q --tb=short > gpurun_out/${TAG}_pytest_group.log 2>&1
echo "pytest group rc=$?"; tail -5 gpurun_out/${TAG}_pytest_group.log
fi
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    sp = d.get("single_process_api") or {}
    print(sys.argv[2], "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "frac", round(d["roofline"]["frac"], 4),
          "e2e", e.get("value") and round(e["value"]), "e2e_ms", e.get("ms_per_step") and round(e["ms_per_step"], 1),
          "ranks_identical", d.get("ranks_identical"), "parity", (d.get("parity") or {}).get("max_abs_err"),
          "single_process", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in sp.items() if k in ("ms_per_step", "value", "e2e_ms_per_step", "e2e_value")})
except Exception as e:
    print(sys.argv[2], "no line", e)
PY
}
if [ "$ONLY_SWEEP" != "1" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/${TAG}_torchrun_N4000.json 2> gpurun_out/${TAG}_torchrun_N4000.err
echo "torchrun rc=$?"; line gpurun_out/${TAG}_torchrun_N4000.json "torchrun N=4000 gpus=$NG"; tail -3 gpurun_out/${TAG}_torchrun_N4000.err
fi
if [ "$ONLY_SWEEP" != "1" ]; then
timeout 900 python bench.py --gpus $NG --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_inlib_N4000.json 2> gpurun_out/${TAG}_inlib_N4000.err
echo "in-library rc=$?"; line gpurun_out/${TAG}_inlib_N4000.json "single-process N=4000 gpus=$NG"; tail -3 gpurun_out/${TAG}_inlib_N4000.err
fi
if [ "${SWEEP:-0}" = "1" ]; then
  for n in ${SWEEP_SIZES:-1000 2000}; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --norb $n --steps 5 --warmup 3 --no-cpu-baseline --no-single-process > gpurun_out/${TAG}_torchrun_N$n.json 2> gpurun_out/${TAG}_torchrun_N$n.err
    line gpurun_out/${TAG}_torchrun_N$n.json "torchrun N=$n gpus=$NG"
  done
  for wl in sn2 hubbard24; do
    timeout 600 python bench.py --gpus $NG --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_inlib_$wl.json 2> gpurun_out/${TAG}_inlib_$wl.err
    line gpurun_out/${TAG}_inlib_$wl.json "single-process $wl gpus=$NG"
  done
fi

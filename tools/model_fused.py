"""Pipe-time model of the fused DMMA kernel: from the host-only schedule (qcdmet_schedule_info)
the number of CTAs, whole waves on 148 SMs and the DMMA-pipe time of one CTA, against the kernel times in
profiles/.  No device needed:  python tools/model_fused.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qcdmet_b200 import _lib, workloads as wl   # noqa: E402

SMS, GHZ = 148, 1.965
print("| Norb | CTAs | waves | stages/CTA | DMMA time of a CTA (us) | whole waves (ms) | measured kernel (ms) | model/measured | per-wave gap (us) |")
print("|---|---|---|---|---|---|---|---|---|")
for n in (200, 400, 576, 1000, 2000, 4000):
    w = wl.sweep(n)
    info = _lib.schedule_info(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.row_idx)
    pairs, nj, n_ot = info["block_pairs"], info["nj"], info["column_tiles"]
    stages = n_ot * ((n - w.num_pairs + 15) // 16)
    # per stage and SM sub-partition: 2 warps x (4 k-steps x 4 row fragments x nj column fragments) DMMA x 16 cycles
    t_cta = stages * 2 * 4 * 4 * nj * 16 / (GHZ * 1e9)
    waves = -(-pairs // SMS)
    measured = None
    for f in (f"r1_bench_dmma_v4_N{n}.json", f"r1_bench_dmma_v3_N{n}.json"):
        try:
            measured = json.load(open(os.path.join(ROOT, "profiles", f)))["roofline"]["kernel_ms"]
            break
        except OSError:
            pass
    model = waves * t_cta * 1e3
    print(f"| {n} | {pairs} | {pairs / SMS:.2f} -> {waves} | {stages} | {t_cta * 1e6:.1f} | {model:.3f} | {measured:.3f} | "
          f"{model / measured:.2f} | {(measured - model) / waves * 1e3:.0f} |")

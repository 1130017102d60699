#!/usr/bin/env python
"""bench.py -- umat-fit Jacobian throughput (columns/s) on N B200s, vs the reference on host cores.

    python bench.py --gpus N --steps K --warmup W                (this framework)
    python bench.py --impl reference --gpus N --steps K --warmup W    (reference CPU kernel)

A step = one full (Q, P) Jacobian of the synthetic response sweep: Norb orbitals,
Nocc = Norb/2, 4-site fragments tiling all sites, P = 2.5 Norb perturbations; default Norb = 4000,
P = 10^4 (BASELINE.json north_star target).  `value` = columns/s with (C, e), perturbation list and
fragment rows resident in HBM (eigensolve excluded, as the metric defines); `e2e` = the same through
the reference-facing host API (H0 on the host -> H2D, cuSOLVER syevd, kernels, D2H of J).
At N > 1 (torchrun) the columns are sharded over the ranks (strong scaling: the Jacobian is fixed)
and the slices are all-gathered with NCCL inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os

# torchrun exports OMP_NUM_THREADS=1 to its children; the reference arm / CPU baseline (rank 0 only) must
# use every host core (libgomp reads the variable when it is first loaded, i.e. before `import torch`).
# The other ranks do no heavy host work (rank 0 builds the inputs and broadcasts them): keep them at one
# thread so that 8 ranks do not oversubscribe the box with spinning BLAS threads.
_RANK0 = int(os.environ.get("RANK", "0")) == 0
os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1) if _RANK0 else "1"
if not _RANK0:
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

print_json = None
METRIC = "umat_jacobian_columns_per_s"
UNIT = "columns/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sweep",
                    choices=["sweep", "hubbard24", "hubbard24_bath", "hubbard24_ti", "hubbard24_ti_bath", "polyyne", "sn2"],
                    help="sweep: BASELINE configs[4] (default, the north-star target at --norb 4000); hubbard24*: "
                         "configs[1] 24x24 Hubbard, 2x2 impurities (non-TI / TI, impurity-only / impurity+bath rows); "
                         "polyyne / sn2: the shapes of configs[2] / configs[3]")
    ap.add_argument("--norb", type=int, default=int(os.environ.get("QCDMET_BENCH_NORB", 4000)))
    ap.add_argument("--nterms", type=int, default=None, help="default 2.5*norb")
    ap.add_argument("--kernel", default="auto", choices=["auto", "general", "dmma", "factorised"])
    ap.add_argument("--cpu-sample-cols", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--check-cols", type=int, default=8, help="columns checked against the oracle (0 = off)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except ValueError:
                continue
            for n, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# peaks
# ------------------------------------------------------------------------------------------
def fp64_peak():
    """FP64 ceiling for the roofline.  MEASURED_PEAKS.json (driver-written) carries only HBM and
    bf16; the FP64 DMMA ceiling is measured by tools/fp64_peaks.cu on this pool's B200 and kept in
    profiles/fp64_peaks.json.  Fallback: 37 TFLOP/s (148 SMs x 64 FP64 FMA/clk x 1.965 GHz)."""
    path = os.path.join(ROOT, "profiles", "fp64_peaks.json")
    try:
        d = json.load(open(path))
        return float(d["dmma_tflops_sustained"]), float(d.get("dmma_tflops_w8", 0.0)), "measured (profiles/fp64_peaks.json: DMMA.8x8x4 issue rate, sustained)"
    except (OSError, KeyError, ValueError):
        return 37.2, 37.2, "fallback (148 SM x 64 DFMA/clk x 1.965 GHz; profiles/fp64_peaks.json absent)"


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference kernel (oracle/_ref) on host cores
# ------------------------------------------------------------------------------------------
_T_EIG = {}


def cpu_reference_rate(w, sample_cols, repeats=1):
    """columns/s of lib/rhf_response.cc on the host: T(P_cpu) - T(0) isolates the response loop
    from dsyev (the Nterms = 0 call is legal)."""
    from oracle import ref_loader
    if not ref_loader.available("mkl"):
        return None
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    sample_cols = min(sample_cols, w.nterms)
    hs = w.h1start[:sample_cols + 1]
    hr, hc = w.h1row[:hs[-1]], w.h1col[:hs[-1]]
    if w.norb not in _T_EIG:       # the Nterms = 0 call: dsyev + 1-RDM only (measured once)
        small = np.diag(np.arange(8.0))
        ref_loader.rhf_response(8, 4, hs[:1], hr[:0], hc[:0], small, "mkl")   # library/thread-pool warm-up
        best0 = None
        for _ in range(2 if w.norb <= 2000 else 1):
            t0 = time.perf_counter()
            ref_loader.rhf_response(w.norb, w.num_pairs, hs[:1], hr[:0], hc[:0], w.h0, "mkl")
            t = time.perf_counter() - t0
            best0 = t if best0 is None else min(best0, t)
        _T_EIG[w.norb] = best0
    t_eig = _T_EIG[w.norb]
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        ref_loader.rhf_response(w.norb, w.num_pairs, hs, hr, hc, w.h0, "mkl")
        t = time.perf_counter() - t0
        best = t if best is None else min(best, t)
    dt = max(best - t_eig, 1e-9)
    return {"value": sample_cols / dt, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"first {sample_cols} of {w.nterms} columns at Norb={w.norb}; T(P)-T(0), "
                      f"dsyev+1RDM T(0)={t_eig:.2f}s excluded; MKL (torch) build of lib/rhf_response.cc, "
                      f"OMP_NUM_THREADS={os.environ['OMP_NUM_THREADS']}",
            "eigensolve_s": t_eig, "response_s": dt}


def default_sample_cols(norb):
    cores = os.cpu_count() or 1
    # about 10-30 s of CPU work: reference cost 1.5 N^3 flop/column (a5+a6), ~25 GF/s per core assumed
    per_col = 1.5 * norb ** 3 / (25e9 * cores)
    cols = int(max(cores, min(20.0 / max(per_col, 1e-9), 2048)))
    cols = max(cores, cols // cores * cores)
    while norb * norb * cols >= 2 ** 31 or 8.0 * norb * norb * cols > 8e9:
        cols -= cores
    return max(cols, 1)


def log(msg):
    if os.environ.get("QCDMET_BENCH_VERBOSE", "1") != "0":
        print(f"[bench rank {os.environ.get('RANK', '0')} t={time.perf_counter() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


_T0 = time.perf_counter()


def make_workload(args, placeholder_h0=False):
    from qcdmet_b200 import workloads
    if args.workload == "sweep":
        if placeholder_h0:      # non-zero ranks: H0 arrives by broadcast from rank 0 (skips the N^3 host set-up)
            m = 4
            w = workloads.make_workload(f"sweep_N{args.norb}", args.norb, args.norb // 2, [m] * (args.norb // m),
                                        nterms=args.nterms, h0=np.zeros((args.norb, args.norb)))
        else:
            w = workloads.sweep(args.norb, args.nterms)
        desc = (f"response sweep Norb={w.norb} Nocc={w.num_pairs} P={w.nterms} "
                f"(4-site fragments, impurity-only rows, Q={w.nresid})")
    elif args.workload.startswith("hubbard24"):
        ti, bath = "_ti" in args.workload, args.workload.endswith("_bath")
        w = workloads.hubbard24(trans_inv=ti, fit_imp_bath=bath)
        desc = (f"2D Hubbard 24x24 half filling (PBC x APBC), 2x2 impurities, {'TI' if ti else 'non-TI'}, "
                f"{'impurity+bath' if bath else 'impurity-only'} rows: Norb=576 Nocc=288 P={w.nterms} Q={w.nresid}")
    elif args.workload == "polyyne":
        w = workloads.polyyne_like()
        desc = f"polyyne-like shape Norb=168 Nocc=36, 12 fragments of 14: P={w.nterms} Q={w.nresid}"
    else:
        w = workloads.sn2_like()
        desc = f"sn2-like shape Norb=321 Nocc=58, fragments [52, 24x10, 29]: P={w.nterms} Q={w.nresid}"
    return w, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    w, desc = make_workload(args)
    cols = args.cpu_sample_cols or default_sample_cols(w.norb)
    rates = []
    info = None
    warm = min(args.warmup, 1)    # each call is tens of seconds of CPU work: one warm-up is enough
    for i in range(warm + args.steps):
        info = cpu_reference_rate(w, cols)
        if info is None:
            print_json({"impl": "reference", "unavailable": "oracle/_ref/libqcdmet_ref_mkl.so not built"})
            return
        if i >= warm:
            rates.append(info)
    cols = min(cols, w.nterms)
    tot_cols = cols * len(rates)
    tot_t = sum(r["response_s"] for r in rates)
    value = tot_cols / tot_t
    info = dict(info, value=value)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 0,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / len(rates),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc, "sample_cols_per_step": cols},
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print_json(line)


# ------------------------------------------------------------------------------------------
# this framework
# ------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from qcdmet_b200 import _lib, sharding, workloads

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    log("start")
    bcast = world > 1 and args.workload == "sweep"
    w, desc = make_workload(args, placeholder_h0=(bcast and rank != 0))
    P, Q = w.nterms, w.nresid
    log(f"workload built: {desc}")
    # column shard, aligned to whole fragments (10 parameters each) so the ranks do equal work
    shards = sharding.column_shards(P, world, 10)
    lo, hi = shards[rank]
    per = shards[0][1] - shards[0][0]
    assert world == 1 or all(e_ - b_ == per for b_, e_ in shards), \
        "bench shards need P divisible into equal whole-fragment slices"

    # setup only: the metric takes (C, e) as given (eigensolve excluded).  Multi-rank: rank 0 diagonalises and
    # broadcasts, so every rank works on bit-identical orbitals.
    if bcast:
        n = w.norb
        h0_t = torch.from_numpy(np.ascontiguousarray(w.h0)).cuda() if rank == 0 else torch.empty((n, n), dtype=torch.float64, device="cuda")
        if rank == 0:
            e_t, c_t = torch.linalg.eigh(h0_t)          # set-up only (cuSOLVER via torch): not part of any timing
            c_t = c_t.contiguous()
        else:
            c_t = torch.empty((n, n), dtype=torch.float64, device="cuda")
            e_t = torch.empty(n, dtype=torch.float64, device="cuda")
        for t in (h0_t, c_t, e_t):
            dist.broadcast(t, 0)
        w.h0, c, e = h0_t.cpu().numpy(), c_t.cpu().numpy(), e_t.cpu().numpy()
        del h0_t, c_t, e_t
    elif w.norb >= 1000:
        e_t, c_t = torch.linalg.eigh(torch.from_numpy(np.ascontiguousarray(w.h0)).cuda())   # set-up only
        c, e = c_t.cpu().numpy(), e_t.cpu().numpy()
        del c_t, e_t
    else:
        e, c = np.linalg.eigh(w.h0)
    log("orbitals ready")
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx,
                     device=local)
    plan.set_kernel({"auto": _lib.KERNEL_AUTO, "general": _lib.KERNEL_GENERAL, "dmma": _lib.KERNEL_DMMA,
                     "factorised": _lib.KERNEL_FACTORISED}[args.kernel])
    plan.set_columns(lo, hi)
    plan.set_orbitals(c, e)

    gathered = mine = None
    fused = False
    if world > 1:
        # preferred: fused compute + all-gather -- every rank maps every rank's full Jacobian (CUDA IPC) and
        # the kernel stores each element into all of them over NVLink; the block-pair work list (symmetric
        # mode) is split evenly over the ranks.  Needs peer mapping; otherwise fall back to column shards +
        # NCCL all-gather.
        try:
            if os.environ.get("QCDMET_BENCH_NO_FUSED"):
                raise _lib.QcdmetError(0, "fused all-gather disabled by QCDMET_BENCH_NO_FUSED")
            plan.set_columns(0, P)
            handles = [None] * world
            dist.all_gather_object(handles, plan.export_jacobian())
            plan.attach_peers(rank, handles)
            fused = True
        except _lib.QcdmetError as exc:
            print(f"[rank {rank}] fused all-gather unavailable ({exc}); falling back to NCCL all-gather", file=sys.stderr)
        flags = [None] * world
        dist.all_gather_object(flags, fused)
        fused = all(flags)
        if fused:
            lo, hi = 0, P
            sync_t = torch.zeros(1, device="cuda")
        else:
            plan.detach_peers()
            plan.set_columns(lo, hi)
            # in-place all-gather: each rank's kernels write straight into its slot of the gathered buffer
            gathered = torch.empty(world * per * Q, dtype=torch.float64, device="cuda")
            mine = gathered[rank * per * Q:(rank + 1) * per * Q]
            plan.set_output_device(mine.data_ptr(), mine.numel())
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")

    def step():
        flush.zero_()                       # evict L2 between steps (256 MB > 126 MB L2)
        plan.compute()
        if fused:
            dist.all_reduce(sync_t)          # cross-rank barrier on the stream: all peer stores have landed
        elif world > 1:
            dist.all_gather_into_tensor(gathered, mine)

    log(f"plan ready (fused={fused})")
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    log("warm-up done")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()          # after rank 0's sampler start-up, so that all ranks enter together
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    log(f"timed region done: {ms_total / args.steps:.2f} ms/step")
    if os.environ.get("QCDMET_BENCH_DEBUG"):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); plan.compute(); t1 = time.perf_counter(); plan.sync(); t2 = time.perf_counter()
        if fused:
            dist.all_reduce(sync_t)
        elif world > 1:
            dist.all_gather_into_tensor(gathered, mine)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        print(f"[rank {rank}] launch {1e3*(t1-t0):.2f} ms, kernel {1e3*(t2-t1):.2f} ms, gather {1e3*(t3-t2):.2f} ms",
              file=sys.stderr, flush=True)
    clocks = sampler.stop() if rank == 0 else None
    # dominant kernel: events recorded by the library on the launching stream around that launch
    for _ in range(max(1, min(args.steps, 3))):
        flush.zero_()
        plan.compute()
        plan.sync()
        kernel_ms.append(plan.kernel_time_ms())
    k_ms = float(np.mean(kernel_ms))
    flops = plan.flops
    launches = plan.launches
    kernel_used = plan.kernel_used
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = P / (ms_step * 1e-3)

    # parity gate on a few columns (oracle as the checker only)
    check = None
    if rank == 0 and args.check_cols > 0:
        from oracle import response_oracle as ro
        ncheck = min(args.check_cols, hi - lo)
        jac = plan.jacobian()
        ref = ro.jacobian_from_orbitals(c, e, w.num_pairs, w.h1start[:ncheck + 1], w.h1row[:w.h1start[ncheck]],
                                        w.h1col[:w.h1start[ncheck]], w.row_start, w.rows, w.row_idx)
        check = float(np.abs(jac[:, :ncheck] - ref).max())

    # end to end through the reference-facing host API (rank-local shard; host buffers)
    e2e = None
    if not args.no_e2e:
        h0_pinned = torch.from_numpy(np.ascontiguousarray(w.h0)).pin_memory()
        need_out = (hi - lo) if (not fused or rank == 0) else 0
        out_pinned = torch.empty(Q * max(need_out, 1), dtype=torch.float64).pin_memory()
        h0_np = h0_pinned.numpy()
        out_np = out_pinned.numpy()[:Q * need_out].reshape((Q, need_out), order="F")

        def e2e_step():
            t_a = time.perf_counter()
            plan.set_h0(h0_np)              # H2D + cuSOLVER syevd + panels
            if os.environ.get("QCDMET_BENCH_DEBUG") and rank == 0:
                print(f"[e2e] set_h0 (H2D + syevd + panels) {1e3 * (time.perf_counter() - t_a):.1f} ms", file=sys.stderr)
            plan.compute()
            if fused:
                dist.all_reduce(sync_t)
                if rank != 0:
                    torch.cuda.synchronize()
                    return                  # the least-squares update lives on rank 0's host
            plan.jacobian(out=out_np)       # D2H (blocking)

        for _ in range(min(args.warmup, 2)):
            e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        nst = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(nst):
            e2e_step()
        torch.cuda.synchronize()
        t_e2e = (time.perf_counter() - t0) / nst
        if world > 1:
            t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_e2e = float(t.item())
        e2e = {"value": P / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(8 * w.norb * w.norb),
               "d2h_bytes_per_step": int(8 * Q * need_out), "ms_per_step": 1e3 * t_e2e,
               "includes": "H2D of H0 from pinned host memory, cuSOLVER syevd, panel build, Jacobian kernels, "
                           "D2H into pinned host memory (rank's slice; fused multi-GPU mode: the full Jacobian on rank 0)"}

    if rank == 0:
        peak, peak_burst, peak_src = fp64_peak()
        achieved = flops / (k_ms * 1e-3) * 1e-12 if k_ms > 0 else None
        # SURVEY.md 8(d) per-perturbation count W_col = 2 S nv no + (2 nnz+1) nv no + 2 no sum r_f^2 + 2 sum r_f^2
        nv = w.norb - w.num_pairs
        nnz = np.diff(w.h1start[lo:hi + 1]).astype(float)
        w_cols = float(np.sum(2.0 * w.nrows * nv * w.num_pairs + (2.0 * nnz + 1.0) * nv * w.num_pairs
                              + 2.0 * w.num_pairs * Q + 2.0 * Q))
        if fused:
            w_cols /= world      # this rank executed 1/world of the block-pair list
        cpu = None
        if not args.no_cpu_baseline:
            cpu = cpu_reference_rate(w, args.cpu_sample_cols or default_sample_cols(w.norb))
        traffic = alg_bytes = None
        try:
            key = (f"sweep_N{w.norb}_P{P}_g{world}" if args.workload == "sweep" else f"{args.workload}_g{world}")
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(key)
            if tr:
                traffic, alg_bytes = tr["traffic_bytes"], tr["algorithmic_bytes"]
        except (OSError, ValueError, KeyError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc,
                       "kernel": {1: "general", 2: "dmma", 4: "factorised"}.get(kernel_used, str(kernel_used)),
                       "sharding": ("single GPU" if world == 1 else
                                    f"{world} ranks, block-pair work list split evenly, fused all-gather: each Jacobian element "
                                    "stored into every rank's buffer by NVLink peer stores, stream barrier (1-element NCCL all-reduce) "
                                    "inside the timed region" if fused else
                                    f"{world} x {per} columns (whole fragments), NCCL all-gather of the slices inside the timed region"),
                       "l2": "256 MB buffer zeroed between steps (L2 flush), inside the timed region"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "algorithmic_bytes": alg_bytes,
                         "kernel_ms": k_ms, "flops_per_launch": flops,
                         "peak_source": peak_src,
                         "algorithmic_equiv_tflops": (w_cols / (k_ms * 1e-3) * 1e-12) if k_ms > 0 else None,
                         "note": "achieved = FP64 flops actually executed by the dominant kernel on this rank / its "
                                 "CUDA-event time (site-factorised algorithm, SURVEY 8d: executed flops disclosed); "
                                 "algorithmic_equiv = the per-perturbation count W_col x columns / the same time"},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches * args.steps),
            "clocks": clocks,
            "parity_check_max_abs_err": check,
        }
        print_json(line)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: everything libraries print (e.g. NCCL's version banner, written
    # to fd 1 from C) is routed to stderr while the benchmark runs.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real_stdout, "w")
    global print_json
    print_json = lambda line: (out.write(json.dumps(line) + "\n"), out.flush())
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- umat-fit Jacobian throughput (columns/s) on N B200s, vs the reference on host cores.

    python bench.py --gpus N --steps K --warmup W                (this framework)
    python bench.py --impl reference --gpus N --steps K --warmup W    (reference CPU kernel)

A step = one full (Q, P) Jacobian of the synthetic response sweep: Norb orbitals,
Nocc = Norb/2, 4-site fragments tiling all sites, P = 2.5 Norb perturbations; default Norb = 4000,
P = 10^4 (BASELINE.json north_star target).  `value` = columns/s with (C, e), perturbation list and
fragment rows resident in HBM (eigensolve excluded, as the metric defines); `e2e` = the same through
the reference-facing host API (H0 on the host -> H2D, ONE cuSOLVER syevd, kernels, D2H of J).

Multi-GPU, two launch modes, same kernels:
  * torchrun (one process per GPU, what the driver launches for N > 1): the block-pair work list is split
    over the ranks and every Jacobian element is stored into every rank's buffer by NVLink peer stores
    (fused all-gather); the stream barrier is a 1-element NCCL all-reduce.
  * plain `python bench.py --gpus N` (one process): the same thing inside the library
    (qcdmet_plan_set_devices: the drop-in ctypes boundary spans the GPUs, no torch.distributed).
    Under torchrun rank 0 additionally times this single-process path (`single_process_api`).
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
    ap.add_argument("--no-parity", action="store_true", help="skip the parity gate against the compiled reference")
    ap.add_argument("--no-single-process", action="store_true",
                    help="torchrun only: do not time the in-library multi-GPU path on rank 0 as well")
    ap.add_argument("--lm", action="store_true", help="also time one device-side Levenberg-Marquardt step (f-3)")
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
def committed_fp64_peak():
    """The FP64 DMMA ceiling committed in round 1 (tools/fp64_peaks.cu on this pool's B200): kept beside the
    in-run measurement (qcdmet_measure_fp64_peak) so that the two can be compared."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "fp64_peaks.json")))
        return float(d["dmma_tflops_sustained"])
    except (OSError, KeyError, ValueError):
        return None


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference kernel (oracle/_ref) on host cores
# ------------------------------------------------------------------------------------------
_T_EIG = {}


def spread_columns(nterms, ncols):
    """ncols distinct column indices spread evenly over [0, nterms) (first and last included)."""
    ncols = max(1, min(ncols, nterms))
    return np.unique(np.linspace(0, nterms - 1, ncols).round().astype(np.int64))


def cpu_reference_rate(w, cols, repeats=1, keep_response=False):
    """columns/s of lib/rhf_response.cc on the host for the perturbations `cols` (index list): T(P_cpu) - T(0)
    isolates the response loop from dsyev (the Nterms = 0 call is legal).  With keep_response
    the (len(cols), N, N) response of the last call is returned under "response" (parity gate)."""
    from oracle import ref_loader
    if not ref_loader.available("mkl"):
        return None
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    cols = np.asarray(cols, dtype=np.int64)
    lens = (w.h1start[cols + 1] - w.h1start[cols]).astype(np.int64)
    hs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    take = np.concatenate([np.arange(w.h1start[c], w.h1start[c + 1]) for c in cols]) if len(cols) else np.zeros(0, int)
    hr, hc = w.h1row[take], w.h1col[take]
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
    best, resp = None, None
    for _ in range(repeats):
        t0 = time.perf_counter()
        resp, _ = ref_loader.rhf_response(w.norb, w.num_pairs, hs, hr, hc, w.h0, "mkl")
        t = time.perf_counter() - t0
        best = t if best is None else min(best, t)
    dt = max(best - t_eig, 1e-9)
    out = {"value": len(cols) / dt, "unit": UNIT, "cores": cores, "kind": "reference",
           "sample": f"{len(cols)} of {w.nterms} columns spread over [0, P) at Norb={w.norb}; T(P)-T(0), "
                     f"dsyev+1RDM T(0)={t_eig:.2f}s excluded; MKL (torch) build of lib/rhf_response.cc with "
                     f"g++ -O3 -march=x86-64-v3 -fopenmp (the reference's CMake uses -march=native), "
                     f"OMP_NUM_THREADS={os.environ['OMP_NUM_THREADS']}",
           "eigensolve_s": t_eig, "response_s": dt}
    if keep_response:
        out["response"] = resp
    return out


def default_sample_cols(norb):
    cores = os.cpu_count() or 1
    # about 10-30 s of CPU work: reference cost 1.5 N^3 flop/column (a5+a6), ~25 GF/s per core assumed
    per_col = 1.5 * norb ** 3 / (25e9 * cores)
    cols = int(max(cores, min(20.0 / max(per_col, 1e-9), 2048)))
    cols = max(cores, cols // cores * cores)
    while norb * norb * cols >= 2 ** 31 or 8.0 * norb * norb * cols > 8e9:
        cols -= cores
    return max(cols, 1)


def jacobian_columns_from_response(w, resp):
    """Jacobian columns assembled from full dD_p by the rule of src/dmet.py:496-520 (fragment blocks,
    column-major per fragment)."""
    if w.rows is not None:
        from oracle import response_oracle as ro
        imp_size = [int(np.sum(np.abs(c))) for c in (w.imp_clust[:1] if w.trans_inv else w.imp_clust)]
        return ro.jacobian_from_response(resp, w.imp_clust, imp_size, True, w.dmet_orbs)
    ra, rb = [], []
    for f in range(len(w.row_start) - 1):
        idx = w.row_idx[w.row_start[f]:w.row_start[f + 1]]
        a, b = np.meshgrid(idx, idx, indexing="ij")
        ra.append(a.reshape(-1, order="F"))
        rb.append(b.reshape(-1, order="F"))
    ra, rb = np.concatenate(ra), np.concatenate(rb)
    return np.stack([resp[k][ra, rb] for k in range(resp.shape[0])], axis=1)


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
    ncols = args.cpu_sample_cols or default_sample_cols(w.norb)
    cols = spread_columns(w.nterms, ncols)
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
    tot_cols = len(cols) * len(rates)
    tot_t = sum(r["response_s"] for r in rates)
    value = tot_cols / tot_t
    info = dict(info, value=value)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 0,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / len(rates),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc, "sample_cols_per_step": int(len(cols))},
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print_json(line)


# ------------------------------------------------------------------------------------------
# this framework
# ------------------------------------------------------------------------------------------
class _DevArray:
    """Raw device memory as a __cuda_array_interface__ object (for torch.as_tensor: NCCL collectives straight on
    the library's buffers)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class SweepInts:
    """Duck type of the reference's localintegrals object for the response path: what
    qcdmethelper / dmet read from it when doSCF is False."""

    def __init__(self, w):
        self.Norbs, self.Nelec, self.ERIinMEM, self.TI_OK = w.norb, 2 * w.num_pairs, False, True
        self._fock = w.h0

    def loc_rhf_fock(self):
        return self._fock

    def loc_oei(self):
        return self._fock


def time_python_api(w, nst, reuse_output):
    """End to end through the reference's Python call: dmet.rdm_differences_derivative(x) -> (Q, P) numpy array.
    Includes flat2square, Fock + umat on the host, H2D, syevd, kernels, D2H.  x changes every call, so the
    eigensolve is never served from the helper's cache."""
    from qcdmet_b200 import dmet as dmet_mod
    d = dmet_mod.dmet(SweepInts(w), w.imp_clust, w.trans_inv, method='ED', SCmethod='LSTSQ',
                      fitImpBath=w.fit_imp_bath)
    if w.fit_imp_bath:
        d.dmetOrbs = w.dmet_orbs
    d.helper.reuse_output = reuse_output
    p = d.helper.Nterms
    if p != w.nterms:
        return None                       # truncated perturbation list: not reachable through the class API
    x = np.zeros(p)
    times = []
    for k in range(nst + 1):
        x[0] = 1e-9 * (k + 1)
        t0 = time.perf_counter()
        jac = d.rdm_differences_derivative(x)
        times.append(time.perf_counter() - t0)
        assert jac.shape == (w.nresid, p)
    for plan in d.helper._plans.values():
        plan.close()
    return float(np.mean(times[1:]))


def run_b200(args):
    import torch
    import torch.distributed as dist
    from qcdmet_b200 import _lib, sharding, workloads  # noqa: F401

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        gloo = dist.new_group(backend="gloo")       # host-side barriers that do not occupy the GPUs
        assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    inlib = args.gpus if world == 1 else 1           # devices spanned by ONE plan inside this process
    ngpu = max(world, inlib)

    log("start")
    bcast = world > 1 and args.workload == "sweep"
    w, desc = make_workload(args, placeholder_h0=(bcast and rank != 0))
    P, Q = w.nterms, w.nresid
    log(f"workload built: {desc}")
    # column shard, aligned to whole fragments (10 parameters each) so the ranks do equal work
    shards = sharding.column_shards(P, world, 10)
    lo, hi = shards[rank]
    per = shards[0][1] - shards[0][0]
    equal_shards = all(e_ - b_ == per for b_, e_ in shards)     # only the NCCL all-gather fallback needs equal slices

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
    kernel_id = {"auto": _lib.KERNEL_AUTO, "general": _lib.KERNEL_GENERAL, "dmma": _lib.KERNEL_DMMA,
                 "factorised": _lib.KERNEL_FACTORISED}[args.kernel]
    plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx,
                     device=local, devices=(inlib if inlib > 1 else None))
    plan.set_kernel(kernel_id)
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
            assert equal_shards, "the NCCL all-gather fallback needs P divisible into equal whole-fragment slices"
            plan.set_columns(lo, hi)
            # in-place all-gather: each rank's kernels write straight into its slot of the gathered buffer
            gathered = torch.empty(world * per * Q, dtype=torch.float64, device="cuda")
            mine = gathered[rank * per * Q:(rank + 1) * per * Q]
            plan.set_output_device(mine.data_ptr(), mine.numel())

    def finish_step():
        if fused:
            dist.all_reduce(sync_t)          # cross-rank barrier on the stream: all peer stores have landed
        elif world > 1:
            dist.all_gather_into_tensor(gathered, mine)

    def step():
        plan.flush_l2()                     # evict L2 between steps (256 MB > 126 MB L2), on the plan's stream(s)
        plan.compute()
        finish_step()

    log(f"plan ready (fused={fused}, devices in this process={plan.num_devices})")
    for _ in range(args.warmup):
        step()
    plan.sync()
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
    ev1.record()            # single-process multi-GPU: the first device's stream has waited for every device
    torch.cuda.synchronize()
    plan.sync()
    if world > 1:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    log(f"timed region done: {ms_total / args.steps:.2f} ms/step")
    clocks = sampler.stop() if rank == 0 else None
    # dominant kernel: events recorded by the library on the launching stream around that launch
    for _ in range(max(1, min(args.steps, 3))):
        plan.flush_l2()
        plan.compute()
        finish_step()
        plan.sync()
        kernel_ms.append(plan.kernel_time_ms())
    torch.cuda.synchronize()
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

    # every rank / device must hold a bit-identical Jacobian after the fused all-gather
    ranks_identical = None
    ck = plan.checksum()
    if world > 1 and fused:
        cks = [None] * world
        dist.all_gather_object(cks, ck[:3], group=gloo)
        ranks_identical = all(np.array(x).tobytes() == np.array(cks[0]).tobytes() for x in cks)
    elif plan.num_devices > 1:
        ranks_identical = ck[3]

    # in-run FP64 ceiling: the same process, right after the timed region
    peak_inrun = _lib.measure_fp64_peak(local, 0.5) if rank == 0 else None

    # parity gate + CPU baseline: the compiled reference (oracle/_ref) on columns spread over [0, P)
    parity = cpu = None
    if rank == 0 and not (args.no_cpu_baseline and args.no_parity):
        cols = spread_columns(P, args.cpu_sample_cols or default_sample_cols(w.norb))
        cols = cols[(cols >= lo) & (cols < hi)]
        cpu = cpu_reference_rate(w, cols, keep_response=not args.no_parity)
        if cpu is not None and not args.no_parity:
            resp = cpu.pop("response")
            jref = jacobian_columns_from_response(w, resp)
            del resp
            err = 0.0
            for k, cidx in enumerate(cols):
                got = plan.jacobian_columns(int(cidx - lo), int(cidx - lo) + 1)
                err = max(err, float(np.abs(got[:, 0] - jref[:, k]).max()))
            parity = {"cols": int(len(cols)), "first_col": int(cols[0]), "last_col": int(cols[-1]),
                      "max_abs_err": err, "tolerance": 1e-10, "ok": bool(err <= 1e-10),
                      "checker": "compiled reference (oracle/_ref, unmodified lib/rhf_response.cc, own dsyev), "
                                 "Jacobian assembled by the rule of src/dmet.py:496-520"}
            log(f"parity vs compiled reference on {len(cols)} spread columns: max|dJ| = {err:.3e}")
        if args.no_cpu_baseline:
            cpu = None
    if world > 1:
        dist.barrier(group=gloo)

    # ---------------------------------------------------------------- end to end through the host API
    e2e = None
    breakdown = None
    if not args.no_e2e:
        n = w.norb
        h0_np = _lib.pinned_empty((n, n), order="C")
        h0_np[...] = w.h0
        shm = None
        c0, c1 = 0, hi - lo
        use_shm = False
        if world > 1 and fused:
            # needs 8 Q P bytes of /dev/shm (tmpfs); otherwise rank 0 copies the whole Jacobian itself
            flag = [False]
            if rank == 0:
                try:
                    st = os.statvfs("/dev/shm")
                    flag = [bool(st.f_bavail * st.f_frsize > 1.1 * 8 * Q * P) and not os.environ.get("QCDMET_BENCH_NO_SHM")]
                except OSError:
                    flag = [False]
            dist.broadcast_object_list(flag, 0, group=gloo)
            use_shm = bool(flag[0])
        if use_shm:
            # ONE shared page-locked host buffer for the whole Jacobian: rank r ships columns [P r / N, P (r+1) / N)
            # over its own PCIe link; rank 0 (where the least-squares update lives) reads the result in place
            from multiprocessing import shared_memory
            names = [None]
            if rank == 0:
                shm = shared_memory.SharedMemory(create=True, size=8 * Q * P)
                names = [shm.name]
            dist.broadcast_object_list(names, 0, group=gloo)
            if rank != 0:
                shm = shared_memory.SharedMemory(name=names[0])
            out_full = np.ndarray((Q, P), dtype=np.float64, buffer=shm.buf, order="F")
            torch.cuda.check_error(torch.cuda.cudart().cudaHostRegister(out_full.ctypes.data, 8 * Q * P, 0))
            c0, c1 = P * rank // world, P * (rank + 1) // world
            out_np = out_full[:, c0:c1]
        elif world > 1 and fused:
            if rank != 0:
                c1 = 0                           # no shared buffer: rank 0 ships the whole Jacobian over one link
            out_np = _lib.pinned_empty((Q, c1 - c0), order="F")
        else:
            out_np = _lib.pinned_empty((Q, hi - lo), order="F")
        if world > 1 and fused:
            cptr, eptr = plan.orbitals_device()
            c_dev = torch.as_tensor(_DevArray(cptr, n * n), device="cuda")
            e_dev = torch.as_tensor(_DevArray(eptr, n), device="cuda")

        def e2e_step():
            if world > 1 and fused:
                if rank == 0:
                    plan.set_h0(h0_np)          # H2D + ONE cuSOLVER syevd + rank 0's panels
                dist.broadcast(c_dev, 0)        # bit-identical orbitals on every rank (128 MB over NVLink)
                dist.broadcast(e_dev, 0)
                torch.cuda.synchronize()
                if rank != 0:
                    plan.commit_orbitals()      # panels from the received orbitals
            else:
                plan.set_h0(h0_np)              # H2D + cuSOLVER syevd + panels (all devices of the plan)
            plan.compute()
            finish_step()
            if c1 > c0:
                plan.jacobian_columns(c0, c1, out=out_np)  # D2H (blocking; every rank / device its own columns)
            if world > 1:
                dist.barrier(group=gloo)        # the consumer has the whole Jacobian; nobody overwrites it earlier

        for _ in range(min(args.warmup, 2)):
            e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=gloo)
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
        breakdown = plan.timings() if rank == 0 else None
        d2h = int(8 * Q * P) if (world > 1 and fused) else int(8 * Q * (c1 - c0))
        e2e = {"value": P / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(8 * n * n),
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e,
               "breakdown_ms": breakdown,
               "includes": "H2D of H0 from page-locked host memory, ONE cuSOLVER syevd (first rank / device; (C, e) "
                           "sent to the others over NVLink), panel build, Jacobian kernels, D2H of the whole Jacobian "
                           "into page-locked host memory (every rank / device ships its 1/N of the columns over its "
                           "own PCIe link into one buffer readable by rank 0), barrier"}
        e2e["shared_host_buffer"] = use_shm
        if use_shm:
            # the shared buffer must hold the same Jacobian as the device: spot-check on rank 0
            if rank == 0:
                chk = plan.jacobian_columns(P - 3, P)
                e2e["shared_buffer_ok"] = bool(np.array_equal(chk, out_full[:, P - 3:P]))
            dist.barrier(group=gloo)
            torch.cuda.cudart().cudaHostUnregister(out_full.ctypes.data)
            del out_np, out_full
            shm.close()
            if rank == 0:
                shm.unlink()
        # the reference's own Python call (plain numpy arrays in and out), single process
        if world == 1 and rank == 0 and args.workload in ("sweep", "hubbard24", "hubbard24_bath", "polyyne", "sn2"):
            if inlib > 1:
                _lib.set_devices(inlib)
            try:
                t_np = time_python_api(w, nst, reuse_output=False)
                t_pin = time_python_api(w, nst, reuse_output=True)
            finally:
                if inlib > 1:
                    _lib.set_devices(1)
            if t_np is not None:
                e2e["python_api"] = {
                    "call": "qcdmet_b200.dmet.rdm_differences_derivative(x) -> (Q, P) numpy array",
                    "fresh_numpy_output": {"ms_per_step": 1e3 * t_np, "value": P / t_np},
                    "reused_pinned_output": {"ms_per_step": 1e3 * t_pin, "value": P / t_pin},
                    "includes": "flat2square, Fock + umat on the host, H2D from pageable memory, syevd, panels, "
                                "kernels, D2H (fresh_numpy_output: into a new pageable array per call, as the "
                                "reference returns; reused_pinned_output: helper.reuse_output = True)"}

    # ---------------------------------------------------------------- f-3: one device-side LM step
    lm = None
    if args.lm and rank == 0 and world == 1:
        rng = np.random.default_rng(5)
        r = rng.standard_normal(Q) * 1e-3
        plan.compute()
        plan.sync()
        t0 = time.perf_counter()
        plan.lm_step(r, 1e-3)
        t_first = time.perf_counter() - t0
        t0 = time.perf_counter()
        plan.lm_step(r, 1e-2)
        t_retry = time.perf_counter() - t0
        lm = {"normal_matrix_plus_cholesky_ms": 1e3 * t_first, "retry_other_lambda_ms": 1e3 * t_retry,
              "devices": plan.num_devices,
              "note": "qcdmet_plan_lm_step on the device-resident Jacobian: J^T J (DGEMM block rows, one per device) "
                      "+ J^T r + equilibrated Cholesky; a retry with another lambda reuses J^T J"}

    # ---------------------------------------------------------------- torchrun: the same job through ONE process
    single = None
    if world > 1 and not args.no_single_process and args.workload == "sweep":
        dist.barrier(group=gloo)                 # the other ranks idle on the host while rank 0 drives all GPUs
        if rank == 0:
            try:
                sp = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx,
                               device=0, devices=world)
                sp.set_kernel(kernel_id)
                out_sp = _lib.pinned_empty((Q, P), order="F")
                h0_sp = _lib.pinned_empty((w.norb, w.norb), order="C")
                h0_sp[...] = w.h0
                sp.set_orbitals(c, e)
                for _ in range(2):
                    sp.flush_l2(); sp.compute()
                sp.sync()
                nst = max(1, min(args.steps, 3))
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.device(0):
                    a0.record()
                    for _ in range(nst):
                        sp.flush_l2(); sp.compute()
                    a1.record()
                    sp.sync()
                    torch.cuda.synchronize()
                    ms_sp = a0.elapsed_time(a1) / nst
                sp.set_h0(h0_sp); sp.compute(); sp.jacobian(out=out_sp)
                t0 = time.perf_counter()
                for _ in range(nst):
                    sp.set_h0(h0_sp); sp.compute(); sp.jacobian(out=out_sp)
                t_sp = (time.perf_counter() - t0) / nst
                ck_sp = sp.checksum()
                single = {"devices": sp.num_devices, "value": P / (ms_sp * 1e-3), "ms_per_step": ms_sp,
                          "e2e": {"value": P / t_sp, "ms_per_step": 1e3 * t_sp, "breakdown_ms": sp.timings()},
                          "devices_identical": ck_sp[3],
                          "note": "one process, one qcdmet_plan spanning all GPUs (qcdmet_plan_set_devices): the path "
                                  "behind qcdmet_jacobian / dmet.rdm_differences_derivative with QCDMET_B200_DEVICES=N; "
                                  "timed on rank 0 after the torchrun measurement while the other ranks idle"}
                sp.close()
            except _lib.QcdmetError as exc:
                single = {"unavailable": str(exc)}
        dist.barrier(group=gloo)

    if rank == 0:
        peak_committed = committed_fp64_peak()
        peak = peak_inrun or peak_committed or 37.2
        achieved = flops / (k_ms * 1e-3) * 1e-12 if k_ms > 0 else None
        # SURVEY.md 8(d) per-perturbation count W_col = 2 S nv no + (2 nnz+1) nv no + 2 no sum r_f^2 + 2 sum r_f^2
        nv = w.norb - w.num_pairs
        nnz = np.diff(w.h1start[lo:hi + 1]).astype(float)
        w_cols = float(np.sum(2.0 * w.nrows * nv * w.num_pairs + (2.0 * nnz + 1.0) * nv * w.num_pairs
                              + 2.0 * w.num_pairs * Q + 2.0 * Q))
        if fused or plan.num_devices > 1:
            w_cols /= ngpu       # this rank / device executed 1/N of the block-pair list
        traffic = alg_bytes = traffic_src = None
        try:
            key = (f"sweep_N{w.norb}_P{P}_g{ngpu}" if args.workload == "sweep" else f"{args.workload}_g{ngpu}")
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(key)
            if tr:
                traffic, alg_bytes = tr["traffic_bytes"], tr["algorithmic_bytes"]
                traffic_src = tr["source"] + " (ncu --set full capture, not measured in this run)"
        except (OSError, ValueError, KeyError):
            pass
        if world > 1:
            shard_desc = (f"{world} ranks (torchrun, one process per GPU), block-pair work list split evenly, fused all-gather: "
                          "each Jacobian element stored into every rank's buffer by NVLink peer stores, stream barrier "
                          "(1-element NCCL all-reduce) inside the timed region" if fused else
                          f"{world} x {per} columns (whole fragments), NCCL all-gather of the slices inside the timed region")
        elif plan.num_devices > 1:
            shard_desc = (f"ONE process, one plan spanning {plan.num_devices} GPUs (qcdmet_plan_set_devices): block-pair work "
                          "list split evenly, every Jacobian element stored into every device's buffer by NVLink peer "
                          "stores, event barrier between the devices' streams inside the timed region")
        else:
            shard_desc = "single GPU"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ngpu, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc,
                       "kernel": {1: "general", 2: "dmma", 4: "factorised"}.get(kernel_used, str(kernel_used)),
                       "sharding": shard_desc,
                       "l2": "256 MB buffer overwritten between steps (L2 flush), inside the timed region"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "algorithmic_bytes": alg_bytes,
                         "kernel_ms": k_ms, "flops_per_launch": flops,
                         "peak_source": ("measured in this run, after the timed region, by qcdmet_measure_fp64_peak "
                                         "(DMMA.8x8x4 issue rate, 0.5 s)" if peak_inrun else
                                         "profiles/fp64_peaks.json (round 1, tools/fp64_peaks.cu)"),
                         "peak_committed": peak_committed,
                         "algorithmic_equiv_tflops": (w_cols / (k_ms * 1e-3) * 1e-12) if k_ms > 0 else None,
                         "note": "achieved = FP64 flops actually executed by the dominant kernel on this rank / its "
                                 "CUDA-event time (site-factorised algorithm, SURVEY 8d: executed flops disclosed); "
                                 "algorithmic_equiv = the per-perturbation count W_col x columns / the same time"},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches * args.steps),
            "clocks": clocks,
            "parity": parity,
            "ranks_identical": ranks_identical,
            "jacobian_checksum": {"sum": ck[0], "sumsq": ck[1], "xor_bits": f"{ck[2]:016x}"},
        }
        if single is not None:
            line["single_process_api"] = single
        if lm is not None:
            line["lm_step"] = lm
        if plan.info:
            line["config"]["note"] = plan.info
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

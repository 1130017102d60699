"""Debug probe (multi-rank): where does the per-step time go when compute and NCCL all-gather alternate?"""
import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from qcdmet_b200 import _lib, sharding, workloads
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = workloads.sweep(int(sys.argv[1]) if len(sys.argv) > 1 else 2000)
P, Q = w.nterms, w.nresid
shards = sharding.column_shards(P, world, 10); lo, hi = shards[rank]; per = hi - lo
e, c = np.linalg.eigh(w.h0)
plan = _lib.Plan(w.norb, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx, device=local)
plan.set_columns(lo, hi); plan.set_orbitals(c, e)
gathered = torch.empty(world * per * Q, dtype=torch.float64, device="cuda")
mine = gathered[rank * per * Q:(rank + 1) * per * Q]
plan.set_output_device(mine.data_ptr(), mine.numel())
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
def timeit(name, fn, n=3):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"[rank {rank}] {name}: {e0.elapsed_time(e1)/n:.2f} ms/step (events), {(t1-t0)*1e3/n:.2f} ms (wall)", flush=True)
    dist.barrier()
timeit("compute", lambda: plan.compute())
timeit("compute+gather", lambda: (plan.compute(), dist.all_gather_into_tensor(gathered, mine)))
timeit("zero+compute+gather", lambda: (flush.zero_(), plan.compute(), dist.all_gather_into_tensor(gathered, mine)))
timeit("zero+compute", lambda: (flush.zero_(), plan.compute()))
timeit("gather", lambda: dist.all_gather_into_tensor(gathered, mine))
proc = None
if rank == 0:
    proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader", "-lms", "100", "-i", "0"], stdout=subprocess.DEVNULL)
    time.sleep(0.5)
dist.barrier()
timeit("with nvidia-smi polling: zero+compute+gather", lambda: (flush.zero_(), plan.compute(), dist.all_gather_into_tensor(gathered, mine)))
timeit("with nvidia-smi polling: compute", lambda: plan.compute())
if proc: proc.terminate()
dist.destroy_process_group()

"""Aggregate host<->device DMA ceiling of the node: every rank copies 50 MB pinned buffers to / from its GPU in a bare
cudaMemcpyAsync loop (torch tensors over pf_host_alloc memory), all ranks at once.  This is the ceiling the end-to-end
float64 path of bench.py runs into at N = 8 (182 MB of DMA per pair).
usage: python -m torch.distributed.run --nproc-per-node N tools/pcie_ceiling.py   (PF_HOST_ALLOC_WC=1: write-combined)"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from papteam_opticalflow_b200 import _lib

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = _lib.lib()
NB = 50 * 1024 * 1024
def pinned():
    p = lib.pf_host_alloc(NB)
    a = np.frombuffer((C.c_ubyte * NB).from_address(p), dtype=np.uint8)
    a[:] = 1
    return torch.from_numpy(a)
hs = [pinned() for _ in range(4)]
ds = [torch.empty(NB, dtype=torch.uint8, device="cuda") for _ in range(4)]
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
def run(mode, reps=24):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s_in): ds[i % 2].copy_(hs[i % 2], non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s_out): hs[2 + i % 2].copy_(ds[2 + i % 2], non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1: dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    n = reps * NB * (2 if mode == "both" else 1)
    return world * n / dt.item() / 1e9
for mode in ("h2d", "d2h", "both"):
    run(mode, 4)
    g = run(mode)
    if rank == 0:
        print("%d GPUs, %s, pinned%s: %.1f GB/s aggregate (%.1f per GPU)" % (world, mode, " write-combined" if os.environ.get("PF_HOST_ALLOC_WC") == "1" else "", g, g / world), flush=True)
if world > 1:
    dist.destroy_process_group()

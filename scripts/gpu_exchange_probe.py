"""N ranks: GPU time of the slab halo exchange alone (pack, NCCL send/recv, unpack) and of its parts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from gamecraft_b200 import mesher
rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = mesher.Context(lr, stream.cuda_stream)
world = mesher.World(ctx, 36, 36)
hb = world.halo_bytes()
halo = {k: torch.empty(hb, dtype=torch.uint8, device="cuda") for k in ("su", "sd", "ru", "rd")}
top, bottom = 33, 2
def pack():
    if rank + 1 < ws: world.pack_halo(top, 31, halo["su"].data_ptr())
    if rank > 0: world.pack_halo(bottom, 0, halo["sd"].data_ptr())
def xfer():
    ops = []
    if rank + 1 < ws: ops += [dist.P2POp(dist.isend, halo["su"], rank + 1), dist.P2POp(dist.irecv, halo["ru"], rank + 1)]
    if rank > 0: ops += [dist.P2POp(dist.isend, halo["sd"], rank - 1), dist.P2POp(dist.irecv, halo["rd"], rank - 1)]
    for w in dist.batch_isend_irecv(ops): w.wait()
def unpack():
    if rank + 1 < ws: world.unpack_halo(top + 1, 0, halo["ru"].data_ptr())
    if rank > 0: world.unpack_halo(bottom - 1, 31, halo["rd"].data_ptr())
def timed(f, n=50):
    for _ in range(5): f()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter(); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, (time.perf_counter() - t) / n * 1e6
for name, f in (("pack", pack), ("nccl send/recv", xfer), ("unpack", unpack), ("all three", lambda: (pack(), xfer(), unpack()))):
    g, h = timed(f)
    if rank in (0, 1): print("rank %d  %-16s GPU %.1f us  host %.1f us per call (halo %d KB)" % (rank, name, g, h, hb // 1024), flush=True)
dist.destroy_process_group()

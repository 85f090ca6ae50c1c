"""Where does SlabStencil.vjp spend its time?  (torchrun, N ranks; CUDA events per phase, medians.)"""
import os, sys, time, statistics, json
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
N = 16384
def laplacian(x):
    return x[1, 0] + x[0, -1] - 4 * x[0, 0] + x[0, 1] + x[-1, 0]
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from kernex_b200 import dist as kd
slab = kd.SlabStencil(laplacian, kernel_size=(3, 3), relative=True, rows_per_rank=N, width=N, device=dev)
slab.fill_random(torch.Generator(device=dev).manual_seed(rank))
o = slab.step(); g = torch.randn_like(o)
for _ in range(3): slab.vjp(g)
torch.cuda.synchronize(); dist.barrier()
L = slab.layout
rows = []
for _ in range(10):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    t0 = time.perf_counter()
    ev[0].record()
    dext = torch.empty_like(slab.ext)
    if L.halo_top: dext[:L.halo_top].zero_()
    if L.halo_bot: dext[L.halo_top + L.rows:].zero_()
    ev[1].record()
    for p, pm in slab.maps:
        gp = g[p.out0:p.out1]
        if gp.shape[0] and not p.needs_halo and p.in0 == L.halo_top and p.in1 == L.halo_top + L.rows:
            pm.vjp_input(gp, out=dext[p.in0:p.in1])
    ev[2].record()
    for p, pm in slab.maps:
        gp = g[p.out0:p.out1]
        if gp.shape[0] == 0 or (not p.needs_halo and p.in0 == L.halo_top and p.in1 == L.halo_top + L.rows): continue
        dext[p.in0:p.in1] += pm.vjp_input(gp)
    ev[3].record()
    kd.exchange_halo_grads(dext, L, None)
    ev[4].record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    rows.append([ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(4)] + [(t1 - t0) * 1e6])
med = [statistics.median(r[i] for r in rows) for i in range(5)]
allr = [None] * world
dist.all_gather_object(allr, {"rank": rank, "alloc_zero_us": med[0], "interior_us": med[1], "strips_us": med[2], "exchange_us": med[3], "host_us": med[4],
                              "pieces": [(p.in0, p.in1, p.out0, p.out1, p.needs_halo) for p, _ in slab.maps]})
if rank == 0: print(json.dumps(allr))
dist.barrier(); slab.close(); dist.destroy_process_group()

"""Dev probe (torchrun, one rank per GPU): the aggregate end-to-end rate of dtsdf_voxelize_host over ALL ranks of a
box for several expansion-thread counts per rank and for plain copies - which split of the shared host cores the
ranks should take.  Rank 0 prints one line per setting (time = max over ranks, barrier on both sides).
usage: python -m torch.distributed.run --nproc-per-node N tools/e2e_ranks_probe.py"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
capi = pkg.capi
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
F, R = 8500, 32
headers, depth = pkg.synth.make_uniform_batch(F, seed0=1234 + 7919 * rank, box=(100, 60, 220, 190))
offsets, packed_len, max_px = capi.pack_plan(headers)
h_packed = capi.PinnedBuffer(packed_len, np.float32); h_packed.array[:] = depth.reshape(-1)
h_vol = capi.PinnedBuffer((F, 3, R, R, R), np.float32)
h_maxl = capi.PinnedBuffer(F, np.float32); h_mid = capi.PinnedBuffer((F, 3), np.float32)
vox = pkg.Voxelizer(device=local)
tok = torch.zeros(1, device="cuda")
def barrier():
    if world > 1:
        dist.all_reduce(tok)
    torch.cuda.synchronize()
def run(label):
    def call():
        vox.voxelize_packed(headers, offsets, h_packed.array, out_vol=h_vol.array, out_max_l=h_maxl.array, out_mid_p=h_mid.array)
    call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        call()
    dt = torch.tensor([(time.perf_counter() - t0) / 4], device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    barrier()
    if rank == 0:
        print("%-34s %.1f ms per call  %.0f volumes/s over %d rank(s)" % (label, dt.item() * 1e3, F * world / dt.item(), world), flush=True)
if rank == 0:
    print("host threads", os.cpu_count(), "ranks", world, flush=True)
vox.set_option(capi.OPT_PACKED_D2H, 0); run("plain copies")
vox.set_option(capi.OPT_PACKED_D2H, 1)
for nt in (1, 2, 3, 4, 6, 8):
    vox.set_option(capi.OPT_HOST_THREADS, nt)
    run("packed, %d expansion threads/rank" % nt)
if world > 1:
    dist.destroy_process_group()

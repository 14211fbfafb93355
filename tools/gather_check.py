"""Multi-process check of the gathers of volumes into rank 0 (run under torchrun, one rank per GPU):
dense NCCL gather, peer-store gather (gather.PeerVolumeBuffer) and packed gather
(gather.PackedVolumeGather) all deliver every rank's local volumes, bit for bit, ragged shard sizes included.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/gather_check.py [frames_per_rank]
"""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = graft.load_package()
    gmod = importlib.import_module(pkg.__name__ + ".gather")
    base = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    counts = [base + 17 * r for r in range(world)]                      # ragged shards
    frames = pkg.synth.make_frames(range(5000 + 1000 * rank, 5000 + 1000 * rank + counts[rank]))
    headers, offsets, packed, max_px = pkg.pack_frames(frames)
    d = [torch.from_numpy(a).to(dev) for a in (packed, offsets, headers)]
    local_vol, _, _ = pkg.voxelize_device(d[0], d[1], d[2], max_px)

    g = gmod.PackedVolumeGather(counts, 32, root=0)
    for rep in range(3):                                               # reuse of the shared buffers
        if rank == 0:
            g.mine.zero_()
            pkg.voxelize_device(d[0], d[1], d[2], max_px, out_vol=g.mine)
        else:
            g.pack(local_vol)
        full = g.gather()
    nbytes = g.bytes_on_nvlink()
    # every rank sends its local volumes the plain way for the comparison
    refs = {}
    for r in range(world):
        if rank == 0:
            ref = local_vol if r == 0 else torch.empty((counts[r], 3, 32, 32, 32), dtype=torch.float32, device=dev)
            if r != 0:
                dist.recv(ref, src=r)
            refs[r] = ref
        elif rank == r:
            dist.send(local_vol, dst=0)

    def equal_everywhere(full):
        return all(bool(torch.equal(full[sum(counts[:r]):sum(counts[:r + 1])].view(torch.int32), refs[r].view(torch.int32)))
                   for r in range(world))
    ok = equal_everywhere(full) if rank == 0 else True
    # the same in parts (gather_part / finish_parts): one, three and seven parts, the shared buffers reused
    for parts in (1, 3, 7):
        if rank == 0:
            g.full.zero_()
            pkg.voxelize_device(d[0], d[1], d[2], max_px, out_vol=g.mine)
        for k in range(parts):
            g.gather_part(k, parts, None if rank == 0 else local_vol)
        full = g.finish_parts()
        torch.cuda.synchronize()
        if rank == 0:
            ok = ok and equal_everywhere(full)
    g.close()
    if rank == 0:
        dense = sum(counts[1:]) * 3 * 32 ** 3 * 4
        print("packed gather: %d ranks, %s frames, %d bytes pulled over NVLink (dense: %d, %.1fx), equal: %s"
              % (world, counts, nbytes, dense, dense / max(nbytes, 1), ok), flush=True)
        assert ok
        print("gather_check ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

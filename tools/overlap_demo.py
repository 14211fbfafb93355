"""BASELINE.json configs[4]: on-the-fly D-TSDF feeding a training step, voxelisation overlapped with
the step.  The consumer is a stand-in 3-D CNN (the reference's DenseNet, 3D_CNN/network.py, is out of
scope: it is cuDNN work); what is measured is what the voxeliser adds to the step time.

    python tools/overlap_demo.py [--frames 4096] [--batch 32] [--steps 100]
"""
import argparse, importlib, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=4096)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--augment", action="store_true", help="random stretch + rotation per batch (augmented voxeliser)")
args = ap.parse_args()
pkg = g.load_package()
lmod = importlib.import_module(pkg.__name__ + ".loader")
dev = torch.device("cuda:0")
headers, depth = pkg.synth.make_uniform_batch(args.frames, seed0=11)
offsets, packed_len, max_px = pkg.capi.pack_plan(headers)
gt = np.random.default_rng(1).uniform(-150, 150, size=(args.frames, 63)).astype(np.float32)
res = lmod.ResidentFrames(headers, offsets, depth.reshape(-1), gt, dev, max_px)

net = torch.nn.Sequential(
    torch.nn.Conv3d(3, 16, 3, padding=1), torch.nn.ReLU(), torch.nn.MaxPool3d(2),
    torch.nn.Conv3d(16, 48, 3, padding=1), torch.nn.ReLU(), torch.nn.MaxPool3d(2),
    torch.nn.Conv3d(48, 96, 3, padding=1), torch.nn.ReLU(), torch.nn.MaxPool3d(2),
    torch.nn.Flatten(), torch.nn.Linear(96 * 64, 1024), torch.nn.ReLU(), torch.nn.Linear(1024, 63)).to(dev)
opt = torch.optim.SGD(net.parameters(), lr=1e-3)

def train_step(tsdf, target):
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.mse_loss(net(tsdf), target)
    loss.backward()
    opt.step()
    return loss

def timed(fn, steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(steps)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3

# (a) the step alone, on a batch that is already there
loader = lmod.OnTheFlyLoader(res, batch_size=args.batch, shuffle=True, seed=0, drop_last=True, augment=args.augment,
                             augment_center=(0.0, 0.0, -380.0))
tsdf0, jn0, _, _ = next(iter(loader))
tsdf0, jn0 = tsdf0.clone(), jn0.clone()
def step_only(n):
    for _ in range(n):
        train_step(tsdf0, jn0)
step_only(10)
ms_step = timed(step_only, args.steps)

# (b) voxelise, then step, on one stream (no overlap)
idx_all = torch.randperm(args.frames, device=dev)
def serial(n):
    for k in range(n):
        idx = idx_all[(k * args.batch) % (args.frames - args.batch):][:args.batch]
        v, ml, mp, jn = res.voxelize(idx)
        train_step(v, jn)
serial(10)
ms_serial = timed(serial, args.steps)

# (c) the loader: batch k+1 voxelised on a side stream during step k
def overlapped(n):
    done = 0
    while done < n:
        for tsdf, jn, ml, mp in loader:
            train_step(tsdf, jn)
            done += 1
            if done >= n:
                break
overlapped(10)
ms_overlap = timed(overlapped, args.steps)
print(json.dumps({"batch": args.batch, "augment": bool(args.augment), "ms_step_alone": ms_step, "ms_voxelise_then_step": ms_serial,
                  "ms_overlapped_loader": ms_overlap,
                  "voxeliser_cost_serial_ms": ms_serial - ms_step, "voxeliser_cost_overlapped_ms": ms_overlap - ms_step}))

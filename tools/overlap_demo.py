"""BASELINE.json configs[4]: on-the-fly D-TSDF feeding the training step of 3D_CNN/network.py, voxelisation
overlapped with the step.  The consumer is the reference's DenseNet and optimiser step rebuilt in
tools/consumer_net.py (harness, pinned to the unmodified 3D_CNN/network.py by tests/test_consumer_network.py; cuDNN
work, not product code) or, with --net standin, a three-convolution stand-in; what is measured is what the
voxeliser adds to the step time.

    python tools/overlap_demo.py [--frames 4096] [--batch 16] [--steps 100] [--augment] [--net densenet|standin]
"""
import argparse, importlib, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=4096)
ap.add_argument("--batch", type=int, default=16, help="the reference trains with 16 (3D_CNN/train.py:36)")
ap.add_argument("--net", choices=["densenet", "standin"], default="densenet")
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

if args.net == "densenet":
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import consumer_net
    net = consumer_net.DenseNet().to(dev)
    train_step = consumer_net.make_training_step(net, lr=0.005)
else:
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
print(json.dumps({"net": args.net, "batch": args.batch, "augment": bool(args.augment), "ms_step_alone": ms_step, "ms_voxelise_then_step": ms_serial,
                  "ms_overlapped_loader": ms_overlap,
                  "voxeliser_cost_serial_ms": ms_serial - ms_step, "voxeliser_cost_overlapped_ms": ms_overlap - ms_step}))

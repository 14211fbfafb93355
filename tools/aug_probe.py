"""Dev probe: throughput of the augmented voxeliser (dtsdf_voxelize_aug_batch) next to the plain one."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
amod = importlib.import_module(pkg.__name__ + ".augment")
F = 8500
headers, depth = pkg.synth.make_uniform_batch(F, seed0=3)
offsets, packed_len, max_px = pkg.capi.pack_plan(headers)
dev = torch.device("cuda:0")
args = (torch.from_numpy(depth.reshape(-1)).to(dev), torch.from_numpy(offsets).to(dev), torch.from_numpy(headers).to(dev))
aug = torch.from_numpy(amod.make_augmentations(F, np.random.default_rng(0), centers=np.tile([0, 0, -380.0], (F, 1)))).to(dev)
out = (torch.empty((F, 3, 32, 32, 32), device=dev), torch.empty(F, device=dev), torch.empty((F, 3), device=dev))
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t0 = timeit(lambda: pkg.voxelize_device(*args, max_px, out_vol=out[0], out_max_l=out[1], out_mid_p=out[2]))
t1 = timeit(lambda: amod.voxelize_aug_device(*args, aug, max_px, out_vol=out[0], out_max_l=out[1], out_mid_p=out[2]))
print("plain %.3f ms (%.2f M volumes/s)   augmented %.3f ms (%.2f M volumes/s)" % (t0, F / t0 / 1e3, t1, F / t1 / 1e3))
t2 = timeit(lambda: amod.voxelize_aug_device(*args, aug, 16, out_vol=out[0], out_max_l=out[1], out_mid_p=out[2]))
print("augmented, crop left in global memory (staging hint 16 px: more CTAs per SM, lookups through L1): %.3f ms" % t2)

"""Dev probe: DataProcess.process() semantics on the device - the sampled voxeliser (one launch, no clouds) next
to the plain pixel-bbox launch over the same 8,500 frames.  Under ncu: -k regex:dtsdf_fused -s 8 -c 1."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
if os.environ.get("PROBE_LIB"):                      # an experimental build (tools/exp_bench.py build NAME ...)
    pkg.capi.LIB_PATH = os.environ["PROBE_LIB"]
F, P = 8500, 6000
headers, depth = pkg.synth.make_uniform_batch(F, seed0=3)
offsets, packed_len, max_px = pkg.capi.pack_plan(headers)
dev = torch.device("cuda:0")
args = (torch.from_numpy(depth.reshape(-1)).to(dev), torch.from_numpy(offsets).to(dev), torch.from_numpy(headers).to(dev))
n_valid = pkg.count_valid_device(*args).cpu().numpy()
np.random.seed(1)
idx = torch.from_numpy(pkg.capi.resample_indices(n_valid, P)).to(dev)
out = (torch.empty((F, 3, 32, 32, 32), device=dev), torch.empty(F, device=dev), torch.empty((F, 3), device=dev))
def timeit(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t0 = timeit(lambda: pkg.voxelize_device(*args, max_px, out_vol=out[0], out_max_l=out[1], out_mid_p=out[2]))
t1 = timeit(lambda: pkg.voxelize_sampled_device(*args, idx, max_px, out_vol=out[0], out_max_l=out[1], out_mid_p=out[2]))
print("pixel bbox %.3f ms (%.2f M volumes/s)   sampled bbox %.3f ms (%.2f M volumes/s)   mean n_valid %.0f" %
      (t0, F / t0 / 1e3, t1, F / t1 / 1e3, n_valid.mean()))

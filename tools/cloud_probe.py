"""Dev probe: the three kernels of process() with clouds (point_cloud_kernel, resample_extrema_kernel, fused
voxeliser in extrema mode) timed one by one over 8,500 resident frames."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
capi = pkg.capi
lib = capi.load_library()
F, P = 8500, 6000
headers, depth = pkg.synth.make_uniform_batch(F, seed0=3)
offsets, packed_len, max_px = capi.pack_plan(headers)
dev = torch.device("cuda:0")
d_packed, d_off, d_hdr = (torch.from_numpy(depth.reshape(-1)).to(dev), torch.from_numpy(offsets).to(dev), torch.from_numpy(headers).to(dev))
n_px = int(headers[0, 4] - headers[0, 2]) * int(headers[0, 5] - headers[0, 3])
d_points = torch.empty((F * n_px, 3), dtype=torch.float64, device=dev)
d_poff = (torch.arange(F, dtype=torch.int64, device=dev) * n_px).contiguous()
d_nvalid = torch.empty((F,), dtype=torch.int32, device=dev)
d_ext = torch.empty((F, 6), dtype=torch.float32, device=dev)
d_pc = torch.empty((F, P, 3), dtype=torch.float64, device=dev)
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def pc():
    capi.check(lib.dtsdf_point_cloud_batch(F, ctypes.c_double(241.42), d_packed.data_ptr(), d_off.data_ptr(), d_hdr.data_ptr(),
                                           d_points.data_ptr(), d_poff.data_ptr(), d_nvalid.data_ptr(), stream))
pc()
n_valid = d_nvalid.cpu().numpy()
np.random.seed(1)
d_idx = torch.from_numpy(capi.resample_indices(n_valid, P)).to(dev)
def rs():
    capi.check(lib.dtsdf_resample_extrema_batch(F, P, d_points.data_ptr(), d_poff.data_ptr(), d_idx.data_ptr(), d_pc.data_ptr(),
                                                d_ext.data_ptr(), stream))
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def sc():
    capi.check(lib.dtsdf_sample_cloud_batch(F, ctypes.c_double(241.42), P, d_packed.data_ptr(), d_off.data_ptr(), d_hdr.data_ptr(),
                                            d_idx.data_ptr(), int(max_px), d_pc.data_ptr(), d_ext.data_ptr(), None, stream))
t_sc = timeit(sc)
print("sample_cloud_kernel %.3f ms (%.2f TB/s on crop + index + sampled cloud)" % (t_sc, F * (n_px * 4 + P * 28 + 24) / t_sc / 1e9))
t_pc, t_rs = timeit(pc), timeit(rs)
nv = float(n_valid.mean())
b_pc = F * (n_px * 4 + nv * 24)
b_rs = F * (P * (24 + 4) + P * 24 + 24)
print("point_cloud_kernel %.3f ms (%.0f valid px/frame, %.2f TB/s)   resample_extrema_kernel %.3f ms (%.2f TB/s)" %
      (t_pc, nv, b_pc / t_pc / 1e9, t_rs, b_rs / t_rs / 1e9))

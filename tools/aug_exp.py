"""Dev tool: time build variants of the augmented voxeliser and compare their output bit for bit with the main build.

    python tools/exp_bench.py build NAME [-DFLAG ...]     # here (no GPU): tools/_exp/NAME.so
    python tools/aug_exp.py NAME [NAME ...]               # on the GPU box ('main' = the in-tree library, always first)
"""
import ctypes, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
amod = importlib.import_module(pkg.__name__ + ".augment")
F = int(os.environ.get("EXP_FRAMES", 8500))
headers, depth = pkg.synth.make_uniform_batch(F, seed0=3)
offsets, packed_len, max_px = pkg.capi.pack_plan(headers)
dev = torch.device("cuda:0")
d_packed, d_off, d_hdr = torch.from_numpy(depth.reshape(-1)).to(dev), torch.from_numpy(offsets).to(dev), torch.from_numpy(headers).to(dev)
aug = torch.from_numpy(amod.make_augmentations(F, np.random.default_rng(0), centers=np.tile([0, 0, -380.0], (F, 1)))).to(dev)
cfg = pkg.capi.default_config(32, 0, 0)
vp = ctypes.c_void_p
ref = None
for name in ["main"] + [n for n in sys.argv[1:] if n != "main"]:
    path = pkg.capi.LIB_PATH if name == "main" else os.path.join(ROOT, "tools", "_exp", name + ".so")
    fn = ctypes.CDLL(path).dtsdf_voxelize_aug_batch
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.POINTER(pkg.capi.Config), ctypes.c_int32, vp, vp, vp, vp, ctypes.c_int32, vp, ctypes.c_int32, vp, vp, vp, vp, vp]
    out = (torch.empty((F, 3, 32, 32, 32), device=dev), torch.empty(F, device=dev), torch.empty((F, 3), device=dev))
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        rc = fn(ctypes.byref(cfg), F, d_packed.data_ptr(), d_off.data_ptr(), d_hdr.data_ptr(), aug.data_ptr(), max_px, None, 0,
                out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), None, stream)
        assert rc == 0, rc
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    same = ""
    if ref is None:
        ref = out
    else:
        same = "   bit-identical to main: %s" % all(torch.equal(a.view(torch.int32), b.view(torch.int32)) for a, b in zip(ref, out))
    print("%-24s %.4f ms  %.2f M volumes/s%s" % (name, ms, F / ms / 1e3, same), flush=True)

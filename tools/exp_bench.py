"""Dev tool: build an experimental variant of the library (-D flags) and time the fused kernel.

    python tools/exp_bench.py build NAME [-DFLAG ...]     # here (no GPU): tools/_exp/NAME.so
    python tools/exp_bench.py run NAME [NAME ...]         # on the GPU box: ms per launch
"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
EXP = os.path.join(ROOT, "tools", "_exp")


def build(name, flags):
    import __graft_entry__ as g
    os.makedirs(EXP, exist_ok=True)
    srcs = [os.path.join(g.CSRC, n) for n in g.CUDA_SOURCES]
    cmd = [g._find("nvcc")] + g.NVCC_FLAGS + flags + srcs + ["-o", os.path.join(EXP, name + ".so")]
    subprocess.run(cmd, check=True, cwd=ROOT)


def run(names, frames=8500, steps=20):
    import numpy as np, torch
    R = int(os.environ.get("EXP_RES", "32"))
    frames = int(os.environ.get("EXP_FRAMES", frames))
    import __graft_entry__ as g
    pkg = g.load_package()
    box = tuple(int(v) for v in os.environ.get("EXP_BOX", "100,60,220,190").split(","))
    headers, depth = pkg.synth.make_uniform_batch(frames, seed0=1234, box=box)
    bw, bh = box[2] - box[0], box[3] - box[1]
    alg = frames * (24 + 4 * bw * bh + 12 * R ** 3 + 16)
    offsets, packed_len, max_px = pkg.capi.pack_plan(headers)
    dev = torch.device("cuda:0")
    d_packed = torch.from_numpy(depth.reshape(-1)).to(dev)
    d_off = torch.from_numpy(offsets).to(dev)
    d_hdr = torch.from_numpy(headers).to(dev)
    d_vol = torch.empty((frames, 3, R, R, R), dtype=torch.float32, device=dev)
    d_ml = torch.empty((frames,), dtype=torch.float32, device=dev)
    d_mp = torch.empty((frames, 3), dtype=torch.float32, device=dev)
    cfg = pkg.capi.default_config(R, 0, 0)
    max_px = int(os.environ.get('EXP_MAXPX', max_px))
    for name in names:
        path = os.path.join(EXP, name + ".so") if name != "main" else pkg.capi.LIB_PATH
        lib = ctypes.CDLL(path)
        fn = lib.dtsdf_voxelize_batch
        fn.restype = ctypes.c_int
        vp = ctypes.c_void_p
        fn.argtypes = [ctypes.POINTER(pkg.capi.Config), ctypes.c_int32, vp, vp, vp, vp, ctypes.c_int32, vp, vp, vp, vp]
        stream = torch.cuda.current_stream().cuda_stream

        def step():
            rc = fn(ctypes.byref(cfg), frames, d_packed.data_ptr(), d_off.data_ptr(), d_hdr.data_ptr(), None, max_px,
                    d_vol.data_ptr(), d_ml.data_ptr(), d_mp.data_ptr(), stream)
            assert rc == 0, rc
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print("R=%d %-28s %.4f ms/launch  %.2f M volumes/s  %.0f GB/s algorithmic (crop %dx%d)" % (R, name, ms, frames / ms / 1e3, alg / ms / 1e6, bw, bh), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2], sys.argv[3:])
    else:
        run(sys.argv[2:])

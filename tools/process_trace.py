"""Dev probe: the process() host call (dtsdf_process_host, volumes only, pinned buffers) with DTSDF_TRACE=1 -
where the submitting thread's time goes.  usage: python tools/process_trace.py [frames]"""
import os, sys, time
import numpy as np
os.environ["DTSDF_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
capi = pkg.capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
frames = pkg.synth.make_frames(range(n), box=pkg.synth.DEFAULT_BOX)
hdr, off, pk, _ = pkg.pack_frames(frames)
pin = capi.PinnedBuffer(pk.shape, np.float32)
pin.array[...] = pk
out = capi.PinnedBuffer((n, 3, 32, 32, 32), np.float32)
vox = pkg.Voxelizer(device=0)
np.random.seed(3)
for rep in range(4):
    t0 = time.perf_counter()
    vox.process_packed(hdr, off, pin.array, 6000, want_points=False, out_vol=out.array)
    dt = time.perf_counter() - t0
    print("call %d: %.2f ms = %.0f frames/s" % (rep, dt * 1e3, n / dt), flush=True)
print("pageable input and output (what a caller with plain NumPy arrays gets):")
out_pg = np.empty((n, 3, 32, 32, 32), dtype=np.float32)
out_pg[...] = 0
for rep in range(3):
    t0 = time.perf_counter()
    vox.process_packed(hdr, off, pk, 6000, want_points=False, out_vol=out_pg)
    dt = time.perf_counter() - t0
    print("call %d: %.2f ms = %.0f frames/s" % (rep, dt * 1e3, n / dt), flush=True)
print("frame records in, NumPy arrays out (Voxelizer.process_frames: packs into its page-locked staging area):")
for rep in range(3):
    t0 = time.perf_counter()
    vox.process_frames(frames, 6000, want_points=False)
    dt = time.perf_counter() - t0
    print("call %d: %.2f ms = %.0f frames/s" % (rep, dt * 1e3, n / dt), flush=True)
nv = np.full(n, 8811, dtype=np.int32)
t0 = time.perf_counter(); capi.resample_indices(nv, 6000); dt = time.perf_counter() - t0
print("index draw alone: %.2f ms" % (dt * 1e3))

"""Dev probe: end-to-end rate of dtsdf_voxelize_host for a few host-thread counts / transfer modes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
capi = pkg.capi
if os.environ.get('DTSDF_LIB'):
    capi.LIB_PATH = os.path.join(ROOT, 'tools', '_exp', os.environ['DTSDF_LIB'] + '.so')
F, R = 8500, 32
headers, depth = pkg.synth.make_uniform_batch(F, seed0=1234, box=(100, 60, 220, 190))
offsets, packed_len, max_px = capi.pack_plan(headers)
h_packed = capi.PinnedBuffer(packed_len, np.float32); h_packed.array[:] = depth.reshape(-1)
h_vol = capi.PinnedBuffer((F, 3, R, R, R), np.float32)
h_maxl = capi.PinnedBuffer(F, np.float32); h_mid = capi.PinnedBuffer((F, 3), np.float32)
vox = pkg.Voxelizer(device=0)
def run(label):
    for _ in range(2):
        vox.voxelize_packed(headers, offsets, h_packed.array, out_vol=h_vol.array, out_max_l=h_maxl.array, out_mid_p=h_mid.array)
    t0 = time.perf_counter()
    for _ in range(5):
        vox.voxelize_packed(headers, offsets, h_packed.array, out_vol=h_vol.array, out_max_l=h_maxl.array, out_mid_p=h_mid.array)
    dt = (time.perf_counter() - t0) / 5
    print("%-24s %.1f ms  %.0f volumes/s" % (label, dt * 1e3, F / dt), flush=True)
print("host threads", os.cpu_count(), flush=True)
vox.set_option(capi.OPT_PACKED_D2H, 0); run("plain")
vox.set_option(capi.OPT_PACKED_D2H, 1)
for nt in (4, 8, 12, 14, 16):
    vox.set_option(capi.OPT_HOST_THREADS, nt)
    run("packed, %d threads" % nt)

for mb in (16, 32, 128):
    vox.set_option(capi.OPT_HOST_THREADS, 0)
    vox.set_option(capi.OPT_CHUNK_MB, mb)
    run("packed, default threads, %d MB chunks" % mb)

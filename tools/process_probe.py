"""Dev probe: throughput of the DataProcess.process() path (pre/process.py:13-28) through the batched host
entry points - point_cloud_kernel (f64 back-projection + compaction), host-drawn resampling indices,
resample_extrema_kernel, fused voxeliser in extrema mode - next to the pixel-bbox path.  Prints JSON."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
n = int(os.environ.get("PROBE_FRAMES", "2000"))
frames = pkg.synth.make_frames(range(n), box=pkg.synth.DEFAULT_BOX)
vox = pkg.Voxelizer(device=0)
res = {"frames": n, "crop": "120x130", "point_num": 6000}
def rate(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return n * reps / (time.perf_counter() - t0)
np.random.seed(0)
res["process_frames_with_point_clouds_per_s"] = rate(lambda: vox.process_frames(frames, 6000, want_points=True))
res["process_frames_volumes_only_per_s"] = rate(lambda: vox.process_frames(frames, 6000, want_points=False))
res["process_frames_two_calls_numpy_draws_volumes_only_per_s"] = rate(
    lambda: vox.process_frames_two_calls(frames, 6000, want_points=False), reps=1)
hdr, off, pk, _ = pkg.pack_frames(frames)
out = pkg.capi.PinnedBuffer((n, 3, 32, 32, 32), np.float32)
res["process_packed_volumes_only_pinned_out_per_s"] = rate(
    lambda: vox.process_packed(hdr, off, pk, 6000, want_points=False, out_vol=out.array))
t0 = time.perf_counter()
pkg.capi.resample_indices(np.full(n, 8811, dtype=np.int32), 6000)
res["index_draw_in_library_only_per_s"] = n / (time.perf_counter() - t0)
res["voxelize_frames_pixel_bbox_per_s"] = rate(lambda: vox.voxelize_frames(frames))
# where the time goes in process_frames: the np.random index draw alone
def draw():
    for _ in range(n):
        np.random.randint(0, 8811, size=6000)
res["index_draw_only_per_s"] = rate(draw)
# the per-frame drop-in calls, as the reference's own frame loop makes them (pre/read_MSRA.py:98-111, pre/time_test.py:27)
sys.path.insert(0, os.path.join(g.PKG_DIR, "pre"))
import process as dp_mod, tsdf_numba as tn_mod
k = 300
recs = [{"header": f["header"], "data": f["depth"]} for f in frames[:k]]
tn_mod.cal_tsdf_cuda(recs[0])
t0 = time.perf_counter()
for r in recs:
    tn_mod.cal_tsdf_cuda(r)
res["cal_tsdf_cuda_single_frame_calls_per_s"] = k / (time.perf_counter() - t0)
gt = np.zeros(63, dtype=np.float32)
dp_mod.DataProcess(frames[0], gt, 6000, aug=False).process()
t0 = time.perf_counter()
for f in frames[:k]:
    dp_mod.DataProcess(f, gt, 6000, aug=False).process()
res["DataProcess_process_single_frame_calls_per_s"] = k / (time.perf_counter() - t0)
print(json.dumps(res))

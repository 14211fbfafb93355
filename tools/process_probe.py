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
res["voxelize_frames_pixel_bbox_per_s"] = rate(lambda: vox.voxelize_frames(frames))
# where the time goes in process_frames: the np.random index draw alone
def draw():
    for _ in range(n):
        np.random.randint(0, 8811, size=6000)
res["index_draw_only_per_s"] = rate(draw)
print(json.dumps(res))

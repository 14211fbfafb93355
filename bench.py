#!/usr/bin/env python
"""Benchmark of the depth-crop -> D-TSDF hot path (BASELINE.json: D-TSDF volumes/s, 32^3 x 3).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU reference arm (oracle port)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W      # N > 1, one rank per GPU

A "step" is one pass of the fused voxeliser over one batch of synthetic MSRA-shaped frames
(BASELINE.json configs[1]: 8,500 frames of 120x130 px per GPU, R = 32).  Frames shard across
ranks with no data-path collective (weak scaling: every rank voxelises its own 8,500 frames).

Printed JSON (one line, rank 0):
  value      volumes/s over all ranks, inputs resident in HBM, K launches timed with CUDA events
             on the launching stream, max over ranks
  e2e        the same metric through the host entry point of the C ABI (dtsdf_voxelize_host):
             pinned HOST buffers in and out, H2D + kernel + D2H inside the timed region
  roofline   HBM roofline of the fused kernel: algorithmic bytes per launch / measured launch time
             against MEASURED_PEAKS.json
  cpu_baseline  the C oracle port of the reference loop on the host cores (N = 1, rank 0 only); beside it
             the reference's own numba.cuda kernels driven per frame on this GPU (oracle/_ref)
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "dtsdf_volumes_per_sec"
UNIT = "volumes/s"
FRAMES_PER_GPU = 8500          # BASELINE.json configs[1]
BOX = (100, 60, 220, 190)      # 120 x 130 crop, BASELINE.md section 3


def algorithmic_bytes(headers, R):
    """24 (header) + 4*bw*bh (crop read once) + 12*R^3 (volume) + 16 (max_l, mid_p)."""
    bw = (headers[:, 4] - headers[:, 2]).astype(np.int64)
    bh = (headers[:, 5] - headers[:, 3]).astype(np.int64)
    return int(np.sum(24 + 4 * bw * bh + 12 * R ** 3 + 16))


def workload_config(args, world):
    return {
        "workload": "%s: %d synthetic MSRA-shaped frames per GPU, crop %dx%d, R=%d -> 3x%d^3 f32"
                    % ("ragged crops (80-150 x 80-140 px; the size below is the default, not used)" if args.ragged else
                       "configs[1]" if (args.frames == FRAMES_PER_GPU and args.res == 32) else
                       ("configs[2] (full MSRA-sized set)" if args.frames * world >= 76000 else
                        ("configs[3] (resolution sweep)" if args.res != 32 else "custom size")),
                       args.frames, BOX[2] - BOX[0], BOX[3] - BOX[1], args.res, args.res),
        "frames_per_gpu": args.frames,
        "frames_total": args.frames * world,
        "voxel_res": args.res,
        "layout": "cxyz" if args.layout == 0 else "czyx",
        "bbox": "valid pixels (tsdf_numba.cal_tsdf_cuda semantics)",
        "sharding": "frames, contiguous per rank, no data-path collective",
        "l2": "inputs+outputs per step (%.2f GB) exceed the 126 MB L2; no explicit flush"
              % ((args.frames * (15600 * 4 + 12 * args.res ** 3)) / 1e9),
    }


class ClockSampler:
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        names = {
            "nvmlClocksEventReasonSwPowerCap": "sw_power_cap",
            "nvmlClocksThrottleReasonSwPowerCap": "sw_power_cap",
            "nvmlClocksThrottleReasonHwSlowdown": "hw_slowdown",
            "nvmlClocksThrottleReasonHwThermalSlowdown": "hw_thermal_slowdown",
            "nvmlClocksThrottleReasonSwThermalSlowdown": "sw_thermal_slowdown",
            "nvmlClocksThrottleReasonHwPowerBrakeSlowdown": "hw_power_brake",
        }
        masks = {getattr(nv, k): v for k, v in names.items() if hasattr(nv, k)}
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for m, name in masks.items():
                    if r & m:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        self._stop = threading.Event()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(res, layout, frames):
    """dram bytes per launch of the fused kernel from the committed ncu capture, if any, scaled from the
    capture's frame count to this run's (the traffic is per frame: fixed-shape crops)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        e = json.load(open(p)).get("R%d_L%d" % (res, layout))
        if not e:
            return None
        return e["dram_bytes_per_launch"] * frames / float(e.get("frames", frames))
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the C oracle port of the reference's loop, all host threads
# ------------------------------------------------------------------------------------------------
def host_threads():
    """Hardware threads this process may run on.  Passed to the oracle explicitly: torchrun exports
    OMP_NUM_THREADS=1 to its ranks, which would leave the reference arm on one core for N > 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_port_rate(headers, offsets, packed, R, layout, seconds, nthreads=0):
    """Times oracle_dtsdf_batch on a bounded sample of about `seconds` of CPU work: whole passes over
    (a prefix of) the batch, repeated.  Returns (volumes/s, threads, frames per pass, passes, seconds)."""
    from oracle import c_oracle
    nthreads = nthreads or host_threads()
    n = headers.shape[0]
    probe = min(n, 64)
    px = int(offsets[probe]) if probe < n else packed.size
    t0 = time.perf_counter()
    _, _, _, used = c_oracle.dtsdf_batch(headers[:probe], offsets[:probe], packed[:px], R=R, layout=layout,
                                         nthreads=nthreads)
    dt = max(time.perf_counter() - t0, 1e-6)
    sample = int(min(n, max(probe, seconds * probe / dt)))
    px = int(offsets[sample]) if sample < n else packed.size
    out = np.empty((sample, 3, R, R, R), dtype=np.float32)
    passes = 0
    t0 = time.perf_counter()
    while True:
        c_oracle.dtsdf_batch(headers[:sample], offsets[:sample], packed[:px], R=R, layout=layout, nthreads=nthreads,
                             out=out)
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= seconds or passes >= 1000:
            break
    return sample * passes / dt, used, sample, passes, dt


def python_loop_rate(headers, offsets, packed, R):
    """The reference's own form of the loop (interpreted Python over 32^3 voxels), one frame."""
    from oracle import dtsdf_oracle as orc
    h = headers[0]
    n = int(h[4] - h[2]) * int(h[5] - h[3])
    depth = packed[int(offsets[0]):int(offsets[0]) + n]
    t0 = time.perf_counter()
    pmin, pmax = orc.bbox_from_pixels(h, depth)
    vox_ori, voxel_len, trunc, _, _ = orc.grid_params(pmax, pmin, R)
    orc.tsdf_cal_loop(h, depth, vox_ori, voxel_len, trunc, R)
    return 1.0 / (time.perf_counter() - t0)


def numba_gpu_rate(headers, offsets, packed, frames=256):
    """The reference's own GPU path on this box: its two numba kernels (compiled from the unmodified
    pre/tsdf_numba.py into oracle/_ref by oracle/make_numba_ptx.py) driven per frame exactly as
    cal_tsdf_cuda does (pre/tsdf_numba.py:119-161: 2 launches, 2 H2D, 2 D2H, a host reduction).
    Only R = 32, layout (c,z,y,x): the kernel is hard-wired to both."""
    try:
        from oracle import numba_ptx_ref
        if not numba_ptx_ref.available():
            return {"unavailable": "oracle/_ref PTX not built"}
        import torch
        ref = numba_ptx_ref.NumbaPtxReference(device=torch.cuda.current_device())
        n = min(frames, headers.shape[0])
        recs = []
        for i in range(n):
            h = headers[i]
            px = int(h[4] - h[2]) * int(h[5] - h[3])
            recs.append({"header": h, "data": np.array(packed[int(offsets[i]):int(offsets[i]) + px])})
        for r in recs[:8]:
            ref.cal_tsdf_cuda(r)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for r in recs:
            ref.cal_tsdf_cuda(r)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return {"value": n / dt, "unit": UNIT, "frames": n,
                "how": "reference numba kernels (numba %s PTX of the unmodified source, driver-JITed for this GPU), "
                       "one frame per call through the steps of cal_tsdf_cuda, 1 host thread" % ref.manifest["numba"]}
    except Exception as e:  # a comparator, never fatal
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from __graft_entry__ import build_oracle, load_package
    build_oracle()
    pkg = load_package()
    headers, depth = pkg.synth.make_uniform_batch(min(args.frames, 4096), seed0=1234, box=BOX)    # the first frames of rank 0's batch
    offsets = (np.arange(headers.shape[0], dtype=np.int64) * depth.shape[1])
    packed = depth.reshape(-1)
    from oracle import c_oracle
    # size one step for ~0.25 s
    rate, used, _, _, _ = cpu_port_rate(headers, offsets, packed, args.res, args.layout, 0.5)
    per_step = int(max(32, min(headers.shape[0], rate * 0.25)))
    px = int(offsets[per_step]) if per_step < headers.shape[0] else packed.size
    out = np.empty((per_step, 3, args.res, args.res, args.res), dtype=np.float32)

    def step():
        c_oracle.dtsdf_batch(headers[:per_step], offsets[:per_step], packed[:px], R=args.res,
                             layout=args.layout, out=out, nthreads=host_threads())

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": dict(workload_config(args, world), frames_timed_per_step=per_step,
                                            note="the CPU arm times a bounded sample of the workload: %d frames per step "
                                                 "(the first frames of the same batch), not all %d" % (per_step, args.frames)),
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": used, "kind": "port",
            "sample": "%d frames per step x %d steps of the same workload; C restatement of the reference's "
                      "Python voxel loop (oracle/dtsdf_oracle.c, gcc -O2, OpenMP over frames, %d threads); "
                      "the reference itself is interpreted Python and cannot travel to this box"
                      % (per_step, args.steps, used),
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------------------------------
# gather of every rank's volumes into rank 0 (BASELINE.json configs[2]) - three ways
# ------------------------------------------------------------------------------------------------
NVLINK_PEER_GBS = 770.0        # measured peer copy per direction on this pool (B200_PROFILING.md)


def run_gather_legs(pkg, dist, dev, rank, world, F, R, cfg, inputs, d_vol, d_maxl, d_mid, barrier, max_over_ranks):
    """voxelise + gather, timed per step with CUDA events on every rank, max over ranks.
      nccl    kernel into local HBM, then dist.gather of the dense volumes (ncclSend/Recv into rank 0)
      p2p     the voxeliser's own stores land in rank 0's IPC-mapped buffer over NVLink, then a barrier
      packed  every rank packs its volumes (2-bit codes + literal triples, lossless) in its own HBM; rank 0's
              expansion kernel pulls the bytes in use over NVLink with TMA bulk copies while it writes the dense
              volumes (gather.PackedVolumeGather): transfer and expansion are one kernel
    Every leg's result on rank 0 is compared with the dense NCCL gather over ALL frames."""
    import importlib
    import torch
    gmod = importlib.import_module(pkg.__name__ + ".gather")
    d_packed, d_off, d_hdr, max_px = inputs
    vol_bytes = F * 3 * R ** 3 * 4
    dense_bytes = (world - 1) * vol_bytes
    peak, _ = measured_peak()

    def voxelise(out):
        pkg.voxelize_device(d_packed, d_off, d_hdr, max_px, cfg, out_vol=out, out_max_l=d_maxl, out_mid_p=d_mid)

    def timed(fn, reps, warm=2):
        for _ in range(warm):
            fn()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(reps):
            fn()
        g1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(g0.elapsed_time(g1)) / reps
        barrier()
        return ms

    def all_ok(flag):
        t = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() == 1.0)

    res = {"frames_total": F * world, "dense_bytes_to_rank0": int(dense_bytes),
           "timed": "one step = voxelise every rank's %d frames + gather to rank 0; CUDA events, max over ranks" % F}
    # ---- nccl: the dense gather; its result is the reference for the two others
    bufs = [torch.empty_like(d_vol) for _ in range(world)] if rank == 0 else None

    def nccl_step():
        voxelise(d_vol)
        dist.gather(d_vol, bufs, dst=0)
    ms = timed(nccl_step, reps=3, warm=1)
    res["nccl"] = {"value": F * world / (ms * 1e-3), "unit": UNIT, "ms": ms,
                   "rank0_ingress_gbs": dense_bytes / (ms * 1e-3) / 1e9}

    # ---- p2p: fused peer stores
    peer = gmod.PeerVolumeBuffer(F * world, R, root=0)

    def p2p_step():
        voxelise(peer.mine)
    ms = timed(p2p_step, reps=3, warm=1)
    peer.complete()
    same = True
    if rank == 0:
        same = all(bool(torch.equal(peer.full[r * F:(r + 1) * F].view(torch.int32), bufs[r].view(torch.int32)))
                   for r in range(world))
    res["p2p"] = {"value": F * world / (ms * 1e-3), "unit": UNIT, "ms": ms, "matches_nccl_gather_all_frames": all_ok(same),
                  "rank0_ingress_gbs": dense_bytes / (ms * 1e-3) / 1e9}
    peer.close()
    del peer
    torch.cuda.empty_cache()

    # ---- packed: pack on every rank, pull + expand on rank 0
    g = gmod.PackedVolumeGather([F] * world, R, root=0)

    def packed_step():
        if rank == 0:
            voxelise(g.mine)
        else:
            voxelise(d_vol)
            g.pack(d_vol)
        g.gather()
    ms = timed(packed_step, reps=5, warm=2)
    nvlink_bytes = g.bytes_on_nvlink()
    same = True
    if rank == 0:
        same = all(bool(torch.equal(g.full[r * F:(r + 1) * F].view(torch.int32), bufs[r].view(torch.int32)))
                   for r in range(world))
    res["packed"] = {"value": F * world / (ms * 1e-3), "unit": UNIT, "ms": ms,
                     "matches_nccl_gather_all_frames": all_ok(same), "bytes_pulled_over_nvlink": int(nvlink_bytes),
                     "rank0_ingress_gbs": nvlink_bytes / (ms * 1e-3) / 1e9,
                     "rank0_hbm_write_gbs": world * vol_bytes / (ms * 1e-3) / 1e9}
    # ---- packed, in two parts: the root expands the first half of every rank while the others voxelise and pack
    #      their second half
    PARTS = 2

    def voxelise_rows(out, f0, f1):
        pkg.voxelize_device(d_packed, d_off[f0:f1], d_hdr[f0:f1], max_px, cfg, out_vol=out[f0:f1], out_max_l=d_maxl[f0:f1],
                            out_mid_p=d_mid[f0:f1])

    def parts_step():
        if rank == 0:
            voxelise(g.mine)
            for k in range(PARTS):
                g.gather_part(k, PARTS)
        else:
            for k in range(PARTS):
                f0, f1 = g.part(rank, k, PARTS)
                voxelise_rows(d_vol, f0, f1)
                g.gather_part(k, PARTS, d_vol)
        g.finish_parts()
    if rank == 0:
        g.full.zero_()
    ms = timed(parts_step, reps=5, warm=2)
    same = True
    if rank == 0:
        same = all(bool(torch.equal(g.full[r * F:(r + 1) * F].view(torch.int32), bufs[r].view(torch.int32)))
                   for r in range(world))
    res["packed_in_parts"] = {"value": F * world / (ms * 1e-3), "unit": UNIT, "ms": ms, "parts": PARTS,
                              "matches_nccl_gather_all_frames": all_ok(same),
                              "rank0_hbm_write_gbs": world * vol_bytes / (ms * 1e-3) / 1e9}
    g.close()
    del g, bufs
    torch.cuda.empty_cache()
    # what bounds a gather into ONE GPU: its NVLink ingress (dense) and, for any form, its own HBM writes
    hbm_ms = world * vol_bytes / (peak * 1e9) * 1e3
    res["bound"] = {
        "rank0_hbm_write_ms": hbm_ms, "rank0_hbm_write_volumes_per_s": F * world / (hbm_ms * 1e-3),
        "dense_nvlink_ingress_ms": dense_bytes / (NVLINK_PEER_GBS * 1e9) * 1e3,
        "packed_nvlink_ingress_ms": nvlink_bytes / (NVLINK_PEER_GBS * 1e9) * 1e3,
        "note": "every volume of every rank is written into rank 0's HBM once (%.1f GB at the measured copy peak of "
                "%.0f GB/s): no gather into one GPU can beat ~%.1f M volumes/s, about one GPU's own voxelising rate; "
                "NVLink ingress at the measured %.0f GB/s peer-copy rate bounds the dense forms far below that"
                % (world * vol_bytes / 1e9, peak, F * world / hbm_ms / 1e3, NVLINK_PEER_GBS),
        "packed_frac_of_hbm_write_bound": hbm_ms / min(res["packed"]["ms"], res["packed_in_parts"]["ms"])}
    res["best"] = max(("nccl", "p2p", "packed", "packed_in_parts"), key=lambda k: res[k]["value"])
    return res


# ------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations, one short measurement each (device-resident, CUDA events)
# ------------------------------------------------------------------------------------------------
def run_secondary(pkg, dev, rank, world, args, max_over_ranks, barrier):
    import importlib
    import torch
    capi = pkg.capi
    amod = importlib.import_module(pkg.__name__ + ".augment")
    peak, _ = measured_peak()
    out = {}

    def timed(fn, reps=5, warm=3):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1)) / reps
        barrier()
        return ms

    def entry(name, frames, ms, alg, extra=None):
        e = {"frames_per_gpu": frames, "ms": ms, "value": frames * world / (ms * 1e-3), "unit": UNIT,
             "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": alg / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": alg}}
        if extra:
            e.update(extra)
        out[name] = e

    F = args.frames
    headers, depth = pkg.synth.make_uniform_batch(F, seed0=4321 + 7919 * rank, box=BOX)
    offsets, packed_len, max_px = capi.pack_plan(headers)
    d_packed = torch.from_numpy(depth.reshape(-1)).to(dev)
    d_off, d_hdr = torch.from_numpy(offsets).to(dev), torch.from_numpy(headers).to(dev)
    d_maxl = torch.empty((F,), dtype=torch.float32, device=dev)
    d_mid = torch.empty((F, 3), dtype=torch.float32, device=dev)

    # ---- configs[3]: resolution sweep (same crops; B = 24 + 4 bw bh + 12 R^3 + 16 per frame)
    for R, n in ((48, F), (64, min(F, 4000))):
        cfg = capi.default_config(R, args.layout, pkg.BBOX_PIXELS)
        vol = torch.empty((n, 3, R, R, R), dtype=torch.float32, device=dev)
        ms = timed(lambda: pkg.voxelize_device(d_packed, d_off[:n], d_hdr[:n], max_px, cfg, out_vol=vol,
                                               out_max_l=d_maxl[:n], out_mid_p=d_mid[:n]))
        entry("R%d" % R, n, ms, algorithmic_bytes(headers[:n], R), {"config": "configs[3]: same crops, R=%d" % R})
        del vol
        torch.cuda.empty_cache()

    R = 32
    cfg = capi.default_config(R, args.layout, pkg.BBOX_PIXELS)
    vol = torch.empty((F, 3, R, R, R), dtype=torch.float32, device=dev)

    # ---- configs[4] kernel: augmented D-TSDF (per-frame stretch + rotation), joints mapped in the same launch
    aug = torch.from_numpy(amod.make_augmentations(F, np.random.default_rng(rank), centers=np.tile([0, 0, -380.0], (F, 1)))).to(dev)
    ms = timed(lambda: amod.voxelize_aug_device(d_packed, d_off, d_hdr, aug, max_px, cfg, out_vol=vol, out_max_l=d_maxl,
                                                out_mid_p=d_mid))
    entry("augmented", F, ms, algorithmic_bytes(headers, R) + 84 * F,
          {"config": "configs[4] kernel: dtsdf_voxelize_aug_batch, random stretch [2/3,3/2] + rotation per frame"})
    del aug

    # ---- the semantics the reference runs in production: DataProcess.process() (bounding box from the
    #      resampled 6000-point cloud, pre/process.py:13-28).  Device kernels alone, then the host call.
    P = 6000
    lib = capi.load_library()
    import ctypes
    n_px = int(headers[0, 4] - headers[0, 2]) * int(headers[0, 5] - headers[0, 3])
    d_points = torch.empty((F * n_px, 3), dtype=torch.float64, device=dev)
    d_poff = (torch.arange(F, dtype=torch.int64, device=dev) * n_px).contiguous()
    d_nvalid = torch.empty((F,), dtype=torch.int32, device=dev)
    d_ext = torch.empty((F, 6), dtype=torch.float32, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def point_clouds():
        capi.check(lib.dtsdf_point_cloud_batch(F, ctypes.c_double(241.42), d_packed.data_ptr(), d_off.data_ptr(),
                                               d_hdr.data_ptr(), d_points.data_ptr(), d_poff.data_ptr(),
                                               d_nvalid.data_ptr(), stream))
    point_clouds()
    n_valid = d_nvalid.cpu().numpy()
    np.random.seed(1 + rank)
    d_idx = torch.from_numpy(capi.resample_indices(n_valid, P)).to(dev)
    cfg_e = capi.default_config(R, args.layout, pkg.BBOX_EXTREMA)

    def process_device():
        point_clouds()
        capi.check(lib.dtsdf_resample_extrema_batch(F, P, d_points.data_ptr(), d_poff.data_ptr(), d_idx.data_ptr(), None,
                                                    d_ext.data_ptr(), stream))
        pkg.voxelize_device(d_packed, d_off, d_hdr, max_px, cfg_e, aux=d_ext, out_vol=vol, out_max_l=d_maxl, out_mid_p=d_mid)
    ms = timed(process_device)
    alg = algorithmic_bytes(headers, R) + int(np.sum(4 * n_px + 24 * n_valid.astype(np.int64) + 28 * P + 24))
    entry("process_semantics_device_with_clouds", F, ms, alg,
          {"config": "bbox from the resampled 6000-point cloud, clouds materialised (what a caller that wants the "
                     "sampled points pays): point_cloud_kernel (f64) + resample_extrema_kernel + fused kernel (extrema "
                     "mode), index vectors resident",
           "bytes": "fused 455,656 + crop 4 bw bh + cloud 24 n_valid written + gather 24+4 per sample + 24"})
    del d_points
    torch.cuda.empty_cache()

    # the sampled cloud (what process() returns and read_MSRA.py stores) without the full one: two launches,
    # hand_points[rand_idx] straight from the depth map + fused kernel in extrema mode
    d_pc = torch.empty((F, P, 3), dtype=torch.float64, device=dev)

    def process_sampled_cloud():
        capi.check(lib.dtsdf_sample_cloud_batch(F, ctypes.c_double(241.42), P, d_packed.data_ptr(), d_off.data_ptr(),
                                                d_hdr.data_ptr(), d_idx.data_ptr(), int(max_px), d_pc.data_ptr(),
                                                d_ext.data_ptr(), None, stream))
        pkg.voxelize_device(d_packed, d_off, d_hdr, max_px, cfg_e, aux=d_ext, out_vol=vol, out_max_l=d_maxl, out_mid_p=d_mid)
    if max_px <= 22528:
        ms = timed(process_sampled_cloud)
        entry("process_semantics_device_with_sampled_clouds", F, ms,
              algorithmic_bytes(headers, R) + int(np.sum(4 * n_px + 0 * n_valid.astype(np.int64) + (4 + 24) * P + 24)),
              {"config": "bbox from the resampled 6000-point cloud, the SAMPLED clouds written (float64[6000,3] per frame, what "
                         "process() returns), the full cloud never built: dtsdf_sample_cloud_batch + fused kernel (extrema mode)",
               "bytes": "fused 455,656 + crop 4 bw bh + index 4 and point 24 per sample + 24"})
    del d_pc
    torch.cuda.empty_cache()

    # the same semantics in ONE launch when the sampled cloud itself is not wanted (3D_CNN/dataset.py never
    # reads it): ranks of the valid pixels, a bitmap of the sampled ranks and the extrema of the sample inside
    # the voxeliser (DTSDF_BBOX_SAMPLED) - no cloud is ever written
    cfg_s = capi.default_config(R, args.layout, pkg.BBOX_PIXELS)

    def process_sampled():
        pkg.voxelize_sampled_device(d_packed, d_off, d_hdr, d_idx, max_px, cfg_s, out_vol=vol, out_max_l=d_maxl,
                                    out_mid_p=d_mid)
    ms_count = timed(lambda: pkg.count_valid_device(d_packed, d_off, d_hdr, out=d_nvalid))
    ms = timed(process_sampled)
    entry("process_semantics_device", F, ms, algorithmic_bytes(headers, R) + 4 * P * F,
          {"config": "bbox from the resampled 6000-point cloud inside the voxeliser launch "
                     "(dtsdf_voxelize_sampled_batch), index vectors resident",
           "bytes": "fused 455,656 + 4 x 6000 index bytes per frame",
           "count_valid_ms": ms_count,
           "count_valid": "dtsdf_count_valid_batch over the same frames (what the index draw needs first; reads the crops)"})
    del d_idx, vol
    torch.cuda.empty_cache()

    # the host call: crops in, volumes out, index vectors drawn from np.random's MT19937 by the library
    vox = pkg.Voxelizer(device=dev.index, voxel_res=R)
    n_host = min(F, 4000)
    h_out = capi.PinnedBuffer((n_host, 3, R, R, R), np.float32)
    n_in = int(offsets[n_host]) if n_host < F else depth.size
    h_in = capi.PinnedBuffer((n_in,), np.float32)              # pinned both ways, like the e2e leg of the main line
    h_in.array[...] = depth.reshape(-1)[:n_in]
    packed_h = h_in.array
    np.random.seed(7 + rank)
    vox.process_packed(headers[:n_host], offsets[:n_host], packed_h, P, want_points=False, out_vol=h_out.array)
    barrier()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        vox.process_packed(headers[:n_host], offsets[:n_host], packed_h, P, want_points=False, out_vol=h_out.array)
    dt = max_over_ranks(time.perf_counter() - t0) / reps
    barrier()
    # what bounds it: the caller's random stream is ONE sequential MT19937 (each frame's 6000 draws start where
    # the previous frame's rejection sampling stopped): one thread has to find those starting points in order
    # (nothing written), helper threads then produce the draws; beyond that the dense stores of the volumes
    # into host memory, as for the plain host call (e2e.bound)
    t0 = time.perf_counter()
    for _ in range(reps):
        capi.resample_indices(n_valid[:n_host], P)
    dt_rng = max_over_ranks(time.perf_counter() - t0) / reps
    barrier()
    out["process_semantics_host_call"] = {
        "frames_per_gpu": n_host, "ms": dt * 1e3, "value": n_host * world / dt, "unit": "frames/s",
        "config": "Voxelizer.process_packed(want_points=False) = dtsdf_process_host: host crops in, volumes out (both pinned), "
                  "resampling indices drawn in C from the caller's np.random MT19937 state (bit-compatible with the "
                  "reference's per-frame np.random.randint calls, pre/process.py:75-80); dtsdf_count_valid_batch + "
                  "dtsdf_voxelize_sampled_batch per chunk, the crops uploaded and counted chunk by chunk behind the draws",
        "bound": {"what": "the index draw alone (dtsdf_resample_indices: one thread scans the sequential stream for "
                          "where each frame's draws begin, up to four helpers write them)",
                  "ms": dt_rng * 1e3, "value": n_host * world / dt_rng, "unit": "frames/s",
                  "ns_per_draw": dt_rng / (n_host * P) * 1e9},
        "frac": dt_rng / dt}
    vox.close()
    h_out.free()
    h_in.free()

    # ---- ragged crops (every frame its own 80-150 x 80-140 px bounding box)
    n_r = F
    frames = pkg.synth.make_frames(range(200000 + n_r * rank, 200000 + n_r * (rank + 1)))
    rh, ro, rp, rmax = pkg.pack_frames(frames)
    del frames
    dr = (torch.from_numpy(rp).to(dev), torch.from_numpy(ro).to(dev), torch.from_numpy(rh).to(dev))
    vol = torch.empty((n_r, 3, R, R, R), dtype=torch.float32, device=dev)
    ms = timed(lambda: pkg.voxelize_device(dr[0], dr[1], dr[2], rmax, cfg, out_vol=vol, out_max_l=d_maxl[:n_r],
                                           out_mid_p=d_mid[:n_r]))
    entry("ragged", n_r, ms, algorithmic_bytes(rh, R), {"config": "every frame its own 80-150 x 80-140 px crop, R=32"})
    return out


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from __graft_entry__ import build_cuda, load_package

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the D-TSDF path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stdout_fd = None
    if world > 1:
        # stdout carries the JSON line only: NCCL's own banner ("NCCL version ...", printed by the library on some
        # boxes whatever NCCL_DEBUG_FILE says) goes to stderr - file descriptor 1 points there until the line is printed
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/dtsdf_bench_nccl.%h.%p.log")
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    build_cuda()
    pkg = load_package()
    capi = pkg.capi
    R, F = args.res, args.frames

    # ---- workload: F fixed-shape synthetic frames per rank, in pinned host memory
    if args.ragged:
        # every frame its own bounding box (80-150 x 80-140 px)
        frames = pkg.synth.make_frames(range(100000 + F * rank, 100000 + F * (rank + 1)))
        headers, offsets, depth, max_px = pkg.pack_frames(frames)
        packed_len = depth.size
        del frames
    else:
        headers, depth = pkg.synth.make_uniform_batch(F, seed0=1234 + 7919 * rank, box=BOX)
        offsets, packed_len, max_px = capi.pack_plan(headers)
        assert packed_len == depth.size
    h_packed = capi.PinnedBuffer(packed_len, np.float32)
    h_packed.array[:] = depth.reshape(-1)
    del depth
    do_e2e = args.e2e_steps > 0
    h_vol = capi.PinnedBuffer((F, 3, R, R, R), np.float32) if do_e2e else None
    h_maxl = capi.PinnedBuffer(F, np.float32)
    h_mid = capi.PinnedBuffer((F, 3), np.float32)
    alg_bytes = algorithmic_bytes(headers, R)

    d_packed = torch.from_numpy(h_packed.array).to(dev)
    d_off = torch.from_numpy(offsets).to(dev)
    d_hdr = torch.from_numpy(headers).to(dev)
    d_vol = torch.empty((F, 3, R, R, R), dtype=torch.float32, device=dev)
    d_maxl = torch.empty((F,), dtype=torch.float32, device=dev)
    d_mid = torch.empty((F, 3), dtype=torch.float32, device=dev)
    cfg = capi.default_config(R, args.layout, pkg.BBOX_PIXELS)

    def step():
        pkg.voxelize_device(d_packed, d_off, d_hdr, max_px, cfg, out_vol=d_vol, out_max_l=d_maxl,
                            out_mid_p=d_mid)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)

    # ---- device-resident timing
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler:
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    launches = capi.launch_count() - launches0
    local_ms = e0.elapsed_time(e1)
    barrier()
    ms = max_over_ranks(local_ms)
    value = F * world * args.steps / (ms * 1e-3)

    # ---- gather of the volumes to rank 0 (BASELINE.json configs[2]: "... with volume all-gather to rank 0")
    gather = None
    if world > 1 and not args.no_gather:
        gather = run_gather_legs(pkg, dist, dev, rank, world, F, R, cfg, (d_packed, d_off, d_hdr, max_px),
                                 d_vol, d_maxl, d_mid, barrier, max_over_ranks)
    secondary = None
    if not args.no_secondary:
        del d_vol
        torch.cuda.empty_cache()
        secondary = run_secondary(pkg, dev, rank, world, args, max_over_ranks, barrier)
        d_vol = torch.empty((F, 3, R, R, R), dtype=torch.float32, device=dev)
        step()                                              # the end-to-end leg compares against it

    # ---- end to end through the host entry point (pinned host buffers, H2D + kernel + D2H)
    vox = pkg.Voxelizer(device=local_rank, voxel_res=R)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))       # (--e2e-steps 0 skips the end-to-end leg)

    def e2e_step():
        vox.voxelize_packed(headers, offsets, h_packed.array, layout=args.layout, bbox_mode=pkg.BBOX_PIXELS,
                            out_vol=h_vol.array, out_max_l=h_maxl.array, out_mid_p=h_mid.array)

    def time_e2e():
        for _ in range(2):
            e2e_step()
        barrier()
        with sampler:
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        barrier()
        return F * world * e2e_steps / max_over_ranks(dt)

    # plain device->host copies of the float volumes first (PCIe-bound), then the library's default:
    # volumes cross PCIe packed and are expanded into the host buffer by the host threads
    e2e_plain = e2e_value = pcie_d2h = ok = None
    e2e_extra = {}
    h2d = int(h_packed.array.nbytes + headers.nbytes + offsets.nbytes)
    d2h = int(F * 3 * R ** 3 * 4 + h_maxl.array.nbytes + h_mid.array.nbytes)
    if do_e2e:
        vox.set_option(capi.OPT_PACKED_D2H, 0)
        e2e_plain = time_e2e()
        vox.set_option(capi.OPT_PACKED_D2H, 1)
        vox.set_option(capi.OPT_LITERAL_PUSH, 1)
        e2e_push = time_e2e()
        vox.set_option(capi.OPT_LITERAL_PUSH, 0)
        e2e_value = time_e2e()
        pcie_d2h = vox.stat(capi.STAT_D2H_BYTES)
        # the e2e result is the real thing: compare ALL of it with the device-resident run
        ok = True
        for i0 in range(0, F, 1000):
            ok = ok and bool(torch.equal(torch.from_numpy(h_vol.array[i0:i0 + 1000]).to(dev).view(torch.int32),
                                         d_vol[i0:i0 + 1000].view(torch.int32)))
        # ---- what bounds it, measured in this run: (1) the host's stores of the dense volumes (the expansion
        # alone, same threads, same destination, all ranks at once), (2) the crops over PCIe (host -> device alone)
        lazy, _, _ = vox.voxelize_packed_lazy(headers, offsets, h_packed.array, layout=args.layout, bbox_mode=pkg.BBOX_PIXELS)
        lazy.expand(out=h_vol.array)
        best = float("inf")                            # a bound: the fastest of five repetitions (box noise is one-sided)
        for _ in range(5):
            barrier()
            t0 = time.perf_counter()
            lazy.expand(out=h_vol.array)
            best = min(best, max_over_ranks(time.perf_counter() - t0))
        expand_rate = F * world / best
        barrier()
        packed_host_bytes = lazy.nbytes
        lazy.close()
        torch.cuda.synchronize()
        t_h = torch.from_numpy(h_packed.array)
        d_tmp = torch.empty_like(d_packed)
        d_tmp.copy_(t_h, non_blocking=True)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            d_tmp.copy_(t_h, non_blocking=True)          # (page-locked by dtsdf_host_alloc: a plain DMA)
        c1.record()
        torch.cuda.synchronize()
        h2d_ms = max_over_ranks(c0.elapsed_time(c1)) / 3
        barrier()
        del d_tmp
        h2d_rate = F * world / (h2d_ms * 1e-3)

        def lazy_step():
            lz, _, _ = vox.voxelize_packed_lazy(headers, offsets, h_packed.array, layout=args.layout, bbox_mode=pkg.BBOX_PIXELS)
            lz.close()
        for _ in range(2):
            lazy_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            lazy_step()
        packed_result_rate = F * world * e2e_steps / max_over_ranks(time.perf_counter() - t0)
        barrier()
        bound_v = min(expand_rate, h2d_rate)
        e2e_extra = {
            "bound": {"value": bound_v, "unit": UNIT,
                      "what": "host stores of the dense volumes" if expand_rate <= h2d_rate else "crops over PCIe (H2D)",
                      "host_expansion_alone_value": expand_rate,
                      "host_expansion_alone_gbs": expand_rate * 3 * R ** 3 * 4 / 1e9,
                      "h2d_alone_value": h2d_rate, "h2d_alone_gbs": h_packed.array.nbytes * world / (h2d_ms * 1e-3) / 1e9,
                      "how": "measured in this run on all %d rank(s) at once: dtsdf_expand_frames of the same %d packed volumes "
                             "into the same pinned destination with the call's thread count; one pinned host -> device copy of "
                             "the crops" % (world, F)},
            "frac": e2e_value / bound_v,
            "push_kernel_value": e2e_push,
            "packed_result_value": packed_result_rate,
            "packed_result": "dtsdf_voxelize_host_packed: the volumes stay packed in host memory (%.1f KB per volume instead of "
                             "%.0f KB), expanded on demand by dtsdf_expand_frames" % (packed_host_bytes / F / 1e3, 3 * R ** 3 * 4 / 1e3),
        }

    if rank == 0:
        peak, peak_src = measured_peak()
        kernel_s = (local_ms * 1e-3) / args.steps
        achieved = alg_bytes / kernel_s / 1e9
        # the process() host call is bound by the index draw OR, like the plain call, by the host's dense stores
        # (measured above with all ranks expanding at once): state both, the fraction is against the tighter one
        hc = (secondary or {}).get("process_semantics_host_call")
        if hc and e2e_extra.get("bound"):
            stores = e2e_extra["bound"]["host_expansion_alone_value"]
            hc["bound"]["index_draw_value"] = hc["bound"]["value"]
            hc["bound"]["host_stores_value"] = stores
            hc["bound"]["value"] = min(hc["bound"]["value"], stores)
            hc["bound"]["what"] += "; the host's dense stores of the volumes (e2e.bound.host_expansion_alone_value, same " \
                                   "run); value = the lower of the two"
            hc["frac"] = hc["value"] / hc["bound"]["value"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps if do_e2e else 0, "api": "dtsdf_voxelize_host (C ABI, pinned host buffers)",
                    "matches_device_run_all_frames": ok,
                    "transfer": "volumes packed on the GPU (2-bit codes + literals, lossless), brought back by the copy "
                                "engine, expanded into the caller's buffer by host threads",
                    "d2h_bytes_on_pcie_per_step": pcie_d2h, "plain_copy_value": e2e_plain, **e2e_extra},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None if args.ragged else ncu_traffic(R, args.layout, F),
                         "traffic_source": "profiles/roofline_traffic.json: ncu --set full capture of this kernel "
                                           "(dram__bytes_read.sum + dram__bytes_write.sum per launch), scaled to this run's "
                                           "frame count; not re-measured inside this run",
                         "peak_source": peak_src, "kernel": "dtsdf_fused_kernel<%d,%d>" % (R, args.layout),
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": kernel_s * 1e3},
            "clocks": sampler.summary(),
        }
        if gather is not None:
            line["with_gather_to_rank0"] = gather
        if secondary is not None:
            line["secondary"] = secondary
        if world == 1 and not args.no_cpu_baseline:
            from __graft_entry__ import build_oracle
            build_oracle()
            rate, used, sample, passes, dt = cpu_port_rate(headers, offsets, h_packed.array, R, args.layout,
                                                           args.cpu_seconds)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": used, "kind": "port",
                "sample": "%d passes over the first %d frames of the same batch, %.1f s; C restatement of the "
                          "reference's Python voxel loop (oracle/dtsdf_oracle.c, gcc -O2, OpenMP over frames, "
                          "%d threads)" % (passes, sample, dt, used),
                "python_loop_volumes_per_s_per_core": python_loop_rate(headers, offsets, h_packed.array, R),
            }
            if R == 32:
                line["cpu_baseline"]["reference_numba_kernels_on_this_gpu"] = numba_gpu_rate(
                    headers, offsets, h_packed.array)
        if stdout_fd is not None:
            sys.stdout.flush()
            os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
    barrier()
    vox.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU")
    ap.add_argument("--res", type=int, default=32, choices=[32, 48, 64])
    ap.add_argument("--layout", type=int, default=0, choices=[0, 1], help="0 = (c,x,y,z), 1 = (c,z,y,x)")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: skip the voxelise + gather-to-rank-0 legs")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short measurements of the other configurations")
    ap.add_argument("--e2e-steps", type=int, default=10, help="0 skips the end-to-end leg (sizes beyond host memory)")
    ap.add_argument("--ragged", action="store_true", help="every frame its own bounding box instead of 120x130")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

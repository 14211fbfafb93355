// Hand-written sm_100a kernels for the depth-crop -> projective directional TSDF path.
//
// Reference behaviour (paths relative to the reference root):
//   bounding box over valid pixels ... pre/tsdf_numba.py:75-116,140-141
//   grid parameters ................... pre/process.py:91-96 == pre/tsdf_numba.py:142-147
//   per-voxel D-TSDF .................. pre/process.py:140-198 == pre/tsdf_numba.py:15-72
//   point cloud / resample / extrema .. pre/process.py:30-68, 82, 102-120
//   joint normalisation ............... pre/joint_nor.py:8-18, 3D_CNN/train.py:236-244
//
// Arithmetic contract: every float operation of the voxel path is one IEEE-754
// binary32 round-to-nearest operation in the reference's order (the reference's
// intermediates are np.float32 scalars under NumPy >= 2), so all of it is written
// with __fadd_rn/__fmul_rn/__fdiv_rn, which the compiler never contracts into FMAs.
// The file is also compiled with -fmad=false.
//
// Design: ONE CTA (8 warps) PER FRAME, grid = number of frames.  The hardware block
// scheduler balances ragged batches, and - measured on B200 - CTAs that come and go keep the
// HBM write stream 15-40 % faster than persistent CTAs marching in lock-step (a store-only
// variant of this kernel: 0.46 ms one CTA per frame vs 0.73 ms persistent for 8,500 frames).
// Two or three CTAs are resident per SM - as many as keep the volumes under construction inside the
// L2 (launch_fused) - so the prologue of one frame hides behind the voxel loops of the others.
//   0  thread 0: frame record, mbarrier init, ONE TMA bulk copy (cp.async.bulk) of the crop
//      global -> shared memory
//   1  bounding-box scan of the staged crop, one warp per crop row (LDS.128), which also
//      rewrites "no depth" pixels (|d| < 1) as NaN in place; warp reduction with
//      redux.sync.{min,max}.f32, block reduction through shared memory            [barrier]
//   2  grid parameters (every thread); projection tables at[s][f], yt (the projection
//      separates: pix_x depends on (x,z), pix_y on (y,z))                          [barrier]
//   3  voxel loop, two interleaved passes, warps free-running
//
// Voxel loop.  A thread owns a group of 4 consecutive voxels along the fastest output
// axis F (z for (c,x,y,z), x for (c,z,y,x)): one 16-byte store per channel.  Within a
// warp, lane = (slow-axis row, F-group): R/4 lanes cover one full F-row (128 B for R=32)
// and 32/(R/4) consecutive slow-axis rows share the warp; the warp then sweeps the middle
// axis (y).  The lane's table row at[s][f0..f0+3] and its voxel z centres stay in
// registers for the sweep; the per-iteration table read is a broadcast; the depth gather
// of a warp is a random gather (2.7 wavefronts on average, independent of the row pitch);
// the stores of a warp instruction are whole 128-byte lines.
//   pass 1 classifies every voxel with ~6 instructions: outside the crop / no depth
//     -> 0; farther than trunc from the surface along z -> +-1 (|v_z - w_z| > trunc
//     alone forces the Euclidean truncation); otherwise "near".  Sectors (32 B, an even/odd
//     lane pair) without near voxels are stored at once; the others are appended to a
//     per-warp list in shared memory (ballot + popc) and left to pass 2, so every byte
//     crosses the SM -> L2 write path once and no sector is ever written in halves (a
//     half-written sector costs L2 a fill from DRAM: measured 1.06 ms vs 0.76 ms).
//   pass 2 runs at the end of a row group on whole batches of 32 listed groups: every lane
//     takes one group, runs the full per-axis evaluation two voxels at a time on the packed
//     FP32 pipe (FFMA2/FMUL2) and stores the group's 3 x 16 bytes.
// Divisions: x/focal and x/trunc have per-launch / per-frame constant divisors; they
// are evaluated as q = a*y, then one or two residual corrections q += fma(-b,q,a)*y
// with y = rcp_rn(b) (Markstein).  With a correctly rounded reciprocal and a faithful
// q the last correction is the correctly rounded quotient; for the default focal
// 241.42f the one-correction form was checked exhaustively over every dividend with a
// biased exponent of 120..175 (the kernel uses 127..166; tests/test_division.py).  Frames or pixels outside the proven exponent ranges
// take the generic IEEE path (__fdiv_rn).  Results are bit-identical either way.
// The path is a streaming, write-dominated map with a small gather: HBM-bound, no
// tensor cores (there is no contraction to feed them).
#include "dtsdf_internal.h"
#include "dtsdf_ptx.cuh"

#include <math.h>
#include <stddef.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>

namespace dtsdf {

std::atomic<long long> g_launch_count{0};

namespace {

using namespace ptx;

constexpr int NT = 256;                 // threads per CTA
constexpr int NWARP = NT / 32;
constexpr float SENTINEL = 99999.0f;    // pre/tsdf_numba.py:88-89
constexpr int MAGIC_I = 0x4B400000;     // bit pattern of 12582912.0f = 1.5 * 2^23
constexpr float MAGIC_F = 12582912.0f;  // int_as_float(MAGIC_I + k) == MAGIC_F + k for |k| < 2^22
constexpr float THR_SQ = 1.00000011920928955078125f;   // nextafter(1,2): sqrt_rn(s) > 1 <=> s > THR_SQ
constexpr float BIG_DEPTH = 1099511627776.0f;          // 2^40: depth range of the FMA division proof
constexpr int OUTSIDE = -(1 << 30);                    // table sentinel: projection outside the crop
constexpr unsigned DEFAULT_FOCAL_BITS = 0x43716B85u;   // 241.42f
constexpr unsigned FULL = 0xffffffffu;
constexpr int QNAN_BITS = 0x7fc00000;

// ---------------------------------------------------------------- exact division by a fixed divisor
// a / b with y = rcp_rn(b): one (ONE=true) or two residual corrections.  Preconditions
// (checked by the callers): b and a/b well inside the normal range, see file header.
template <bool ONE>
__device__ __forceinline__ float div_fixed(float a, float b, float y) {
    float q = __fmul_rn(a, y);
    q = __fmaf_rn(__fmaf_rn(-b, q, a), y, q);
    if (!ONE) q = __fmaf_rn(__fmaf_rn(-b, q, a), y, q);
    return q;
}

// A 4-byte read of the staged crop through its shared-memory address: where a kernel keeps the crop as a
// pointer that is shared OR global at run time, the compiler can only emit generic LD.E with 64-bit address arithmetic.
__device__ __forceinline__ float lds_f32(unsigned saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}

// ---------------------------------------------------------------- shared memory
template <int R, int LAYOUT, int NW>
struct __align__(16) FrameShared {
    static constexpr int YS = (LAYOUT == 0) ? R : R + 1;   // (c,z,y,x): yt rows are read down a column
    int   at[R * R];                // [s][f]: 4*(pix_x - left), OUTSIDE when not in the crop
    int   yt[R * YS];               // CXYZ [y][z], CZYX [z][y]: 4*bw*(pix_y - top), or OUTSIDE
    short yr[R * YS];               // same indexing: pix_y - top (pass 2 only)
    float vx[R], vy[R], vz[R];      // voxel centres per axis
    float red[NW][8];            // per-warp partial min xyz / max xyz (6 used)
    // per-warp work lists (groups with near-surface voxels): what one slow-axis row group can
    // add plus the remainder (< 32) carried over from the previous one
    // a carried remainder (< 32) plus what one step of the sweep can add: a unit of the block map
    // (96 groups) or one slow-axis row group of the row map
    static constexpr int LIST_CAP = 32 + ((32 % (R / 4)) != 0 ? 96 : 32 * (R / NW));
    unsigned short list[NW][LIST_CAP];
    unsigned long long bar;         // crop TMA completion barrier
    unsigned int nanflag;           // the staged crop holds a NaN: take the unstaged path
    unsigned int pad_;
    float nodepth[4];               // nodepth[3] sits right below the crop: target of "outside" gathers (NaN)
};

struct Frame {
    int l, t, r, b, bw, bh, n, n_pad;
    const float* crop;
    bool staged;
};

// per-frame constants of the voxel evaluation (everything else comes from the kernel parameters)
struct VoxConst {
    int   ix;        // MAGIC_I + left:  int_as_float(ix + (px-left)) == MAGIC_F + px
    int   iy;        // MAGIC_I + top
    float trunc;
    float yt;        // rcp_rn(trunc)
    float t_far;     // |v_z - w_z| above this forces truncation; +inf disables the shortcut
};

// frame record (pre/tsdf_numba.py:122-129)
__device__ __forceinline__ Frame load_frame(const BatchArgs& a, int f) {
    const int* h = a.headers + 6 * (size_t)f;
    Frame fr;
    fr.l = __ldg(h + 2); fr.t = __ldg(h + 3); fr.r = __ldg(h + 4); fr.b = __ldg(h + 5);
    fr.bw = fr.r - fr.l;
    fr.bh = fr.b - fr.t;
    const int lim = 1 << 22;
    const bool bad = fr.bw <= 0 || fr.bh <= 0 || fr.bw > 32767 || fr.bh > 32767 ||
                     (long long)fr.bw * fr.bh > (1ll << 26) ||      // byte offsets stay below 2^28
                     fr.l <= -lim || fr.t <= -lim || fr.r >= lim || fr.b >= lim;
    if (bad) { fr.bw = 0; fr.bh = 0; fr.r = fr.l; fr.b = fr.t; }      // empty pixel ranges: every voxel is outside
    fr.n = fr.bw * fr.bh;
    fr.n_pad = (fr.n + 3) & ~3;
    fr.crop = a.depth + __ldg(a.offsets + f);
    fr.staged = fr.n_pad > 0 && fr.n_pad <= a.cap_floats;     // CTA-uniform
    return fr;
}

// ---------------------------------------------------------------- bounding box scan
// pre/tsdf_numba.py:83-96 without evaluating every pixel.  q = RN(d/f) is monotone in d,
// X = RN(q*(x-cx)) is monotone in q for a fixed column and Y = RN(-q*(y-cy)) for a fixed row,
// so the extremes of X (Y) over any set of pixels of one column (row) are attained at the
// set's smallest and largest valid depth: the scan only tracks min/max valid depth per column
// (in registers, 4 columns per lane and 128-column panel) and per lane and row, and
// back-projects those.  The result is the same min/max the reference computes over all valid
// pixels.  Invalid pixels (|d| < 1, pre/tsdf_numba.py:87) become NaN, which fminf/fmaxf skip;
// a row without valid pixels yields +-inf extremes whose quotient is NaN as well.
// One warp per crop row, four rows in flight per warp.
// SMEM: `crop` is the staged copy in shared memory; the scan writes the NaN-for-no-depth form
//       back in place (the voxel loop then needs no validity test) and reports whether the crop
//       holds NaNs of its own (`nansum`).  Otherwise `crop` is global memory, read-only.
// VEC:  rows start 16-byte aligned (bw % 4 == 0): a lane owns 4 adjacent columns (one 128-bit
//       access); otherwise columns lane, lane+32, lane+64, lane+96 of the panel.
// EXACT: IEEE division; otherwise the fixed-divisor form, valid while |d| < BIG_DEPTH (the
//       caller checks the depth range afterwards and rescans with EXACT if needed).
template <bool ONE, bool EXACT>
__device__ __forceinline__ float div_focal(float d, float focal, float yf) {
    return EXACT ? __fdiv_rn(d, focal) : div_fixed<ONE>(d, focal, yf);
}

template <bool ONE, bool EXACT, bool VEC, bool SMEM, int NW>
__device__ __forceinline__ void bbox_scan(const Frame& fr, float* crop, float focal, float yf, int cx, int cy,
                                          float (&mn)[3], float (&mx)[3], float2& nansum) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float kx = MAGIC_F + (float)cx;
    const float ky = MAGIC_F + (float)cy;
    const float qnan = __int_as_float(QNAN_BITS);
    float dlo = INFINITY, dhi = -INFINITY;           // valid depth extremes (Z = -d)
    for (int panel = 0; panel < fr.bw; panel += 128) {
        const int c0 = panel + (VEC ? 4 * lane : lane);          // first column of this lane
        constexpr int CS = VEC ? 1 : 32;                         // column stride of this lane
        bool inb[4];
        float cmin[4], cmax[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            inb[j] = c0 + CS * j < fr.bw;
            cmin[j] = INFINITY; cmax[j] = -INFINITY;
        }
        auto load_row = [&](int row, float (&d)[4]) {
            const float* p = crop + (size_t)row * fr.bw + c0;
            if (VEC) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (inb[0]) v = SMEM ? *reinterpret_cast<const float4*>(p)     // bw % 4 == 0: all four in range
                                     : __ldg(reinterpret_cast<const float4*>(p));
                d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) d[j] = inb[j] ? (SMEM ? p[32 * j] : __ldg(p + 32 * j)) : 0.0f;
            }
        };
        // one row: this lane's (up to) 4 pixels
        auto row_step = [&](int row, const float (&d)[4]) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = fabsf(d[j]) < 1.0f ? qnan : d[j];
                cmin[j] = fminf(cmin[j], v[j]); cmax[j] = fmaxf(cmax[j], v[j]);
            }
            if (SMEM) {
                // NaNs of the crop itself poison the running sums (so may inf - inf: harmless,
                // the frame then takes the general path)
                nansum = __fadd2_rn(nansum, __fadd2_rn(make_float2(d[0], d[1]), make_float2(d[2], d[3])));
                float* p = crop + (size_t)row * fr.bw + c0;
                if (VEC) {
                    if (inb[0]) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (inb[j]) p[32 * j] = v[j];
                }
            }
            const float rmin = fminf(fminf(fminf(v[0], INFINITY), v[1]), fminf(v[2], v[3]));
            const float rmax = fmaxf(fmaxf(fmaxf(v[0], -INFINITY), v[1]), fmaxf(v[2], v[3]));
            const float fy = __fsub_rn(__int_as_float(MAGIC_I + row + fr.t), ky);   // (y - cy), exact
            const float y1 = __fmul_rn(-div_focal<ONE, EXACT>(rmin, focal, yf), fy);
            const float y2 = __fmul_rn(-div_focal<ONE, EXACT>(rmax, focal, yf), fy);
            // a row without valid pixels: rmin = +inf, rmax = -inf.  EXACT: the products are
            // +-inf or NaN; fixed-divisor form: NaN (inf - inf in the residual).  Mask the former.
            const bool has = !EXACT || rmin <= rmax;
            mn[1] = fminf(mn[1], has ? fminf(y1, y2) : qnan);
            mx[1] = fmaxf(mx[1], has ? fmaxf(y1, y2) : qnan);
        };
        int row = warp;
        for (; row + 3 * NW < fr.bh; row += 4 * NW) {
            float d[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) load_row(row + u * NW, d[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) row_step(row + u * NW, d[u]);
        }
        for (; row < fr.bh; row += NW) {
            float d[4];
            load_row(row, d);
            row_step(row, d);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // a column without valid pixels gives NaN products (see above), which min/max skip
            const bool has = !EXACT || cmin[j] <= cmax[j];
            const float fx = __fsub_rn(__int_as_float(MAGIC_I + c0 + CS * j + fr.l), kx);   // (x - cx)
            const float x1 = __fmul_rn(div_focal<ONE, EXACT>(cmin[j], focal, yf), fx);
            const float x2 = __fmul_rn(div_focal<ONE, EXACT>(cmax[j], focal, yf), fx);
            mn[0] = fminf(mn[0], has ? fminf(x1, x2) : qnan);
            mx[0] = fmaxf(mx[0], has ? fmaxf(x1, x2) : qnan);
            dlo = fminf(dlo, cmin[j]); dhi = fmaxf(dhi, cmax[j]);
        }
    }
    if (dlo <= dhi) { mn[2] = fminf(mn[2], -dhi); mx[2] = fmaxf(mx[2], -dlo); }
}

// scan + warp reduction; lane 0 leaves the warp's partial in shared memory.  Returns true if
// this thread saw a NaN in the staged crop (SMEM only).
template <bool ONE, bool EXACT, bool SMEM, int NW>
__device__ __forceinline__ bool bbox_partial(const Frame& fr, float* crop, const BatchArgs& a, float yf,
                                             float (*red)[8]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mn[3], mx[3];
    float2 nansum = make_float2(0.f, 0.f);
    mn[0] = mn[1] = mn[2] = SENTINEL;
    mx[0] = mx[1] = mx[2] = -SENTINEL;
    if ((fr.bw & 3) == 0) bbox_scan<ONE, EXACT, true, SMEM, NW>(fr, crop, a.focal, yf, a.center_x, a.center_y, mn, mx, nansum);
    else                  bbox_scan<ONE, EXACT, false, SMEM, NW>(fr, crop, a.focal, yf, a.center_x, a.center_y, mn, mx, nansum);
#pragma unroll
    for (int c = 0; c < 3; ++c) { mn[c] = warp_min(mn[c]); mx[c] = warp_max(mx[c]); }
    if (lane == 0) {
        *reinterpret_cast<float4*>(&red[warp][0]) = make_float4(mn[0], mn[1], mn[2], mx[0]);
        *reinterpret_cast<float2*>(&red[warp][4]) = make_float2(mx[1], mx[2]);
    }
    const float t = __fadd_rn(nansum.x, nansum.y);
    return t != t;
}

// block reduction of the partials (every thread gets the result)
template <int NW>
__device__ __forceinline__ void bbox_reduce(const float (*red)[8], float (&mn)[3], float (&mx)[3]) {
    mn[0] = mn[1] = mn[2] = SENTINEL;
    mx[0] = mx[1] = mx[2] = -SENTINEL;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const float4 p = *reinterpret_cast<const float4*>(&red[w][0]);
        const float2 q = *reinterpret_cast<const float2*>(&red[w][4]);
        mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
        mx[0] = fmaxf(mx[0], p.w); mx[1] = fmaxf(mx[1], q.x); mx[2] = fmaxf(mx[2], q.y);
    }
}

// "no depth" -> NaN over the staged crop, for the bounding-box modes that do not scan it.
// Returns true if this thread saw a NaN of the crop's own.
template <int NW>
__device__ __forceinline__ bool sanitize_crop(float* crop_s, int n_pad) {
    float4* c4 = reinterpret_cast<float4*>(crop_s);
    const int n4 = n_pad >> 2;
    const float qnan = __int_as_float(QNAN_BITS);
    float2 nansum = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int i = threadIdx.x; i < n4; i += (32 * NW)) {
        float4 v = c4[i];
        nansum = __fadd2_rn(nansum, __fadd2_rn(make_float2(v.x, v.y), make_float2(v.z, v.w)));
        v.x = fabsf(v.x) < 1.0f ? qnan : v.x;
        v.y = fabsf(v.y) < 1.0f ? qnan : v.y;
        v.z = fabsf(v.z) < 1.0f ? qnan : v.z;
        v.w = fabsf(v.w) < 1.0f ? qnan : v.w;
        c4[i] = v;
    }
    const float t = __fadd_rn(nansum.x, nansum.y);
    return t != t;
}

// ---------------------------------------------------------------- sampled bounding box (DTSDF_BBOX_SAMPLED)
// DataProcess.process() without ever materialising a cloud (pre/process.py:13-28): the bounding box of
// hand_points[rand_idx] where hand_points are the back-projected pixels with d != 0 in row-major order
// (pre/process.py:30-68, float64) and rand_idx the caller's index vector (:70-84).
//   * rank of a pixel = number of valid pixels before it: counts per contiguous piece of the crop, one block
//     scan, then ballot / popc within a warp step;
//   * the sampled ranks are a bitmap in shared memory (one atomicOr per index; ranks beyond the bitmap are
//     handled by further windows over the same data - not needed below 43,008 valid pixels at R = 32);
//   * X = RN64(P / f) with P = RN64(u * d), u = (col + left) - w/2, is monotone in P, Y = RN64(-Q / f) in
//     Q = RN64(v * d), v = (row + top) - h/2, and float32 rounding is monotone too: the per-axis extremes over the
//     sample (pre/process.py:102-120) are those of P, Q and d, compared as order-preserving integers (-0 < +0,
//     like the FMNMX of resample_extrema_kernel); NaNs are skipped as fminf/fmaxf skip them.  FOUR float64
//     divisions per frame remain.
//   * sampled_bbox_fast: per sampled pixel the products u * d and v * d are formed EXACTLY as two floats each
//     (hi = RN32(u d), lo = fma(u, d, -hi): u, v have at most 11 significant bits, d 24) and compared as 64-bit
//     integers of their ordered bits; the extremes stay in registers - no float64 in the loop, no atomics.  This
//     needs finite depths and none so small that lo would underflow (|d| >= 2^-90), and n + 64 bits of bitmap;
//   * sampled_bbox_general: any other frame; every sampled pixel is multiplied out in float64.
// Result in shared memory (FrameShared::red[0]): {max x, y, z, min x, y, z} - what resample_extrema_kernel
// produces from the materialised cloud, bit for bit.  Scratch is the projection-table area of FrameShared (not
// yet in use).
__device__ __forceinline__ long long dkey(double x) {
    const long long b = __double_as_longlong(x);
    return b ^ ((b >> 63) & 0x7fffffffffffffffll);            // an involution
}
__device__ __forceinline__ double dkey_inv(long long k) { return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffll)); }
__device__ __forceinline__ int fkey(float x) {
    const int b = __float_as_int(x);
    return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float fkey_inv(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }


template <int NW>
struct SampledScratch {
    long long kred[NW][4];        // per warp: min P, max P, min Q, max Q
    float     fred[NW][2];        // per warp: min d, max d
    int       wsum[NW];
    int       flag[NW];           // per warp: its piece holds an infinite depth
    unsigned  bitmap[1];          // ... to the end of the scratch area
};

struct KeyRange {
    long long pmin, pmax, qmin, qmax;
    float dmin, dmax;
    __device__ __forceinline__ void init() {
        pmin = qmin = 0x7fffffffffffffffll;
        pmax = qmax = -0x7fffffffffffffffll - 1;
        dmin = INFINITY; dmax = -INFINITY;
    }
    __device__ __forceinline__ void p(double P) { if (P == P) { const long long k = dkey(P); pmin = min(pmin, k); pmax = max(pmax, k); } }
    __device__ __forceinline__ void q(double Q) { if (Q == Q) { const long long k = dkey(Q); qmin = min(qmin, k); qmax = max(qmax, k); } }
};

// block reduction of the per-thread ranges and the four divisions; ends with a barrier
template <int NW>
__device__ __forceinline__ void sampled_finish(KeyRange kr, SampledScratch<NW>* sc, float* out6, double focal) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kr.pmin = min(kr.pmin, __shfl_xor_sync(FULL, kr.pmin, o)); kr.pmax = max(kr.pmax, __shfl_xor_sync(FULL, kr.pmax, o));
        kr.qmin = min(kr.qmin, __shfl_xor_sync(FULL, kr.qmin, o)); kr.qmax = max(kr.qmax, __shfl_xor_sync(FULL, kr.qmax, o));
    }
    const float dmin = warp_min(kr.dmin), dmax = warp_max(kr.dmax);
    if (lane == 0) {
        sc->kred[warp][0] = kr.pmin; sc->kred[warp][1] = kr.pmax; sc->kred[warp][2] = kr.qmin; sc->kred[warp][3] = kr.qmax;
        sc->fred[warp][0] = dmin; sc->fred[warp][1] = dmax;
    }
    __syncthreads();
    if (tid < 4) {
        // tid 0: max x (largest P)  1: min x (smallest P)  2: max y (smallest Q, negated)  3: min y (largest Q, negated)
        const int pair = tid < 2 ? 0 : 2;
        long long kmn = sc->kred[0][pair], kmx = sc->kred[0][pair + 1];
        for (int i = 1; i < NW; ++i) { kmn = min(kmn, sc->kred[i][pair]); kmx = max(kmx, sc->kred[i][pair + 1]); }
        const bool is_max = tid == 0 || tid == 2;
        float v = is_max ? -INFINITY : INFINITY;             // an empty sample leaves the fminf / fmaxf identities
        if (kmn <= kmx) {
            const double num = tid == 0 ? dkey_inv(kmx) : (tid == 1 ? dkey_inv(kmn) : (tid == 2 ? -dkey_inv(kmn) : -dkey_inv(kmx)));
            v = __double2float_rn(__ddiv_rn(num, focal));
        }
        out6[(is_max ? 0 : 3) + (tid >> 1)] = v;             // {max x, max y, max z, min x, min y, min z}
    } else if (tid == 4) {
        float lo = sc->fred[0][0], hi = sc->fred[0][1];
        for (int i = 1; i < NW; ++i) { lo = fminf(lo, sc->fred[i][0]); hi = fmaxf(hi, sc->fred[i][1]); }
        out6[2] = -lo;            // max z = -(min d)
        out6[5] = -hi;
    }
    __syncthreads();
}

// marks the ranks [w0, w0 + wn) that the index vector holds; eight loads in flight per thread
template <int NTH>
__device__ __forceinline__ void mark_samples(const int* idx, int point_num, int w0, int wn, unsigned* bitmap) {
    for (int k0 = threadIdx.x; k0 < point_num; k0 += 8 * NTH) {
        int v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = k0 + j * NTH < point_num ? idx[k0 + j * NTH] : -1;      // (shared or global)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const unsigned r = (unsigned)(v[j] - w0);                    // out-of-range indices set nothing
            if (v[j] >= 0 && r < (unsigned)wn) atomicOr(&bitmap[r >> 5], 1u << (r & 31));
        }
    }
}

// what sampled_bbox_fast expects to find (the kernel does this while the crop is still on its way, ahead of a
// barrier it has anyway): a cleared bitmap for the n ranks a crop of n pixels can have (+ two spare words)
template <int NW>
__device__ __forceinline__ void sampled_scratch_init(unsigned char* smem_raw, int n) {
    constexpr int NTH = 32 * NW;
    SampledScratch<NW>* sc = reinterpret_cast<SampledScratch<NW>*>(smem_raw);
    for (int i = threadIdx.x; i < (n + 31) / 32 + 2; i += NTH) sc->bitmap[i] = 0u;
}

// Needs n + 64 <= CAP_BITS and sampled_scratch_init + a barrier before it.
// Returns false (CTA-uniform, no result written) when the crop holds an infinite or a tiny (< 2^-90) depth.
template <int NW, bool SMEM, int CROP_OFF, int OUT6_OFF>
__device__ __noinline__ bool sampled_bbox_fast(const float* __restrict__ crop_g, int n, int bw, int bh,
                                               int l, int t, int w, int h, double focal,
                                               const int* idx, int point_num) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NTH = 32 * NW;
    SampledScratch<NW>* sc = reinterpret_cast<SampledScratch<NW>*>(smem_raw);
    const float* crop = SMEM ? reinterpret_cast<const float*>(smem_raw + CROP_OFF) : crop_g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    if (abs(w) >= (1 << 20) || abs(h) >= (1 << 20)) return false;      // (u, v would not be exact float32 values)
    // the sampled ranks (all below the number of valid pixels, hence below n; anything else marks nothing)
    mark_samples<NTH>(idx, point_num, 0, n, sc->bitmap);
    // a warp owns one contiguous piece of the crop, a whole number of 32-pixel steps
    const int S = ((n + NTH - 1) / NTH) * 32;
    const int s0 = min(n, warp * S), s1 = min(n, s0 + S);
    int cnt = 0;
    float amax = 0.0f, amin = INFINITY;
#pragma unroll 4
    for (int q = s0; q < s1; q += 32) {
        const int p = q + lane;
        const float d = p < s1 ? (SMEM ? crop[p] : __ldg(crop + p)) : 0.0f;
        cnt += d != 0.0f;
        amax = fmaxf(amax, fabsf(d));                          // (skips NaN)
        amin = fminf(amin, d != 0.0f ? fabsf(d) : INFINITY);
    }
    cnt = __reduce_add_sync(FULL, cnt);
    // infinite depths (0 * inf = NaN) and depths so small that the low half of a two-float product would be lost
    // to underflow go to the general form
    const unsigned anyinf = __ballot_sync(FULL, amax == INFINITY || amin < 0x1p-90f);
    if (lane == 0) { sc->wsum[warp] = cnt; sc->flag[warp] = anyinf != 0u; }
    __syncthreads();
    int base = 0, inf = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        if (i < warp) base += sc->wsum[i];
        inf |= sc->flag[i];
    }
    if (inf) return false;
    // P = u d and Q = v d are exact as TWO floats: u, v are integers or half-integers below 2^10 and d has 24 bits, so
    // hi = RN32(u d) and lo = fma(u, d, -hi) hold the product exactly (no underflow: |d| >= 2^-90 here), and since
    // rounding is monotone the order of the exact products is the lexicographic order of (hi, lo) - compared as one
    // 64-bit integer of their ordered bits (-0 < +0 as in dkey).  The extremes of the sample stay in registers: no
    // per-column / per-row tables, no atomics (round 2, third form of this pass: four shared-memory atomics per
    // sampled pixel before, 1.015 ms per 8,500 frames).
    auto key2 = [](float hi, float lo) -> long long {
        return (long long)(((unsigned long long)(unsigned)fkey(hi) << 32) | (unsigned)(fkey(lo) ^ (int)0x80000000));
    };
    long long kpmin = 0x7fffffffffffffffll, kpmax = -0x7fffffffffffffffll - 1, kqmin = kpmin, kqmax = kpmax;
    float sdmin = INFINITY, sdmax = -INFINITY;
    const float half_wf = (float)w * 0.5f, half_hf = (float)h * 0.5f;
    {
        {
            const int w0 = 0, wn = n;
            int rb = base - w0;
            int row = (s0 + lane) / bw, col = (s0 + lane) - row * bw;
            const bool wide = bw >= 32;                        // a 32-pixel step then crosses at most one row end
            // four steps per iteration, every load ahead of the atomics that follow (the compiler cannot move
            // a shared-memory load across them itself): depth, ballot and rank of all four, then the four
            // bitmap words, then the updates
            constexpr int U = 4;
#pragma unroll 1
            for (int q = s0; q < s1; q += 32 * U) {
                float d[U];
                int r[U];
                unsigned word[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int p = q + 32 * u + lane;
                    d[u] = p < s1 ? (SMEM ? crop[p] : __ldg(crop + p)) : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const unsigned bal = __ballot_sync(FULL, d[u] != 0.0f);
                    r[u] = rb + __popc(bal & lt);
                    rb += __popc(bal);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool inwin = d[u] != 0.0f && (unsigned)r[u] < (unsigned)wn;
                    word[u] = inwin ? sc->bitmap[r[u] >> 5] : 0u;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (((word[u] >> (r[u] & 31)) & 1u) && d[u] == d[u]) {
                        const float uf = __fsub_rn((float)(col + l), half_wf), vf = __fsub_rn((float)(row + t), half_hf);
                        const float ph = __fmul_rn(uf, d[u]), qh = __fmul_rn(vf, d[u]);
                        const long long kp = key2(ph, __fmaf_rn(uf, d[u], -ph)), kq = key2(qh, __fmaf_rn(vf, d[u], -qh));
                        kpmin = min(kpmin, kp); kpmax = max(kpmax, kp);
                        kqmin = min(kqmin, kq); kqmax = max(kqmax, kq);
                        sdmin = fminf(sdmin, d[u]); sdmax = fmaxf(sdmax, d[u]);
                    }
                    col += 32;
                    if (wide) { if (col >= bw) { col -= bw; ++row; } }
                    else      { while (col >= bw) { col -= bw; ++row; } }
                }
            }
        }
    }
    // back to float64: hi + lo is the exact product, which is what RN64(u d) is (35 significant bits)
    auto exact = [](long long k) -> double {
        const float hi = fkey_inv((int)(k >> 32)), lo = fkey_inv((int)((unsigned)k ^ 0x80000000u));
        return lo == 0.0f ? (double)hi : __dadd_rn((double)hi, (double)lo);      // (-0 + +0 would lose the sign of a zero product)
    };
    KeyRange kr;
    kr.init();
    if (kpmin <= kpmax) {
        kr.p(exact(kpmin)); kr.p(exact(kpmax));
        kr.q(exact(kqmin)); kr.q(exact(kqmax));
        kr.dmin = sdmin; kr.dmax = sdmax;
    }
    sampled_finish<NW>(kr, sc, reinterpret_cast<float*>(smem_raw + OUT6_OFF), focal);
    return true;
}

// any shape, any depth: every thread owns one contiguous run of the crop (odd length: the strided shared-memory
// reads are conflict-free) and multiplies its sampled pixels out
template <int NW, bool SMEM, int CROP_OFF, int OUT6_OFF, int CAP_BITS>
__device__ __noinline__ void sampled_bbox_general(const float* __restrict__ crop_g, int n, int bw,
                                                  int l, int t, int w, int h, double focal,
                                                  const int* __restrict__ idx, int point_num) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NTH = 32 * NW;
    SampledScratch<NW>* sc = reinterpret_cast<SampledScratch<NW>*>(smem_raw);
    const float* crop = SMEM ? reinterpret_cast<const float*>(smem_raw + CROP_OFF) : crop_g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = ((n + NTH - 1) / NTH) | 1;
    const int p0 = min(n, tid * L), p1 = min(n, p0 + L);
    int cnt = 0;
    for (int p = p0; p < p1; ++p) cnt += (SMEM ? crop[p] : __ldg(crop + p)) != 0.0f;
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += v;
    }
    __syncthreads();                                           // (the fast attempt's readers of wsum are done)
    if (lane == 31) sc->wsum[warp] = inc;
    __syncthreads();
    int base = inc - cnt, total = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        const int c = sc->wsum[i];
        if (i < warp) base += c;
        total += c;
    }
    KeyRange kr;
    kr.init();
    const double half_w = (double)w / 2.0, half_h = (double)h / 2.0;
    for (int w0 = 0; w0 < total; w0 += CAP_BITS) {
        const int wn = min(CAP_BITS, total - w0);
        for (int i = tid; i < (wn + 31) / 32; i += NTH) sc->bitmap[i] = 0u;
        __syncthreads();
        mark_samples<NTH>(idx, point_num, w0, wn, sc->bitmap);
        __syncthreads();
        if (base < w0 + wn && base + cnt > w0) {
            int r = base - w0;
            int row = p0 / bw, col = p0 - row * bw;
            for (int p = p0; p < p1; ++p) {
                const float d = SMEM ? crop[p] : __ldg(crop + p);
                if (d != 0.0f) {
                    if ((unsigned)r < (unsigned)wn && ((sc->bitmap[r >> 5] >> (r & 31)) & 1u)) {
                        const double dd = (double)d;
                        kr.p(__dmul_rn(((double)col + (double)l) - half_w, dd));
                        kr.q(__dmul_rn(((double)row + (double)t) - half_h, dd));
                        kr.dmin = fminf(kr.dmin, d); kr.dmax = fmaxf(kr.dmax, d);
                    }
                    ++r;
                }
                if (++col == bw) { col = 0; ++row; }
            }
        }
        __syncthreads();
    }
    sampled_finish<NW>(kr, sc, reinterpret_cast<float*>(smem_raw + OUT6_OFF), focal);
}

// ---------------------------------------------------------------- voxel evaluation
struct Tsdf3 { float x, y, z; };

// Generic IEEE evaluation of one valid voxel: pre/process.py:172-194 == pre/tsdf_numba.py:43-68.
// sz = v_z + d (= v_z - w_z).
__device__ __noinline__ Tsdf3 eval_generic(float d, float fx, float fy, float vx, float vy, float vz,
                                           float sz, float focal, float trunc) {
    const float c1 = __fdiv_rn(d, focal);
    const float wx = __fmul_rn(fx, c1);
    const float wy = __fmul_rn(fy, c1);
    float tx = __fdiv_rn(fabsf(__fsub_rn(vx, wx)), trunc);
    float ty = __fdiv_rn(fabsf(__fsub_rn(vy, wy)), trunc);
    float tz = __fdiv_rn(fabsf(sz), trunc);
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(tx, tx), __fmul_rn(ty, ty)), __fmul_rn(tz, tz));
    // sqrt_rn(s) > 1 <=> s > nextafter(1,2); below it every component is <= 1 already, so
    // min(t,1) (pre/process.py:187-189) is the identity there.
    if (s > THR_SQ) { tx = 1.0f; ty = 1.0f; tz = 1.0f; }
    if (-d > vz) { tx = -tx; ty = -ty; tz = -tz; }
    Tsdf3 o; o.x = tx; o.y = ty; o.z = tz;
    return o;
}

// Two voxels at a time on the packed FP32 pipe of sm_100 (FFMA2 / FMUL2: two IEEE binary32
// round-to-nearest operations per instruction), with the fixed-divisor divisions.  Same
// operations in the same order as eval_generic, bit-identical inside the proven ranges.
// The quotients keep the sign of their dividends (RN division is odd); |.| and the final sign
// (-d > v_z  <=>  v_z + d < 0) are one LOP3 per output.
template <bool ONE>
__device__ __forceinline__ float2 div_fixed2(float2 a, float2 nb, float2 y) {      // nb = -b
    float2 q = __fmul2_rn(a, y);
    q = __ffma2_rn(__ffma2_rn(nb, q, a), y, q);
    if (!ONE) q = __ffma2_rn(__ffma2_rn(nb, q, a), y, q);
    return q;
}

// A caution that cost a round: ptxas (12.9) CONTRACTS packed operations.  mul.rn.f32x2 followed by add.rn.f32x2
// becomes one FFMA2 - the explicit .rn that protects the scalar forms does not protect the x2 forms, with
// -fmad=false or without - and an fma.rn.f32x2 by +-1 is first simplified to an add and then contracted the same
// way.  So the packed pipe is used only where no product can reach an addition: sums of values that are not
// products, and fused multiply-adds that ARE fused in the specification (the residual corrections of
// div_fixed2).  Every product that is then added or subtracted - w = f * c1 -> v - w, the squares -> their
// sum - is a scalar mul.rn feeding a scalar add.rn, which PTX guarantees are kept apart
// (tests/test_sass_contraction.py checks the compiled code for it).
template <bool DEFF>
__device__ __forceinline__ void eval_pair(float2 d, float2 fx, float2 fy, float2 vx, float2 vy, float2 sz,
                                          float focal, float yf, float trunc, float ytr,
                                          float (&o0)[2], float (&o1)[2], float (&o2)[2]) {
    const float2 nf = make_float2(-focal, -focal), yf2 = make_float2(yf, yf);
    const float2 nt = make_float2(-trunc, -trunc), yt2 = make_float2(ytr, ytr);
    const float2 c1 = div_fixed2<DEFF>(d, nf, yf2);
    // v - w with w = f * c1: product and difference each rounded once
    const float2 dx = make_float2(__fsub_rn(vx.x, __fmul_rn(fx.x, c1.x)), __fsub_rn(vx.y, __fmul_rn(fx.y, c1.y)));
    const float2 dy = make_float2(__fsub_rn(vy.x, __fmul_rn(fy.x, c1.x)), __fsub_rn(vy.y, __fmul_rn(fy.y, c1.y)));
    const float2 tx = div_fixed2<false>(dx, nt, yt2);      // signed; |t| below
    const float2 ty = div_fixed2<false>(dy, nt, yt2);
    const float2 tz = div_fixed2<false>(sz, nt, yt2);
    const float ss[2] = {__fadd_rn(__fadd_rn(__fmul_rn(tx.x, tx.x), __fmul_rn(ty.x, ty.x)), __fmul_rn(tz.x, tz.x)),
                         __fadd_rn(__fadd_rn(__fmul_rn(tx.y, tx.y), __fmul_rn(ty.y, ty.y)), __fmul_rn(tz.y, tz.y))};
    const float szz[2] = {sz.x, sz.y};
    const float txx[2] = {tx.x, tx.y}, tyy[2] = {ty.x, ty.y}, tzz[2] = {tz.x, tz.y};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int sg = __float_as_int(szz[i]) & 0x80000000;
        const bool sat = ss[i] > THR_SQ;
        const int one = 0x3F800000 | sg;
        o0[i] = __int_as_float(sat ? one : ((__float_as_int(txx[i]) & 0x7fffffff) | sg));
        o1[i] = __int_as_float(sat ? one : ((__float_as_int(tyy[i]) & 0x7fffffff) | sg));
        o2[i] = __int_as_float(sat ? one : ((__float_as_int(tzz[i]) & 0x7fffffff) | sg));
    }
}

// byte offset into the crop; negative <=> either table entry is OUTSIDE; reads as "no depth"
template <bool STAGED>
__device__ __forceinline__ float gather_depth(const float* __restrict__ crop, int off) {
    if (STAGED) {
        // the word right below the staged crop reads as "no depth" (FrameShared::nodepth[3])
        off = max(off, -4);
        return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(crop) + off);
    }
    if (off < 0) return 0.0f;
    return __ldg(reinterpret_cast<const float*>(reinterpret_cast<const char*>(crop) + off));
}

// Classification of one voxel (pass 1).
//   value 0        : projection outside the crop or |depth| < 1 (pre/process.py:159,167)
//   value +-1      : farther than trunc from the surface along z alone: t_z > 1, so the
//                    Euclidean truncation fires (pre/process.py:180-185) and the result is +-1
//                    by the side of the surface (:191-194)
//   near           : needs the full evaluation (value 0 for now)
// SAN: the staged crop was sanitised (|d| < 1 -> NaN, FrameShared::nodepth[3] is NaN too), so
// "no depth" makes both comparisons false and needs no test of its own.
template <bool SAN>
__device__ __forceinline__ void classify(float d, float vz, float t_far, float& out, bool& near) {
    const float sz = __fadd_rn(vz, d);                      // v_z - w_z, w_z = -d
    const float one = __int_as_float((__float_as_int(sz) & 0x80000000) | 0x3F800000);
    if (SAN) {
        out = fabsf(sz) > t_far ? one : 0.0f;
        near = near || fabsf(sz) <= t_far;
    } else {
        const bool ok = !(fabsf(d) < 1.0f);
        const bool far = fabsf(sz) > t_far;
        out = (ok && far) ? one : 0.0f;
        near = near || (ok && !far);
    }
}

// Lane geometry of the voxel loop for an R^3 grid.
//   Row map (R = 32, 64): lane = (slow-axis row, F-group); R/4 lanes cover one F-row, 32/(R/4) consecutive
//   slow-axis rows share the warp, which sweeps the middle axis.
//   Block map (R = 48, where an F-row is 12 groups and 32 lanes would cover 2.67 rows): the unit of work
//   is a block of MB = 8 consecutive middle-axis rows of one slow-axis row = 96 groups = 1,536 contiguous
//   bytes per channel, swept in T = 3 warp steps of 32 consecutive groups (512 B: whole lines, all lanes
//   busy).  With the row map only 24 of 32 lanes work and pass 1 costs 1.33x the instructions per voxel:
//   the R = 48 kernel was then SM-bound (0.88 of the copy peak at any occupancy, while its store-only
//   ablation ran at 0.99), see DESIGN.md.
template <int R, int NW>
struct LaneMap {
    static constexpr int LPR = R / 4;               // lanes per F-row (groups of 4 voxels)
    static constexpr bool BLOCKED = (32 % LPR) != 0;
    static constexpr int ROWS = BLOCKED ? 1 : 32 / LPR;   // slow-axis rows per warp (row map)
    static constexpr int ACTIVE = BLOCKED ? 32 : ROWS * LPR;
    static constexpr int NSG = R / ROWS;            // slow-axis row groups per frame
    static constexpr int MSTEPS = R / NW;           // middle-axis positions per warp and row group (row map)
    static constexpr int MB = 8, T = (MB * LPR) / 32, NMB = R / MB;      // block map
    static_assert(R % 4 == 0 && R % ROWS == 0 && R <= 64, "unsupported grid size");
    static_assert(BLOCKED ? ((MB * LPR) % 32 == 0 && R % MB == 0 && LPR % 2 == 0 && LPR <= 16)
                          : (R % NW == 0 && NSG <= 32), "work-list entries are 16 bits");
    // work-list entry -> (slow-axis index s, middle-axis index m, first voxel f0 of the group)
    __device__ static __forceinline__ void decode(int ent, int& s, int& m, int& f0) {
        if (BLOCKED) {                              // s (6 bits) | m (6 bits) | F-group (4 bits)
            s = ent >> 10; m = (ent >> 4) & 63; f0 = 4 * (ent & 15);
        } else {                                    // row group (5 bits) | m (6 bits) | lane (5 bits)
            const int L = ent & 31;
            m = (ent >> 5) & 63;
            s = (ent >> 11) * ROWS + L / LPR;
            f0 = 4 * (L % LPR);
        }
    }
};

// pass 2: full evaluation of the n (<= 32) listed groups list[0..n)
template <int R, int LAYOUT, int NW, bool STAGED, bool FAST, bool DEFF>
__device__ __forceinline__ void eval_listed(const float* __restrict__ crop,
                                            const FrameShared<R, LAYOUT, NW>* __restrict__ sh, const VoxConst& k,
                                            const BatchArgs& a, float yf, float* __restrict__ out,
                                            const unsigned short* list, int n) {
    using LM = LaneMap<R, NW>;
    constexpr int YS = FrameShared<R, LAYOUT, NW>::YS;
    constexpr int CH = R * R * R;
    const int lane = threadIdx.x & 31;
    if (lane < n) {
        const int ent = list[lane];
        int s, m, f0;
        LM::decode(ent, s, m, f0);
        const int4 A = *reinterpret_cast<const int4*>(&sh->at[s * R + f0]);
        const int a4[4] = {A.x, A.y, A.z, A.w};
        int b4[4], row4[4];
        float vx4[4], vz4[4];
        const float vy = sh->vy[m];
        if (LAYOUT == 0) {
            const int4 B = *reinterpret_cast<const int4*>(&sh->yt[m * YS + f0]);
            b4[0] = B.x; b4[1] = B.y; b4[2] = B.z; b4[3] = B.w;
            const short4 rr = *reinterpret_cast<const short4*>(&sh->yr[m * YS + f0]);
            row4[0] = rr.x; row4[1] = rr.y; row4[2] = rr.z; row4[3] = rr.w;
            const float4 z4 = *reinterpret_cast<const float4*>(&sh->vz[f0]);
            vz4[0] = z4.x; vz4[1] = z4.y; vz4[2] = z4.z; vz4[3] = z4.w;
            const float x1 = sh->vx[s];
            vx4[0] = vx4[1] = vx4[2] = vx4[3] = x1;
        } else {
            const int b1 = sh->yt[s * YS + m];
            const int r1 = sh->yr[s * YS + m];
            const float z1 = sh->vz[s];
            const float4 x4 = *reinterpret_cast<const float4*>(&sh->vx[f0]);
            vx4[0] = x4.x; vx4[1] = x4.y; vx4[2] = x4.z; vx4[3] = x4.w;
#pragma unroll
            for (int j = 0; j < 4; ++j) { b4[j] = b1; row4[j] = r1; vz4[j] = z1; }
        }
        const float kx = MAGIC_F + (float)a.center_x, ky = MAGIC_F + (float)a.center_y;
        float d4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) d4[j] = gather_depth<STAGED>(crop, a4[j] + b4[j]);
        float o0[4], o1[4], o2[4];
#pragma unroll
        for (int h = 0; h < 4; h += 2) {
            float sz[2], fx[2], fy[2];
            bool ok[2], far[2], big[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int j = h + i;
                sz[i] = __fadd_rn(vz4[j], d4[j]);
                // sanitised staged crop: "no depth" is NaN, and real NaNs never get here
                ok[i] = STAGED ? (d4[j] == d4[j]) : !(fabsf(d4[j]) < 1.0f);
                far[i] = fabsf(sz[i]) > k.t_far;
                big[i] = fabsf(d4[j]) >= BIG_DEPTH;          // (NaN = no depth: not evaluated, not big)
                fx[i] = __fsub_rn(__int_as_float(k.ix + (a4[j] >> 2)), kx);   //  (pix_x - cx)
                fy[i] = __fsub_rn(ky, __int_as_float(k.iy + row4[j]));        // -(pix_y - cy)
            }
            const bool near0 = ok[0] && !far[0], near1 = ok[1] && !far[1];
            float p0[2] = {0.0f, 0.0f}, p1[2] = {0.0f, 0.0f}, p2[2] = {0.0f, 0.0f};
            if (near0 || near1) {
                if (FAST && !big[0] && !big[1]) {
                    eval_pair<DEFF>(make_float2(d4[h], d4[h + 1]), make_float2(fx[0], fx[1]),
                                    make_float2(fy[0], fy[1]), make_float2(vx4[h], vx4[h + 1]),
                                    make_float2(vy, vy), make_float2(sz[0], sz[1]), a.focal, yf, k.trunc, k.yt,
                                    p0, p1, p2);
                } else {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const Tsdf3 t = eval_generic(d4[h + i], fx[i], fy[i], vx4[h + i], vy, vz4[h + i], sz[i],
                                                     a.focal, k.trunc);
                        p0[i] = t.x; p1[i] = t.y; p2[i] = t.z;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float one = __int_as_float((__float_as_int(sz[i]) & 0x80000000) | 0x3F800000);
                const bool nr = ok[i] && !far[i];
                const float rest = (ok[i] && far[i]) ? one : 0.0f;
                o0[h + i] = nr ? p0[i] : rest;
                o1[h + i] = nr ? p1[i] : rest;
                o2[h + i] = nr ? p2[i] : rest;
            }
        }
        float* p = out + ((size_t)(s * R + m) * R + f0);
        st_f4(p, make_float4(o0[0], o0[1], o0[2], o0[3]));
        st_f4(p + CH, make_float4(o1[0], o1[1], o1[2], o1[3]));
        st_f4(p + 2 * CH, make_float4(o2[0], o2[1], o2[2], o2[3]));
    }
}

// One frame's voxel loop (both passes).  The warps run independently.  Warp w handles the
// middle-axis positions m = w, w + 8, ... of every slow-axis row group, so every warp sees every
// part of the volume (balanced pass-2 work).  Pass 2 runs at the end of each row group on whole
// batches of 32 listed groups; the remainder moves to the front of the list.
template <int R, int LAYOUT, int NW, bool STAGED, bool FAST, bool DEFF>
__device__ __forceinline__ void voxel_passes(const float* __restrict__ crop,
                                             FrameShared<R, LAYOUT, NW>* __restrict__ sh, const VoxConst& k,
                                             const BatchArgs& a, float yf, float* __restrict__ out,
                                             int part, int nparts) {
    using LM = LaneMap<R, NW>;
    constexpr int YS = FrameShared<R, LAYOUT, NW>::YS;
    constexpr int CH = R * R * R;
    static_assert(!LM::BLOCKED, "the block map has its own sweep (voxel_passes_blocked)");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool active = lane < LM::ACTIVE;
    const int fg = lane % LM::LPR;
    const int sr = active ? lane / LM::LPR : LM::ROWS - 1;
    const int f0 = 4 * fg;
    const float t_far = k.t_far;
    unsigned short* list = sh->list[warp];
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0;

    float4 vzf = make_float4(0.f, 0.f, 0.f, 0.f);
    if (LAYOUT == 0) vzf = *reinterpret_cast<const float4*>(&sh->vz[f0]);

#pragma unroll 1
    for (int sg = part * (LM::NSG / nparts), sg_end = sg + LM::NSG / nparts; sg < sg_end; ++sg) {
        const int s = sg * LM::ROWS + sr;
        const int4 A = *reinterpret_cast<const int4*>(&sh->at[s * R + f0]);
        if (LAYOUT != 0) {
            const float z1 = sh->vz[s];
            vzf = make_float4(z1, z1, z1, z1);
        }
        float* const obase = out + ((size_t)s * R * R + f0);
        const int ebase = (sg << 11) | lane;
#pragma unroll
        for (int ks = 0; ks < LM::MSTEPS; ks += 2) {
            const int m0 = warp + NW * ks, m1 = m0 + NW;
            int4 B0, B1;
            if (LAYOUT == 0) {
                B0 = *reinterpret_cast<const int4*>(&sh->yt[m0 * YS + f0]);
                B1 = *reinterpret_cast<const int4*>(&sh->yt[m1 * YS + f0]);
            } else {
                const int y0 = sh->yt[s * YS + m0], y1 = sh->yt[s * YS + m1];
                B0 = make_int4(y0, y0, y0, y0);
                B1 = make_int4(y1, y1, y1, y1);
            }
            const float d00 = gather_depth<STAGED>(crop, A.x + B0.x);
            const float d01 = gather_depth<STAGED>(crop, A.y + B0.y);
            const float d02 = gather_depth<STAGED>(crop, A.z + B0.z);
            const float d03 = gather_depth<STAGED>(crop, A.w + B0.w);
            const float d10 = gather_depth<STAGED>(crop, A.x + B1.x);
            const float d11 = gather_depth<STAGED>(crop, A.y + B1.y);
            const float d12 = gather_depth<STAGED>(crop, A.z + B1.z);
            const float d13 = gather_depth<STAGED>(crop, A.w + B1.w);
            float4 v0, v1;
            bool n0 = false, n1 = false;
            classify<STAGED>(d00, vzf.x, t_far, v0.x, n0);
            classify<STAGED>(d01, vzf.y, t_far, v0.y, n0);
            classify<STAGED>(d02, vzf.z, t_far, v0.z, n0);
            classify<STAGED>(d03, vzf.w, t_far, v0.w, n0);
            classify<STAGED>(d10, vzf.x, t_far, v1.x, n1);
            classify<STAGED>(d11, vzf.y, t_far, v1.y, n1);
            classify<STAGED>(d12, vzf.z, t_far, v1.z, n1);
            classify<STAGED>(d13, vzf.w, t_far, v1.w, n1);
            // Sectors (32 bytes = the groups of an even/odd lane pair) with near voxels are written
            // by pass 2 only, whole: every byte crosses the SM -> L2 write path once and no sector
            // is ever written in halves (a half-written sector costs L2 a fill from DRAM).
            unsigned bal0 = __ballot_sync(FULL, n0 && active);
            unsigned bal1 = __ballot_sync(FULL, n1 && active);
            bal0 |= ((bal0 & 0x55555555u) << 1) | ((bal0 & 0xAAAAAAAAu) >> 1);
            bal1 |= ((bal1 & 0x55555555u) << 1) | ((bal1 & 0xAAAAAAAAu) >> 1);
            n0 = (bal0 >> lane) & 1u;
            n1 = (bal1 >> lane) & 1u;
            if (active && !n0) {
                float* p0 = obase + m0 * R;
                st_f4(p0, v0); st_f4(p0 + CH, v0); st_f4(p0 + 2 * CH, v0);
            }
            if (active && !n1) {
                float* p1 = obase + m1 * R;
                st_f4(p1, v1); st_f4(p1 + CH, v1); st_f4(p1 + 2 * CH, v1);
            }
            if (n0) list[cnt + __popc(bal0 & lt)] = (unsigned short)(ebase | (m0 << 5));
            cnt += __popc(bal0);
            if (n1) list[cnt + __popc(bal1 & lt)] = (unsigned short)(ebase | (m1 << 5));
            cnt += __popc(bal1);
        }
        // ---- pass 2 on whole batches (warp-uniform control flow)
        if (cnt >= 32) {
            __syncwarp();                 // list entries and the pass-1 stores are ordered before pass 2
            int done = 0;
            do {
                eval_listed<R, LAYOUT, NW, STAGED, FAST, DEFF>(crop, sh, k, a, yf, out, list + done, 32);
                done += 32;
            } while (cnt - done >= 32);
            const int rem = cnt - done;
            unsigned short carry = 0;
            if (lane < rem) carry = list[done + lane];
            __syncwarp();
            if (lane < rem) list[lane] = carry;
            cnt = rem;
        }
    }
    __syncwarp();
    eval_listed<R, LAYOUT, NW, STAGED, FAST, DEFF>(crop, sh, k, a, yf, out, list, cnt);
}

// The same two passes with the block map (LaneMap::BLOCKED, R = 48).  Warp w takes the units
// u = w, w + NW, ... with u = s * NMB + (block of 8 middle-axis rows); in step t of a unit lane l owns
// group 32 t + l of the block: row (32 t + l) / LPR, F-group (32 t + l) % LPR - fixed per lane and step,
// so the table rows at[s][f0..f0+3] are loaded once per unit and step.  Pass 2 runs at the end of a unit.
template <int R, int LAYOUT, int NW, bool STAGED, bool FAST, bool DEFF>
__device__ __forceinline__ void voxel_passes_blocked(const float* __restrict__ crop,
                                                     FrameShared<R, LAYOUT, NW>* __restrict__ sh, const VoxConst& k,
                                                     const BatchArgs& a, float yf, float* __restrict__ out,
                                                     int part, int nparts) {
    using LM = LaneMap<R, NW>;
    constexpr int YS = FrameShared<R, LAYOUT, NW>::YS;
    constexpr int CH = R * R * R;
    constexpr int T = LM::T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float t_far = k.t_far;
    unsigned short* list = sh->list[warp];
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0;
    int rt[T], f0t[T];
    float4 vzt[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int gi = 32 * t + lane;
        rt[t] = gi / LM::LPR;
        f0t[t] = 4 * (gi % LM::LPR);
        vzt[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (LAYOUT == 0) vzt[t] = *reinterpret_cast<const float4*>(&sh->vz[f0t[t]]);
    }
    const int u_per_part = (R / nparts) * LM::NMB;       // whole slow-axis rows per part
#pragma unroll 1
    for (int u = part * u_per_part + warp, u_end = (part + 1) * u_per_part; u < u_end; u += NW) {
        const int s = u / LM::NMB, m_base = (u - s * LM::NMB) * LM::MB;
        float* const obase = out + (size_t)s * R * R;
        float zs = 0.0f;
        if (LAYOUT != 0) zs = sh->vz[s];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int m = m_base + rt[t], f0 = f0t[t];
            const int4 A = *reinterpret_cast<const int4*>(&sh->at[s * R + f0]);
            int4 B;
            float4 vzf;
            if (LAYOUT == 0) {
                B = *reinterpret_cast<const int4*>(&sh->yt[m * YS + f0]);
                vzf = vzt[t];
            } else {
                const int y0 = sh->yt[s * YS + m];
                B = make_int4(y0, y0, y0, y0);
                vzf = make_float4(zs, zs, zs, zs);
            }
            const float d0 = gather_depth<STAGED>(crop, A.x + B.x);
            const float d1 = gather_depth<STAGED>(crop, A.y + B.y);
            const float d2 = gather_depth<STAGED>(crop, A.z + B.z);
            const float d3 = gather_depth<STAGED>(crop, A.w + B.w);
            float4 v;
            bool n = false;
            classify<STAGED>(d0, vzf.x, t_far, v.x, n);
            classify<STAGED>(d1, vzf.y, t_far, v.y, n);
            classify<STAGED>(d2, vzf.z, t_far, v.z, n);
            classify<STAGED>(d3, vzf.w, t_far, v.w, n);
            // sectors (even/odd lane pairs: LPR is even, a pair never straddles rows) with near voxels
            // are written by pass 2 only, whole
            unsigned bal = __ballot_sync(FULL, n);
            bal |= ((bal & 0x55555555u) << 1) | ((bal & 0xAAAAAAAAu) >> 1);
            n = (bal >> lane) & 1u;
            if (!n) {
                float* p = obase + m * R + f0;
                st_f4(p, v); st_f4(p + CH, v); st_f4(p + 2 * CH, v);
            } else {
                list[cnt + __popc(bal & lt)] = (unsigned short)((s << 10) | (m << 4) | (f0 >> 2));
            }
            cnt += __popc(bal);
        }
        if (cnt >= 32) {
            __syncwarp();
            int done = 0;
            do {
                eval_listed<R, LAYOUT, NW, STAGED, FAST, DEFF>(crop, sh, k, a, yf, out, list + done, 32);
                done += 32;
            } while (cnt - done >= 32);
            const int rem = cnt - done;
            unsigned short carry = 0;
            if (lane < rem) carry = list[done + lane];
            __syncwarp();
            if (lane < rem) list[lane] = carry;
            cnt = rem;
        }
    }
    __syncwarp();
    eval_listed<R, LAYOUT, NW, STAGED, FAST, DEFF>(crop, sh, k, a, yf, out, list, cnt);
}

// ---------------------------------------------------------------- fused kernel
// CTA shape: 8 warps, up to three CTAs per SM (80 registers, no spills).  16 warps x 2 CTAs/SM (64
// registers) was measured as well: the same speed as 8 warps x 2 CTAs/SM - what counts is how many
// frames are in flight, see launch_fused.
// SPLIT: CTAs per frame.  A frame can be cut in SPLIT slabs along the slow axis, one CTA each (adjacent
// block indices: the second crop load hits L2); every CTA redoes the cheap prologue (crop, bounding box,
// tables), nothing is exchanged.  Measured per 4,000 frames: R = 48: 0.929 ms unsplit, 0.903 ms with
// two or three CTAs per frame (kept: 2); R = 64: 2.04 ms unsplit at one CTA per SM, 2.22 ms with two
// CTAs per frame, 2.10 ms with four (kept: 1 - fewer concurrent write streams win there); R = 32 (per
// 8,500 frames): 0.616 ms unsplit at two CTAs per SM, 0.619 ms with two CTAs per frame at three per SM (kept: 1).
template <int R, int LAYOUT>
struct CtaShape {
    static constexpr int NW = 8;
    static constexpr int MIN_CTAS = 3;
    static constexpr int SPLIT = 1;
};

// R = 48 (block map): 288 units of 8 middle-axis rows per frame, 36 per warp
template <int LAYOUT>
struct CtaShape<48, LAYOUT> {
    static constexpr int NW = 8;
    static constexpr int MIN_CTAS = 2;
    static constexpr int SPLIT = 2;
};

// SAMPLED: the instantiation that serves DTSDF_BBOX_SAMPLED (and only that mode).  A kernel of its own so that the
// other modes compile exactly as they would without it - the extra live values and out-of-line calls of the
// sampled prologue cost the R = 64 kernel 8 % through spills when they shared one.
template <int R, int LAYOUT, bool DEFF, bool SAMPLED>
__global__ void __launch_bounds__(32 * CtaShape<R, LAYOUT>::NW, CtaShape<R, LAYOUT>::MIN_CTAS)
dtsdf_fused_kernel(const __grid_constant__ BatchArgs a) {
    constexpr int NW = CtaShape<R, LAYOUT>::NW;
    using FS = FrameShared<R, LAYOUT, NW>;
    constexpr int YS = FS::YS;
    static_assert(sizeof(FS) % 16 == 0 && offsetof(FS, nodepth) + 16 == sizeof(FS), "the pad must touch the crop");
    static_assert(LaneMap<R, NW>::BLOCKED || LaneMap<R, NW>::MSTEPS % 2 == 0, "the sweep is unrolled by two");
    static_assert(LaneMap<R, NW>::NSG % CtaShape<R, LAYOUT>::SPLIT == 0 && R % CtaShape<R, LAYOUT>::SPLIT == 0,
                  "a frame splits into whole row groups");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FS* sh = reinterpret_cast<FS*>(smem_raw);
    float* crop_s = reinterpret_cast<float*>(smem_raw + sizeof(FS));
    constexpr int SPLIT = CtaShape<R, LAYOUT>::SPLIT;
    const int tid = threadIdx.x;
    const int f = blockIdx.x / SPLIT, part = blockIdx.x % SPLIT;
    const bool writer = part == 0;                    // max_l, mid_p, joints: once per frame

    // ---- 0. frame record; thread 0 starts the TMA bulk copy of the crop into shared memory
    const Frame fr = load_frame(a, f);
    // DTSDF_BBOX_SAMPLED: the frame's index vector rides behind the crop on the same barrier when both fit
    const int* idx_g = SAMPLED ? a.sample_idx + (size_t)f * a.point_num : nullptr;
    const bool idx_staged = SAMPLED && fr.staged && (a.point_num & 3) == 0 &&
                            fr.n_pad + a.point_num <= a.cap_floats && (reinterpret_cast<uintptr_t>(idx_g) & 15) == 0;
    if (tid == 0) {
        mbar_init(&sh->bar, 1);
        fence_mbar_init();
        if (fr.staged) {
            mbar_expect_tx(&sh->bar, (uint32_t)fr.n_pad * 4u + (idx_staged ? (uint32_t)a.point_num * 4u : 0u));
            bulk_g2s(crop_s, fr.crop, (uint32_t)fr.n_pad * 4u, &sh->bar);
            if (idx_staged) bulk_g2s(crop_s + fr.n_pad, idx_g, (uint32_t)a.point_num * 4u, &sh->bar);
        }
        sh->nanflag = 0u;
        sh->nodepth[0] = sh->nodepth[1] = sh->nodepth[2] = sh->nodepth[3] = __int_as_float(QNAN_BITS);
    }
    const float yf = __frcp_rn(a.focal);
    const bool focal_ok = a.focal >= 0x1p-30f && a.focal <= 0x1p30f;
    const bool scan = !SAMPLED && a.bbox_mode == 0;  // DTSDF_BBOX_PIXELS
    constexpr bool sampled = SAMPLED;                // DTSDF_BBOX_SAMPLED
    constexpr unsigned SCRATCH_BITS = (unsigned)((offsetof(FS, vx) - offsetof(SampledScratch<NW>, bitmap)) * 8);
    static_assert(offsetof(FS, at) == 0 && offsetof(FS, vx) > offsetof(SampledScratch<NW>, bitmap) + 1024, "scratch overlays the tables");
    const bool scan_exact = !(DEFF || focal_ok);     // no proven fixed-divisor form for this focal
    // the common shape of a sampled frame: per-column / per-row extremes and one bitmap window (prepared now,
    // while the crop is on its way)
    const bool sampled_fast = sampled && fr.n + 64 <= (int)SCRATCH_BITS && fr.bw < (1 << 20) && fr.bh < (1 << 20) &&
                              abs(fr.l) < (1 << 20) && abs(fr.t) < (1 << 20);          // (u, v exact in float32)
    if constexpr (SAMPLED) {
        if (sampled_fast) sampled_scratch_init<NW>(smem_raw, fr.n);
    }
    __syncthreads();                                  // the mbarrier is initialised

    // ---- 1. bounding-box scan (pre/tsdf_numba.py:75-116) of the staged crop, which also turns
    //         "no depth" into NaN in place; crops too large for shared memory are scanned in
    //         global memory and left alone
    bool sawnan = false;
    if (fr.staged) {
        mbar_wait(&sh->bar, 0);
        if constexpr (SAMPLED) {   // (ends with a barrier: every thread is done reading the crop before it is sanitised)
            const int iw = __ldg(a.headers + 6 * (size_t)f), ih = __ldg(a.headers + 6 * (size_t)f + 1);
            const int* idx = idx_staged ? reinterpret_cast<const int*>(crop_s + fr.n_pad) : idx_g;
            bool done = false;
            if (sampled_fast)
                done = sampled_bbox_fast<NW, true, (int)sizeof(FS), (int)offsetof(FS, red)>(
                    nullptr, fr.n, fr.bw, fr.bh, fr.l, fr.t, iw, ih, a.focal64, idx, a.point_num);
            if (!done)
                sampled_bbox_general<NW, true, (int)sizeof(FS), (int)offsetof(FS, red), (int)SCRATCH_BITS>(
                    nullptr, fr.n, fr.bw, fr.l, fr.t, iw, ih, a.focal64, idx, a.point_num);
        }
        if (!scan)           sawnan = sanitize_crop<NW>(crop_s, fr.n_pad);
        else if (scan_exact) sawnan = bbox_partial<DEFF, true, true, NW>(fr, crop_s, a, yf, sh->red);
        else                 sawnan = bbox_partial<DEFF, false, true, NW>(fr, crop_s, a, yf, sh->red);
    } else if (scan) {
        if (scan_exact) bbox_partial<DEFF, true, false, NW>(fr, const_cast<float*>(fr.crop), a, yf, sh->red);
        else            bbox_partial<DEFF, false, false, NW>(fr, const_cast<float*>(fr.crop), a, yf, sh->red);
    } else if constexpr (SAMPLED) {
        const int iw = __ldg(a.headers + 6 * (size_t)f), ih = __ldg(a.headers + 6 * (size_t)f + 1);
        const int* idx = idx_g;
        bool done = false;
        if (sampled_fast)
            done = sampled_bbox_fast<NW, false, (int)sizeof(FS), (int)offsetof(FS, red)>(
                fr.crop, fr.n, fr.bw, fr.bh, fr.l, fr.t, iw, ih, a.focal64, idx, a.point_num);
        if (!done)
            sampled_bbox_general<NW, false, (int)sizeof(FS), (int)offsetof(FS, red), (int)SCRATCH_BITS>(
                fr.crop, fr.n, fr.bw, fr.l, fr.t, iw, ih, a.focal64, idx, a.point_num);
    }
    if (sawnan) sh->nanflag = 1u;
    __syncthreads();      // (A) bbox partials, sanitised crop and the NaN flag are in

    // ---- 2. bounding box -> grid parameters
    float ox, oy, oz, voxel_len, trunc;
    float jmid0, jmid1, jmid2, jmax_l;      // what the joint normalisation uses
    if (a.bbox_mode == 2) {       // DTSDF_BBOX_GRID: pre/process.py:122 arguments
        const float* g = a.aux + 5 * (size_t)f;
        ox = __ldg(g); oy = __ldg(g + 1); oz = __ldg(g + 2);
        voxel_len = __ldg(g + 3); trunc = __ldg(g + 4);
        // centre of the cube: vox_ori + (R-1)/2 * voxel_len (informational only)
        const float hl = __fmul_rn(voxel_len, 0.5f * (float)(R - 1));
        jmax_l = __fmul_rn(voxel_len, (float)R);
        jmid0 = __fadd_rn(ox, hl); jmid1 = __fadd_rn(oy, hl); jmid2 = __fadd_rn(oz, hl);
        if (tid == 0 && writer) {
            if (a.max_l) a.max_l[f] = jmax_l;
            if (a.mid_p) {
                a.mid_p[3 * (size_t)f + 0] = jmid0;
                a.mid_p[3 * (size_t)f + 1] = jmid1;
                a.mid_p[3 * (size_t)f + 2] = jmid2;
            }
        }
    } else {
        float mn[3], mx[3];
        if (scan) {
            bbox_reduce<NW>(sh->red, mn, mx);
            // the fixed-divisor scan is exact for |d| < 2^40 only: rescan with IEEE divisions
            // if the frame holds anything larger (block-uniform, never on real data; the
            // sanitised crop scans like the original)
            if (!scan_exact && !(fabsf(mn[2]) < BIG_DEPTH && fabsf(mx[2]) < BIG_DEPTH)) {
                __syncthreads();
                if (fr.staged) bbox_partial<DEFF, true, true, NW>(fr, crop_s, a, yf, sh->red);
                else           bbox_partial<DEFF, true, false, NW>(fr, const_cast<float*>(fr.crop), a, yf, sh->red);
                __syncthreads();
                bbox_reduce<NW>(sh->red, mn, mx);
            }
        } else if (sampled) {     // DTSDF_BBOX_SAMPLED: the same six numbers, just reduced in shared memory
            const float* g = &sh->red[0][0];
            mx[0] = g[0]; mx[1] = g[1]; mx[2] = g[2];
            mn[0] = g[3]; mn[1] = g[4]; mn[2] = g[5];
        } else {                  // DTSDF_BBOX_EXTREMA: {max xyz, min xyz}, pre/process.py:120
            const float* g = a.aux + 6 * (size_t)f;
            mx[0] = __ldg(g); mx[1] = __ldg(g + 1); mx[2] = __ldg(g + 2);
            mn[0] = __ldg(g + 3); mn[1] = __ldg(g + 4); mn[2] = __ldg(g + 5);
        }
        // grid parameters (pre/process.py:91-96 == pre/tsdf_numba.py:142-147)
        const float mid0 = __fmul_rn(__fadd_rn(mx[0], mn[0]), 0.5f);
        const float mid1 = __fmul_rn(__fadd_rn(mx[1], mn[1]), 0.5f);
        const float mid2 = __fmul_rn(__fadd_rn(mx[2], mn[2]), 0.5f);
        const float max_l = fmaxf(fmaxf(__fsub_rn(mx[0], mn[0]), __fsub_rn(mx[1], mn[1])),
                                  __fsub_rn(mx[2], mn[2]));
        voxel_len = __fdiv_rn(max_l, (float)R);
        trunc = __fmul_rn(voxel_len, a.trunc_mul);
        const float half_l = __fmul_rn(max_l, 0.5f);
        const float half_v = __fmul_rn(voxel_len, 0.5f);
        ox = __fadd_rn(__fsub_rn(mid0, half_l), half_v);
        oy = __fadd_rn(__fsub_rn(mid1, half_l), half_v);
        oz = __fadd_rn(__fsub_rn(mid2, half_l), half_v);
        jmax_l = max_l; jmid0 = mid0; jmid1 = mid1; jmid2 = mid2;
        if (tid == 0 && writer) {
            if (a.max_l) a.max_l[f] = max_l;
            if (a.mid_p) {
                a.mid_p[3 * (size_t)f + 0] = mid0;
                a.mid_p[3 * (size_t)f + 1] = mid1;
                a.mid_p[3 * (size_t)f + 2] = mid2;
            }
        }
    }

    // ---- joint normalisation in the same launch: (gt - mid_p)/max_l + 0.5 in float32
    //      (pre/joint_nor.py:8-18; 3D_CNN/train.py:236-244 with the clamp)
    if (a.joints != nullptr && tid < 63 && writer) {
        const int c = tid % 3;
        float g = __ldg(a.joints + 63 * (size_t)f + tid);
        if ((a.joint_flags & 2) && c == 2) g = -g;
        const float m = c == 0 ? jmid0 : (c == 1 ? jmid1 : jmid2);
        float v = __fadd_rn(__fdiv_rn(__fsub_rn(g, m), jmax_l), 0.5f);
        if (a.joint_flags & 1) v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
        a.out_joints[63 * (size_t)f + tid] = v;
    }

    // ---- per-axis voxel centres and the two projection tables
    //      (pre/process.py:150-159 == pre/tsdf_numba.py:26-32,36)
    if (tid < R) {
        const float sft = __fmul_rn((float)tid, voxel_len);
        sh->vx[tid] = __fadd_rn(ox, sft);
        sh->vy[tid] = __fadd_rn(oy, sft);
        sh->vz[tid] = __fadd_rn(oz, sft);
    }
    {
        // entry e: zi = e % R is the z index, ti = e / R the x (at) / y (yt) index.  With
        // (32 * NW) a multiple of R the z index, and with it the division, is the same for every
        // entry of a thread.
        constexpr bool HOIST = ((32 * NW) % R) == 0;
        float coeff_h = 0.0f;
        if (HOIST) coeff_h = __fdiv_rn(-a.focal, __fadd_rn(oz, __fmul_rn((float)(tid % R), voxel_len)));
#pragma unroll 2
        for (int e = tid; e < R * R; e += (32 * NW)) {
            const int ti = e / R, zi = e % R;
            const float coeff = HOIST ? coeff_h
                                      : __fdiv_rn(-a.focal, __fadd_rn(oz, __fmul_rn((float)zi, voxel_len)));
            const float sft = __fmul_rn((float)ti, voxel_len);
            const float fpx = __fadd_rn(__fmul_rn(__fadd_rn(ox, sft), coeff), (float)a.center_x);
            const float fpy = __fadd_rn(__fmul_rn(-__fadd_rn(oy, sft), coeff), (float)a.center_y);
            const int px = __float2int_rz(fpx);          // int(): truncation toward zero
            const int py = __float2int_rz(fpy);
            const bool okx = fabsf(fpx) < 2e9f && px >= fr.l && px < fr.r;   // non-finite -> outside
            const bool oky = fabsf(fpy) < 2e9f && py >= fr.t && py < fr.b;   // (bad header: empty ranges)
            // at[s][f]: CXYZ s = x, f = z; CZYX s = z, f = x.   yt: CXYZ [y][z]; CZYX [z][y]
            const int ia = (LAYOUT == 0) ? ti * R + zi : zi * R + ti;
            const int iy = (LAYOUT == 0) ? ti * YS + zi : zi * YS + ti;
            sh->at[ia] = okx ? 4 * (px - fr.l) : OUTSIDE;
            sh->yt[iy] = oky ? 4 * fr.bw * (py - fr.t) : OUTSIDE;
            sh->yr[iy] = oky ? (short)(py - fr.t) : (short)0;
        }
    }

    // the fixed-divisor divisions are proven for these ranges (file header); anything else
    // takes the generic IEEE path
    const float ax = fabsf(ox), ay = fabsf(oy), az = fabsf(oz);
    const bool fast = (DEFF || focal_ok) && trunc >= 0x1p-30f && trunc <= 0x1p30f &&
                      voxel_len >= 0x1p-30f && voxel_len <= 0x1p30f &&
                      (ax == 0.0f || (ax >= 0x1p-60f && ax <= 0x1p30f)) &&
                      (ay == 0.0f || (ay >= 0x1p-60f && ay <= 0x1p30f)) &&
                      (az == 0.0f || (az >= 0x1p-60f && az <= 0x1p30f));
    VoxConst k;
    k.ix = MAGIC_I + fr.l;
    k.iy = MAGIC_I + fr.t;
    k.trunc = trunc;
    k.yt = __frcp_rn(trunc);
    // margin 2^-10 on the far test: t_z = RN(|s_z|/trunc) > 1 + 2^-11, t_z^2 > THR_SQ
    k.t_far = fast ? __fmul_rn(trunc, 1.0009765625f) : INFINITY;
    float* out = a.vol + (size_t)f * 3 * R * R * R;

    __syncthreads();      // (C) tables are complete
    // a staged crop that holds NaNs of its own was sanitised into ambiguity: read the original
    const bool staged = fr.staged && sh->nanflag == 0u;

    // ---- 3. voxel loop
    if constexpr (LaneMap<R, NW>::BLOCKED) {
        if (fast) {
            if (staged) voxel_passes_blocked<R, LAYOUT, NW, true, true, DEFF>(crop_s, sh, k, a, yf, out, part, SPLIT);
            else        voxel_passes_blocked<R, LAYOUT, NW, false, true, DEFF>(fr.crop, sh, k, a, yf, out, part, SPLIT);
        } else {
            if (staged) voxel_passes_blocked<R, LAYOUT, NW, true, false, DEFF>(crop_s, sh, k, a, yf, out, part, SPLIT);
            else        voxel_passes_blocked<R, LAYOUT, NW, false, false, DEFF>(fr.crop, sh, k, a, yf, out, part, SPLIT);
        }
    } else {
        if (fast) {
            if (staged) voxel_passes<R, LAYOUT, NW, true, true, DEFF>(crop_s, sh, k, a, yf, out, part, SPLIT);
            else        voxel_passes<R, LAYOUT, NW, false, true, DEFF>(fr.crop, sh, k, a, yf, out, part, SPLIT);
        } else {
            if (staged) voxel_passes<R, LAYOUT, NW, true, false, DEFF>(crop_s, sh, k, a, yf, out, part, SPLIT);
            else        voxel_passes<R, LAYOUT, NW, false, false, DEFF>(fr.crop, sh, k, a, yf, out, part, SPLIT);
        }
    }
}

template <int R, int LAYOUT, bool DEFF, bool SAMPLED = false>
cudaError_t launch_fused(const BatchArgs& args_in, int max_crop_px, cudaStream_t stream) {
    if constexpr (!SAMPLED) {
        if (args_in.bbox_mode == 3) return launch_fused<R, LAYOUT, DEFF, true>(args_in, max_crop_px, stream);
    }
    BatchArgs args = args_in;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int smem_optin = 0;
    e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    if ((reinterpret_cast<uintptr_t>(args.vol) & 15) != 0) return cudaErrorMisalignedAddress;
    if ((reinterpret_cast<uintptr_t>(args.depth) & 15) != 0) return cudaErrorMisalignedAddress;
    constexpr int NW = CtaShape<R, LAYOUT>::NW;
    const int fixed = (int)sizeof(FrameShared<R, LAYOUT, NW>);
    // staging capacity: the largest crop of the batch when known (the occupancy then follows from
    // it: three CTAs per SM up to ~63 KB crops, two up to ~100 KB, one above); oversize crops take the direct path.
    long long want = max_crop_px > 0 ? (((long long)max_crop_px + 3) & ~3ll) * 4 : (long long)smem_optin;
    if (args.bbox_mode == 3 && max_crop_px > 0) want += (((long long)args.point_num + 3) & ~3ll) * 4;   // the index vector
    long long avail = (long long)smem_optin - fixed - 1024;
    if (want > avail) want = avail;
    if (want < 16) want = 16;
    want &= ~15ll;
    args.cap_floats = (int)(want / 4);
    size_t smem = (size_t)fixed + (size_t)want;
    auto kern = dtsdf_fused_kernel<R, LAYOUT, DEFF, SAMPLED>;
    static std::atomic<unsigned> smem_ok{0u};
    e = allow_max_smem(kern, dev, smem_optin, smem_ok);
    if (e != cudaSuccess) return e;
    // Frames in flight.  Every resident CTA has a volume (12 R^3 bytes) under construction; the stores
    // of all of them sit in L2 until they are written back.  While that footprint fits in L2 the write
    // stream reaches DRAM more efficiently (fewer concurrent streams, longer runs - the exact mechanism
    // was not isolated; the policy rests on the measurement): for R = 32 on B200 (126 MB L2),
    // 2 CTAs/SM = 296 volumes = 116 MB -> 0.618 ms per 8,500 frames, 3 CTAs/SM = 174 MB -> 0.642 ms,
    // 1 CTA/SM -> 0.78 ms (nothing left to hide a frame's prologue).  So: at most as many CTAs per SM as
    // keep the footprint inside L2, but never fewer than two; the limit is imposed by padding the
    // dynamic shared memory request.  Where not even two volumes per SM fit (R = 48, 64) the occupancy
    // is whatever shared memory allows, except at R = 64: a frame's voxel loop is then so long (3.1 MB
    // of stores) that its prologue no longer needs hiding and ONE CTA per SM is fastest - 4,000
    // frames: 2.04 ms vs 2.26-2.30 ms with two or three (16 warps x 1: 2.14 ms; R = 48: 0.99 ms with
    // one, two or three alike).
    {
        int occ = 0, sms = 0, l2 = 0, smem_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * NW, smem);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        if (e != cudaSuccess) return e;
        const long long vol_bytes = 12ll * R * R * R;
        const int fit = (int)((long long)l2 / ((long long)sms * vol_bytes));
        constexpr int SPLIT = CtaShape<R, LAYOUT>::SPLIT;
        // (with SPLIT CTAs per frame, SPLIT CTAs per SM are one volume per SM under construction)
        const int target = fit >= 2 ? std::min(occ, fit * SPLIT) : (R >= 64 ? std::min(occ, SPLIT) : occ);
        if (target >= 1 && target < occ) {
            // the largest request that still lets `target` CTAs share an SM (1 KB per CTA is reserved)
            long long padded = ((long long)smem_sm / target - 1024) & ~15ll;
            if (padded > smem_optin) padded = smem_optin;
            if (padded > (long long)smem) smem = (size_t)padded;
        }
        args.cap_floats = (int)((smem - (size_t)fixed) / 4) & ~3;      // the padding stages larger crops for free
    }
    kern<<<(unsigned)args.n_frames * CtaShape<R, LAYOUT>::SPLIT, (32 * NW), smem, stream>>>(args);   // SPLIT CTAs per frame
    ++g_launch_count;
    return cudaGetLastError();
}

// ---------------------------------------------------------------- valid-pixel count
// n_valid[f] = pixels with d != 0 (the rows pre/process.py:61-66 keeps): what the caller needs to draw the
// resampling indices (pre/process.py:70-84) before the sampled launch.  One CTA per frame, 128-bit loads.
__global__ void __launch_bounds__(NT) count_valid_kernel(const float* __restrict__ depth,
                                                         const long long* __restrict__ offsets,
                                                         const int* __restrict__ headers, int* __restrict__ n_valid) {
    __shared__ int wsum[NWARP];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* h = headers + 6 * (size_t)f;
    const int bw = h[4] - h[2], bh = h[5] - h[3];
    const int n = (bw > 0 && bh > 0) ? bw * bh : 0;
    const float* crop = depth + offsets[f];                  // 16-byte aligned (packed batch format)
    int cnt = 0;
    for (int i = tid; i < (n >> 2); i += NT) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(crop) + i);
        cnt += (v.x != 0.0f) + (v.y != 0.0f) + (v.z != 0.0f) + (v.w != 0.0f);
    }
    if (tid < (n & 3)) cnt += __ldg(crop + (n & ~3) + tid) != 0.0f;
    cnt = __reduce_add_sync(FULL, cnt);
    if (lane == 0) wsum[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        int s = 0;
#pragma unroll
        for (int i = 0; i < NWARP; ++i) s += wsum[i];
        n_valid[f] = s;
    }
}

// ---------------------------------------------------------------- point cloud (f64)
// a / b in float64 for a per-launch divisor b with y = rcp_rn(b): q = a y and two residual corrections
// q += fma(-b, q, a) y.  The first brings q within half an ulp (and a hair) of a / b whatever the rounding of a y
// did; for a faithful q and a correctly rounded reciprocal the correction is the correctly rounded quotient
// (Markstein 1990; the one divisor class the theorem excepts - a significand of all ones - takes the IEEE path,
// like every dividend outside 2^-900 .. 2^900, zeros included: their residuals would lose the sign / overflow).
// tests/native/ddivision_check.c runs the same five operations against the host's IEEE division on 2 x 10^8
// dividends, half of them built to put the quotient within a few units of 2^-53 ulp of a rounding boundary.
__device__ __forceinline__ double ddiv_fixed(double a, double b, double y, bool fast_b) {
    const double m = fabs(a);
    if (!(fast_b && m >= 0x1p-900 && m <= 0x1p900)) return __ddiv_rn(a, b);
    double q = __dmul_rn(a, y);
    q = __fma_rn(__fma_rn(-b, q, a), y, q);
    q = __fma_rn(__fma_rn(-b, q, a), y, q);
    return q;
}

// pre/process.py:30-68: X = ((u+left) - w/2)*d / f, Y = -(((v+top) - h/2)*d) / f, Z = -d in
// float64, keeping the pixels with d != 0 in row-major order.  One CTA per frame.  Every warp owns a contiguous
// piece of the crop: it counts its valid pixels, ONE block barrier turns the eight counts into output offsets, and
// the warps then write their pieces on their own (ballot + popc per 32 pixels) - round 1 synchronised the block
// three times per 256 pixels and paid two IEEE float64 divisions per point (0.95 TB/s).  The crop is read twice
// (count, then write): a crop of up to PC_STAGE_FLOATS pixels is staged in shared memory by ONE TMA bulk copy (three
// CTAs per SM) and both passes read it there; larger crops are read from global memory, eight steps of loads in flight.
constexpr int PC_STAGE_FLOATS = 18 * 1024;          // 72 KB: the MSRA crops (<= 150 x 140) almost always fit
__global__ void __launch_bounds__(NT) point_cloud_kernel(int n_frames, double focal,
                                                         const float* __restrict__ depth,
                                                         const long long* __restrict__ offsets,
                                                         const int* __restrict__ headers,
                                                         double* __restrict__ points,
                                                         const long long* __restrict__ point_offsets,
                                                         int* __restrict__ n_valid) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int warp_cnt[NWARP];
    __shared__ unsigned long long bar;
    const int f = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* h = headers + 6 * (size_t)f;
    const int w = h[0], hh = h[1], l = h[2], t = h[3], r = h[4], b = h[5];
    const int bw = r - l, bh = b - t;
    const int n = (bw > 0 && bh > 0) ? bw * bh : 0;
    const float* crop_g = depth + offsets[f];                // 16-byte aligned, padded to 4 floats (packed batch)
    const int n_pad = (n + 3) & ~3;
    const bool staged = n > 0 && n_pad <= PC_STAGE_FLOATS;   // (CTA-uniform)
    if (staged) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
            mbar_expect_tx(&bar, (uint32_t)n_pad * 4u);
            bulk_g2s(smem_raw, crop_g, (uint32_t)n_pad * 4u, &bar);
        }
        __syncthreads();
        mbar_wait(&bar, 0);
    }
    // (reads through the generic pointer: explicit shared-space loads were measured here, 0.646 vs 0.578 ms)
    const float* crop = staged ? reinterpret_cast<const float*>(smem_raw) : crop_g;
    double* out = points ? points + 3 * point_offsets[f] : nullptr;
    // the warp's piece: steps of 32 pixels [s0, s1)
    const int steps = (n + 31) >> 5, per_warp = (steps + NWARP - 1) / NWARP;
    const int s0 = min(warp * per_warp, steps), s1 = min(s0 + per_warp, steps);
    int cnt = 0;
#pragma unroll 8
    for (int st = s0; st < s1; ++st) {
        const int pos = 32 * st + lane;
        const float d = pos < n ? crop[pos] : 0.0f;
        cnt += __popc(__ballot_sync(FULL, d != 0.0f));                   // any(row) of pre/process.py:63
    }
    if (lane == 0) warp_cnt[warp] = cnt;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int i = 0; i < NWARP; ++i) {
        const int c = warp_cnt[i];
        if (i < warp) before += c;
        total += c;
    }
    if (tid == 0) n_valid[f] = total;
    if (out == nullptr || s0 >= s1) return;
    const double half_w = (double)w / 2.0, half_h = (double)hh / 2.0;
    const double yf = __drcp_rn(focal);
    const unsigned long long fb = (unsigned long long)__double_as_longlong(focal);
    const int fe = (int)((fb >> 52) & 0x7ff);
    const bool fast_b = fe > 1023 - 100 && fe < 1023 + 100 && (fb & 0xfffffffffffffull) != 0xfffffffffffffull;
    const unsigned lt = (1u << lane) - 1u;
    int pos = 32 * s0 + lane;
    int row = pos / bw, col = pos - row * bw;
    // eight steps per iteration, their loads issued together (an unstaged crop: one load per step and nothing else in
    // flight left the warps waiting on memory 63 % of the time - ncu, long scoreboard)
    constexpr int U = 8;
    for (int st = s0; st < s1; st += U) {
        float dv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = pos + 32 * u;
            dv[u] = (st + u < s1 && p < n) ? crop[p] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (st + u >= s1) break;                                      // (warp-uniform)
            const float d = dv[u];
            const bool keep = d != 0.0f;
            const unsigned m = __ballot_sync(FULL, keep);
            if (keep) {
                // (staging the 32 points in shared memory and storing them as one contiguous run was measured: no gain)
                const size_t dst = (size_t)(before + __popc(m & lt));
                const double dd = (double)d;
                const double uu = ((double)col + (double)l) - half_w;
                const double vv = ((double)row + (double)t) - half_h;
                out[3 * dst + 0] = ddiv_fixed(__dmul_rn(uu, dd), focal, yf, fast_b);
                out[3 * dst + 1] = ddiv_fixed(-__dmul_rn(vv, dd), focal, yf, fast_b);
                out[3 * dst + 2] = -dd;
            }
            before += __popc(m);
            col += 32;
            while (col >= bw) { col -= bw; ++row; }
        }
        pos += 32 * U;
    }
}

// pre/process.py:82 (gather by the caller's index vector) + :102-120 (per-axis extrema,
// z ignoring exact zeros, rounded to float32; rounding is monotonic so it commutes with
// min/max).  One CTA per frame.
__global__ void __launch_bounds__(NT) resample_extrema_kernel(int point_num,
                                                              const double* __restrict__ points,
                                                              const long long* __restrict__ point_offsets,
                                                              const int* __restrict__ indices,
                                                              double* __restrict__ out_pc,
                                                              float* __restrict__ out_extrema) {
    __shared__ float red[NWARP][8];
    const int f = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* src = points + 3 * point_offsets[f];
    const int* idx = indices ? indices + (size_t)f * point_num : nullptr;
    double* dst = out_pc ? out_pc + (size_t)f * point_num * 3 : nullptr;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    // four samples per thread and iteration: the four index loads, then the twelve gathers, in flight together (one
    // sample at a time is a chain of two dependent loads per iteration: latency-bound)
    constexpr int U = 4;
    for (int k0 = tid; k0 < point_num; k0 += U * NT) {
        int i[U];
        double p[U][3];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = k0 + u * NT;
            i[u] = k < point_num ? (idx ? idx[k] : k) : -1;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i[u] >= 0) { p[u][0] = src[3 * (size_t)i[u]]; p[u][1] = src[3 * (size_t)i[u] + 1]; p[u][2] = src[3 * (size_t)i[u] + 2]; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i[u] < 0) continue;
            const int k = k0 + u * NT;
            const double x = p[u][0], y = p[u][1], z = p[u][2];
            if (dst) { dst[3 * (size_t)k] = x; dst[3 * (size_t)k + 1] = y; dst[3 * (size_t)k + 2] = z; }
            const float xf = __double2float_rn(x), yf = __double2float_rn(y), zf = __double2float_rn(z);
            mn[0] = fminf(mn[0], xf); mx[0] = fmaxf(mx[0], xf);
            mn[1] = fminf(mn[1], yf); mx[1] = fmaxf(mx[1], yf);
            if (z != 0.0) { mn[2] = fminf(mn[2], zf); mx[2] = fmaxf(mx[2], zf); }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { mn[c] = warp_min(mn[c]); mx[c] = warp_max(mx[c]); }
    if (lane == 0) {
        red[warp][0] = mx[0]; red[warp][1] = mx[1]; red[warp][2] = mx[2];
        red[warp][3] = mn[0]; red[warp][4] = mn[1]; red[warp][5] = mn[2];
    }
    __syncthreads();
    if (tid < 6) {
        float v = red[0][tid];
        for (int w = 1; w < NWARP; ++w) v = tid < 3 ? fmaxf(v, red[w][tid]) : fminf(v, red[w][tid]);
        out_extrema[6 * (size_t)f + tid] = v;
    }
}

// ---------------------------------------------------------------- sampled cloud in one kernel
// DataProcess.point_cloud() + set_length() + max_min_point() (pre/process.py:30-84, 102-120) when only the SAMPLED
// cloud is wanted: hand_points[rand_idx] is written straight from the depth map, the full cloud (24 B per valid
// pixel, written by point_cloud_kernel and gathered from by resample_extrema_kernel: 211 KB + 1.75 sectors per
// sample) never exists.  One CTA per frame:
//   0  the crop is staged by one TMA bulk copy (crops beyond the staging area are read in place);
//   1  per 32-pixel word: the mask of valid pixels (d != 0) and, after one scan by warp 0, the number of valid
//      pixels before the word - the rank of a pixel is that count plus a popc;
//   2  sample k wants the pixel of rank idx[k]: its word by a table over the ranks that are multiples of 32 and a
//      short walk, the rank-th set bit of the word (five halvings), back-projection in float64 exactly as point_cloud_kernel does it, the point goes to
//      out_pc[k]; extrema as resample_extrema_kernel (float32-rounded, z ignoring exact zeros).
// Bit for bit what the two kernels produce.  Crops of more than SC_MAX_WORDS * 32 pixels are not handled here
// (the launcher says so and the caller takes the two-kernel route).
constexpr int SC_STAGE_FLOATS = 16 * 1024;          // 64 KB staged crop
constexpr int SC_MAX_WORDS = 704;                   // 22,528 pixels (the MSRA crops are below 150 x 140)
constexpr int SC_SMEM = SC_STAGE_FLOATS * 4;

__global__ void __launch_bounds__(NT) sample_cloud_kernel(double focal, int point_num,
                                                          const float* __restrict__ depth,
                                                          const long long* __restrict__ offsets,
                                                          const int* __restrict__ headers,
                                                          const int* __restrict__ indices,
                                                          double* __restrict__ out_pc,
                                                          float* __restrict__ out_extrema,
                                                          int* __restrict__ n_valid_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red[NWARP][8];
    __shared__ unsigned long long bar;
    __shared__ int total_s;
    __shared__ unsigned vmask[SC_MAX_WORDS];                  // valid pixels of each 32-pixel word
    __shared__ int wpre[SC_MAX_WORDS + 1];                    // valid pixels before the word
    __shared__ short start[SC_MAX_WORDS + 2];                 // the word that holds rank 32 q
    const int f = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* h = headers + 6 * (size_t)f;
    const int w = h[0], hh = h[1], l = h[2], t = h[3], r = h[4], b = h[5];
    const int bw = r - l, bh = b - t;
    const int n = (bw > 0 && bh > 0) ? bw * bh : 0;
    const int words = (n + 31) >> 5;
    if (words > SC_MAX_WORDS) {                               // (CTA-uniform) a crop the caller's hint did not announce:
        if (tid < 6) out_extrema[6 * (size_t)f + tid] = __int_as_float(QNAN_BITS);      // no result, nothing overrun
        if (tid == 0 && n_valid_out) n_valid_out[f] = -1;
        return;
    }
    const float* crop_g = depth + offsets[f];
    const int n_pad = (n + 3) & ~3;
    const bool staged = n > 0 && n_pad <= SC_STAGE_FLOATS;
    if (staged) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
            mbar_expect_tx(&bar, (uint32_t)n_pad * 4u);
            bulk_g2s(smem_raw, crop_g, (uint32_t)n_pad * 4u, &bar);
        }
        __syncthreads();
        mbar_wait(&bar, 0);
    }
    const float* crop = staged ? reinterpret_cast<const float*>(smem_raw) : crop_g;
    // ---- 1. valid masks and counts per word: a THREAD per word (32 pixels = eight 16-byte reads, in a per-lane
    //      rotated order so that a warp's reads spread over the banks), each thread a run of consecutive words, then
    //      one block scan of the runs.  (A warp per word - one ballot per 32 pixels - cost 30 % of the kernel's
    //      instructions: a warp instruction per word where a thread's would do.)
    __shared__ int wtot[NWARP];
    const int wpt = (words + NT - 1) / NT;                    // words per thread: 1..3
    const int w0 = min(tid * wpt, words), w1 = min(w0 + wpt, words);
    int mine = 0;
    for (int wd = w0; wd < w1; ++wd) {
        unsigned m = 0u;
        const int p0 = 32 * wd;
        if (staged && p0 + 32 <= n_pad) {                     // whole word inside the staged (16-byte aligned, zero-padded) crop
            const float4* q4 = reinterpret_cast<const float4*>(smem_raw) + 8 * wd;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int j = (i + lane) & 7;
                const float4 v = q4[j];
                const unsigned nib = (v.x != 0.0f ? 1u : 0u) | (v.y != 0.0f ? 2u : 0u) | (v.z != 0.0f ? 4u : 0u) | (v.w != 0.0f ? 8u : 0u);
                m |= nib << (4 * j);
            }
            if (p0 + 32 > n) m &= (1u << (n - p0)) - 1u;      // (padding is zero anyway)
        } else {
            for (int i = 0; i < 32 && p0 + i < n; ++i) m |= (crop[p0 + i] != 0.0f ? 1u : 0u) << i;
        }
        vmask[wd] = m;                                        // any(row) of pre/process.py:63
        mine += __popc(m);
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    int run = incl - mine;
#pragma unroll
    for (int i = 0; i < NWARP; ++i) run += i < warp ? wtot[i] : 0;
    for (int wd = w0; wd < w1; ++wd) {
        const int c = __popc(vmask[wd]);
        wpre[wd] = run;
        // a word holds at most 32 valid pixels, hence at most one rank that is a multiple of 32: it starts there
        const int q = (run + 31) >> 5;
        if (32 * q < run + c) start[q] = (short)wd;
        run += c;
    }
    if (tid == NT - 1) total_s = run;
    __syncthreads();
    const int total = total_s;
    if (tid == 0 && n_valid_out) n_valid_out[f] = total;
    // ---- 2. the samples
    const double half_w = (double)w / 2.0, half_h = (double)hh / 2.0;
    const double yf = __drcp_rn(focal);
    const unsigned long long fb = (unsigned long long)__double_as_longlong(focal);
    const int fe = (int)((fb >> 52) & 0x7ff);
    const bool fast_b = fe > 1023 - 100 && fe < 1023 + 100 && (fb & 0xfffffffffffffull) != 0xfffffffffffffull;
    const int* idx = indices ? indices + (size_t)f * point_num : nullptr;
    double* dst = out_pc ? out_pc + (size_t)f * point_num * 3 : nullptr;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    // four samples per thread and iteration, their index loads issued together: one sample at a time left the warps
    // waiting on the index load (ncu: long scoreboard)
    const float inv_bw = __frcp_rn((float)bw);
    constexpr int U = 8;
    for (int k0 = tid; k0 < point_num; k0 += U * NT) {
        int rank[U], wd[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = k0 + u * NT;
            rank[u] = k < point_num ? (idx ? __ldg(idx + k) : k) : -1;
            wd[u] = 0;
        }
        // the last word with wpre <= rank: from the word that holds the multiple of 32 below the rank, a few words on
        // (one or two in a hand; a binary search over all words was 24 % of the kernel's instructions)
#pragma unroll
        for (int u = 0; u < U; ++u)
            if ((unsigned)rank[u] < (unsigned)total) wd[u] = start[rank[u] >> 5];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if ((unsigned)rank[u] < (unsigned)total)
                while (wd[u] + 1 < words && wpre[wd[u] + 1] <= rank[u]) ++wd[u];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = k0 + u * NT;
            if (k >= point_num) continue;
            if ((unsigned)rank[u] >= (unsigned)total) {       // not a point of this frame: nothing to gather
                if (dst) { dst[3 * (size_t)k] = 0.0; dst[3 * (size_t)k + 1] = 0.0; dst[3 * (size_t)k + 2] = 0.0; }
                continue;
            }
            // the (rank - wpre[word])-th set bit of the word, by halving (__fns is a loop over the bits; a shortcut
            // for words with all 32 pixels valid only added divergence: 0.642 vs 0.626 ms)
            unsigned m = vmask[wd[u]];
            int want = rank[u] - wpre[wd[u]], bit = 0;
#pragma unroll
            for (int span = 16; span > 0; span >>= 1) {
                const int c = __popc(m & ((1u << span) - 1u));
                if (want >= c) { want -= c; m >>= span; bit += span; }
            }
            const int pos = 32 * wd[u] + bit;
            // pos / bw for pos < 2^15: float quotient, off by at most one
            int row = __float2int_rz(__fmul_rz((float)pos, inv_bw));
            if (row * bw > pos) --row;
            if ((row + 1) * bw <= pos) ++row;
            const int col = pos - row * bw;
            const double dd = (double)crop[pos];
            const double uu = ((double)col + (double)l) - half_w;
            const double vv = ((double)row + (double)t) - half_h;
            const double x = ddiv_fixed(__dmul_rn(uu, dd), focal, yf, fast_b);
            const double y = ddiv_fixed(-__dmul_rn(vv, dd), focal, yf, fast_b);
            const double z = -dd;
            if (dst) { dst[3 * (size_t)k] = x; dst[3 * (size_t)k + 1] = y; dst[3 * (size_t)k + 2] = z; }
            const float xf = __double2float_rn(x), yff = __double2float_rn(y), zf = __double2float_rn(z);
            mn[0] = fminf(mn[0], xf); mx[0] = fmaxf(mx[0], xf);
            mn[1] = fminf(mn[1], yff); mx[1] = fmaxf(mx[1], yff);
            if (z != 0.0) { mn[2] = fminf(mn[2], zf); mx[2] = fmaxf(mx[2], zf); }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { mn[c] = warp_min(mn[c]); mx[c] = warp_max(mx[c]); }
    if (lane == 0) {
        red[warp][0] = mx[0]; red[warp][1] = mx[1]; red[warp][2] = mx[2];
        red[warp][3] = mn[0]; red[warp][4] = mn[1]; red[warp][5] = mn[2];
    }
    __syncthreads();
    if (tid < 6) {
        float v = red[0][tid];
        for (int ww = 1; ww < NWARP; ++ww) v = tid < 3 ? fmaxf(v, red[ww][tid]) : fminf(v, red[ww][tid]);
        out_extrema[6 * (size_t)f + tid] = v;
    }
}

// (gt - mid_p)/max_l + 0.5 in float32, optional clamp to [0,1]
// (pre/joint_nor.py:8-18; 3D_CNN/train.py:236-244 does the same on float32 tensors).
__global__ void normalize_joints_kernel(int total, const float* __restrict__ joints,
                                        const float* __restrict__ max_l,
                                        const float* __restrict__ mid_p, int clamp01,
                                        float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int f = i / 63, j = i - f * 63;
    float v = __fadd_rn(__fdiv_rn(__fsub_rn(joints[i], mid_p[3 * f + j % 3]), max_l[f]), 0.5f);
    if (clamp01) v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    out[i] = v;
}

// ---------------------------------------------------------------- augmented D-TSDF
// Specification: oracle/dtsdf_oracle.c, oracle_dtsdf_aug_frame (the
// reference's own augmentation cannot run, pre/process.py:252).  p' = (p - c) . A + c; the grid is
// laid over the augmented hand, every voxel centre is taken back to the camera frame with Ainv
// (a sum of three per-axis table entries), projected and looked up as in pre/process.py:150-175, and the
// distance vector is measured in the augmented frame: (v - w) . A.  The projection no longer separates
// per axis, so pass 1 projects every voxel itself; it is the training-time path, not the bench's.
__device__ __forceinline__ void xform(float px, float py, float pz, const float* __restrict__ M,
                                      const float* __restrict__ c, float (&q)[3]) {
    const float dx = __fsub_rn(px, c[0]), dy = __fsub_rn(py, c[1]), dz = __fsub_rn(pz, c[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float t = __fmul_rn(dx, M[k]);
        t = __fmaf_rn(dy, M[3 + k], t);
        t = __fmaf_rn(dz, M[6 + k], t);
        q[k] = __fadd_rn(t, c[k]);
    }
}

// CTA shape of the augmented kernel: 8 warps, three CTAs per SM (the staged crop allows no more).  Measured on B200,
// 8,500 default frames: 256 x 3 1.306 ms; 512 x 2 (64 registers, 140 bytes spilled) 1.453 ms; 384 x 2 with the
// round-2 experiments 1.49 ms - with two CTAs per SM nothing covers a CTA's prologue (crop wait, bounding box, tables).
#ifndef DTSDF_AUG_NT
#define DTSDF_AUG_NT 256
#define DTSDF_AUG_MIN_CTAS 3
#endif
constexpr int AUG_NT = DTSDF_AUG_NT, AUG_NW = AUG_NT / 32, AUG_MIN_CTAS = DTSDF_AUG_MIN_CTAS;

// Per-frame constants of the augmented voxel loop.
struct AugFrame {
    const float* crop;            // staged (shared) or global crop
    unsigned crop_s32;            // the staged crop's address in the shared window (STAGED only)
    int l, t, bw, bh;
    int cx, cy;
    float focal, trunc;
    float yf, ytr;                // rcp_rn(focal), rcp_rn(trunc): the fixed-divisor divisions of pass 2
    bool fast_f, fast_t;          // ... are proven for this focal (241.42f) / this trunc range
    float t_far;                  // |v_z - w_z| above this forces the Euclidean truncation; +inf: no shortcut
};

// Per-axis tables of the separable voxel -> camera map (specification: oracle/dtsdf_oracle.c, step 3):
//   v_k = tx[x][k] + (ty[y][k] + tz[z][k]),   tx[i] = (va_x[i] - c_x) * Ainv[0][.],  ty likewise with row 1,
//   tz[i][k] = fma(va_z[i] - c_z, Ainv[2][k], c_k);   va_axis[i]: the voxel centres in the augmented frame.
template <int R>
struct __align__(16) AugTables {
    float4 tx[R], ty[R], tz[R];
};

template <int R, int LAYOUT>
__device__ __forceinline__ void aug_group_index(int g, int& f0, int& m, int& s) {
    constexpr int Q = R / 4;
    f0 = 4 * (g % Q); m = (g / Q) % R; s = g / (Q * R);
}

// Camera-frame centres of the four voxels of a group (two per instruction on the packed FP32 pipe).
template <int R, int LAYOUT>
__device__ __forceinline__ void aug_camera4(const AugTables<R>* __restrict__ T, int f0, int m, int s,
                                            float (&vx)[4], float (&vy)[4], float (&vz)[4]) {
    if (LAYOUT == 0) {                       // x = s, y = m, z = f0 + j
        const float4 a = T->tx[s], b = T->ty[m];
        const float2 ax = make_float2(a.x, a.x), ay = make_float2(a.y, a.y), az = make_float2(a.z, a.z);
        const float2 bx = make_float2(b.x, b.x), by = make_float2(b.y, b.y), bz = make_float2(b.z, b.z);
#pragma unroll
        for (int h = 0; h < 4; h += 2) {
            const float4 c0 = T->tz[f0 + h], c1 = T->tz[f0 + h + 1];
            const float2 qx = __fadd2_rn(ax, __fadd2_rn(bx, make_float2(c0.x, c1.x)));
            const float2 qy = __fadd2_rn(ay, __fadd2_rn(by, make_float2(c0.y, c1.y)));
            const float2 qz = __fadd2_rn(az, __fadd2_rn(bz, make_float2(c0.z, c1.z)));
            vx[h] = qx.x; vx[h + 1] = qx.y; vy[h] = qy.x; vy[h + 1] = qy.y; vz[h] = qz.x; vz[h + 1] = qz.y;
        }
    } else {                                 // x = f0 + j, y = m, z = s
        const float4 b = T->ty[m], c = T->tz[s];
        const float px = __fadd_rn(b.x, c.x), py = __fadd_rn(b.y, c.y), pz = __fadd_rn(b.z, c.z);
        const float2 qx = make_float2(px, px), qy = make_float2(py, py), qz = make_float2(pz, pz);
#pragma unroll
        for (int h = 0; h < 4; h += 2) {
            const float4 a0 = T->tx[f0 + h], a1 = T->tx[f0 + h + 1];
            const float2 rx = __fadd2_rn(make_float2(a0.x, a1.x), qx);
            const float2 ry = __fadd2_rn(make_float2(a0.y, a1.y), qy);
            const float2 rz = __fadd2_rn(make_float2(a0.z, a1.z), qz);
            vx[h] = rx.x; vx[h + 1] = rx.y; vy[h] = ry.x; vy[h + 1] = ry.y; vz[h] = rz.x; vz[h + 1] = rz.y;
        }
    }
}

// coeff = -focal / v_z (pre/process.py:154).  SAFE is a per-frame fact (aug_grid_is_safe): every voxel of the frame
// has 2^-60 <= |v_z| <= 2^60, and 2^-20 <= |focal| <= 2^20.  The division is then the straight-line part of what the
// compiler emits for __fdiv_rn - reciprocal seed (MUFU.RCP), one Newton step, quotient, one residual correction -
// without the range check and the branch to a slow path that every IEEE division carries (four per voxel group: the
// compiler cannot schedule across them).  Equal to __fdiv_rn: tests/native/rcp_division_check.cu runs EVERY divisor
// of the range on the GPU, for the default focal and 64 other dividends (the seed is a hardware approximation).
template <bool SAFE>
__device__ __forceinline__ float aug_div_coeff(float nf, float vz) {
    if (!SAFE) return __fdiv_rn(nf, vz);
    const float r0 = rcp_approx(vz);
    const float r1 = __fmaf_rn(r0, __fmaf_rn(-vz, r0, 1.0f), r0);
    const float q0 = __fmul_rn(nf, r1);
    return __fmaf_rn(r1, __fmaf_rn(-vz, q0, nf), q0);
}

// Projection and depth lookup of one voxel (pre/process.py:154-167), branch-free.  Returns the depth, or a
// value with |d| < 1 ("no depth": the voxel stays 0) when the projection is not finite or outside the crop.
// STAGED: the crop is in shared memory and the word right below it is 0.0f - the target of "outside".
template <bool STAGED, bool SAFE>
__device__ __forceinline__ float aug_lookup(const AugFrame& k, float vx, float vy, float vz, int& px, int& py) {
    const float coeff = aug_div_coeff<SAFE>(-k.focal, vz);
    const float fpx = __fadd_rn(__fmul_rn(vx, coeff), (float)k.cx);
    const float fpy = __fadd_rn(__fmul_rn(-vy, coeff), (float)k.cy);
    // the specification drops |fp| >= 2e9 (or NaN) before int(): the conversion saturates, so anything
    // that large falls outside the crop by itself; only NaN (which converts to 0) needs its own test -
    // and in a SAFE frame (finite v, finite non-zero coeff) no NaN can arise
    px = __float2int_rz(fpx);
    py = __float2int_rz(fpy);
    const unsigned ux = (unsigned)(px - k.l), uy = (unsigned)(py - k.t);
    if (STAGED && SAFE) {                                               // (outside: the zero word below the crop)
        int idx;                                                        // one predicate, one select
        asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\tsetp.lt.and.u32 p, %3, %4, p;\n\tselp.b32 %0, %5, -1, p;\n\t}"
            : "=r"(idx) : "r"(ux), "r"((unsigned)k.bw), "r"(uy), "r"((unsigned)k.bh), "r"((int)(uy * (unsigned)k.bw + ux)));
        return lds_f32(k.crop_s32 + 4u * (unsigned)idx);
    }
    bool in = (ux < (unsigned)k.bw) & (uy < (unsigned)k.bh);
    if (!SAFE) {
        const float fs = __fadd_rn(fpx, fpy);
        in = in & (fs == fs);
    }
    if (STAGED) {                                                       // (outside: the zero word below the crop)
        const unsigned off = in ? 4u * (uy * (unsigned)k.bw + ux) : 0xfffffffcu;
        return lds_f32(k.crop_s32 + off);
    }
    return in ? __ldg(k.crop + ((int)uy * k.bw + (int)ux)) : 0.0f;
}

// d / focal of pass 2: the fixed-divisor form (div_fixed, file header) inside its proven range - 1 <= |d| < 2^40
// with the default focal - and IEEE division otherwise.  Bit-identical either way.
__device__ __forceinline__ float aug_div_focal(const AugFrame& k, float d) {
    if (k.fast_f && fabsf(d) < BIG_DEPTH) return div_fixed<true>(d, k.focal, k.yf);     // (|d| >= 1: valid depth)
    return __fdiv_rn(d, k.focal);
}

// Pass 1: every voxel of the group is 0, +-1 by the far-surface shortcut, or "near".
template <int R, int LAYOUT, bool STAGED, bool SAFE>
__device__ __forceinline__ bool aug_classify_group(const AugFrame& k, const AugTables<R>* __restrict__ T, int g, float4& val) {
    int f0, m, s;
    aug_group_index<R, LAYOUT>(g, f0, m, s);
    float vx[4], vy[4], vz[4], o[4];
    aug_camera4<R, LAYOUT>(T, f0, m, s, vx, vy, vz);
    bool near = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int px, py;
        const float d = aug_lookup<STAGED, SAFE>(k, vx[j], vy[j], vz[j], px, py);
        const float sz = __fadd_rn(vz[j], d);                      // v_z - w_z, w_z = -d
        const bool valid = !(fabsf(d) < 1.0f);
        const bool far = fabsf(sz) > k.t_far && fabsf(d) < BIG_DEPTH;
        o[j] = (valid && far) ? __int_as_float((__float_as_int(sz) & 0x80000000) | 0x3F800000) : 0.0f;
        near = near || (valid && !far);
    }
    val = make_float4(o[0], o[1], o[2], o[3]);
    return near;
}

// Pass 2, one voxel, any operands: the specification's steps in its own order (oracle/dtsdf_oracle.c), the divisions
// in their fixed-divisor forms where their range tests allow.
__device__ __forceinline__ void aug_eval_voxel(const AugFrame& k, const float* __restrict__ A, float vx, float vy, float vz,
                                               float d, int px, int py, float& t0, float& t1, float& t2) {
    t0 = 0.0f; t1 = 0.0f; t2 = 0.0f;
    if (fabsf(d) < 1.0f) return;
    // the distance vector in the augmented frame: (v - w) . A, the products in the order of xform()
    const float c1 = aug_div_focal(k, d);
    const float dx = __fsub_rn(vx, __fmul_rn((float)(px - k.cx), c1));
    const float dy = __fsub_rn(vy, __fmul_rn((float)(-(py - k.cy)), c1));
    const float dz = __fadd_rn(vz, d);                                                 // v_z - w_z, w_z = -d
    float e[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        e[q] = __fmul_rn(dx, A[q]);
        e[q] = __fmaf_rn(dy, A[3 + q], e[q]);
        e[q] = __fmaf_rn(dz, A[6 + q], e[q]);
    }
    // the three divisions by trunc behind ONE range test (div_fixed's, file header, on the smallest and the
    // largest of the three magnitudes; a zero component takes the IEEE branch with the rest of its voxel)
    if (k.fast_t && fmin3(fabsf(e[0]), fabsf(e[1]), fabsf(e[2])) >= 0x1p-60f && fmax3(fabsf(e[0]), fabsf(e[1]), fabsf(e[2])) <= 0x1p60f) {
        t0 = fabsf(div_fixed<false>(e[0], k.trunc, k.ytr));                 // |a| / b == |a / b| under RN
        t1 = fabsf(div_fixed<false>(e[1], k.trunc, k.ytr));
        t2 = fabsf(div_fixed<false>(e[2], k.trunc, k.ytr));
    } else {
        t0 = fabsf(__fdiv_rn(e[0], k.trunc)); t1 = fabsf(__fdiv_rn(e[1], k.trunc)); t2 = fabsf(__fdiv_rn(e[2], k.trunc));
    }
    const float ss = __fadd_rn(__fadd_rn(__fmul_rn(t0, t0), __fmul_rn(t1, t1)), __fmul_rn(t2, t2));
    // sqrt_rn(ss) > 1 <=> ss > nextafter(1,2); then min(t,1) is the identity below it,
    // except for NaN (kept) - same as the specification's comparisons
    if (ss > THR_SQ) { t0 = 1.0f; t1 = 1.0f; t2 = 1.0f; }
    if (1.0f < t0) t0 = 1.0f;
    if (1.0f < t1) t1 = 1.0f;
    if (1.0f < t2) t2 = 1.0f;
    if (-d > vz) { t0 = -t0; t1 = -t1; t2 = -t2; }
}

// Pass 2: the specification's evaluation of all four voxels of a listed group.
template <int R, int LAYOUT, bool STAGED, bool SAFE>
__device__ __forceinline__ void aug_eval_group(const AugFrame& k, const AugTables<R>* __restrict__ T,
                                               const float* __restrict__ A, int g, float* __restrict__ out) {
    constexpr int CH = R * R * R;
    int f0, m, s;
    aug_group_index<R, LAYOUT>(g, f0, m, s);
    float vx[4], vy[4], vz[4];
    aug_camera4<R, LAYOUT>(T, f0, m, s, vx, vy, vz);
    float o0[4], o1[4], o2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int px, py;
        const float d = aug_lookup<STAGED, SAFE>(k, vx[j], vy[j], vz[j], px, py);
        aug_eval_voxel(k, A, vx[j], vy[j], vz[j], d, px, py, o0[j], o1[j], o2[j]);
    }
    float* p = out + 4 * (size_t)g;
    st_f4(p, make_float4(o0[0], o0[1], o0[2], o0[3]));
    st_f4(p + CH, make_float4(o1[0], o1[1], o1[2], o1[3]));
    st_f4(p + 2 * CH, make_float4(o2[0], o2[1], o2[2], o2[3]));
}

// Both passes over one frame.
template <int R, int LAYOUT, bool STAGED, bool SAFE>
__device__ __forceinline__ void aug_voxel_passes(const AugFrame& k, const AugTables<R>* __restrict__ T,
                                                 const float* __restrict__ A,
                                                 unsigned short* __restrict__ list, float* __restrict__ out) {
    constexpr int CH = R * R * R;
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0;
#pragma unroll 1
    for (int g = tid; g < CH / 4; g += AUG_NT) {
        float4 val;
        bool near = aug_classify_group<R, LAYOUT, STAGED, SAFE>(k, T, g, val);
        // a 32-byte sector (an even/odd lane pair) is written once, whole: by pass 1 or by pass 2
        unsigned bal = __ballot_sync(FULL, near);
        bal |= ((bal & 0x55555555u) << 1) | ((bal & 0xAAAAAAAAu) >> 1);
        near = (bal >> lane) & 1u;
        if (!near) {
            float* p = out + 4 * (size_t)g;
            st_f4(p, val); st_f4(p + CH, val); st_f4(p + 2 * CH, val);
        } else {
            list[cnt + __popc(bal & lt)] = (unsigned short)g;
        }
        cnt += __popc(bal);
        if (cnt >= 32) {
            __syncwarp();
            aug_eval_group<R, LAYOUT, STAGED, SAFE>(k, T, A, (int)list[lane], out);
            const int rem = cnt - 32;
            unsigned short carry = 0;
            if (lane < rem) carry = list[32 + lane];
            __syncwarp();
            if (lane < rem) list[lane] = carry;
            cnt = rem;
        }
    }
    __syncwarp();
    if (lane < cnt) aug_eval_group<R, LAYOUT, STAGED, SAFE>(k, T, A, (int)list[lane], out);
}

// Far-surface shortcut of the augmented path.  The specification's distance vector is (v - w) . A, v the camera-
// frame point it projects, and |u . A| >= smin(A) |u| >= smin(A) |u_z| for any u.  Lower bound for smin(A):
// |det A| / (s1 s2) >= 2 |det A| / |A|_F^2.  If
//     |v_z - w_z| > t_far = (trunc (1 + 2^-8) + slack) / smin_lo + drift
// then |(v - w) . A| / trunc > 1 with a margin far above the float32 rounding of the specification's own
// evaluation (slack: absolute rounding of coordinates of the frame's magnitude, 2^-17 relative; the
// part of the rounding that grows with |d| is relative to |v_z - w_z| itself and below 2^-12 while
// smin_lo > 2^-6 |A|_F, required here), so the Euclidean truncation fires and the voxel is +-1 by the
// sign of v_z - w_z (pre/process.py:180-194).  drift and the test on |Ainv.A - I| date from the first form of the
// specification, which measured from the voxel centre v' itself - v = (v' - c).Ainv + c is its preimage only as far
// as the caller's Ainv is A^-1; they are kept: more margin, and a map whose Ainv is unrelated to A runs without the
// shortcut.  Anything unusual (non-finite or huge entries, a nearly singular A, tiny or huge trunc) switches the
// shortcut off: t_far = +inf.
__device__ __forceinline__ float aug_far_threshold(const float* __restrict__ A, const float* __restrict__ Ainv,
                                                   const float* __restrict__ c, float trunc, float max_l,
                                                   float mid0, float mid1, float mid2) {
    const float det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
                      A[2] * (A[3] * A[7] - A[4] * A[6]);
    float fa = 0.0f, fi = 0.0f, e = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) { fa += A[i] * A[i]; fi += Ainv[i] * Ainv[i]; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float p = Ainv[3 * i] * A[j] + Ainv[3 * i + 1] * A[3 + j] + Ainv[3 * i + 2] * A[6 + j];
            e = fmaxf(e, fabsf(p - (i == j ? 1.0f : 0.0f)));
        }
    const float na = sqrtf(fa), ni = sqrtf(fi);
    float smin_lo = 2.0f * fabsf(det) / fa * 0.999f;
    {
        // smallest eigenvalue of B = A^T A in closed form (trigonometric solution of the cubic); its
        // float32 error is ~1e-6 of the LARGEST eigenvalue, so it is used only when it is at least
        // 2^-6 of the trace (relative error then < 1e-4, covered by the factor 0.998)
        const float b00 = A[0] * A[0] + A[3] * A[3] + A[6] * A[6], b11 = A[1] * A[1] + A[4] * A[4] + A[7] * A[7];
        const float b22 = A[2] * A[2] + A[5] * A[5] + A[8] * A[8];
        const float b01 = A[0] * A[1] + A[3] * A[4] + A[6] * A[7], b02 = A[0] * A[2] + A[3] * A[5] + A[6] * A[8];
        const float b12 = A[1] * A[2] + A[4] * A[5] + A[7] * A[8];
        const float q = (b00 + b11 + b22) * (1.0f / 3.0f);
        const float p1 = b01 * b01 + b02 * b02 + b12 * b12;
        const float p2 = (b00 - q) * (b00 - q) + (b11 - q) * (b11 - q) + (b22 - q) * (b22 - q) + 2.0f * p1;
        float emin = q;
        if (p2 > 1e-12f * q * q) {
            const float pp = sqrtf(p2 * (1.0f / 6.0f)), ip = 1.0f / pp;
            const float c00 = (b00 - q) * ip, c11 = (b11 - q) * ip, c22 = (b22 - q) * ip;
            const float c01 = b01 * ip, c02 = b02 * ip, c12 = b12 * ip;
            float r = 0.5f * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) + c02 * (c01 * c12 - c11 * c02));
            r = fminf(1.0f, fmaxf(-1.0f, r));
            const float phi = acosf(r) * (1.0f / 3.0f);
            emin = q + 2.0f * pp * cosf(phi + 2.0943951023931953f);
        }
        if (emin > 0x1p-6f * 3.0f * q) smin_lo = fmaxf(smin_lo, sqrtf(emin) * 0.998f);
    }
    // What the shortcut tests is the z component alone, and for a vector u with a given u_z the smallest
    // |u . A| is |u_z| / |column z of A^-1| (minimum of the quadratic form u (A A^T) u^T with one coordinate
    // fixed) - never below smin(A), and exactly 1 for the stretch-in-x,y-then-rotate maps of the reference
    // (show/visualize.py:106-114: A = diag(sx, sy, 1) . R), whose near band is then as narrow as the plain
    // kernel's instead of 1/min(sx, sy) times wider.  Column z of A^-1 = (row 0 of A) x (row 1 of A) / det.
    float kappa_lo = smin_lo;
    {
        const float k0 = A[1] * A[5] - A[2] * A[4], k1 = A[2] * A[3] - A[0] * A[5], k2 = A[0] * A[4] - A[1] * A[3];
        const float cn = sqrtf(k0 * k0 + k1 * k1 + k2 * k2);
        // (each cross-product term carries a float32 error of ~2^-22 |A|_F^2: used only when the result is
        // well above it)
        if (cn > 0x1p-8f * fa && fabsf(det) > 0x1p-12f * na * fa) kappa_lo = fmaxf(kappa_lo, fabsf(det) / cn * 0.998f);
    }
    const float size = fabsf(c[0]) + fabsf(c[1]) + fabsf(c[2]) + fabsf(mid0) + fabsf(mid1) + fabsf(mid2) +
                       2.0f * max_l + 1.0f;
    const float reach = fabsf(mid0 - c[0]) + fabsf(mid1 - c[1]) + fabsf(mid2 - c[2]) + max_l;
    const float drift = 3.0f * e * reach / smin_lo;
    const float slack = 0x1p-17f * (1.0f + na) * (1.0f + ni) * size;
    const float t_far = ((trunc * 1.00390625f + slack) / kappa_lo + drift) * 1.001f;
    const bool ok = smin_lo > 0x1p-6f * na && na < 0x1p10f && ni < 0x1p10f && size < 0x1p20f &&
                    trunc >= 0x1p-20f && trunc <= 0x1p20f && e < 0x1p-10f && t_far < 0x1p30f;   // NaN -> false
    return ok ? t_far : INFINITY;
}

// Is the whole grid of the frame inside the range aug_div_coeff<true> is proven for, and free of NaN sources?
// The per-axis table entries (AugTables) are monotone in the voxel index - every step of their evaluation is a
// correctly rounded monotone operation - and so is the table sum: the extremes over the grid of each camera-frame
// component are EXACTLY the sums of the per-axis extremes, which sit at the first and last voxel of each axis.
template <int R>
__device__ __forceinline__ bool aug_grid_is_safe(float ox, float oy, float oz, float voxel_len, const float* __restrict__ Ainv,
                                                 const float* __restrict__ c, float focal) {
    const float o[3] = {ox, oy, oz};
    bool ok = fabsf(focal) >= 0x1p-20f && fabsf(focal) <= 0x1p20f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float lo[3], hi[3];
#pragma unroll
        for (int axis = 0; axis < 3; ++axis) {
            const float d0 = __fsub_rn(__fadd_rn(o[axis], __fmul_rn(0.0f, voxel_len)), c[axis]);
            const float d1 = __fsub_rn(__fadd_rn(o[axis], __fmul_rn((float)(R - 1), voxel_len)), c[axis]);
            const float e0 = axis < 2 ? __fmul_rn(d0, Ainv[3 * axis + k]) : __fmaf_rn(d0, Ainv[6 + k], c[k]);
            const float e1 = axis < 2 ? __fmul_rn(d1, Ainv[3 * axis + k]) : __fmaf_rn(d1, Ainv[6 + k], c[k]);
            ok = ok && e0 == e0 && e1 == e1;
            lo[axis] = fminf(e0, e1); hi[axis] = fmaxf(e0, e1);
        }
        const float vlo = __fadd_rn(lo[0], __fadd_rn(lo[1], lo[2])), vhi = __fadd_rn(hi[0], __fadd_rn(hi[1], hi[2]));
        ok = ok && fabsf(vlo) <= 0x1p60f && fabsf(vhi) <= 0x1p60f;
        if (k == 2) ok = ok && (vlo >= 0x1p-60f || vhi <= -0x1p-60f);
    }
    return ok;
}

template <int R, int LAYOUT>
__global__ void __launch_bounds__(AUG_NT, AUG_MIN_CTAS) dtsdf_aug_kernel(const __grid_constant__ BatchArgs a,
                                                       const float* __restrict__ aug_all) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red[AUG_NW][8];
    __shared__ float aug[21];
    __shared__ unsigned long long bar;
    __shared__ AugTables<R> tab;
    __shared__ float t_far_s;
    __shared__ int safe_s;
    // the staged crop starts 16 bytes into the dynamic shared memory; the word right below it is 0.0f
    // ("no depth"), where lookups that fall outside the crop are pointed
    float* crop_s = reinterpret_cast<float*>(smem_raw + 16);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.x;
    const Frame fr = load_frame(a, f);
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        if (fr.staged) {
            mbar_expect_tx(&bar, (uint32_t)fr.n_pad * 4u);
            bulk_g2s(crop_s, fr.crop, (uint32_t)fr.n_pad * 4u, &bar);
        }
        *reinterpret_cast<float4*>(smem_raw) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid < 21) aug[tid] = __ldg(aug_all + 21 * (size_t)f + tid);
    __syncthreads();
    if (fr.staged) mbar_wait(&bar, 0);
    const float* __restrict__ crop = fr.staged ? crop_s : fr.crop;
    const unsigned crop_s32 = smem_u32(crop_s);
    const float *A = aug, *Ainv = aug + 9, *c = aug + 18;
    const float focal = a.focal;

    // ---- 1. bounding box of the augmented points of all valid pixels: two pixels of a row per thread and
    //      step, back-projection and transform on the packed FP32 pipe (FMUL2 / FFMA2: the same IEEE
    //      operations, two at a time)
    float mn[3] = {SENTINEL, SENTINEL, SENTINEL}, mx[3] = {-SENTINEL, -SENTINEL, -SENTINEL};
    {
        const float yf = __frcp_rn(focal);
        const bool deff = __float_as_uint(focal) == DEFAULT_FOCAL_BITS;
        const float2 nc2 = make_float2(-c[2], -c[2]);
        for (int row = warp; row < fr.bh; row += AUG_NW) {
            const float fy = (float)(row + fr.t - a.center_y);
            for (int col = lane; col < fr.bw; col += 64) {
                const int i0 = row * fr.bw + col;
                const float d0 = fr.staged ? lds_f32(crop_s32 + 4u * (unsigned)i0) : __ldg(fr.crop + i0);
                const float d1 = col + 32 < fr.bw ? (fr.staged ? lds_f32(crop_s32 + 4u * (unsigned)(i0 + 32)) : __ldg(fr.crop + i0 + 32)) : 0.0f;
                const bool v0 = !(fabsf(d0) < 1.0f), v1 = !(fabsf(d1) < 1.0f);
                if (!(v0 || v1)) continue;
                float fx0 = (float)(col + fr.l - a.center_x), fx1 = (float)(col + 32 + fr.l - a.center_x);
                // a pixel without depth takes its partner's place: the pair then holds the valid pixel twice and
                // the extremes need no per-pixel guard (three-input minimum / maximum over both lanes of the pair)
                const float e0 = v0 ? d0 : d1, e1 = v1 ? d1 : d0;
                fx0 = v0 ? fx0 : fx1; fx1 = v1 ? fx1 : fx0;
                float2 q;
                if (deff && fabsf(e0) < BIG_DEPTH && fabsf(e1) < BIG_DEPTH)
                    q = div_fixed2<true>(make_float2(e0, e1), make_float2(-focal, -focal), make_float2(yf, yf));
                else
                    q = make_float2(__fdiv_rn(e0, focal), __fdiv_rn(e1, focal));
                // X, Y are products that are then added to: scalar mul.rn, scalar add.rn (ptxas contracts the
                // packed forms, see eval_pair); from dx on, every product is inside a specified fma
                const float2 dx = make_float2(__fsub_rn(__fmul_rn(q.x, fx0), c[0]), __fsub_rn(__fmul_rn(q.y, fx1), c[0]));
                const float2 dy = make_float2(__fsub_rn(__fmul_rn(-q.x, fy), c[1]), __fsub_rn(__fmul_rn(-q.y, fy), c[1]));
                const float2 dz = __fadd2_rn(make_float2(-e0, -e1), nc2);
                float2 pa[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float2 t = __fmul2_rn(dx, make_float2(A[k], A[k]));
                    t = __ffma2_rn(dy, make_float2(A[3 + k], A[3 + k]), t);
                    t = __ffma2_rn(dz, make_float2(A[6 + k], A[6 + k]), t);
                    pa[k] = __fadd2_rn(t, make_float2(c[k], c[k]));
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    mn[k] = fmin3(mn[k], pa[k].x, pa[k].y); mx[k] = fmax3(mx[k], pa[k].x, pa[k].y);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { mn[k] = warp_min(mn[k]); mx[k] = warp_max(mx[k]); }
    if (lane == 0) {
        *reinterpret_cast<float4*>(&red[warp][0]) = make_float4(mn[0], mn[1], mn[2], mx[0]);
        *reinterpret_cast<float2*>(&red[warp][4]) = make_float2(mx[1], mx[2]);
    }
    __syncthreads();
    bbox_reduce<AUG_NW>(red, mn, mx);

    // ---- 2. grid parameters (pre/process.py:91-96)
    const float mid0 = __fmul_rn(__fadd_rn(mx[0], mn[0]), 0.5f);
    const float mid1 = __fmul_rn(__fadd_rn(mx[1], mn[1]), 0.5f);
    const float mid2 = __fmul_rn(__fadd_rn(mx[2], mn[2]), 0.5f);
    const float max_l = fmaxf(fmaxf(__fsub_rn(mx[0], mn[0]), __fsub_rn(mx[1], mn[1])), __fsub_rn(mx[2], mn[2]));
    const float voxel_len = __fdiv_rn(max_l, (float)R);
    const float trunc = __fmul_rn(voxel_len, a.trunc_mul);
    const float half_l = __fmul_rn(max_l, 0.5f), half_v = __fmul_rn(voxel_len, 0.5f);
    const float ox = __fadd_rn(__fsub_rn(mid0, half_l), half_v);
    const float oy = __fadd_rn(__fsub_rn(mid1, half_l), half_v);
    const float oz = __fadd_rn(__fsub_rn(mid2, half_l), half_v);
    if (tid == 0) {
        if (a.max_l) a.max_l[f] = max_l;
        if (a.mid_p) {
            a.mid_p[3 * (size_t)f + 0] = mid0; a.mid_p[3 * (size_t)f + 1] = mid1; a.mid_p[3 * (size_t)f + 2] = mid2;
        }
    }
    // augmented, normalised joints: (T(gt; A, c) - mid_p)/max_l + 0.5
    if (a.joints != nullptr && tid < 21) {
        const float* g = a.joints + 63 * (size_t)f + 3 * tid;
        float gz = __ldg(g + 2);
        if (a.joint_flags & 2) gz = -gz;
        float ga[3];
        xform(__ldg(g), __ldg(g + 1), gz, A, c, ga);
        const float m[3] = {mid0, mid1, mid2};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float v = __fadd_rn(__fdiv_rn(__fsub_rn(ga[k], m[k]), max_l), 0.5f);
            if (a.joint_flags & 1) v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
            a.out_joints[63 * (size_t)f + 3 * tid + k] = v;
        }
    }

    // ---- 3. per-axis tables of the voxel -> camera map (the specification's step 3), then the voxel loop:
    //      a thread owns 4 consecutive voxels along the fastest output axis (a warp: 512 contiguous bytes per
    //      channel).  Two passes per warp as in the fused kernel: pass 1 classifies (table sum, projection,
    //      lookup, far-surface shortcut) and stores the sectors that hold only 0 / +-1; groups in sectors with
    //      near-surface voxels go to a per-warp list and are evaluated 32 at a time by pass 2 with all lanes
    //      busy (the specification's arithmetic, IEEE divisions).
    constexpr int CH = R * R * R;
    static_assert((CH / 4) % AUG_NT == 0 && CH / 4 <= 65536 && 3 * R <= AUG_NT, "groups per frame: whole warps, 16-bit list entries");
    if (tid < 3 * R) {
        const int axis = tid / R, i = tid - axis * R;
        const float o = axis == 0 ? ox : (axis == 1 ? oy : oz);
        const float va = __fadd_rn(o, __fmul_rn((float)i, voxel_len));
        const float dd = __fsub_rn(va, c[axis]);
        if (axis == 0)      tab.tx[i] = make_float4(__fmul_rn(dd, Ainv[0]), __fmul_rn(dd, Ainv[1]), __fmul_rn(dd, Ainv[2]), 0.0f);
        else if (axis == 1) tab.ty[i] = make_float4(__fmul_rn(dd, Ainv[3]), __fmul_rn(dd, Ainv[4]), __fmul_rn(dd, Ainv[5]), 0.0f);
        else                tab.tz[i] = make_float4(__fmaf_rn(dd, Ainv[6], c[0]), __fmaf_rn(dd, Ainv[7], c[1]),
                                                    __fmaf_rn(dd, Ainv[8], c[2]), 0.0f);
    }
    __shared__ unsigned short glist[AUG_NW][64];
    AugFrame k;
    k.crop = crop; k.crop_s32 = smem_u32(crop_s);
    asm volatile("" : "+r"(k.crop_s32));      // kept in a register: rebuilt from the CTA's shared window per group otherwise
 k.l = fr.l; k.t = fr.t; k.bw = max(fr.bw, 0); k.bh = max(fr.bh, 0);   // bad header: empty crop
    k.cx = a.center_x; k.cy = a.center_y; k.focal = focal; k.trunc = trunc;
    k.yf = __frcp_rn(focal); k.ytr = __frcp_rn(trunc);
    k.fast_f = __float_as_uint(focal) == DEFAULT_FOCAL_BITS;
    k.fast_t = trunc >= 0x1p-20f && trunc <= 0x1p20f;
    // (one thread, not one of those filling the tables: a few hundred instructions that every thread used to repeat)
    if (tid == AUG_NT - 1) {
        t_far_s = aug_far_threshold(A, Ainv, c, trunc, max_l, mid0, mid1, mid2);
        safe_s = aug_grid_is_safe<R>(ox, oy, oz, voxel_len, Ainv, c, focal) ? 1 : 0;
    }
    __syncthreads();          // the tables are complete
    k.t_far = t_far_s;
    float* out = a.vol + (size_t)f * 3 * CH;
    if (!fr.staged)                     aug_voxel_passes<R, LAYOUT, false, false>(k, &tab, A, glist[warp], out);
    else if (safe_s)                    aug_voxel_passes<R, LAYOUT, true, true>(k, &tab, A, glist[warp], out);
    else                                aug_voxel_passes<R, LAYOUT, true, false>(k, &tab, A, glist[warp], out);
}

template <int R, int LAYOUT>
cudaError_t launch_aug_t(const BatchArgs& args_in, const float* aug, int max_crop_px, cudaStream_t stream) {
    BatchArgs args = args_in;
    int dev = 0, smem_optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    if ((reinterpret_cast<uintptr_t>(args.vol) & 15) != 0 || (reinterpret_cast<uintptr_t>(args.depth) & 15) != 0)
        return cudaErrorMisalignedAddress;
    auto kern = dtsdf_aug_kernel<R, LAYOUT>;
    // what is left for the staged crop beside the kernel's static shared memory (tables, lists: 3.5-7 KB by R)
    static std::atomic<long long> static_bytes{-1};
    long long fixed = static_bytes.load(std::memory_order_relaxed);
    if (fixed < 0) {
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, kern);
        if (e != cudaSuccess) return e;
        fixed = (long long)fa.sharedSizeBytes;
        static_bytes.store(fixed, std::memory_order_relaxed);
    }
    long long want = max_crop_px > 0 ? (((long long)max_crop_px + 3) & ~3ll) * 4 : (long long)smem_optin;
    const long long avail = (long long)smem_optin - fixed - 1024 - 16;
    if (want > avail) want = avail;
    if (want < 16) want = 16;
    want &= ~15ll;
    args.cap_floats = (int)(want / 4);
    static std::atomic<unsigned> smem_ok{0u};
    e = allow_max_smem(kern, dev, smem_optin, smem_ok);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)args.n_frames, AUG_NT, (size_t)want + 16, stream>>>(args, aug);          // + the zero word below the crop
    ++g_launch_count;
    return cudaGetLastError();
}

}  // namespace

// ---------------------------------------------------------------- launchers
cudaError_t launch_voxelize(int voxel_res, int layout, BatchArgs args, int max_crop_px,
                            cudaStream_t stream) {
    if (args.n_frames <= 0) return cudaSuccess;
    // the one-correction division by focal is proven for 241.42f only (exhaustive test)
    const bool deff = __builtin_bit_cast(unsigned, args.focal) == DEFAULT_FOCAL_BITS;
#define DTSDF_CASE(RR)                                                                              \
    case RR:                                                                                        \
        if (layout == 0)                                                                            \
            return deff ? launch_fused<RR, 0, true>(args, max_crop_px, stream)                      \
                        : launch_fused<RR, 0, false>(args, max_crop_px, stream);                    \
        return deff ? launch_fused<RR, 1, true>(args, max_crop_px, stream)                          \
                    : launch_fused<RR, 1, false>(args, max_crop_px, stream);
    switch (voxel_res) {
        DTSDF_CASE(32)
        DTSDF_CASE(48)
        DTSDF_CASE(64)
        default:
            return cudaErrorInvalidValue;
    }
#undef DTSDF_CASE
}

cudaError_t launch_voxelize_aug(int voxel_res, int layout, BatchArgs args, const float* aug, int max_crop_px,
                                cudaStream_t stream) {
    if (args.n_frames <= 0) return cudaSuccess;
    switch (voxel_res) {
        case 32: return layout == 0 ? launch_aug_t<32, 0>(args, aug, max_crop_px, stream) : launch_aug_t<32, 1>(args, aug, max_crop_px, stream);
        case 48: return layout == 0 ? launch_aug_t<48, 0>(args, aug, max_crop_px, stream) : launch_aug_t<48, 1>(args, aug, max_crop_px, stream);
        case 64: return layout == 0 ? launch_aug_t<64, 0>(args, aug, max_crop_px, stream) : launch_aug_t<64, 1>(args, aug, max_crop_px, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_count_valid(int n_frames, const float* depth, const long long* offsets, const int* headers,
                               int* n_valid, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(depth) & 15) != 0) return cudaErrorMisalignedAddress;
    count_valid_kernel<<<n_frames, NT, 0, stream>>>(depth, offsets, headers, n_valid);
    ++g_launch_count;
    return cudaGetLastError();
}

cudaError_t launch_point_cloud(int n_frames, double focal, const float* depth,
                               const long long* offsets, const int* headers, double* points,
                               const long long* point_offsets, int* n_valid,
                               cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    // (the opt-in limit is a per-device attribute of the kernel)
    int dev = 0, smem_optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    static std::atomic<unsigned> smem_ok{0u};
    if (e == cudaSuccess) e = allow_max_smem(point_cloud_kernel, dev, smem_optin, smem_ok);
    if (e != cudaSuccess) return e;
    point_cloud_kernel<<<n_frames, NT, PC_STAGE_FLOATS * 4, stream>>>(n_frames, focal, depth, offsets, headers, points,
                                                                      point_offsets, n_valid);
    ++g_launch_count;
    return cudaGetLastError();
}

int sample_cloud_max_pixels() { return SC_MAX_WORDS * 32; }

cudaError_t launch_sample_cloud(int n_frames, double focal, int point_num, const float* depth, const long long* offsets,
                                const int* headers, const int* indices, double* out_pc, float* out_extrema,
                                int* n_valid, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    // (the opt-in limit is a per-device attribute of the kernel)
    int dev = 0, smem_optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    static std::atomic<unsigned> smem_ok{0u};
    if (e == cudaSuccess) e = allow_max_smem(sample_cloud_kernel, dev, smem_optin, smem_ok);
    if (e != cudaSuccess) return e;
    sample_cloud_kernel<<<n_frames, NT, SC_SMEM, stream>>>(focal, point_num, depth, offsets, headers, indices, out_pc,
                                                           out_extrema, n_valid);
    ++g_launch_count;
    return cudaGetLastError();
}

cudaError_t launch_resample_extrema(int n_frames, int point_num, const double* points,
                                    const long long* point_offsets, const int* indices,
                                    double* out_pc, float* out_extrema, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    resample_extrema_kernel<<<n_frames, NT, 0, stream>>>(point_num, points, point_offsets, indices,
                                                         out_pc, out_extrema);
    ++g_launch_count;
    return cudaGetLastError();
}

cudaError_t launch_normalize_joints(int n_frames, const float* joints, const float* max_l,
                                    const float* mid_p, int clamp01, float* out,
                                    cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    const int total = n_frames * 63;
    normalize_joints_kernel<<<(total + 255) / 256, 256, 0, stream>>>(total, joints, max_l, mid_p,
                                                                     clamp01, out);
    ++g_launch_count;
    return cudaGetLastError();
}

}  // namespace dtsdf

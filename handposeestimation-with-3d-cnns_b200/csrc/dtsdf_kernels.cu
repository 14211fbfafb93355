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
// Design (one CTA works on one frame at a time, persistent over the batch):
//   1. thread 0 starts a TMA bulk copy (cp.async.bulk, completion on an mbarrier) of
//      the frame's depth crop global -> shared memory;
//   2. all threads scan the crop in shared memory for the bounding box; warp
//      reduction with redux.sync.{min,max}.f32 (CREDUX, new on sm_100a), block
//      reduction through shared memory;
//   3. every thread derives the grid parameters; the block fills two R x R tables of
//      projected pixel columns / rows (the projection separates: pix_x depends on
//      (x,z) only, pix_y on (y,z) only), so the voxel loop does no projection maths;
//   4. voxel loop: each thread owns 4 consecutive voxels along the fastest output
//      axis, gathers their depth pixels from shared memory, evaluates the truncated
//      signed directional distances and writes three 16-byte streaming stores (one per
//      channel); a warp store covers 512 contiguous bytes.
// The path is a streaming, write-dominated map with a small gather: HBM-bound, no
// tensor cores (there is no contraction to feed them).
#include "dtsdf_internal.h"

#include <math.h>

namespace dtsdf {

long long g_launch_count = 0;

namespace {

constexpr int NT = 256;                 // threads per CTA
constexpr int NWARP = NT / 32;
constexpr float SENTINEL = 99999.0f;    // pre/tsdf_numba.py:88-89
constexpr int MAGIC_I = 0x4B400000;     // bit pattern of 12582912.0f = 1.5 * 2^23
constexpr float MAGIC_F = 12582912.0f;  // int_as_float(MAGIC_I + k) == MAGIC_F + k for |k| < 2^22

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on the mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ float warp_min(float v) {
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float warp_max(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float* p, float a, float b, float c, float d) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// ---------------------------------------------------------------- shared memory
template <int R>
struct __align__(16) FrameShared {
    float red[NWARP][8];            // per-warp partial min xyz / max xyz (6 used)
    float vx[R], vy[R], vz[R];      // voxel centres per axis
    short pxt[R * R];               // projected column - left, or -1 when outside the crop
    short pyt[R * R];               // projected row - top,    or -1
    unsigned long long bar;         // TMA completion barrier
};

struct Frame {
    int l, t, r, b, bw, n;
};

// ---------------------------------------------------------------- bounding box scan
// pre/tsdf_numba.py:83-96.  `crop` is shared memory when STAGED, global otherwise.
template <bool STAGED>
__device__ __forceinline__ void bbox_scan(const float* __restrict__ crop, const Frame& fr,
                                          float focal, int cx, int cy, float (&mn)[3],
                                          float (&mx)[3]) {
    // row = pos / bw via multiply-high; exact for pos*bw < 2^32 (bw, bh <= 32767 is
    // checked by the caller and n = bw*bh < 2^30, so the error term pos*(m*bw - 2^32)
    // stays below 2^32 as m*bw - 2^32 < bw).
    const unsigned magic = fr.bw > 1 ? (unsigned)((0x100000000ull + (unsigned)fr.bw - 1) / (unsigned)fr.bw) : 0u;
    const float kx = MAGIC_F + (float)cx;
    const float ky = MAGIC_F + (float)cy;
    for (int pos = threadIdx.x; pos < fr.n; pos += NT) {
        const float d = STAGED ? crop[pos] : __ldg(crop + pos);
        if (!(fabsf(d) < 1.0f)) {
            const int row = fr.bw > 1 ? (int)__umulhi((unsigned)pos, magic) : pos;
            const int col = pos - row * fr.bw;
            const float q = __fdiv_rn(d, focal);
            const float fx = __fsub_rn(__int_as_float(MAGIC_I + col + fr.l), kx);   // (x - cx), exact
            const float fy = __fsub_rn(__int_as_float(MAGIC_I + row + fr.t), ky);   // (y - cy), exact
            const float X = __fmul_rn(q, fx);
            const float Y = __fmul_rn(-q, fy);
            const float Z = -d;
            mn[0] = fminf(mn[0], X); mx[0] = fmaxf(mx[0], X);
            mn[1] = fminf(mn[1], Y); mx[1] = fmaxf(mx[1], Y);
            mn[2] = fminf(mn[2], Z); mx[2] = fmaxf(mx[2], Z);
        }
    }
}

// ---------------------------------------------------------------- voxel evaluation
struct VoxConst {
    int   bw;
    int   ix;        // MAGIC_I + left:  int_as_float(ix + (px-left)) == MAGIC_F + px
    int   iy;        // MAGIC_I + top
    float kx;        // MAGIC_F + cx
    float ky;        // MAGIC_F + cy
    float focal;
    float trunc;
};

// pre/process.py:163-194 == pre/tsdf_numba.py:38-68 for one voxel whose projection
// (pa, pb) = (pix_x - left, pix_y - top) is inside the crop.
template <bool STAGED>
__device__ __forceinline__ void eval_voxel(const float* __restrict__ crop, const VoxConst& k,
                                           int pa, int pb, float vx, float vy, float vz,
                                           float& o0, float& o1, float& o2) {
    const int idx = pb * k.bw + pa;
    const float d = STAGED ? crop[idx] : __ldg(crop + idx);
    if (fabsf(d) < 1.0f) return;
    const float c1 = __fdiv_rn(d, k.focal);
    const float fx = __fsub_rn(__int_as_float(k.ix + pa), k.kx);    //  (pix_x - cx)
    const float fy = __fsub_rn(k.ky, __int_as_float(k.iy + pb));    // -(pix_y - cy)
    const float wx = __fmul_rn(fx, c1);
    const float wy = __fmul_rn(fy, c1);
    const float wz = -d;
    float tx = __fdiv_rn(fabsf(__fsub_rn(vx, wx)), k.trunc);
    float ty = __fdiv_rn(fabsf(__fsub_rn(vy, wy)), k.trunc);
    float tz = __fdiv_rn(fabsf(__fsub_rn(vz, wz)), k.trunc);
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(tx, tx), __fmul_rn(ty, ty)), __fmul_rn(tz, tz));
    // sqrt_rn(s) > 1  <=>  s > nextafter(1,2): the square root of pre/process.py:180 is
    // only compared with 1 (tests/test_oracle_golden.py::test_sqrt_truncation_threshold).
    // Below the threshold every component is already <= 1, so min(t,1) (:187-189) is
    // the identity there.
    if (s > 1.00000011920928955078125f) { tx = 1.0f; ty = 1.0f; tz = 1.0f; }
    if (wz > vz) { tx = -tx; ty = -ty; tz = -tz; }
    o0 = tx; o1 = ty; o2 = tz;
}

template <int R, int LAYOUT, bool STAGED>
__device__ __forceinline__ void voxel_loop(const float* __restrict__ crop,
                                           const FrameShared<R>* __restrict__ sh,
                                           const VoxConst& k, float* __restrict__ out) {
    constexpr int Q = R / 4;                 // 4-voxel groups along the fastest axis
    constexpr int ITEMS = R * R * Q;
    constexpr int CH = R * R * R;            // channel stride
#pragma unroll 1
    for (int it = threadIdx.x; it < ITEMS; it += NT) {
        const int q4 = it % Q;
        const int m = (it / Q) % R;          // middle axis: y in both layouts
        const int o = it / (Q * R);          // slowest axis: x (CXYZ) or z (CZYX)
        short4 pa4, pb4;
        float4 va, vb, vc;                   // voxel centre x / y / z for the 4 voxels
        if (LAYOUT == 0) {                   // out[c][x][y][z]: 4 consecutive z
            pa4 = *reinterpret_cast<const short4*>(&sh->pxt[o * R + 4 * q4]);   // [x][z]
            pb4 = *reinterpret_cast<const short4*>(&sh->pyt[m * R + 4 * q4]);   // [y][z]
            const float x = sh->vx[o], y = sh->vy[m];
            va = make_float4(x, x, x, x);
            vb = make_float4(y, y, y, y);
            vc = *reinterpret_cast<const float4*>(&sh->vz[4 * q4]);
        } else {                             // out[c][z][y][x]: 4 consecutive x
            pa4 = *reinterpret_cast<const short4*>(&sh->pxt[o * R + 4 * q4]);   // [z][x]
            const short pb = sh->pyt[o * R + m];                                // [z][y]
            pb4 = make_short4(pb, pb, pb, pb);
            const float y = sh->vy[m], z = sh->vz[o];
            va = *reinterpret_cast<const float4*>(&sh->vx[4 * q4]);
            vb = make_float4(y, y, y, y);
            vc = make_float4(z, z, z, z);
        }
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
        if ((pa4.x | pb4.x) >= 0) eval_voxel<STAGED>(crop, k, pa4.x, pb4.x, va.x, vb.x, vc.x, r0.x, r1.x, r2.x);
        if ((pa4.y | pb4.y) >= 0) eval_voxel<STAGED>(crop, k, pa4.y, pb4.y, va.y, vb.y, vc.y, r0.y, r1.y, r2.y);
        if ((pa4.z | pb4.z) >= 0) eval_voxel<STAGED>(crop, k, pa4.z, pb4.z, va.z, vb.z, vc.z, r0.z, r1.z, r2.z);
        if ((pa4.w | pb4.w) >= 0) eval_voxel<STAGED>(crop, k, pa4.w, pb4.w, va.w, vb.w, vc.w, r0.w, r1.w, r2.w);
        float* p = out + 4 * it;
        st_stream_f4(p, r0.x, r0.y, r0.z, r0.w);
        st_stream_f4(p + CH, r1.x, r1.y, r1.z, r1.w);
        st_stream_f4(p + 2 * CH, r2.x, r2.y, r2.z, r2.w);
    }
}

// ---------------------------------------------------------------- fused kernel
template <int R, int LAYOUT>
__global__ void __launch_bounds__(NT) dtsdf_fused_kernel(const BatchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameShared<R>* sh = reinterpret_cast<FrameShared<R>*>(smem_raw);
    float* crop_s = reinterpret_cast<float*>(smem_raw + sizeof(FrameShared<R>));
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        mbar_init(&sh->bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t parity = 0;

    for (int f = blockIdx.x; f < a.n_frames; f += gridDim.x) {
        // ---- frame record (pre/tsdf_numba.py:122-129)
        const int* h = a.headers + 6 * (size_t)f;
        Frame fr;
        fr.l = __ldg(h + 2); fr.t = __ldg(h + 3); fr.r = __ldg(h + 4); fr.b = __ldg(h + 5);
        fr.bw = fr.r - fr.l;
        const int bh = fr.b - fr.t;
        const int lim = 1 << 22;
        const bool bad = fr.bw <= 0 || bh <= 0 || fr.bw > 32767 || bh > 32767 ||
                         fr.l <= -lim || fr.t <= -lim || fr.r >= lim || fr.b >= lim;
        fr.n = bad ? 0 : fr.bw * bh;
        const float* crop_g = a.depth + __ldg(a.offsets + f);
        const int n_pad = (fr.n + 3) & ~3;
        const bool staged = n_pad > 0 && n_pad <= a.cap_floats;     // CTA-uniform

        // ---- 1. TMA bulk copy of the crop into shared memory
        if (staged && tid == 0) {
            fence_proxy_async();      // earlier generic-proxy reads of crop_s happen-before the async write
            mbar_expect_tx(&sh->bar, (uint32_t)n_pad * 4u);
            bulk_g2s(crop_s, crop_g, (uint32_t)n_pad * 4u, &sh->bar);
        }
        bool landed = false;

        // ---- 2. bounding box (pre/tsdf_numba.py:75-116,140-141) or caller-supplied
        float ox, oy, oz, voxel_len, trunc;
        if (a.bbox_mode == 2) {       // DTSDF_BBOX_GRID: pre/process.py:122 arguments
            const float* g = a.aux + 5 * (size_t)f;
            ox = __ldg(g); oy = __ldg(g + 1); oz = __ldg(g + 2);
            voxel_len = __ldg(g + 3); trunc = __ldg(g + 4);
            if (tid == 0) {
                if (a.max_l) a.max_l[f] = __fmul_rn(voxel_len, (float)R);
                if (a.mid_p) {
                    // centre of the cube: vox_ori + (R-1)/2 * voxel_len (informational only)
                    const float hl = __fmul_rn(voxel_len, 0.5f * (float)(R - 1));
                    a.mid_p[3 * (size_t)f + 0] = __fadd_rn(ox, hl);
                    a.mid_p[3 * (size_t)f + 1] = __fadd_rn(oy, hl);
                    a.mid_p[3 * (size_t)f + 2] = __fadd_rn(oz, hl);
                }
            }
        } else {
            float mn[3], mx[3];
            if (a.bbox_mode == 0) {   // DTSDF_BBOX_PIXELS
                mn[0] = mn[1] = mn[2] = SENTINEL;
                mx[0] = mx[1] = mx[2] = -SENTINEL;
                if (staged) {
                    mbar_wait(&sh->bar, parity);
                    landed = true;
                    bbox_scan<true>(crop_s, fr, a.focal, a.center_x, a.center_y, mn, mx);
                } else {
                    bbox_scan<false>(crop_g, fr, a.focal, a.center_x, a.center_y, mn, mx);
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) { mn[c] = warp_min(mn[c]); mx[c] = warp_max(mx[c]); }
                if (lane == 0) {
                    *reinterpret_cast<float4*>(&sh->red[warp][0]) = make_float4(mn[0], mn[1], mn[2], mx[0]);
                    *reinterpret_cast<float2*>(&sh->red[warp][4]) = make_float2(mx[1], mx[2]);
                }
                __syncthreads();
#pragma unroll
                for (int w = 0; w < NWARP; ++w) {
                    const float4 p = *reinterpret_cast<const float4*>(&sh->red[w][0]);
                    const float2 q = *reinterpret_cast<const float2*>(&sh->red[w][4]);
                    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
                    mx[0] = fmaxf(mx[0], p.w); mx[1] = fmaxf(mx[1], q.x); mx[2] = fmaxf(mx[2], q.y);
                }
            } else {                  // DTSDF_BBOX_EXTREMA: {max xyz, min xyz}, pre/process.py:120
                const float* g = a.aux + 6 * (size_t)f;
                mx[0] = __ldg(g); mx[1] = __ldg(g + 1); mx[2] = __ldg(g + 2);
                mn[0] = __ldg(g + 3); mn[1] = __ldg(g + 4); mn[2] = __ldg(g + 5);
            }
            // ---- grid parameters (pre/process.py:91-96 == pre/tsdf_numba.py:142-147)
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
            if (tid == 0) {
                if (a.max_l) a.max_l[f] = max_l;
                if (a.mid_p) {
                    a.mid_p[3 * (size_t)f + 0] = mid0;
                    a.mid_p[3 * (size_t)f + 1] = mid1;
                    a.mid_p[3 * (size_t)f + 2] = mid2;
                }
            }
        }

        // ---- 3. per-axis voxel centres and the two projection tables
        //         (pre/process.py:150-159 == pre/tsdf_numba.py:26-32,36)
        if (tid < R) {
            const float s = __fmul_rn((float)tid, voxel_len);
            sh->vx[tid] = __fadd_rn(ox, s);
            sh->vy[tid] = __fadd_rn(oy, s);
            sh->vz[tid] = __fadd_rn(oz, s);
        }
        for (int e = tid; e < R * R; e += NT) {
            const int ea = e / R, eb = e % R;
            const int zi = (LAYOUT == 0) ? eb : ea;      // pxt: [x][z] / [z][x]; pyt: [y][z] / [z][y]
            const int xi = (LAYOUT == 0) ? ea : eb;      // (also the y index for pyt)
            const float vz = __fadd_rn(oz, __fmul_rn((float)zi, voxel_len));
            const float vx = __fadd_rn(ox, __fmul_rn((float)xi, voxel_len));
            const float vy = __fadd_rn(oy, __fmul_rn((float)xi, voxel_len));
            const float coeff = __fdiv_rn(-a.focal, vz);
            const float fpx = __fadd_rn(__fmul_rn(vx, coeff), (float)a.center_x);
            const float fpy = __fadd_rn(__fmul_rn(-vy, coeff), (float)a.center_y);
            const int px = __float2int_rz(fpx);          // int(): truncation toward zero
            const int py = __float2int_rz(fpy);
            const bool okx = !bad && fabsf(fpx) < 2e9f && px >= fr.l && px < fr.r;   // non-finite -> outside
            const bool oky = !bad && fabsf(fpy) < 2e9f && py >= fr.t && py < fr.b;
            sh->pxt[e] = okx ? (short)(px - fr.l) : (short)-1;
            sh->pyt[e] = oky ? (short)(py - fr.t) : (short)-1;
        }
        if (staged && !landed) mbar_wait(&sh->bar, parity);
        if (staged) parity ^= 1u;
        __syncthreads();

        // ---- 4. voxel loop
        VoxConst k;
        k.bw = fr.bw;
        k.ix = MAGIC_I + fr.l;
        k.iy = MAGIC_I + fr.t;
        k.kx = MAGIC_F + (float)a.center_x;
        k.ky = MAGIC_F + (float)a.center_y;
        k.focal = a.focal;
        k.trunc = trunc;
        float* out = a.vol + (size_t)f * 3 * R * R * R;
        if (staged) voxel_loop<R, LAYOUT, true>(crop_s, sh, k, out);
        else        voxel_loop<R, LAYOUT, false>(crop_g, sh, k, out);
        __syncthreads();              // crop_s and the tables are rewritten by the next frame
    }
}

template <int R, int LAYOUT>
cudaError_t launch_fused(const BatchArgs& args_in, int max_crop_px, cudaStream_t stream) {
    BatchArgs args = args_in;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int smem_optin = 0, sms = 0;
    e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    const int fixed = (int)sizeof(FrameShared<R>);
    // staging capacity: the largest crop of the batch when known, capped so at least two
    // CTAs fit per SM when that is enough for the crop; oversize crops take the direct path.
    long long want = max_crop_px > 0 ? (((long long)max_crop_px + 3) & ~3ll) * 4 : (long long)smem_optin;
    long long avail = (long long)smem_optin - fixed - 1024;
    if (want > avail) want = avail;
    if (want < 16) want = 16;
    want &= ~15ll;
    args.cap_floats = (int)(want / 4);
    const size_t smem = (size_t)fixed + (size_t)want;
    auto kern = dtsdf_fused_kernel<R, LAYOUT>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    long long grid = (long long)occ * sms;        // persistent: a whole number of waves
    if (grid > args.n_frames) grid = args.n_frames;
    kern<<<(unsigned)grid, NT, smem, stream>>>(args);
    ++g_launch_count;
    return cudaGetLastError();
}

// ---------------------------------------------------------------- point cloud (f64)
// pre/process.py:30-68: X = ((u+left) - w/2)*d / f, Y = -(((v+top) - h/2)*d) / f, Z = -d in
// float64, keeping the pixels with d != 0 in row-major order.  One CTA per frame; block-wide
// stream compaction with ballot + prefix sums.
__global__ void __launch_bounds__(NT) point_cloud_kernel(int n_frames, double focal,
                                                         const float* __restrict__ depth,
                                                         const long long* __restrict__ offsets,
                                                         const int* __restrict__ headers,
                                                         double* __restrict__ points,
                                                         const long long* __restrict__ point_offsets,
                                                         int* __restrict__ n_valid) {
    __shared__ int warp_cnt[NWARP];
    __shared__ int base_s;
    const int f = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* h = headers + 6 * (size_t)f;
    const int w = h[0], hh = h[1], l = h[2], t = h[3], r = h[4], b = h[5];
    const int bw = r - l, bh = b - t;
    const int n = (bw > 0 && bh > 0) ? bw * bh : 0;
    const float* crop = depth + offsets[f];
    double* out = points ? points + 3 * point_offsets[f] : nullptr;
    const double half_w = (double)w / 2.0, half_h = (double)hh / 2.0;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int p0 = 0; p0 < n; p0 += NT) {
        const int pos = p0 + tid;
        float d = 0.0f;
        if (pos < n) d = __ldg(crop + pos);
        const bool keep = (pos < n) && (d != 0.0f);            // any(row) of pre/process.py:63
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        int before = base_s;
        int total = 0;
#pragma unroll
        for (int i = 0; i < NWARP; ++i) {
            const int c = warp_cnt[i];
            if (i < warp) before += c;
            total += c;
        }
        if (keep && out) {
            const int dst = before + __popc(m & ((1u << lane) - 1u));
            const int row = pos / bw, col = pos - row * bw;
            const double dd = (double)d;
            const double u = ((double)col + (double)l) - half_w;
            const double v = ((double)row + (double)t) - half_h;
            out[3 * (size_t)dst + 0] = __ddiv_rn(__dmul_rn(u, dd), focal);
            out[3 * (size_t)dst + 1] = __ddiv_rn(-__dmul_rn(v, dd), focal);
            out[3 * (size_t)dst + 2] = -dd;
        }
        __syncthreads();
        if (tid == 0) base_s += total;
        __syncthreads();
    }
    if (tid == 0) n_valid[f] = base_s;
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
    for (int k = tid; k < point_num; k += NT) {
        const int i = idx ? idx[k] : k;
        const double x = src[3 * (size_t)i], y = src[3 * (size_t)i + 1], z = src[3 * (size_t)i + 2];
        if (dst) { dst[3 * (size_t)k] = x; dst[3 * (size_t)k + 1] = y; dst[3 * (size_t)k + 2] = z; }
        const float xf = __double2float_rn(x), yf = __double2float_rn(y), zf = __double2float_rn(z);
        mn[0] = fminf(mn[0], xf); mx[0] = fmaxf(mx[0], xf);
        mn[1] = fminf(mn[1], yf); mx[1] = fmaxf(mx[1], yf);
        if (z != 0.0) { mn[2] = fminf(mn[2], zf); mx[2] = fmaxf(mx[2], zf); }
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

}  // namespace

// ---------------------------------------------------------------- launchers
cudaError_t launch_voxelize(int voxel_res, int layout, BatchArgs args, int max_crop_px,
                            cudaStream_t stream) {
    if (args.n_frames <= 0) return cudaSuccess;
#define DTSDF_CASE(RR)                                                                   \
    case RR:                                                                             \
        return layout == 0 ? launch_fused<RR, 0>(args, max_crop_px, stream)              \
                           : launch_fused<RR, 1>(args, max_crop_px, stream);
    switch (voxel_res) {
        DTSDF_CASE(32)
        DTSDF_CASE(48)
        DTSDF_CASE(64)
        default:
            return cudaErrorInvalidValue;
    }
#undef DTSDF_CASE
}

cudaError_t launch_point_cloud(int n_frames, double focal, const float* depth,
                               const long long* offsets, const int* headers, double* points,
                               const long long* point_offsets, int* n_valid,
                               cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    point_cloud_kernel<<<n_frames, NT, 0, stream>>>(n_frames, focal, depth, offsets, headers, points,
                                                    point_offsets, n_valid);
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

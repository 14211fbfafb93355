// Volume codec kernels (sm_100a): the lossless packed form of a D-TSDF volume, see dtsdf_internal.h.
//
// A voxel whose three channels are all +0.0, all +1.0 or all -1.0 (about 90 % of a volume: everything
// outside the crop, without depth, or farther than the truncation distance from the surface,
// pre/process.py:159,167,180-194) becomes a 2-bit code; the rest are "literals" (three floats).
// Two containers share the codes:
//   compact   literals of all frames back to back, per frame x[n], y[n], z[n] (device -> host over PCIe,
//             expanded by dtsdf_decode.cpp): pack_count / pack_scan / pack_write / pack_push;
//   strided   a fixed-stride literal region per frame holding interleaved (x, y, z) triples and a
//             per-frame count (device -> device over NVLink in the gather of volumes to a training
//             rank): pack_strided_kernel writes it in ONE pass over the volume,
//             unpack_strided_kernel - run by the receiving rank, one CTA per frame - pulls a frame's
//             codes and literals straight out of the sender's memory with two TMA bulk copies
//             (cp.async.bulk; peer addresses route over NVLink) and writes the dense planes with whole
//             512-byte warp stores.  Only the bytes in use cross the link: ~47 KB instead of 393 KB.
#include "dtsdf_internal.h"
#include "dtsdf_ptx.cuh"

#include <algorithm>

namespace dtsdf {

namespace {

using namespace ptx;

constexpr int NT = 256;                 // threads per CTA
constexpr int NWARP = NT / 32;
constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------- codes
// One CTA per frame; a thread handles 4 consecutive voxels (one LDG.128 per channel plane, one code byte).
__device__ __forceinline__ unsigned voxel_code(unsigned a, unsigned b, unsigned c) {
    const bool same = a == b && b == c;
    return (same && a == 0u) ? 0u : (same && a == 0x3F800000u) ? 1u : (same && a == 0xBF800000u) ? 2u : 3u;
}

__device__ __forceinline__ unsigned group_codes(const float* __restrict__ vol, int r3, int g, uint4& a,
                                                uint4& b, uint4& c) {
    a = __ldg(reinterpret_cast<const uint4*>(vol) + g);
    b = __ldg(reinterpret_cast<const uint4*>(vol + r3) + g);
    c = __ldg(reinterpret_cast<const uint4*>(vol + 2 * (size_t)r3) + g);
    return voxel_code(a.x, b.x, c.x) | (voxel_code(a.y, b.y, c.y) << 2) | (voxel_code(a.z, b.z, c.z) << 4) |
           (voxel_code(a.w, b.w, c.w) << 6);
}

__device__ __forceinline__ unsigned literal_mask(unsigned byte) {     // bit 2j set <=> voxel j has code 3
    return byte & (byte >> 1) & 0x55u;
}

// ---------------------------------------------------------------- compact container
__global__ void __launch_bounds__(NT) pack_count_kernel(const float* __restrict__ vol, int r3,
                                                        unsigned* __restrict__ counts) {
    __shared__ unsigned wsum[NWARP];
    const float* v = vol + (size_t)blockIdx.x * 3 * r3;
    unsigned n = 0;
    for (int g = threadIdx.x; g < r3 / 4; g += NT) {
        uint4 a, b, c;
        n += __popc(literal_mask(group_codes(v, r3, g, a, b, c)));
    }
    n = __reduce_add_sync(FULL, n);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < NWARP; ++w) t += wsum[w];
        counts[blockIdx.x] = t;
    }
}

__global__ void pack_scan_kernel(const unsigned* __restrict__ counts, int n, unsigned long long* __restrict__ off,
                                 unsigned long long* __restrict__ off_copy) {
    // batches are a few hundred frames: one thread is plenty
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < n; ++i) {
            off[i] = run;
            if (off_copy) off_copy[i] = run;
            run += counts[i];
        }
        off[n] = run;
        if (off_copy) off_copy[n] = run;
    }
}

__global__ void __launch_bounds__(NT) pack_write_kernel(const float* __restrict__ vol, int r3,
                                                        const unsigned long long* __restrict__ off_dev,
                                                        unsigned char* __restrict__ codes, float* __restrict__ lits) {
    __shared__ unsigned wsum[NWARP];
    const int f = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* v = vol + (size_t)f * 3 * r3;
    unsigned char* cd = codes + (size_t)f * (r3 / 4);
    // the frame's literal block: x of its nf literal voxels, then y, then z, in voxel order
    const unsigned long long o0 = off_dev[f];
    const unsigned nf = (unsigned)(off_dev[f + 1] - o0);
    unsigned* lx = reinterpret_cast<unsigned*>(lits) + 3 * o0;
    unsigned* ly = lx + nf;
    unsigned* lz = ly + nf;
    unsigned base = 0;
    for (int g0 = 0; g0 < r3 / 4; g0 += NT) {            // r3/4 is a multiple of NT for R = 32, 48, 64
        const int g = g0 + threadIdx.x;
        uint4 a, b, c;
        const unsigned byte = group_codes(v, r3, g, a, b, c);
        cd[g] = (unsigned char)byte;
        const unsigned m2 = literal_mask(byte);
        const unsigned nl = __popc(m2);
        // block-wide exclusive prefix of the literal voxels, in voxel order
        unsigned incl = nl;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) {
            const unsigned t = wsum[w];
            if (w < warp) before += t;
            total += t;
        }
        if (nl) {
            unsigned r = base + before + incl - nl;
            const unsigned av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((m2 >> (2 * j)) & 1u) { lx[r] = av[j]; ly[r] = bv[j]; lz[r] = cv[j]; ++r; }
        }
        base += total;
        __syncthreads();
    }
}

// Copies the literal stream (3 * lit_off[n] floats, a count only the device knows) from device memory
// to host-mapped memory with whole 16-byte, fully coalesced stores: PCIe likes 128-byte writes, the
// scattered 4-byte stores of pack_write_kernel it does not.
__global__ void __launch_bounds__(NT) pack_push_kernel(const float* __restrict__ lits, const unsigned long long* __restrict__ off,
                                                       int n, float* __restrict__ host_lits) {
    const unsigned long long n16 = (3ull * off[n] + 3ull) / 4ull;          // uint4 units (the buffers are padded)
    const uint4* src = reinterpret_cast<const uint4*>(lits);
    uint4* dst = reinterpret_cast<uint4*>(host_lits);
    for (unsigned long long i = blockIdx.x * (unsigned long long)NT + threadIdx.x; i < n16;
         i += (unsigned long long)gridDim.x * NT)
        dst[i] = src[i];
}

// ---------------------------------------------------------------- strided container
// One pass over the volume: codes, interleaved literal triples at the frame's own region, count.
// The next tile's loads are in flight while this tile's block-wide prefix is computed.
__global__ void __launch_bounds__(NT) pack_strided_kernel(const float* __restrict__ vol, int r3, long long lit_stride,
                                                          unsigned char* __restrict__ codes, unsigned* __restrict__ cnt,
                                                          float* __restrict__ lits) {
    __shared__ unsigned wsum[2][NWARP];
    const int f = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* v = vol + (size_t)f * 3 * r3;
    unsigned char* cd = codes + (size_t)f * (r3 / 4);
    unsigned* lt = reinterpret_cast<unsigned*>(lits + (size_t)f * lit_stride);
    const int ntile = r3 / (4 * NT);                     // r3/4 is a multiple of NT for R = 32, 48, 64
    unsigned base = 0;
    uint4 a, b, c;
    unsigned byte = group_codes(v, r3, threadIdx.x, a, b, c);
    for (int t = 0; t < ntile; ++t) {
        const int g = t * NT + threadIdx.x;
        uint4 na = a, nb = b, nc = c;
        unsigned nbyte = 0;
        if (t + 1 < ntile) nbyte = group_codes(v, r3, g + NT, na, nb, nc);
        cd[g] = (unsigned char)byte;
        const unsigned m2 = literal_mask(byte);
        const unsigned nl = __popc(m2);
        unsigned incl = nl;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned s = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += s;
        }
        if (lane == 31) wsum[t & 1][warp] = incl;
        __syncthreads();               // one barrier per tile: the partial sums are double-buffered
        unsigned before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) {
            const unsigned s = wsum[t & 1][w];
            if (w < warp) before += s;
            total += s;
        }
        if (nl) {
            unsigned r = 3u * (base + before + incl - nl);
            const unsigned av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((m2 >> (2 * j)) & 1u) { lt[r] = av[j]; lt[r + 1] = bv[j]; lt[r + 2] = cv[j]; r += 3; }
        }
        base += total;
        a = na; b = nb; c = nc; byte = nbyte;
    }
    if (threadIdx.x == 0) cnt[f] = base;
}

__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Expansion of one frame.  codes / cnt / lits may live in another GPU's memory (peer mapping).
//   0  thread 0 reads the frame's literal count and issues the two bulk copies: the code bytes and, when
//      they fit the staging buffer, the literal triples in use (rounded up to 16 bytes: the region is
//      padded).  A frame whose literals do not fit reads them from where they are, with plain loads.
//   1  every warp owns a contiguous run of groups (4 voxels = one code byte each): it counts the
//      literals of its run in the staged codes; one barrier; the runs before it give its literal base.
//   2  the warp sweeps its run 32 groups (512 contiguous output bytes per channel) at a time: three ballots
//      give every lane the number of literals in the lanes before it; codes 0/1/2 become 0 / +1 / -1,
//      literals are fetched by index; one 16-byte store per lane and channel.
template <bool STAGED>
__device__ __forceinline__ void unpack_run(const unsigned char* __restrict__ codes_s, const float* __restrict__ lit,
                                           unsigned base, unsigned last, int g0, int g1, int r3, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll 2
    for (int g = g0 + lane; g < g1; g += 32) {
        const unsigned byte = codes_s[g];
        const unsigned m2 = literal_mask(byte);
        const unsigned nl = __popc(m2);
        const unsigned b0 = __ballot_sync(FULL, nl & 1u), b1 = __ballot_sync(FULL, nl & 2u), b2 = __ballot_sync(FULL, nl & 4u);
        unsigned r = base + __popc(b0 & lt) + 2u * __popc(b1 & lt) + 4u * __popc(b2 & lt);
        base += __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
        unsigned x[4], y[4], z[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned c = (byte >> (2 * j)) & 3u;
            unsigned v = (c & 1u) ? 0x3F800000u : ((c & 2u) ? 0xBF800000u : 0u);
            x[j] = y[j] = z[j] = v;
            if (c == 3u) {
                const unsigned q = 3u * min(r, last);          // (a count that contradicts the codes must not fault)
                x[j] = __float_as_uint(lit[q]); y[j] = __float_as_uint(lit[q + 1]); z[j] = __float_as_uint(lit[q + 2]);
                ++r;
            }
        }
        float* p = out + 4 * (size_t)g;
        st_f4(p, make_float4(__uint_as_float(x[0]), __uint_as_float(x[1]), __uint_as_float(x[2]), __uint_as_float(x[3])));
        st_f4(p + r3, make_float4(__uint_as_float(y[0]), __uint_as_float(y[1]), __uint_as_float(y[2]), __uint_as_float(y[3])));
        st_f4(p + 2 * (size_t)r3, make_float4(__uint_as_float(z[0]), __uint_as_float(z[1]), __uint_as_float(z[2]), __uint_as_float(z[3])));
    }
}

__global__ void __launch_bounds__(NT) unpack_strided_kernel(const unsigned char* __restrict__ codes,
                                                            const unsigned* __restrict__ cnt,
                                                            const float* __restrict__ lits, long long lit_stride, int r3,
                                                            int lit_cap_bytes, float* __restrict__ out_all) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long bar;
    __shared__ unsigned wsum[NWARP];
    __shared__ unsigned s_n;
    const int f = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nbytes = r3 / 4;                               // code bytes per frame, a multiple of 16
    unsigned char* codes_s = smem_raw;
    float* lits_s = reinterpret_cast<float*>(smem_raw + nbytes);
    const float* lits_g = lits + (size_t)f * lit_stride;
    if (threadIdx.x == 0) {
        unsigned n = ld_relaxed_sys(cnt + f);
        if (n > (unsigned)r3) n = (unsigned)r3;
        const unsigned lbytes = (12u * n + 15u) & ~15u;
        const bool staged = lbytes <= (unsigned)lit_cap_bytes;
        mbar_init(&bar, 1);
        fence_mbar_init();
        mbar_expect_tx(&bar, (uint32_t)nbytes + (staged ? lbytes : 0u));
        bulk_g2s(codes_s, codes + (size_t)f * nbytes, (uint32_t)nbytes, &bar);
        if (staged && lbytes) bulk_g2s(lits_s, lits_g, lbytes, &bar);
        s_n = n;
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    const unsigned n = s_n;
    const bool staged = ((12u * n + 15u) & ~15u) <= (unsigned)lit_cap_bytes;
    const int gw = nbytes / NWARP;                           // groups per warp: a multiple of 32
    {
        const unsigned* cw = reinterpret_cast<const unsigned*>(codes_s + warp * gw);
        unsigned c = 0;
        for (int i = lane; i < gw / 4; i += 32) {
            const unsigned w = cw[i];
            c += __popc(w & (w >> 1) & 0x55555555u);
        }
        c = __reduce_add_sync(FULL, c);
        if (lane == 0) wsum[warp] = c;
    }
    __syncthreads();
    unsigned base = 0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) base += w < warp ? wsum[w] : 0u;
    const unsigned last = n ? n - 1u : 0u;
    float* out = out_all + (size_t)f * 3 * r3;
    if (staged) unpack_run<true>(codes_s, lits_s, base, last, warp * gw, (warp + 1) * gw, r3, out);
    else        unpack_run<false>(codes_s, lits_g, base, last, warp * gw, (warp + 1) * gw, r3, out);
}

}  // namespace

cudaError_t launch_pack_push(int n_frames, const float* lits, const unsigned long long* lit_off, float* host_lits,
                             cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    pack_push_kernel<<<96, NT, 0, stream>>>(lits, lit_off, n_frames, host_lits);
    ++g_launch_count;
    return cudaGetLastError();
}

cudaError_t launch_pack(int n_frames, int r3, const float* vol, unsigned* counts, unsigned long long* lit_off,
                        unsigned long long* lit_off_copy, unsigned char* codes, float* lits, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    if (r3 % (4 * NT) != 0) return cudaErrorInvalidValue;
    pack_count_kernel<<<n_frames, NT, 0, stream>>>(vol, r3, counts);
    pack_scan_kernel<<<1, 32, 0, stream>>>(counts, n_frames, lit_off, lit_off_copy);
    pack_write_kernel<<<n_frames, NT, 0, stream>>>(vol, r3, lit_off, codes, lits);
    g_launch_count += 3;
    return cudaGetLastError();
}

// Strided container: literal region of frame f at lits + f * lit_stride (floats, a multiple of 4).
cudaError_t launch_pack_strided(int n_frames, int r3, const float* vol, long long lit_stride, unsigned char* codes,
                                unsigned* cnt, float* lits, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    if (r3 % (4 * NT) != 0 || lit_stride < 3ll * r3 || (lit_stride & 3)) return cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(vol) & 15) != 0) return cudaErrorMisalignedAddress;
    pack_strided_kernel<<<n_frames, NT, 0, stream>>>(vol, r3, lit_stride, codes, cnt, lits);
    ++g_launch_count;
    return cudaGetLastError();
}

cudaError_t launch_unpack_strided(int n_frames, int r3, const unsigned char* codes, const unsigned* cnt, const float* lits,
                                  long long lit_stride, float* out, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    if (r3 % (4 * 32 * NWARP) != 0 || lit_stride < 3ll * r3 || (lit_stride & 3)) return cudaErrorInvalidValue;   // whole warp steps per run
    if ((reinterpret_cast<uintptr_t>(out) & 15) != 0 || (reinterpret_cast<uintptr_t>(codes) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(lits) & 15) != 0)
        return cudaErrorMisalignedAddress;
    int dev = 0, smem_optin = 0, smem_sm = 0, sms = 0, l2 = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev)) != cudaSuccess) return e;
    static std::atomic<unsigned> smem_ok{0u};
    if ((e = allow_max_smem(unpack_strided_kernel, dev, smem_optin, smem_ok)) != cudaSuccess) return e;
    // CTAs per SM: as in launch_fused, as many frames in flight as keep the volumes under construction inside
    // the L2, between 1 and 4; what is left of the SM's shared memory stages the literals
    const int nbytes = r3 / 4;
    const long long fit = (long long)l2 / ((long long)sms * 12ll * r3);
    int target = (int)std::max(1ll, std::min(4ll, fit));
    if (fit < 1 && r3 < 64 * 64 * 64) target = 2;
    long long smem = ((long long)smem_sm / target - 1024) & ~15ll;
    if (smem > smem_optin - 256) smem = (smem_optin - 256) & ~15ll;
    const long long cap = smem - nbytes;
    if (cap < 16) return cudaErrorInvalidValue;
    unpack_strided_kernel<<<n_frames, NT, (size_t)smem, stream>>>(codes, cnt, lits, lit_stride, r3, (int)cap, out);
    ++g_launch_count;
    return cudaGetLastError();
}

}  // namespace dtsdf

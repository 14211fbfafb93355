// Dev probe: DRAM write bandwidth of a few store patterns (constant vs incompressible data).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ float4 val(size_t i, int rnd) {
    if (!rnd) return make_float4(1.f, 1.f, 1.f, 1.f);
    unsigned h = (unsigned)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    return make_float4(__uint_as_float(0x3f000000u | (h & 0x7fffff)), __uint_as_float(0x3f000000u | ((h * 3u) & 0x7fffff)),
                       __uint_as_float(0x3f000000u | ((h * 5u) & 0x7fffff)), __uint_as_float(0x3f000000u | ((h * 7u) & 0x7fffff)));
}
// K1: grid-stride, every thread one float4 per step, whole grid sweeps linearly
__global__ void k_linear(float4* out, size_t n4, int rnd) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) out[i] = val(i, rnd);
}
// K3: persistent CTAs, each writes whole frames (frame = fb float4) sequentially
__global__ void k_frames(float4* out, int nframes, int fb, int rnd) {
    for (int f = blockIdx.x; f < nframes; f += gridDim.x) {
        float4* o = out + (size_t)f * fb;
        for (int i = threadIdx.x; i < fb; i += blockDim.x) o[i] = val((size_t)f * fb + i, rnd);
    }
}
// K5: the voxeliser's traffic mix with no arithmetic: one CTA per frame reads its crop (62.4 KB), then
// writes its volume (393 KB); the sum of the crop goes into the first word so the reads stay.
__global__ void k_mix(const float4* crops, int cb, float4* out, int fb, int rnd) {
    __shared__ float red[32];
    const float4* c = crops + (size_t)blockIdx.x * cb;
    float s = 0.f;
    for (int i = threadIdx.x; i < cb; i += blockDim.x) { float4 v = __ldg(c + i); s += v.x + v.y + v.z + v.w; }
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < (int)blockDim.x / 32; ++w) t += red[w];
    float4* o = out + (size_t)blockIdx.x * fb;
    for (int i = threadIdx.x; i < fb; i += blockDim.x) { float4 v = val((size_t)blockIdx.x * fb + i, rnd); v.x += t * 1e-30f; o[i] = v; }
}
int main() {
    const int nframes = 8500, fb = 3 * 32 * 32 * 32 / 4;
    const size_t n4 = (size_t)nframes * fb;
    float4* out; CK(cudaMalloc(&out, n4 * 16));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto run = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; ++i) launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 20; ++i) launch();
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 20;
        printf("%-44s %.4f ms  %.0f GB/s\n", name, ms, n4 * 16 / ms / 1e6);
        CK(cudaGetLastError());
    };
    for (int rnd = 0; rnd < 2; ++rnd) {
        printf("-- data %s\n", rnd ? "incompressible" : "constant");
        run("linear grid-stride 148*8 x 256", [&] { k_linear<<<148 * 8, 256>>>(out, n4, rnd); });
        run("linear one-shot (n4/256 CTAs)", [&] { k_linear<<<(unsigned)((n4 + 255) / 256), 256>>>(out, n4, rnd); });
        run("frames persistent 444 x 256", [&] { k_frames<<<444, 256>>>(out, nframes, fb, rnd); });
        run("frames persistent 148 x 1024", [&] { k_frames<<<148, 1024>>>(out, nframes, fb, rnd); });
        run("frames one CTA per frame 8500 x 256", [&] { k_frames<<<nframes, 256>>>(out, nframes, fb, rnd); });
        run("frames one CTA per frame 8500 x 1024", [&] { k_frames<<<nframes, 1024>>>(out, nframes, fb, rnd); });
    }
    {
        const int cb = 120 * 130 / 4;
        float4* crops; CK(cudaMalloc(&crops, (size_t)nframes * cb * 16));
        CK(cudaMemset(crops, 0, (size_t)nframes * cb * 16));
        printf("-- traffic mix of the voxeliser (62.4 KB read + 393 KB written per CTA), bytes counted: both\n");
        for (int nt : {256, 512, 1024}) {
            char name[64]; snprintf(name, sizeof name, "mix one CTA per frame 8500 x %d", nt);
            for (int i = 0; i < 3; ++i) k_mix<<<nframes, nt>>>(crops, cb, out, fb, 1);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            for (int i = 0; i < 20; ++i) k_mix<<<nframes, nt>>>(crops, cb, out, fb, 1);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 20;
            printf("%-44s %.4f ms  %.0f GB/s\n", name, ms, ((double)n4 * 16 + (double)nframes * cb * 16) / ms / 1e6);
        }
    }
    return 0;
}

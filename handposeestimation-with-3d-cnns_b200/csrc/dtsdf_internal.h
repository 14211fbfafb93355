// Internal interface between the C ABI (dtsdf_capi.cu) and the kernels (dtsdf_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <vector>

namespace dtsdf {

// Arguments of the fused voxeliser, passed by value to the kernel.
struct BatchArgs {
    const float*     depth;      // packed crops
    const long long* offsets;    // [n] start of crop i in floats (multiple of 4)
    const int*       headers;    // [n][6]
    const float*     aux;        // [n][6] extrema, [n][5] grid params, or nullptr
    float*           vol;        // [n][3][R][R][R]
    float*           max_l;      // [n] or nullptr
    float*           mid_p;      // [n][3] or nullptr
    int              n_frames;
    int              cap_floats; // shared-memory staging capacity (floats, multiple of 4)
    int              bbox_mode;
    int              center_x;
    int              center_y;
    float            focal;
    float            trunc_mul;
    const float*     joints;     // [n][63] ground-truth joints to normalise in the same launch, or nullptr
    float*           out_joints; // [n][63]
    int              joint_flags; // bit 0: clamp to [0,1]; bit 1: negate z first (3D_CNN/dataset.py:107)
    // bbox_mode 3 (DTSDF_BBOX_SAMPLED, DataProcess.process() in one launch): the bounding box is that of the
    // point cloud resampled by the caller's index vectors (pre/process.py:30-84,102-120), float64
    const int*       sample_idx = nullptr;   // [n][point_num] ranks into each frame's valid pixels (row-major order)
    int              point_num = 0;
    double           focal64 = 0.0;          // the focal length of the float64 back-projection (pre/process.py:9)
};

// Picks the staging capacity / grid and launches the fused kernel.  Returns a cudaError_t.
cudaError_t launch_voxelize(int voxel_res, int layout, BatchArgs args, int max_crop_px,
                            cudaStream_t stream);

// Augmented D-TSDF (oracle/dtsdf_oracle.c: oracle_dtsdf_aug_frame is the specification).
// aug: [n][21] = A (3x3 row-major), Ainv, c.
cudaError_t launch_voxelize_aug(int voxel_res, int layout, BatchArgs args, const float* aug, int max_crop_px,
                                cudaStream_t stream);

// n_valid[i] = pixels of crop i with d != 0 (the rows pre/process.py:61-66 keeps)
cudaError_t launch_count_valid(int n_frames, const float* depth, const long long* offsets, const int* headers,
                               int* n_valid, cudaStream_t stream);

cudaError_t launch_point_cloud(int n_frames, double focal, const float* depth,
                               const long long* offsets, const int* headers, double* points,
                               const long long* point_offsets, int* n_valid,
                               cudaStream_t stream);

// point_cloud + gather by index + extrema without the full cloud (crops of at most sample_cloud_max_pixels())
int sample_cloud_max_pixels();
cudaError_t launch_sample_cloud(int n_frames, double focal, int point_num, const float* depth, const long long* offsets,
                                const int* headers, const int* indices, double* out_pc, float* out_extrema,
                                int* n_valid, cudaStream_t stream);

cudaError_t launch_resample_extrema(int n_frames, int point_num, const double* points,
                                    const long long* point_offsets, const int* indices,
                                    double* out_pc, float* out_extrema, cudaStream_t stream);

cudaError_t launch_normalize_joints(int n_frames, const float* joints, const float* max_l,
                                    const float* mid_p, int clamp01, float* out,
                                    cudaStream_t stream);

// Volume codec for the device -> host transfer (host entry points).  A voxel whose three channels
// are all +0.0, all +1.0 or all -1.0 (about 90 % of a D-TSDF volume) becomes a 2-bit code 0/1/2,
// anything else code 3 = "literal": its three floats go to the frame's literal block.  Lossless.
//   counts   [n]     literal voxels per frame (device scratch)
//   lit_off  [n+1]   exclusive prefix of counts (device); lit_off_copy: optional second copy
//   codes    [n][R3/4] bytes: voxel 4k+j of the frame in bits 2j..2j+1 of byte k
//   lits     [3 * lit_off[n]] floats; frame f's block starts at float 3 * lit_off[f] and holds, for its
//            nf = lit_off[f+1] - lit_off[f] literal voxels in voxel order, x[nf], y[nf], z[nf]
//            (12 bytes per literal voxel; literals per GROUP of four voxels, the first form, were 64 %
//            more bytes on real volumes).  Buffers need 3 * R3 floats per frame + 16 bytes.
cudaError_t launch_pack(int n_frames, int r3, const float* vol, unsigned* counts,
                        unsigned long long* lit_off, unsigned long long* lit_off_copy,
                        unsigned char* codes, float* lits, cudaStream_t stream);

// literal stream device -> host-mapped memory: 3 * lit_off[n] floats, rounded up to whole 16 bytes
cudaError_t launch_pack_push(int n_frames, const float* lits, const unsigned long long* lit_off,
                             float* host_lits, cudaStream_t stream);

// The strided container of the same codes, for device -> device transfers (the gather of volumes to a
// training rank over NVLink): frame f's literals are interleaved (x, y, z) triples in voxel order at
// lits + f * lit_stride (floats; lit_stride >= 3 * R3, a multiple of 4), cnt[f] of them.  Written in one
// pass over the volume; no prefix over frames, so every frame's region can be read on its own.
cudaError_t launch_pack_strided(int n_frames, int r3, const float* vol, long long lit_stride, unsigned char* codes,
                                unsigned* cnt, float* lits, cudaStream_t stream);
// codes / cnt / lits may be peer memory: the kernel pulls each frame with TMA bulk copies.
cudaError_t launch_unpack_strided(int n_frames, int r3, const unsigned char* codes, const unsigned* cnt,
                                  const float* lits, long long lit_stride, float* out, cudaStream_t stream);

// Host side of the codec (dtsdf_decode.cpp, plain C++): expands one packed frame (its r3/4 code bytes and
// its literal block of nlit voxels: x[nlit], y[nlit], z[nlit]) into the three channel planes at `out`
// (any 4-byte alignment; whole-line non-temporal stores when `out` is 16-byte aligned).
void decode_frame(const unsigned char* codes, const float* lits, size_t nlit, size_t r3, float* out);
const char* decode_isa();

// NumPy's legacy global stream (dtsdf_rng.cpp): `count` draws of np.random.randint(0, high) from the MT19937
// state (key[624], pos as in np.random.get_state()), written to out; the state advances as NumPy's would.
void np_legacy_randint(uint32_t* key, int* pos, uint32_t high, int count, int32_t* out);

// The same draws in two phases, so that only the cheap half is sequential.  The stream is ONE MT19937 and every
// frame's draws start where the previous frame's rejection sampling stopped, so WHERE a frame's draws begin must
// be found in order - but that needs no output: np_legacy_randint_skip steps the generator and counts acceptances
// (tempering + compare + popcount; no compress-store, nothing written) after leaving a snapshot of the state in the
// span; draw_fill then produces a frame's draws from its snapshot, any frame on any thread.
struct alignas(64) DrawSpan {
    uint32_t key[624];        // the state the request started from
    int      pos;
    uint32_t high;            // np.random.randint(0, high, size=count)
    int32_t  count;
};
void np_legacy_randint_skip(uint32_t* key, int* pos, uint32_t high, int count, DrawSpan* span);
void draw_fill(const DrawSpan& span, int32_t* out);

// kernels launched by this process (claimed in bench.py's gpu_launches)
extern std::atomic<long long> g_launch_count;

// The opt-in dynamic shared memory limit of a kernel is a per-device function attribute: raise it to the
// device maximum once per (instantiation, device), not per launch (launches may come from several host
// threads with different crop sizes).
template <typename K>
inline cudaError_t allow_max_smem(K kern, int dev, int smem_optin, std::atomic<unsigned>& done_mask) {
    const unsigned bit = 1u << (dev & 31);
    if (done_mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - (int)fa.sharedSizeBytes);
    if (e == cudaSuccess) done_mask.fetch_or(bit, std::memory_order_release);
    return e;
}


}  // namespace dtsdf

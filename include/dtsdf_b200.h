/*
 * dtsdf_b200.h — C ABI of the B200-native projective directional TSDF (D-TSDF)
 * voxeliser: MSRA depth crop + bounding box -> 3 x R^3 float volume.
 *
 * This is the drop-in boundary for the hot path of
 * Moon0shang/HandPoseEstimation-with-3D-CNNs.  The reference has no FFI of its
 * own (it is pure Python); every entry point below names the reference callable
 * it replaces (paths relative to the reference root).  The Python modules under
 * handposeestimation-with-3d-cnns_b200/pre/ (process.py, tsdf_for.py,
 * tsdf_numba.py) bind these symbols with ctypes and keep the reference's names
 * and signatures; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C types only; no CUDA or torch types in any signature
 *     (`void* stream` is a cudaStream_t, NULL = the legacy default stream);
 *   - every function returns DTSDF_OK (0) or a negative DTSDF_ERR_* code and
 *     never throws; dtsdf_error_string() / dtsdf_last_cuda_error() explain;
 *   - "device" entry points take device pointers owned by the caller, enqueue
 *     work on `stream` and return without synchronising;
 *   - "host" entry points take host pointers, run on the context's own streams
 *     and return when the outputs are complete - on failure, too: no copy and no host
 *     thread of the library touches the caller's buffers after an entry point has
 *     returned, whatever it returned.  They leave the calling thread's current CUDA
 *     device as they found it.  Page-locked buffers (dtsdf_host_alloc) let copies and
 *     kernels overlap; pageable ones work, with the host -> device copies serialised;
 *   - there is no CPU fallback: without a usable CUDA device every compute entry
 *     point fails with DTSDF_ERR_NO_DEVICE / DTSDF_ERR_CUDA.
 *
 * Frame record (reference: pre/read_MSRA.py:155-164, pre/process.py:32-40)
 *   header  int32[6] = {img_w, img_h, left, top, right, bottom}, right/bottom
 *           exclusive: bw = right-left, bh = bottom-top
 *   depth   float32[bw*bh], row-major over the crop, millimetres, 0 = background
 *
 * Packed batch
 *   depth_packed  float32[packed_len]: the crops back to back; crop i starts at
 *                 element crop_offsets[i], which must be a multiple of 4 (16-byte
 *                 aligned, so the crop can be bulk-copied to shared memory by
 *                 TMA), and the buffer must be readable up to the next multiple
 *                 of 4 elements after each crop's end (dtsdf_pack_plan() does
 *                 both);
 *   crop_offsets  int64[n_frames];
 *   headers       int32[n_frames][6].
 */
#ifndef DTSDF_B200_H
#define DTSDF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTSDF_ABI_VERSION 1

/* error codes */
#define DTSDF_OK                 0
#define DTSDF_ERR_INVALID_ARG   -1
#define DTSDF_ERR_UNSUPPORTED   -2   /* voxel_res not in {32,48,64}, unknown layout/mode */
#define DTSDF_ERR_CUDA          -3   /* a CUDA runtime call failed, see dtsdf_last_cuda_error() */
#define DTSDF_ERR_NO_DEVICE     -4   /* no CUDA device / not an sm_100 device */
#define DTSDF_ERR_ALLOC         -5
#define DTSDF_ERR_BAD_HEADER    -6   /* a frame header has bw <= 0, bh <= 0 or a crop > 32767 px wide/high */
#define DTSDF_ERR_INTERNAL      -7   /* a C++ exception inside the library (caught at the boundary) */
#define DTSDF_ERR_EMPTY_CLOUD   -8   /* a frame without a single point where one must be drawn: the reference's
                                        np.random.randint(0, 0) raises ValueError there (pre/process.py:77,80) */

/* volume layouts */
#define DTSDF_LAYOUT_CXYZ 0   /* out[c][x][y][z]: pre/process.py:196-198, pre/tsdf_for.py:118-120,
                                 the layout of the result/ tree and of 3D_CNN/dataset.py:35 */
#define DTSDF_LAYOUT_CZYX 1   /* out[c][z][y][x]: pre/tsdf_numba.py:70-72 */

/* where the per-frame bounding box comes from */
#define DTSDF_BBOX_PIXELS  0  /* min/max over all pixels with |d| >= 1, back-projected with
                                 q=d/f, X=q*(x-cx), Y=-q*(y-cy), Z=-d: pre/tsdf_numba.py:75-116,140-141 */
#define DTSDF_BBOX_EXTREMA 1  /* caller gives aux[i] = {max_x,max_y,max_z,min_x,min_y,min_z}, the
                                 result of max_min_point(): pre/process.py:102-120, pre/tsdf_for.py:23-41 */
#define DTSDF_BBOX_GRID    2  /* caller gives aux[i] = {vox_ori_x,y,z, voxel_len, trunc}: the arguments
                                 of DataProcess.tsdf_cal(): pre/process.py:122 */
#define DTSDF_BBOX_SAMPLED 3  /* the box of the point cloud resampled by the caller's index vectors, found inside the
                                 voxeliser launch: only through dtsdf_voxelize_sampled_batch / dtsdf_process_host */

typedef struct dtsdf_config {
    int32_t voxel_res;   /* R: 32 (pre/process.py:88), 48 or 64 */
    int32_t layout;      /* DTSDF_LAYOUT_* */
    int32_t bbox_mode;   /* DTSDF_BBOX_* */
    int32_t center_x;    /* 160: pre/process.py:133, pre/tsdf_numba.py:9 */
    int32_t center_y;    /* 120: pre/process.py:134, pre/tsdf_numba.py:10 */
    float   focal;       /* 241.42: pre/process.py:7, pre/tsdf_numba.py:8 */
    float   trunc_mul;   /* 3: pre/process.py:95, pre/tsdf_numba.py:146 */
} dtsdf_config;

/* ---- library / diagnostics -------------------------------------------------- */
int         dtsdf_abi_version(void);
const char* dtsdf_error_string(int code);
const char* dtsdf_last_cuda_error(void);          /* thread-local text of the last CUDA failure */
void        dtsdf_default_config(dtsdf_config* cfg);  /* R=32, CXYZ, PIXELS, 160, 120, 241.42, 3 */
int         dtsdf_device_count(void);             /* number of CUDA devices, 0 if none */
/* number of CUDA kernels this library has launched in the calling process */
int64_t     dtsdf_launch_count(void);

/* Lets kernels running on `device` address memory that lives on `peer_device` (NVLink /
 * NVSwitch peer access), so that d_out_vol of dtsdf_voxelize_batch may point into another
 * GPU's HBM: the volumes of a frame shard are then gathered into a training rank by the
 * voxeliser's own stores (nothing to replace in the reference, which is
 * single-GPU).  Idempotent.  DTSDF_ERR_UNSUPPORTED if the two devices have no peer path. */
int dtsdf_enable_peer_access(int32_t device, int32_t peer_device);

/* A device buffer that other PROCESSES on the same node can map (CUDA IPC), for the same purpose
 * across ranks: the root rank allocates the buffer of all volumes and publishes the 64-byte
 * handle; every other rank opens it with its own GPU current and passes a pointer into it as
 * d_out_vol.  dtsdf_shared_open enables peer access to the owning GPU lazily. */
#define DTSDF_IPC_HANDLE_BYTES 64
int dtsdf_shared_alloc(int32_t device, int64_t bytes, void** d_ptr, uint8_t handle[DTSDF_IPC_HANDLE_BYTES]);
int dtsdf_shared_free(int32_t device, void* d_ptr);
int dtsdf_shared_open(int32_t device, const uint8_t handle[DTSDF_IPC_HANDLE_BYTES], void** d_ptr);
int dtsdf_shared_close(int32_t device, void* d_ptr);

/* ---- packing helpers (host, no CUDA) ---------------------------------------- */
/* Computes 16-byte-aligned crop offsets for n_frames headers.  crop_offsets[i] is
 * the start (in floats) of crop i; returns the total packed length in floats
 * (a multiple of 4) in *packed_len and the largest crop (pixels) in *max_crop_px.
 * Replaces nothing in the reference (which handles one .bin file at a time,
 * pre/read_MSRA.py:98-103); it defines the batched input format. */
int dtsdf_pack_plan(int32_t n_frames, const int32_t* headers,
                    int64_t* crop_offsets, int64_t* packed_len, int32_t* max_crop_px);

/* ---- device entry points (caller-owned device memory, asynchronous) ---------- */

/* Fused per-frame pipeline, one launch for the whole batch:
 *   bounding box (per cfg->bbox_mode) -> grid parameters -> per-voxel D-TSDF.
 * Replaces tsdf_numba.cal_tsdf_cuda (pre/tsdf_numba.py:119-161: min_max_kernel
 * :75-116, host reduce + grid parameters :140-147, tsdf_kernel :15-72) and the
 * CPU loop DataProcess.tsdf_f/tsdf_cal (pre/process.py:86-100,122-200 ==
 * pre/tsdf_for.py:6-20,44-122).
 *   d_aux        NULL for DTSDF_BBOX_PIXELS, float[n][6] for EXTREMA, float[n][5] for GRID
 *   max_crop_px  upper bound of bw*bh over the batch (a tuning hint for the shared
 *                memory staging buffer; 0 = unknown; results never depend on it)
 *   d_out_vol    float[n][3][R][R][R] in cfg->layout
 *   d_out_max_l  float[n]     (may be NULL)   side of the voxel cube, pre/process.py:93
 *   d_out_mid_p  float[n][3]  (may be NULL)   centre of the bounding box, pre/process.py:91
 */
int dtsdf_voxelize_batch(const dtsdf_config* cfg, int32_t n_frames,
                         const float* d_depth_packed, const int64_t* d_crop_offsets,
                         const int32_t* d_headers, const float* d_aux, int32_t max_crop_px,
                         float* d_out_vol, float* d_out_max_l, float* d_out_mid_p,
                         void* stream);

/* dtsdf_voxelize_batch plus the joint normalisation of the same frames in the same launch:
 * out = (gt - mid_p)/max_l + 0.5 with the frame's own mid_p / max_l (pre/joint_nor.py:8-18), which
 * the trainer recomputes per sample in a Python loop on GPU tensors (3D_CNN/train.py:236-244).
 *   d_joints      float[n][63] ground truth (21 joints x 3)
 *   joint_flags   DTSDF_JOINT_CLAMP01: clamp to [0,1] (3D_CNN/train.py:241-242);
 *                 DTSDF_JOINT_FLIP_Z: negate every z first (3D_CNN/dataset.py:107)
 *   d_out_joints  float[n][63] */
#define DTSDF_JOINT_CLAMP01 1
#define DTSDF_JOINT_FLIP_Z  2
int dtsdf_voxelize_joints_batch(const dtsdf_config* cfg, int32_t n_frames,
                                const float* d_depth_packed, const int64_t* d_crop_offsets,
                                const int32_t* d_headers, const float* d_aux, int32_t max_crop_px,
                                const float* d_joints, int32_t joint_flags,
                                float* d_out_vol, float* d_out_max_l, float* d_out_mid_p,
                                float* d_out_joints, void* stream);

/* DataProcess(frame, gt, point_num).process() of pre/process.py:13-28 in ONE launch, when the sampled point cloud
 * itself is not wanted: hand_points = the back-projected pixels with d != 0 in row-major order (float64,
 * point_cloud(), :30-68); point_clouds = hand_points[rand_idx] (set_length, :70-84); bounding box of
 * point_clouds rounded to float32 (max_min_point, :102-120); grid parameters and voxel loop (tsdf_f / tsdf_cal,
 * :86-100,122-200).  Neither cloud is ever written: the kernel ranks the valid pixels of the staged crop, marks
 * the sampled ranks in a shared-memory bitmap and reduces the extrema of the sample from the depth map itself.
 * Equal, bit for bit, to dtsdf_point_cloud_batch + dtsdf_resample_extrema_batch + dtsdf_voxelize_batch(EXTREMA).
 *   focal          float64 focal length of the back-projection (241.42, pre/process.py:9); cfg->focal is the
 *                  voxel loop's float32 one (pre/process.py:132)
 *   d_sample_idx   int32[n][point_num]: rand_idx of every frame, each in [0, n_valid_i) (others select nothing)
 * cfg->bbox_mode is ignored.  The image size w, h of the back-projection comes from each header.
 *
 * dtsdf_count_valid_batch: d_out_n_valid[i] = number of pixels with d != 0 - hand_points.shape[0], the upper
 * bound of the draws of set_length - so that the caller can draw rand_idx before the launch. */
int dtsdf_count_valid_batch(int32_t n_frames, const float* d_depth_packed, const int64_t* d_crop_offsets,
                            const int32_t* d_headers, int32_t* d_out_n_valid, void* stream);
int dtsdf_voxelize_sampled_batch(const dtsdf_config* cfg, double focal, int32_t n_frames,
                                 const float* d_depth_packed, const int64_t* d_crop_offsets,
                                 const int32_t* d_headers, const int32_t* d_sample_idx, int32_t point_num,
                                 int32_t max_crop_px, float* d_out_vol, float* d_out_max_l,
                                 float* d_out_mid_p, void* stream);

/* Augmented D-TSDF for training-time augmentation.  Every frame gets its
 * own affine map p' = (p - c) . A + c (row vector times matrix, A = stretch . rotation as in
 * show/visualize.py:106-114): the grid is laid over the AUGMENTED hand, voxel centres are taken back
 * to the camera frame with Ainv, projected and looked up as usual, distances are measured in the
 * augmented frame ((v - w) . A for the camera-frame voxel centre v and surface point w).  There is no reference behaviour to reproduce (data_aug raises AxisError,
 * pre/process.py:252; cut_version voxelises the un-augmented depth, cut_version/pre/process.py:207):
 * the specification is oracle/dtsdf_oracle.c:oracle_dtsdf_aug_frame, matched bit for bit.
 *   d_aug         float[n][21] = A[3][3], Ainv[3][3], c[3]
 *   d_joints / joint_flags / d_out_joints   optional: the ground truth mapped by the same
 *                 augmentation and normalised, (T(gt) - mid_p)/max_l + 0.5
 * cfg->bbox_mode is ignored (the box is over the augmented valid pixels). */
int dtsdf_voxelize_aug_batch(const dtsdf_config* cfg, int32_t n_frames,
                             const float* d_depth_packed, const int64_t* d_crop_offsets,
                             const int32_t* d_headers, const float* d_aug, int32_t max_crop_px,
                             const float* d_joints, int32_t joint_flags,
                             float* d_out_vol, float* d_out_max_l, float* d_out_mid_p,
                             float* d_out_joints, void* stream);

/* Back-projection of every pixel with d != 0 to a float64 point, compacted in
 * row-major order.  Replaces DataProcess.point_cloud (pre/process.py:30-68).
 *   focal           the Python float 241.42 as a double (the reference divides float64
 *                   values by it, pre/process.py:53,59), not the float32 of dtsdf_config
 *   d_points        double[..][3]; frame i's points start at row d_point_offsets[i]
 *                   and need room for bw*bh rows
 *   d_n_valid       int32[n]: number of points written per frame
 */
int dtsdf_point_cloud_batch(int32_t n_frames, double focal,
                            const float* d_depth_packed, const int64_t* d_crop_offsets,
                            const int32_t* d_headers,
                            double* d_points, const int64_t* d_point_offsets,
                            int32_t* d_n_valid, void* stream);

/* Gathers point_num rows per frame by index and reduces the per-axis extrema.
 * Replaces the gather of DataProcess.set_length (pre/process.py:82; the index
 * vector itself comes from the caller's np.random stream, :75-80) and
 * max_min_point (pre/process.py:102-120).
 *   d_indices      int32[n][point_num], each in [0, n_valid_i)
 *   d_out_pc       double[n][point_num][3]  (may be NULL)
 *   d_out_extrema  float[n][6] = {max_x,max_y,max_z,min_x,min_y,min_z}; z ignores exact zeros
 */
int dtsdf_resample_extrema_batch(int32_t n_frames, int32_t point_num,
                                 const double* d_points, const int64_t* d_point_offsets,
                                 const int32_t* d_indices,
                                 double* d_out_pc, float* d_out_extrema, void* stream);

/* The two calls above in one launch when only the SAMPLED cloud is wanted: the points
 * hand_points[rand_idx] (pre/process.py:30-68 followed by :82) are back-projected straight
 * from the depth map - the full cloud is never written - and their extrema reduced
 * (pre/process.py:102-120).  Bit for bit what dtsdf_point_cloud_batch +
 * dtsdf_resample_extrema_batch produce.
 *   d_indices      int32[n][point_num]: ranks among the frame's valid pixels (d != 0, row-major);
 *                  a value outside [0, n_valid_i) yields a zero point and does not enter the extrema
 *   d_out_pc       double[n][point_num][3]  (may be NULL)
 *   d_out_extrema  float[n][6] as above
 *   d_out_n_valid  int32[n] or NULL
 *   max_crop_px    the largest crop of the batch in pixels; beyond DTSDF_SAMPLE_CLOUD_MAX_PX
 *                  the call returns DTSDF_ERR_UNSUPPORTED (use the two calls above)
 */
#define DTSDF_SAMPLE_CLOUD_MAX_PX 22528
int dtsdf_sample_cloud_batch(int32_t n_frames, double focal, int32_t point_num,
                             const float* d_depth_packed, const int64_t* d_crop_offsets,
                             const int32_t* d_headers, const int32_t* d_indices, int32_t max_crop_px,
                             double* d_out_pc, float* d_out_extrema, int32_t* d_out_n_valid,
                             void* stream);

/* ---- lossless volume codec on the device (gather of volumes to a training rank) ----------------
 * A voxel whose three channels are all +0.0, all +1.0 or all -1.0 - outside the crop, no depth, or
 * beyond the truncation distance (pre/process.py:159,167,180-194): about 90 % of a D-TSDF volume -
 * becomes a 2-bit code (0, 1, 2); any other voxel is a "literal" (code 3) whose three floats are kept.
 * The packed form of n frames of R^3 voxels is
 *   d_codes      uint8[n][R^3/4]   voxel 4k+j of a frame in bits 2j..2j+1 of its byte k
 *   d_lit_count  uint32[n]         literal voxels per frame
 *   d_lits       float[n][3*R^3]   per frame, its literals in voxel order as (x, y, z) triples; only
 *                                  the first 3*d_lit_count[f] floats of a frame's region are written
 * (~47 KB in use per frame at R = 32 against 393 KB dense).  It replaces nothing in the reference,
 * which is single-GPU; it exists so that the optional gather of volumes into
 * one rank moves an eighth of the bytes over NVLink.  Both layouts (the codec works on channel planes).
 *
 * dtsdf_pack_volumes_batch     one pass over dense volumes in device memory.
 * dtsdf_unpack_volumes_batch   the inverse, one CTA per frame.  The three packed arrays may live in
 *                              ANOTHER GPU's memory (dtsdf_shared_open / dtsdf_enable_peer_access):
 *                              the kernel pulls each frame's bytes in use over NVLink with TMA bulk
 *                              copies and writes the dense volume into local memory, so transfer and
 *                              expansion are one kernel on the receiving rank.
 * All pointers 16-byte aligned. */
int dtsdf_pack_volumes_batch(int32_t voxel_res, int32_t n_frames, const float* d_vol,
                             uint8_t* d_codes, uint32_t* d_lit_count, float* d_lits, void* stream);
int dtsdf_unpack_volumes_batch(int32_t voxel_res, int32_t n_frames, const uint8_t* d_codes,
                               const uint32_t* d_lit_count, const float* d_lits,
                               float* d_out_vol, void* stream);

/* Joint normalisation (gt - mid_p)/max_l + 0.5 over n frames of 21 joints.
 * Replaces joint_nor.normalize (pre/joint_nor.py:8-18) and DataProcess.normalize
 * (cut_version/pre/process.py:251-263).  clamp01 != 0 also clamps to [0,1] as the
 * trainer does (3D_CNN/train.py:236-244). */
int dtsdf_normalize_joints_batch(int32_t n_frames, const float* d_joints /*[n][63]*/,
                                 const float* d_max_l, const float* d_mid_p,
                                 int32_t clamp01, float* d_out /*[n][63]*/, void* stream);

/* ---- host entry points (host memory in and out, synchronous) ----------------- */
typedef struct dtsdf_ctx dtsdf_ctx;

int  dtsdf_ctx_create(int32_t device, dtsdf_ctx** out);
void dtsdf_ctx_destroy(dtsdf_ctx* ctx);

/* Context options.
 *   DTSDF_OPT_PACKED_D2H   1 (default): dtsdf_voxelize_host brings the volumes back in a lossless
 *                          packed form (2-bit codes for voxels that are 0 / +1 / -1 in all three
 *                          channels, literals for the rest: ~8x fewer PCIe bytes) written by the
 *                          GPU straight into mapped pinned memory, and expands them into the
 *                          caller's buffer with the host threads below (streaming stores).
 *                          Batches whose volumes are under 16 MB always come back as plain
 *                          copies (the packed route costs more than it saves there).
 *                          0: plain device->host copies of the float volumes.
 *   DTSDF_OPT_HOST_THREADS number of host threads for the expansion (default 0 = three quarters of
 *                          the hardware threads, at most 8: more do not add memory bandwidth). */
#define DTSDF_OPT_PACKED_D2H   1
#define DTSDF_OPT_HOST_THREADS 2
#define DTSDF_OPT_CHUNK_MB     3   /* output megabytes per pipeline chunk of dtsdf_voxelize_host (default 64) */
#define DTSDF_OPT_LITERAL_PUSH 4   /* packed transfer: 0 (default) the literal stream of a chunk comes back with one copy-engine
                                      transfer issued once the host knows its size; 1: pushed into mapped pinned
                                      memory by a kernel (no host round trip, SM stores over PCIe) */
#define DTSDF_OPT_TEST_FAIL_CHUNK 100 /* testing only: dtsdf_voxelize_host fails with DTSDF_ERR_CUDA when it reaches
                                         chunk `value`, with the chunks before it in flight (-1, the default: off) */
int dtsdf_ctx_set_option(dtsdf_ctx* ctx, int32_t option, int64_t value);
/* Statistics of the last dtsdf_voxelize_host call.
 *   DTSDF_STAT_D2H_BYTES   bytes that crossed PCIe device -> host (packed: codes + offsets +
 *                          literals + max_l/mid_p; plain: the float volumes + max_l/mid_p) */
#define DTSDF_STAT_D2H_BYTES 1
int dtsdf_ctx_get_stat(dtsdf_ctx* ctx, int32_t stat, int64_t* value);

/* Page-locked host buffers for the host entry points (plain malloc'ed memory
 * works too, only slower). */
int  dtsdf_host_alloc(void** ptr, int64_t bytes);
void dtsdf_host_free(void* ptr);

/* Same computation as dtsdf_voxelize_batch with HOST buffers: the batch is cut
 * into chunks and host->device copy, kernel and device->host copy of successive
 * chunks overlap on three streams.  This is the call the NumPy-facing shims
 * (tsdf_numba.cal_tsdf_cuda, tsdf_for.tsdf_f, DataProcess.tsdf_f/tsdf_cal) make. */
int dtsdf_voxelize_host(dtsdf_ctx* ctx, const dtsdf_config* cfg, int32_t n_frames,
                        const float* h_depth_packed, int64_t packed_len,
                        const int64_t* h_crop_offsets, const int32_t* h_headers,
                        const float* h_aux,
                        float* h_out_vol, float* h_out_max_l, float* h_out_mid_p);

/* dtsdf_voxelize_host that leaves the volumes PACKED in host memory (the lossless form in which they cross
 * PCIe: 2-bit codes + literals, ~47 KB instead of 393 KB per volume at R = 32) behind an opaque handle, for
 * callers that do not need every volume dense at once - a data set that expands the samples of a batch on
 * demand (3D_CNN/dataset.py:35 keeps [N,3,32,32,32] float32 = 30 GB for the full MSRA set in host memory; the
 * packed set is 3.6 GB), or a writer that converts gesture by gesture (pre/read_MSRA.py:119-126).
 *   dtsdf_expand_frames   frames [first, first+count) of the handle -> float[count][3][R][R][R] in the layout
 *                         the volumes were made in; n_threads host threads (0 = default, at most 8)
 *   dtsdf_packed_host_info / _free   size of the handle, release */
typedef struct dtsdf_packed_host dtsdf_packed_host;
int  dtsdf_voxelize_host_packed(dtsdf_ctx* ctx, const dtsdf_config* cfg, int32_t n_frames,
                                const float* h_depth_packed, int64_t packed_len,
                                const int64_t* h_crop_offsets, const int32_t* h_headers,
                                const float* h_aux, dtsdf_packed_host** out,
                                float* h_out_max_l, float* h_out_mid_p);
int  dtsdf_expand_frames(const dtsdf_packed_host* packed, int32_t first, int32_t count, float* h_out,
                         int32_t n_threads);
int  dtsdf_packed_host_info(const dtsdf_packed_host* packed, int32_t* n_frames, int32_t* voxel_res, int64_t* bytes);
void dtsdf_packed_host_free(dtsdf_packed_host* packed);

/* DataProcess.point_cloud for one frame (pre/process.py:30-68).  The compacted
 * cloud also stays resident in the context for dtsdf_resample_extrema_host.
 *   h_out_points  double[bw*bh][3] capacity (may be NULL to only count)
 */
int dtsdf_point_cloud_host(dtsdf_ctx* ctx, double focal, const int32_t* h_header,
                           const float* h_depth, double* h_out_points, int32_t* n_valid);

/* DataProcess.set_length gather + max_min_point on the cloud left in the context
 * by the last dtsdf_point_cloud_host call, or on h_points when it is not NULL
 * (tsdf_for.tsdf_f / DataProcess.tsdf_f are handed an arbitrary cloud).
 *   h_indices      int32[point_num] or NULL (= identity, point_num rows of h_points)
 */
int dtsdf_resample_extrema_host(dtsdf_ctx* ctx, int32_t point_num,
                                const double* h_points, int32_t n_points,
                                const int32_t* h_indices,
                                double* h_out_pc, float* h_out_extrema);

/* DataProcess.process() for a whole batch (pre/process.py:13-28; the frame loop of
 * read_MSRA.main(), pre/read_MSRA.py:98-111), in two calls because the resampling index vectors
 * are drawn by the caller from its np.random stream and depend on the number of valid points:
 *
 *   dtsdf_process_begin_host   uploads the packed crops, back-projects and compacts every
 *                              frame's point cloud (pre/process.py:30-68; it stays resident in
 *                              the context) and returns n_valid per frame;
 *   dtsdf_process_finish_host  gathers point_num rows per frame by the caller's indices
 *                              (pre/process.py:82), reduces the extrema (:102-120) and runs the
 *                              fused voxeliser with them (tsdf_f, :86-100) on the resident crops.
 *     h_indices     int32[n][point_num], each in [0, n_valid_i)
 *     h_out_pc      double[n][point_num][3] (may be NULL)
 *     h_out_vol     float[n][3][R][R][R] in cfg->layout; cfg->bbox_mode is ignored (EXTREMA)
 */
int dtsdf_process_begin_host(dtsdf_ctx* ctx, double focal, int32_t n_frames,
                             const float* h_depth_packed, int64_t packed_len,
                             const int64_t* h_crop_offsets, const int32_t* h_headers,
                             int32_t* h_n_valid);
int dtsdf_process_finish_host(dtsdf_ctx* ctx, const dtsdf_config* cfg, int32_t point_num,
                              const int32_t* h_indices, double* h_out_pc,
                              float* h_out_vol, float* h_out_max_l, float* h_out_mid_p);

/* The caller's global NumPy stream, in C.  DataProcess.set_length draws its resampling indices with
 * np.random.randint on the GLOBAL legacy stream (pre/process.py:75-80), and those indices decide the
 * bounding box and with it every voxel; to stay bit-compatible with a seeded reference run the library
 * consumes exactly the same stream: MT19937 as NumPy steps it, bounded integers by masked rejection
 * (numpy/random/src/distributions: random_bounded_uint64_fill, use_masked = True, 32-bit branch).
 *   mt_key[624], *mt_pos   the state as np.random.get_state() returns it (key, pos); advanced in place, to be
 *                          handed back with np.random.set_state().  (NumPy is a dependency of the reference,
 *                          not part of it: tests pin this against np.random itself.)
 * dtsdf_np_randint         out[0..count) = np.random.randint(0, high, size=count)
 * dtsdf_resample_indices   out_idx[n][point_num]: the index vectors of set_length for n frames in order, given
 *                          each frame's number of points (fewer than point_num: all of them once, then random
 *                          repeats; otherwise point_num draws).  DTSDF_ERR_EMPTY_CLOUD where NumPy would raise. */
int dtsdf_np_randint(uint32_t* mt_key, int32_t* mt_pos, int32_t high, int32_t count, int32_t* out);
int dtsdf_resample_indices(uint32_t* mt_key, int32_t* mt_pos, int32_t n_frames, const int32_t* n_valid,
                           int32_t point_num, int32_t* out_idx);

/* DataProcess(frame, gt, point_num).process() for a whole batch in ONE call (pre/process.py:13-28 inside the
 * frame loop of read_MSRA.main(), pre/read_MSRA.py:98-111): point clouds on the device, resampling indices
 * drawn here from the caller's MT19937 state chunk by chunk while earlier chunks are on the GPU, gather +
 * extrema + fused voxeliser per chunk, volumes back through the packed transfer of dtsdf_voxelize_host.
 *   h_out_pc   double[n][point_num][3] or NULL: the sampled clouds are not brought back - and then never built:
 *              the batch goes through dtsdf_count_valid_batch + dtsdf_voxelize_sampled_batch
 * cfg->bbox_mode is ignored.  Equal to dtsdf_process_begin_host + np.random draws +
 * dtsdf_process_finish_host, bit for bit, state of the stream included. */
int dtsdf_process_host(dtsdf_ctx* ctx, const dtsdf_config* cfg, double focal, int32_t n_frames,
                       const float* h_depth_packed, int64_t packed_len,
                       const int64_t* h_crop_offsets, const int32_t* h_headers,
                       int32_t point_num, uint32_t* mt_key, int32_t* mt_pos,
                       double* h_out_pc, float* h_out_vol, float* h_out_max_l, float* h_out_mid_p);

int dtsdf_normalize_joints_host(dtsdf_ctx* ctx, int32_t n_frames, const float* h_joints,
                                const float* h_max_l, const float* h_mid_p,
                                int32_t clamp01, float* h_out);

#ifdef __cplusplus
}
#endif
#endif /* DTSDF_B200_H */

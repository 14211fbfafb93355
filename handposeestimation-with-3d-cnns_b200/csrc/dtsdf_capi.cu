// C ABI of the D-TSDF voxeliser (see include/dtsdf_b200.h for the contract and the
// reference callables each entry point replaces).  Host-side runtime: argument
// checking, error mapping, the per-process context with its streams and staging
// buffers, and the chunked copy/compute pipeline behind the host entry points.
#include "../../include/dtsdf_b200.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <exception>
#include <new>
#include <thread>
#include <vector>

#include "dtsdf_internal.h"

namespace {

thread_local char g_cuda_err[256] = "";

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", what, cudaGetErrorName(e),
             cudaGetErrorString(e));
    cudaGetLastError();   // clear the sticky-less error state
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return DTSDF_ERR_NO_DEVICE;
    if (e == cudaErrorMemoryAllocation) return DTSDF_ERR_ALLOC;
    return DTSDF_ERR_CUDA;
}

#define CK(call)                                        \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// The C ABI never throws: every entry point that can reach an allocation or a std::thread runs its body
// through this.
template <typename F>
int guarded(F&& body) noexcept {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "out of host memory (std::bad_alloc)");
        return DTSDF_ERR_ALLOC;
    } catch (const std::exception& e) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "C++ exception: %s", e.what());
        return DTSDF_ERR_INTERNAL;
    } catch (...) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "unknown C++ exception");
        return DTSDF_ERR_INTERNAL;
    }
}

// Makes `device` current for the duration of a host entry point and puts the caller's device back on
// every exit (a process that drives several GPUs must not find its current device changed).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err != cudaSuccess) { prev = -1; cudaGetLastError(); }
        err = cudaSetDevice(device);
    }
    ~DeviceGuard() {
        if (prev >= 0) { cudaSetDevice(prev); cudaGetLastError(); }
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define USE_DEVICE(dev)                                                   \
    DeviceGuard device_guard__(dev);                                      \
    if (device_guard__.err != cudaSuccess) return cuda_fail(device_guard__.err, "cudaSetDevice")

int check_cfg(const dtsdf_config* cfg) {
    if (!cfg) return DTSDF_ERR_INVALID_ARG;
    if (cfg->voxel_res != 32 && cfg->voxel_res != 48 && cfg->voxel_res != 64) return DTSDF_ERR_UNSUPPORTED;
    if (cfg->layout != DTSDF_LAYOUT_CXYZ && cfg->layout != DTSDF_LAYOUT_CZYX) return DTSDF_ERR_UNSUPPORTED;
    if (cfg->bbox_mode < DTSDF_BBOX_PIXELS || cfg->bbox_mode > DTSDF_BBOX_GRID) return DTSDF_ERR_UNSUPPORTED;
    if (!(cfg->focal > 0.0f) || !(cfg->trunc_mul > 0.0f)) return DTSDF_ERR_INVALID_ARG;
    return DTSDF_OK;
}

int header_ok(const int32_t* h) {
    const int64_t bw = (int64_t)h[4] - h[2], bh = (int64_t)h[5] - h[3];
    const int64_t lim = 1 << 22;
    if (bw <= 0 || bh <= 0 || bw > 32767 || bh > 32767 || bw * bh > ((int64_t)1 << 26)) return 0;
    if (h[2] <= -lim || h[3] <= -lim || h[4] >= lim || h[5] >= lim) return 0;
    return 1;
}

// grow-only device / pinned buffers
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return DTSDF_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = std::max(bytes, (size_t)4096);
        CK(cudaMalloc(&p, want));
        cap = want;
        return DTSDF_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// mapped pinned host buffer: the GPU writes it over PCIe, the host threads read it
struct MapBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return DTSDF_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = std::max(bytes, (size_t)4096);
        CK(cudaHostAlloc(&p, want, cudaHostAllocMapped | cudaHostAllocPortable));
        cap = want;
        return DTSDF_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// A small persistent thread pool for the host-side expansion of packed volumes.  Jobs are plain records in
// a fixed ring (no allocation on submit: nothing here can throw once the threads run).
struct ExpandJob {
    const void* chunk;          // dtsdf_ctx::ChunkRecord, or (draw = true) a DrawRecord
    int32_t f0, f1;             // frames [f0, f1) of the chunk
    bool draw = false;          // second phase of the resampling index draw instead of an expansion
};
void run_expand_job(const ExpandJob& j);

class Pool {
public:
    static const int CAP = 4096;             // >= chunks in flight x threads
    ~Pool() { stop(); }
    void ensure(int n) {                     // may throw (std::thread): called before anything is in flight
        if ((int)th_.size() == n) return;
        stop();
        quit_ = false;
        th_.reserve(n);
        for (int i = 0; i < n; ++i) th_.emplace_back([this] { run(); });
    }
    // false: the ring is full (the caller runs the job itself)
    bool submit(const ExpandJob& j) noexcept {
        {
            std::lock_guard<std::mutex> g(m_);
            if (count_ == CAP) return false;
            if (j.draw) {                        // the submitting thread waits for draws: they go in front of expansions
                head_ = (head_ + CAP - 1) % CAP;
                ring_[head_] = j;
            } else {
                ring_[(head_ + count_) % CAP] = j;
            }
            ++count_;
        }
        cv_.notify_one();
        return true;
    }
    int size() const { return (int)th_.size(); }
    void stop() {
        { std::lock_guard<std::mutex> g(m_); quit_ = true; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
        th_.clear();
    }
private:
    void run() {
        for (;;) {
            ExpandJob j;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [this] { return quit_ || count_ > 0; });
                if (count_ == 0) return;
                j = ring_[head_];
                head_ = (head_ + 1) % CAP;
                --count_;
            }
            run_expand_job(j);
        }
    }
    std::vector<std::thread> th_;
    ExpandJob ring_[CAP];
    int head_ = 0, count_ = 0;
    std::mutex m_;
    std::condition_variable cv_;
    bool quit_ = false;
};

struct Latch {
    std::mutex m;
    std::condition_variable cv;
    int pending = 0;
    void add(int n) { std::lock_guard<std::mutex> g(m); pending += n; }
    void done() {
        std::lock_guard<std::mutex> g(m);
        if (pending > 0 && --pending == 0) cv.notify_all();
    }
    void reset() {                               // after a failed stream: nothing more will arrive
        std::lock_guard<std::mutex> g(m);
        pending = 0;
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> g(m);
        cv.wait(g, [this] { return pending == 0; });
    }
};

// (the expansion of a packed frame on the host lives in dtsdf_decode.cpp: dtsdf::decode_frame)

}  // namespace

// Volumes left packed in host memory by dtsdf_voxelize_host_packed: per pipeline chunk its code bytes,
// literal offsets and literal stream exactly as they crossed PCIe (dtsdf_internal.h, compact container).
namespace {
// Storage of packed chunks that outlives a handle: a chunk is ~8 MB, and memory fresh from the allocator costs a
// page fault per 4 KB when the expansion threads first write it (a fresh 400 MB result: more time in the kernel's
// page-fault path than in the whole GPU pipeline).  Blocks of freed handles wait here - at most 1 GB of them - for
// the next result; a training loop that packs a set per epoch then never touches fresh memory.
class BlockCache {
public:
    struct Block { unsigned char* p = nullptr; size_t cap = 0; };
    Block take(size_t need) {                      // may throw std::bad_alloc
        {
            std::lock_guard<std::mutex> g(m_);
            for (size_t i = 0; i < free_.size(); ++i)
                if (free_[i].cap >= need && free_[i].cap <= 2 * need + 4096) {
                    Block b = free_[i];
                    free_[i] = free_.back();
                    free_.pop_back();
                    held_ -= b.cap;
                    return b;
                }
        }
        Block b;
        b.cap = (need + 4095) & ~(size_t)4095;
        b.p = new unsigned char[b.cap];
        return b;
    }
    void give(Block b) noexcept {
        if (!b.p) return;
        {
            std::lock_guard<std::mutex> g(m_);
            if (held_ + b.cap <= LIMIT && free_.size() < free_.capacity()) {
                free_.push_back(b);
                held_ += b.cap;
                return;
            }
        }
        delete[] b.p;
    }
    BlockCache() { free_.reserve(4096); }
private:
    static constexpr size_t LIMIT = (size_t)1 << 30;
    std::mutex m_;
    std::vector<Block> free_;
    size_t held_ = 0;
};
BlockCache& blocks() {                             // (never destroyed: handles may be freed during process exit)
    static BlockCache* cache = new BlockCache();
    return *cache;
}
}  // namespace

struct dtsdf_packed_host {
    int32_t n_frames = 0, voxel_res = 0, per_chunk = 1;
    size_t r3 = 0;
    struct Chunk {
        int32_t f0 = 0, nf = 0;
        BlockCache::Block block;                          // one allocation: [loff | codes | lits]
        unsigned long long* loff = nullptr;               // [nf+1]
        unsigned char* codes = nullptr;                   // [nf][r3/4]
        float* lits = nullptr;                            // [3 * loff[nf]]
        void bind(int32_t frames, size_t r3_, unsigned long long total) {      // may throw std::bad_alloc
            const size_t b_loff = ((size_t)frames + 1) * 8, b_codes = ((size_t)frames * (r3_ / 4) + 15) & ~(size_t)15;
            block = blocks().take(b_loff + 8 + b_codes + 12 * (size_t)total + 16);
            loff = reinterpret_cast<unsigned long long*>(block.p);
            codes = block.p + ((b_loff + 15) & ~(size_t)15);
            lits = reinterpret_cast<float*>(codes + b_codes);
        }
    };
    std::vector<Chunk> chunks;
    std::atomic<long long> bytes{0};
    ~dtsdf_packed_host() { for (Chunk& ch : chunks) blocks().give(ch.block); }
};

struct dtsdf_ctx {
    int device = 0;
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr, s_lit = nullptr, s_idx = nullptr;
    static const int NSLOT = 4;                 // chunks in flight
    // packed device -> host transfer
    int opt_packed = 1;
    int opt_threads = 0;                        // 0 = hardware concurrency
    int opt_chunk_mb = 64;
    int opt_push = 0;                           // literal stream by a push kernel instead of the copy engine
    int opt_fail_chunk = -1;                    // testing: the host pipeline fails before submitting this chunk
    Pool pool;
    DevBuf d_cnt[NSLOT], d_loff[NSLOT], d_codes[NSLOT], d_lits[NSLOT];
    MapBuf m_codes[NSLOT], m_loff[NSLOT], m_lits[NSLOT], m_maxl[NSLOT], m_mid[NSLOT];
    Latch latch[NSLOT];
    std::atomic<long long> stat_d2h_bytes{0};
    // what the finalizer thread needs to hand a chunk to the expansion threads (owned by the slot: rewritten only after latch[slot].wait())
    struct ChunkRecord {
        dtsdf_ctx* ctx;
        int slot, jobs;
        int32_t nf;
        size_t r3, vol_floats;
        float* dst;                              // h_vol + f0 * vol_floats
        float* dst_maxl;                         // h_maxl + f0 or null
        float* dst_mid;                          // h_mid + 3 f0 or null
        dtsdf_packed_host::Chunk* keep;          // packed result: the chunk's storage (the volumes are not expanded)
        double* dst_pc;                          // process mode: h_pc + f0 * point_num * 3 or null
        size_t pc_doubles;                       // point_num * 3
        long long fixed_bytes;                   // PCIe bytes of the chunk apart from the literal stream
    } chunk[NSLOT];
    DevBuf d_depth[NSLOT], d_off[NSLOT], d_hdr[NSLOT], d_aux[NSLOT];
    // dtsdf_process_host: per-chunk resampling indices (pinned staging + device), gathered clouds, extrema
    MapBuf m_idx[NSLOT], m_pc[NSLOT];
    DevBuf p_idx[NSLOT], p_pc[NSLOT], p_ext[NSLOT];
    DevBuf d_vol[NSLOT], d_maxl[NSLOT], d_mid[NSLOT];
    cudaEvent_t ev_in[NSLOT] = {}, ev_run[NSLOT] = {}, ev_out[NSLOT] = {}, ev_meta[NSLOT] = {};
    // point-cloud path
    DevBuf d_points, d_idx, d_pc, d_ext, d_misc;
    int32_t last_n_valid = -1;
    // joints
    DevBuf d_j_in, d_j_out, d_j_ml, d_j_mp;
    // batched process(): resident crops, clouds and their tables between begin and finish
    DevBuf b_depth, b_off, b_hdr, b_points, b_poff, b_nvalid, b_idx, b_pc, b_ext, b_vol, b_maxl, b_mid;
    int32_t b_n = 0;
    int32_t b_max_px = 0;
    std::vector<int32_t> b_nvalid_h;
    // dtsdf_process_host without clouds: the crops go up piece by piece (one piece per pipeline chunk) with the
    // valid-pixel counts of a piece coming back behind it, so the first index draw starts after one piece
    MapBuf m_nvalid;
    std::vector<cudaEvent_t> ev_piece;
    int32_t b_pieces = 0;                        // 0: everything was uploaded and counted synchronously
};

extern "C" {

int dtsdf_abi_version(void) { return DTSDF_ABI_VERSION; }

const char* dtsdf_error_string(int code) {
    switch (code) {
        case DTSDF_OK: return "ok";
        case DTSDF_ERR_INVALID_ARG: return "invalid argument";
        case DTSDF_ERR_UNSUPPORTED: return "unsupported voxel resolution, layout or bbox mode";
        case DTSDF_ERR_CUDA: return "CUDA runtime error";
        case DTSDF_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
        case DTSDF_ERR_ALLOC: return "out of memory";
        case DTSDF_ERR_BAD_HEADER: return "frame header with empty or oversize bounding box";
        case DTSDF_ERR_INTERNAL: return "internal error (C++ exception caught at the C boundary)";
        case DTSDF_ERR_EMPTY_CLOUD: return "a frame has no point to draw from (np.random.randint: low >= high)";
        default: return "unknown error";
    }
}

const char* dtsdf_last_cuda_error(void) { return g_cuda_err; }

void dtsdf_default_config(dtsdf_config* cfg) {
    if (!cfg) return;
    cfg->voxel_res = 32;
    cfg->layout = DTSDF_LAYOUT_CXYZ;
    cfg->bbox_mode = DTSDF_BBOX_PIXELS;
    cfg->center_x = 160;
    cfg->center_y = 120;
    cfg->focal = 241.42f;
    cfg->trunc_mul = 3.0f;
}

int dtsdf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int64_t dtsdf_launch_count(void) { return dtsdf::g_launch_count.load(); }

int dtsdf_enable_peer_access(int32_t device, int32_t peer_device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    if (n == 0) return DTSDF_ERR_NO_DEVICE;
    if (device < 0 || device >= n || peer_device < 0 || peer_device >= n) return DTSDF_ERR_INVALID_ARG;
    if (device == peer_device) return DTSDF_OK;
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, device, peer_device));
    if (!can) return DTSDF_ERR_UNSUPPORTED;
    int prev = 0;
    CK(cudaGetDevice(&prev));
    CK(cudaSetDevice(device));
    e = cudaDeviceEnablePeerAccess(peer_device, 0);
    cudaSetDevice(prev);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return DTSDF_OK; }
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
    return DTSDF_OK;
}

static int with_device(int32_t device, int* prev) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    if (n == 0) return DTSDF_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return DTSDF_ERR_INVALID_ARG;
    CK(cudaGetDevice(prev));
    CK(cudaSetDevice(device));
    return DTSDF_OK;
}

int dtsdf_shared_alloc(int32_t device, int64_t bytes, void** d_ptr, uint8_t handle[DTSDF_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == DTSDF_IPC_HANDLE_BYTES, "IPC handle size");
    if (!d_ptr || !handle || bytes <= 0) return DTSDF_ERR_INVALID_ARG;
    *d_ptr = nullptr;
    int prev = 0, rc = with_device(device, &prev);
    if (rc != DTSDF_OK) return rc;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    cudaSetDevice(prev);
    if (e != cudaSuccess) { if (p) cudaFree(p); return cuda_fail(e, "cudaMalloc / cudaIpcGetMemHandle"); }
    memcpy(handle, &h, sizeof(h));
    *d_ptr = p;
    return DTSDF_OK;
}

int dtsdf_shared_free(int32_t device, void* d_ptr) {
    if (!d_ptr) return DTSDF_OK;
    int prev = 0, rc = with_device(device, &prev);
    if (rc != DTSDF_OK) return rc;
    cudaError_t e = cudaFree(d_ptr);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFree");
    return DTSDF_OK;
}

int dtsdf_shared_open(int32_t device, const uint8_t handle[DTSDF_IPC_HANDLE_BYTES], void** d_ptr) {
    if (!d_ptr || !handle) return DTSDF_ERR_INVALID_ARG;
    *d_ptr = nullptr;
    int prev = 0, rc = with_device(device, &prev);
    if (rc != DTSDF_OK) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle");
    *d_ptr = p;
    return DTSDF_OK;
}

int dtsdf_shared_close(int32_t device, void* d_ptr) {
    if (!d_ptr) return DTSDF_OK;
    int prev = 0, rc = with_device(device, &prev);
    if (rc != DTSDF_OK) return rc;
    cudaError_t e = cudaIpcCloseMemHandle(d_ptr);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcCloseMemHandle");
    return DTSDF_OK;
}

int dtsdf_pack_plan(int32_t n_frames, const int32_t* headers, int64_t* crop_offsets,
                    int64_t* packed_len, int32_t* max_crop_px) {
    if (n_frames < 0 || (n_frames > 0 && (!headers || !crop_offsets))) return DTSDF_ERR_INVALID_ARG;
    int64_t off = 0;
    int32_t mx = 0;
    for (int32_t i = 0; i < n_frames; ++i) {
        const int32_t* h = headers + 6 * (size_t)i;
        if (!header_ok(h)) return DTSDF_ERR_BAD_HEADER;
        const int64_t n = ((int64_t)h[4] - h[2]) * ((int64_t)h[5] - h[3]);
        crop_offsets[i] = off;
        off += (n + 3) & ~(int64_t)3;
        if (n > mx) mx = (int32_t)n;
    }
    if (packed_len) *packed_len = off;
    if (max_crop_px) *max_crop_px = mx;
    return DTSDF_OK;
}

// ------------------------------------------------------------------ device entry points
int dtsdf_voxelize_batch(const dtsdf_config* cfg, int32_t n_frames, const float* d_depth_packed,
                         const int64_t* d_crop_offsets, const int32_t* d_headers,
                         const float* d_aux, int32_t max_crop_px, float* d_out_vol,
                         float* d_out_max_l, float* d_out_mid_p, void* stream) {
    return dtsdf_voxelize_joints_batch(cfg, n_frames, d_depth_packed, d_crop_offsets, d_headers, d_aux, max_crop_px,
                                       nullptr, 0, d_out_vol, d_out_max_l, d_out_mid_p, nullptr, stream);
}

int dtsdf_voxelize_joints_batch(const dtsdf_config* cfg, int32_t n_frames, const float* d_depth_packed,
                                const int64_t* d_crop_offsets, const int32_t* d_headers,
                                const float* d_aux, int32_t max_crop_px, const float* d_joints,
                                int32_t joint_flags, float* d_out_vol, float* d_out_max_l,
                                float* d_out_mid_p, float* d_out_joints, void* stream) {
    int rc = check_cfg(cfg);
    if ((d_joints != nullptr) != (d_out_joints != nullptr)) return DTSDF_ERR_INVALID_ARG;
    if (joint_flags & ~3) return DTSDF_ERR_INVALID_ARG;
    if (rc != DTSDF_OK) return rc;
    if (n_frames < 0) return DTSDF_ERR_INVALID_ARG;
    if (n_frames == 0) return DTSDF_OK;
    if (!d_depth_packed || !d_crop_offsets || !d_headers || !d_out_vol) return DTSDF_ERR_INVALID_ARG;
    if (cfg->bbox_mode != DTSDF_BBOX_PIXELS && !d_aux) return DTSDF_ERR_INVALID_ARG;
    dtsdf::BatchArgs a;
    a.depth = d_depth_packed;
    a.offsets = reinterpret_cast<const long long*>(d_crop_offsets);
    a.headers = d_headers;
    a.aux = d_aux;
    a.vol = d_out_vol;
    a.max_l = d_out_max_l;
    a.mid_p = d_out_mid_p;
    a.n_frames = n_frames;
    a.cap_floats = 0;
    a.bbox_mode = cfg->bbox_mode;
    a.center_x = cfg->center_x;
    a.center_y = cfg->center_y;
    a.focal = cfg->focal;
    a.trunc_mul = cfg->trunc_mul;
    a.joints = d_joints;
    a.out_joints = d_out_joints;
    a.joint_flags = joint_flags;
    CK(dtsdf::launch_voxelize(cfg->voxel_res, cfg->layout, a, max_crop_px, (cudaStream_t)stream));
    return DTSDF_OK;
}

int dtsdf_count_valid_batch(int32_t n_frames, const float* d_depth_packed, const int64_t* d_crop_offsets,
                            const int32_t* d_headers, int32_t* d_out_n_valid, void* stream) {
    if (n_frames < 0) return DTSDF_ERR_INVALID_ARG;
    if (n_frames == 0) return DTSDF_OK;
    if (!d_depth_packed || !d_crop_offsets || !d_headers || !d_out_n_valid) return DTSDF_ERR_INVALID_ARG;
    CK(dtsdf::launch_count_valid(n_frames, d_depth_packed, reinterpret_cast<const long long*>(d_crop_offsets), d_headers,
                                 d_out_n_valid, (cudaStream_t)stream));
    return DTSDF_OK;
}

int dtsdf_voxelize_sampled_batch(const dtsdf_config* cfg, double focal, int32_t n_frames,
                                 const float* d_depth_packed, const int64_t* d_crop_offsets,
                                 const int32_t* d_headers, const int32_t* d_sample_idx, int32_t point_num,
                                 int32_t max_crop_px, float* d_out_vol, float* d_out_max_l,
                                 float* d_out_mid_p, void* stream) {
    int rc = check_cfg(cfg);
    if (rc != DTSDF_OK) return rc;
    if (n_frames < 0 || point_num <= 0 || !(focal > 0.0)) return DTSDF_ERR_INVALID_ARG;
    if (n_frames == 0) return DTSDF_OK;
    if (!d_depth_packed || !d_crop_offsets || !d_headers || !d_sample_idx || !d_out_vol) return DTSDF_ERR_INVALID_ARG;
    dtsdf::BatchArgs a;
    a.depth = d_depth_packed;
    a.offsets = reinterpret_cast<const long long*>(d_crop_offsets);
    a.headers = d_headers;
    a.aux = nullptr;
    a.vol = d_out_vol;
    a.max_l = d_out_max_l;
    a.mid_p = d_out_mid_p;
    a.n_frames = n_frames;
    a.cap_floats = 0;
    a.bbox_mode = DTSDF_BBOX_SAMPLED;
    a.center_x = cfg->center_x;
    a.center_y = cfg->center_y;
    a.focal = cfg->focal;
    a.trunc_mul = cfg->trunc_mul;
    a.joints = nullptr;
    a.out_joints = nullptr;
    a.joint_flags = 0;
    a.sample_idx = d_sample_idx;
    a.point_num = point_num;
    a.focal64 = focal;
    CK(dtsdf::launch_voxelize(cfg->voxel_res, cfg->layout, a, max_crop_px, (cudaStream_t)stream));
    return DTSDF_OK;
}

int dtsdf_voxelize_aug_batch(const dtsdf_config* cfg, int32_t n_frames, const float* d_depth_packed,
                             const int64_t* d_crop_offsets, const int32_t* d_headers, const float* d_aug,
                             int32_t max_crop_px, const float* d_joints, int32_t joint_flags,
                             float* d_out_vol, float* d_out_max_l, float* d_out_mid_p, float* d_out_joints,
                             void* stream) {
    int rc = check_cfg(cfg);
    if (rc != DTSDF_OK) return rc;
    if (n_frames < 0) return DTSDF_ERR_INVALID_ARG;
    if (n_frames == 0) return DTSDF_OK;
    if (!d_depth_packed || !d_crop_offsets || !d_headers || !d_out_vol || !d_aug) return DTSDF_ERR_INVALID_ARG;
    if ((d_joints != nullptr) != (d_out_joints != nullptr) || (joint_flags & ~3)) return DTSDF_ERR_INVALID_ARG;
    dtsdf::BatchArgs a;
    a.depth = d_depth_packed;
    a.offsets = reinterpret_cast<const long long*>(d_crop_offsets);
    a.headers = d_headers;
    a.aux = nullptr;
    a.vol = d_out_vol;
    a.max_l = d_out_max_l;
    a.mid_p = d_out_mid_p;
    a.n_frames = n_frames;
    a.cap_floats = 0;
    a.bbox_mode = DTSDF_BBOX_PIXELS;
    a.center_x = cfg->center_x;
    a.center_y = cfg->center_y;
    a.focal = cfg->focal;
    a.trunc_mul = cfg->trunc_mul;
    a.joints = d_joints;
    a.out_joints = d_out_joints;
    a.joint_flags = joint_flags;
    CK(dtsdf::launch_voxelize_aug(cfg->voxel_res, cfg->layout, a, d_aug, max_crop_px, (cudaStream_t)stream));
    return DTSDF_OK;
}

int dtsdf_point_cloud_batch(int32_t n_frames, double focal, const float* d_depth_packed,
                            const int64_t* d_crop_offsets, const int32_t* d_headers,
                            double* d_points, const int64_t* d_point_offsets, int32_t* d_n_valid,
                            void* stream) {
    if (n_frames < 0 || !(focal > 0.0)) return DTSDF_ERR_INVALID_ARG;
    if (n_frames == 0) return DTSDF_OK;
    if (!d_depth_packed || !d_crop_offsets || !d_headers || !d_n_valid) return DTSDF_ERR_INVALID_ARG;
    if (d_points && !d_point_offsets) return DTSDF_ERR_INVALID_ARG;
    CK(dtsdf::launch_point_cloud(n_frames, focal, d_depth_packed,
                                 reinterpret_cast<const long long*>(d_crop_offsets), d_headers, d_points,
                                 reinterpret_cast<const long long*>(d_point_offsets), d_n_valid,
                                 (cudaStream_t)stream));
    return DTSDF_OK;
}

int dtsdf_resample_extrema_batch(int32_t n_frames, int32_t point_num, const double* d_points,
                                 const int64_t* d_point_offsets, const int32_t* d_indices,
                                 double* d_out_pc, float* d_out_extrema, void* stream) {
    if (n_frames < 0 || point_num <= 0) return DTSDF_ERR_INVALID_ARG;
    if (n_frames == 0) return DTSDF_OK;
    if (!d_points || !d_point_offsets || !d_out_extrema) return DTSDF_ERR_INVALID_ARG;
    CK(dtsdf::launch_resample_extrema(n_frames, point_num, d_points,
                                      reinterpret_cast<const long long*>(d_point_offsets), d_indices,
                                      d_out_pc, d_out_extrema, (cudaStream_t)stream));
    return DTSDF_OK;
}

int dtsdf_sample_cloud_batch(int32_t n_frames, double focal, int32_t point_num, const float* d_depth_packed,
                             const int64_t* d_crop_offsets, const int32_t* d_headers, const int32_t* d_indices,
                             int32_t max_crop_px, double* d_out_pc, float* d_out_extrema, int32_t* d_out_n_valid,
                             void* stream) {
    static_assert(DTSDF_SAMPLE_CLOUD_MAX_PX == 22528, "header constant");
    if (n_frames < 0 || point_num <= 0 || !(focal > 0.0)) return DTSDF_ERR_INVALID_ARG;
    if (n_frames == 0) return DTSDF_OK;
    if (!d_depth_packed || !d_crop_offsets || !d_headers || !d_out_extrema) return DTSDF_ERR_INVALID_ARG;
    if (max_crop_px <= 0 || max_crop_px > dtsdf::sample_cloud_max_pixels()) return DTSDF_ERR_UNSUPPORTED;
    if (DTSDF_SAMPLE_CLOUD_MAX_PX != dtsdf::sample_cloud_max_pixels()) return DTSDF_ERR_INTERNAL;
    CK(dtsdf::launch_sample_cloud(n_frames, focal, point_num, d_depth_packed,
                                  reinterpret_cast<const long long*>(d_crop_offsets), d_headers, d_indices, d_out_pc,
                                  d_out_extrema, d_out_n_valid, (cudaStream_t)stream));
    return DTSDF_OK;
}

int dtsdf_normalize_joints_batch(int32_t n_frames, const float* d_joints, const float* d_max_l,
                                 const float* d_mid_p, int32_t clamp01, float* d_out, void* stream) {
    if (n_frames < 0) return DTSDF_ERR_INVALID_ARG;
    if (n_frames == 0) return DTSDF_OK;
    if (!d_joints || !d_max_l || !d_mid_p || !d_out) return DTSDF_ERR_INVALID_ARG;
    CK(dtsdf::launch_normalize_joints(n_frames, d_joints, d_max_l, d_mid_p, clamp01, d_out,
                                      (cudaStream_t)stream));
    return DTSDF_OK;
}

int dtsdf_pack_volumes_batch(int32_t voxel_res, int32_t n_frames, const float* d_vol, uint8_t* d_codes,
                             uint32_t* d_lit_count, float* d_lits, void* stream) {
    if (voxel_res != 32 && voxel_res != 48 && voxel_res != 64) return DTSDF_ERR_UNSUPPORTED;
    if (n_frames < 0) return DTSDF_ERR_INVALID_ARG;
    if (n_frames == 0) return DTSDF_OK;
    if (!d_vol || !d_codes || !d_lit_count || !d_lits) return DTSDF_ERR_INVALID_ARG;
    const int r3 = voxel_res * voxel_res * voxel_res;
    CK(dtsdf::launch_pack_strided(n_frames, r3, d_vol, 3ll * r3, d_codes, d_lit_count, d_lits, (cudaStream_t)stream));
    return DTSDF_OK;
}

int dtsdf_unpack_volumes_batch(int32_t voxel_res, int32_t n_frames, const uint8_t* d_codes,
                               const uint32_t* d_lit_count, const float* d_lits, float* d_out_vol, void* stream) {
    if (voxel_res != 32 && voxel_res != 48 && voxel_res != 64) return DTSDF_ERR_UNSUPPORTED;
    if (n_frames < 0) return DTSDF_ERR_INVALID_ARG;
    if (n_frames == 0) return DTSDF_OK;
    if (!d_codes || !d_lit_count || !d_lits || !d_out_vol) return DTSDF_ERR_INVALID_ARG;
    const int r3 = voxel_res * voxel_res * voxel_res;
    CK(dtsdf::launch_unpack_strided(n_frames, r3, d_codes, d_lit_count, d_lits, 3ll * r3, d_out_vol, (cudaStream_t)stream));
    return DTSDF_OK;
}

// ------------------------------------------------------------------ context
int dtsdf_ctx_create(int32_t device, dtsdf_ctx** out) {
    if (!out) return DTSDF_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    if (n == 0) return DTSDF_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return DTSDF_ERR_INVALID_ARG;
    USE_DEVICE(device);
    int major = 0;
    CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "device %d is sm_%d?, this library is built for sm_100a only",
                 device, major * 10);
        return DTSDF_ERR_NO_DEVICE;
    }
    dtsdf_ctx* c = new (std::nothrow) dtsdf_ctx();
    if (!c) return DTSDF_ERR_ALLOC;
    c->device = device;
    cudaError_t err = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&c->s_run, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&c->s_lit, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&c->s_idx, cudaStreamNonBlocking);
    for (int i = 0; i < dtsdf_ctx::NSLOT && err == cudaSuccess; ++i) {
        err = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming);
        // the host waits on ev_run while its cores expand packed volumes: sleep, do not spin
        if (err == cudaSuccess) err = cudaEventCreateWithFlags(&c->ev_run[i], cudaEventDisableTiming | cudaEventBlockingSync);
        if (err == cudaSuccess) err = cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming);
        if (err == cudaSuccess) err = cudaEventCreateWithFlags(&c->ev_meta[i], cudaEventDisableTiming);
    }
    if (err != cudaSuccess) {
        dtsdf_ctx_destroy(c);
        return cuda_fail(err, "context streams/events");
    }
    *out = c;
    return DTSDF_OK;
}

void dtsdf_ctx_destroy(dtsdf_ctx* c) {
    if (!c) return;
    DeviceGuard guard(c->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < dtsdf_ctx::NSLOT; ++i) {
        c->d_depth[i].release(); c->d_off[i].release(); c->d_hdr[i].release(); c->d_aux[i].release();
        c->d_vol[i].release(); c->d_maxl[i].release(); c->d_mid[i].release();
        c->d_cnt[i].release(); c->d_loff[i].release(); c->d_codes[i].release(); c->d_lits[i].release();
        c->m_codes[i].release(); c->m_loff[i].release(); c->m_lits[i].release(); c->m_maxl[i].release(); c->m_mid[i].release();
        c->m_idx[i].release(); c->m_pc[i].release(); c->p_idx[i].release(); c->p_pc[i].release(); c->p_ext[i].release();
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_run[i]) cudaEventDestroy(c->ev_run[i]);
        if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
        if (c->ev_meta[i]) cudaEventDestroy(c->ev_meta[i]);
    }
    c->d_points.release(); c->d_idx.release(); c->d_pc.release(); c->d_ext.release(); c->d_misc.release();
    c->d_j_in.release(); c->d_j_out.release(); c->d_j_ml.release(); c->d_j_mp.release();
    c->b_depth.release(); c->b_off.release(); c->b_hdr.release(); c->b_points.release(); c->b_poff.release();
    c->b_nvalid.release(); c->b_idx.release(); c->b_pc.release(); c->b_ext.release(); c->b_vol.release();
    c->b_maxl.release(); c->b_mid.release();
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_run) cudaStreamDestroy(c->s_run);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->s_lit) cudaStreamDestroy(c->s_lit);
    if (c->s_idx) cudaStreamDestroy(c->s_idx);
    for (cudaEvent_t e : c->ev_piece) cudaEventDestroy(e);
    c->m_nvalid.release();
    cudaGetLastError();
    delete c;
}

int dtsdf_ctx_set_option(dtsdf_ctx* c, int32_t option, int64_t value) {
    if (!c) return DTSDF_ERR_INVALID_ARG;
    switch (option) {
        case DTSDF_OPT_PACKED_D2H:
            c->opt_packed = value != 0;
            return DTSDF_OK;
        case DTSDF_OPT_HOST_THREADS:
            if (value < 0 || value > 256) return DTSDF_ERR_INVALID_ARG;
            c->opt_threads = (int)value;
            return DTSDF_OK;
        case DTSDF_OPT_CHUNK_MB:
            if (value < 1 || value > 4096) return DTSDF_ERR_INVALID_ARG;
            c->opt_chunk_mb = (int)value;
            return DTSDF_OK;
        case DTSDF_OPT_LITERAL_PUSH:
            c->opt_push = value != 0;
            return DTSDF_OK;
        case DTSDF_OPT_TEST_FAIL_CHUNK:
            c->opt_fail_chunk = value < 0 ? -1 : (int)std::min<int64_t>(value, 1 << 30);
            return DTSDF_OK;
        default:
            return DTSDF_ERR_UNSUPPORTED;
    }
}

int dtsdf_ctx_get_stat(dtsdf_ctx* c, int32_t stat, int64_t* value) {
    if (!c || !value) return DTSDF_ERR_INVALID_ARG;
    if (stat != DTSDF_STAT_D2H_BYTES) return DTSDF_ERR_UNSUPPORTED;
    *value = (int64_t)c->stat_d2h_bytes.load();
    return DTSDF_OK;
}

int dtsdf_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return DTSDF_ERR_INVALID_ARG;
    *ptr = nullptr;
    if (bytes == 0) return DTSDF_OK;
    CK(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault));
    return DTSDF_OK;
}

void dtsdf_host_free(void* ptr) {
    if (ptr) { cudaFreeHost(ptr); cudaGetLastError(); }
}

}  // extern "C"

// A chunk's packed form has arrived in host memory: split its frames over the expansion threads
// (called by the finalizer thread of dtsdf_voxelize_host).  Never throws, never allocates.
static void expand_chunk(const dtsdf_ctx::ChunkRecord* r) {
    dtsdf_ctx* c = r->ctx;
    const unsigned long long* loff = (const unsigned long long*)c->m_loff[r->slot].p;
    c->stat_d2h_bytes += r->fixed_bytes + 12ll * (long long)loff[r->nf];
    // max_l / mid_p came back through the slot's pinned staging buffers (the caller's arrays may be
    // pageable: a device -> host copy straight into them would block the submitting thread)
    if (r->dst_maxl) memcpy(r->dst_maxl, c->m_maxl[r->slot].p, (size_t)r->nf * 4);
    if (r->dst_mid) memcpy(r->dst_mid, c->m_mid[r->slot].p, (size_t)r->nf * 12);
    for (int j = 0; j < r->jobs; ++j) {
        ExpandJob job;
        job.chunk = r;
        job.f0 = (int32_t)((int64_t)r->nf * j / r->jobs);
        job.f1 = (int32_t)((int64_t)r->nf * (j + 1) / r->jobs);
        if (!c->pool.submit(job)) run_expand_job(job);
    }
}

namespace {
// Second phase of the index draw of one pipeline chunk (first phase: HostPipeline::scan_chunk): frame i's
// draws are produced from the snapshot of the stream taken where they begin - any frame on any thread.
struct DrawRecord {
    const dtsdf::DrawSpan* spans;      // [frames of the chunk]
    int32_t point_num;
    int32_t* out;                      // [frames][point_num], pinned staging of the chunk's slot
    Latch* latch;
};
void run_draw_job(const ExpandJob& j) {
    const DrawRecord* r = static_cast<const DrawRecord*>(j.chunk);
    for (int32_t f = j.f0; f < j.f1; ++f) {
        const dtsdf::DrawSpan& sp = r->spans[f];
        int32_t* idx = r->out + (size_t)f * r->point_num;
        const int32_t lead = r->point_num - sp.count;          // pre/process.py:76-79: every point once, then repeats
        for (int32_t k = 0; k < lead; ++k) idx[k] = k;
        dtsdf::draw_fill(sp, idx + lead);
    }
    r->latch->done();
}

void run_expand_job(const ExpandJob& j) {
    if (j.draw) { run_draw_job(j); return; }
    const dtsdf_ctx::ChunkRecord* r = static_cast<const dtsdf_ctx::ChunkRecord*>(j.chunk);
    dtsdf_ctx* c = r->ctx;
    const unsigned char* codes = (const unsigned char*)c->m_codes[r->slot].p;
    const unsigned long long* loff = (const unsigned long long*)c->m_loff[r->slot].p;
    const float* lits = (const float*)c->m_lits[r->slot].p;
    if (r->keep) {                               // packed result: this job's share of the chunk, as it is
        const size_t cb = r->r3 / 4;
        memcpy(r->keep->codes + (size_t)j.f0 * cb, codes + (size_t)j.f0 * cb, (size_t)(j.f1 - j.f0) * cb);
        memcpy(r->keep->lits + 3 * loff[j.f0], lits + 3 * loff[j.f0], (size_t)(loff[j.f1] - loff[j.f0]) * 12);
        if (j.f0 == 0) memcpy(r->keep->loff, loff, (size_t)(r->nf + 1) * 8);
        c->latch[r->slot].done();
        return;
    }
    if (r->dst_pc && j.f1 > j.f0)               // the sampled clouds came back through pinned staging
        memcpy(r->dst_pc + (size_t)j.f0 * r->pc_doubles, (const double*)c->m_pc[r->slot].p + (size_t)j.f0 * r->pc_doubles,
               (size_t)(j.f1 - j.f0) * r->pc_doubles * 8);
    for (int32_t f = j.f0; f < j.f1; ++f)
        dtsdf::decode_frame(codes + (size_t)f * (r->r3 / 4), lits + 3 * loff[f], (size_t)(loff[f + 1] - loff[f]), r->r3,
                            r->dst + (size_t)f * r->vol_floats);
    c->latch[r->slot].done();
}
}  // namespace

// ------------------------------------------------------------------ host pipeline
// The batch is cut into chunks of whole frames; chunk k's host->device copy (s_in),
// kernel (s_run) and device->host copy (s_out) run on their own streams with NSLOT
// chunk buffers in flight, so for large batches the PCIe traffic in both directions
// overlaps the kernel.
//
// Error handling: the submission loop is one function (submit_chunks) whose every failure returns a
// code; its caller then takes the ONE exit path: cancel what has not been handed to the expansion threads,
// join the finalizer, wait for every expansion job that did start, synchronise the three streams - so
// that nothing, neither a DMA nor a host thread, writes into the caller's buffers after the call has
// returned, whatever the outcome.
namespace {

// pre/process.py:70-84 for n frames in order: fewer points than point_num -> all of them once, then random
// repeats; otherwise point_num draws with replacement.  A frame without points makes the reference's
// np.random.randint(0, 0) raise ValueError ("low >= high"): DTSDF_ERR_EMPTY_CLOUD, the stream stays where the
// reference's would (advanced by the frames before it).
int resample_indices(uint32_t* key, int* pos, int32_t n, const int32_t* n_valid, int32_t point_num, int32_t* out) {
    for (int32_t i = 0; i < n; ++i) {
        const int32_t nv = n_valid[i];
        int32_t* idx = out + (size_t)i * point_num;
        if (nv <= 0) return DTSDF_ERR_EMPTY_CLOUD;
        if (nv < point_num) {
            for (int32_t k = 0; k < nv; ++k) idx[k] = k;
            dtsdf::np_legacy_randint(key, pos, (uint32_t)nv, point_num - nv, idx + nv);
        } else {
            dtsdf::np_legacy_randint(key, pos, (uint32_t)nv, point_num, idx);
        }
    }
    return DTSDF_OK;
}

struct HostPipeline {
    dtsdf_ctx* c;
    const dtsdf_config* cfg;
    int32_t n_frames;
    const float* h_depth;
    const int64_t* h_off;
    const int32_t* h_hdr;
    const float* h_aux;
    float* h_vol;
    float* h_maxl;
    float* h_mid;
    int aux_w;
    size_t r3, vol_floats;
    int32_t per_chunk;
    int nchunk;
    bool packed, single;
    int nthreads;
    // DataProcess.process() mode (dtsdf_process_host): crops, tables and point clouds of the whole batch are
    // resident in the context (b_*); per chunk the host draws the resampling indices from the caller's
    // MT19937 state, the device gathers the sampled clouds and reduces the extrema that feed the voxeliser
    bool process_mode = false;
    int32_t point_num = 0;
    uint32_t* mt_key = nullptr;
    int* mt_pos = nullptr;
    double* h_pc = nullptr;
    double focal64 = 0.0;
    // the index draw in two phases (csrc/dtsdf_rng.cpp): this thread finds where every frame's draws begin
    // (sequential, nothing written), the pool threads produce them; the scan runs AHEAD chunks in front of the
    // chunk being submitted, so the pool works on chunk k+1's draws while this thread scans chunk k+2
    static const int AHEAD = 2;                    // chunks scanned ahead of the one being submitted
    std::vector<dtsdf::DrawSpan> spans[AHEAD + 1];
    DrawRecord draw_rec[AHEAD + 1];
    Latch draw_latch[AHEAD + 1];
    int scan_chunk(int k);
    // DTSDF_TRACE=1: where the submitting thread's time went (seconds), printed to stderr at the end of the call
    double tr_scan = 0, tr_draw_wait = 0, tr_slot_wait = 0, tr_piece_wait = 0, tr_submit = 0, tr_exit = 0;
    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    // finalizer: takes the chunks in order, polls the event behind each one's last copy, hands the
    // chunk to the expansion threads
    std::mutex qm;
    std::condition_variable qcv;
    int fq[dtsdf_ctx::NSLOT * 2];            // slots queued for the finalizer (at most NSLOT at a time)
    int fq_head = 0, fq_count = 0;
    bool qdone = false;
    std::atomic<bool> cancel{false};
    std::atomic<bool> failed{false};           // the finalizer saw a CUDA failure (a chunk was dropped)
    std::atomic<bool> no_memory{false};
    dtsdf_packed_host* result = nullptr;       // dtsdf_voxelize_host_packed: the chunks stay packed, in here
    std::thread fin;

    // Two stages per chunk, chunks in order, one thread:
    //   A  codes, literal offsets, max_l / mid_p have landed (ev_meta): the host now knows how many literals
    //      the chunk holds and issues ONE copy-engine transfer of exactly that many bytes on s_lit (a kernel
    //      pushing the device-sized stream into mapped memory with SM stores ran at 20-24 GB/s and was the
    //      longest item on the device side; the copy engine moves it at PCIe speed, and with four chunks in
    //      flight the host round trip in between costs nothing);
    //   B  the literals have landed (ev_run): hand the chunk to the expansion threads.
    void finalizer() {
        cudaSetDevice(c->device);
        constexpr int QN = dtsdf_ctx::NSLOT * 2;
        int qa[QN], qb[QN], a_head = 0, a_n = 0, b_head = 0, b_n = 0;
        int spins = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(qm);
                if (a_n == 0 && b_n == 0) qcv.wait(lk, [&] { return fq_count > 0 || qdone; });
                while (fq_count > 0) {
                    qa[(a_head + a_n++) % QN] = fq[fq_head];
                    fq_head = (fq_head + 1) % QN;
                    --fq_count;
                }
                if (a_n == 0 && b_n == 0 && qdone) return;
            }
            bool progress = false;
            const bool dead = cancel.load(std::memory_order_relaxed);
            if (a_n > 0) {
                const int slot = qa[a_head];
                const cudaError_t e = dead ? cudaErrorUnknown : cudaEventQuery(c->ev_meta[slot]);
                if (e != cudaErrorNotReady) {
                    a_head = (a_head + 1) % QN; --a_n; progress = true;
                    bool ok = e == cudaSuccess;
                    if (ok && c->chunk[slot].keep) {
                        const dtsdf_ctx::ChunkRecord& r = c->chunk[slot];
                        const unsigned long long total = ((const unsigned long long*)c->m_loff[slot].p)[r.nf];
                        try {
                            r.keep->bind(r.nf, r.r3, total);
                            result->bytes += (long long)r.nf * (long long)(r.r3 / 4) + 8ll * (r.nf + 1) + 12ll * (long long)total;
                        } catch (...) {
                            ok = false;
                            no_memory.store(true);
                        }
                    }
                    if (ok && !c->opt_push) {
                        const dtsdf_ctx::ChunkRecord& r = c->chunk[slot];
                        const unsigned long long total = ((const unsigned long long*)c->m_loff[slot].p)[r.nf];
                        if (total > 0)
                            ok = cudaMemcpyAsync(c->m_lits[slot].p, c->d_lits[slot].p, (size_t)total * 12, cudaMemcpyDeviceToHost,
                                                 c->s_lit) == cudaSuccess;
                        ok = ok && cudaEventRecord(c->ev_run[slot], c->s_lit) == cudaSuccess;
                    }
                    // a failed stream or a cancelled call: nothing of this chunk reaches the caller's buffers
                    if (!ok) { cudaGetLastError(); failed.store(true); c->latch[slot].reset(); }
                    else qb[(b_head + b_n++) % QN] = slot;
                }
            }
            if (b_n > 0) {
                const int slot = qb[b_head];
                const cudaError_t e = dead ? cudaErrorUnknown : (c->opt_push ? cudaSuccess : cudaEventQuery(c->ev_run[slot]));
                if (e != cudaErrorNotReady) {
                    b_head = (b_head + 1) % QN; --b_n; progress = true;
                    if (e != cudaSuccess) { cudaGetLastError(); failed.store(true); c->latch[slot].reset(); }
                    else expand_chunk(&c->chunk[slot]);
                }
            }
            if (progress) spins = 0;
            else if (++spins < 64) _mm_pause();      // spin briefly, then yield: the ranks of a node share its cores
            else std::this_thread::yield();
        }
    }
    void enqueue_for_finalizer(int slot) {
        { std::lock_guard<std::mutex> g(qm); fq[(fq_head + fq_count) % (dtsdf_ctx::NSLOT * 2)] = slot; ++fq_count; }
        qcv.notify_one();
    }
    void stop_finalizer() {
        if (fin.joinable()) {
            { std::lock_guard<std::mutex> g(qm); qdone = true; }
            qcv.notify_all();
            fin.join();                  // it first works off (or, cancelled, drops) what is queued
        }
    }

    void configure();          // chunking, transfer route and thread count from the context options
    int submit_chunks();
    int run();
};

// Phase one of chunk k's index draw (pre/process.py:70-84, frames in order on the caller's stream), then the
// pool gets phase two.  With pieces in flight (dtsdf_process_host without clouds) the chunk's valid-pixel counts
// are waited for first.
int HostPipeline::scan_chunk(int k) {
    const int slot = k % dtsdf_ctx::NSLOT, par = k % (AHEAD + 1);
    const int32_t f0 = k * per_chunk;
    const int32_t nf = std::min(per_chunk, n_frames - f0);
    const double t_in = now();
    if (c->b_pieces > 0) {
        CK(cudaEventSynchronize(c->ev_piece[(size_t)k]));
        memcpy(c->b_nvalid_h.data() + f0, (const int32_t*)c->m_nvalid.p + f0, (size_t)nf * 4);
    }
    tr_piece_wait += now() - t_in;
    // the staging buffer of the slot was last read by the host -> device copy of chunk k - NSLOT
    if (k >= dtsdf_ctx::NSLOT) CK(cudaEventSynchronize(c->ev_in[slot]));
    int rc;
    if ((rc = c->m_idx[slot].reserve((size_t)nf * point_num * 4)) != DTSDF_OK) return rc;
    spans[par].resize((size_t)nf);
    for (int32_t i = 0; i < nf; ++i) {
        const int32_t nv = c->b_nvalid_h[(size_t)f0 + i];
        if (nv <= 0) return DTSDF_ERR_EMPTY_CLOUD;       // np.random.randint(0, 0): ValueError; the stream stays here
        dtsdf::np_legacy_randint_skip(mt_key, mt_pos, (uint32_t)nv, nv < point_num ? point_num - nv : point_num,
                                      &spans[par][(size_t)i]);
    }
    DrawRecord& r = draw_rec[par];
    r.spans = spans[par].data(); r.point_num = point_num; r.out = (int32_t*)c->m_idx[slot].p; r.latch = &draw_latch[par];
    const int workers = c->pool.size();
    const int jobs = std::max(1, std::min<int>(nf, workers));
    draw_latch[par].add(jobs);
    for (int j = 0; j < jobs; ++j) {
        ExpandJob job;
        job.chunk = &r; job.draw = true;
        job.f0 = (int32_t)((int64_t)nf * j / jobs);
        job.f1 = (int32_t)((int64_t)nf * (j + 1) / jobs);
        if (workers == 0 || !c->pool.submit(job)) run_expand_job(job);
    }
    tr_scan += now() - t_in;
    return DTSDF_OK;
}

int HostPipeline::submit_chunks() {
    const cudaStream_t st_in = single ? c->s_run : (process_mode ? c->s_idx : c->s_in);
    const cudaStream_t st_out = single ? c->s_run : c->s_out;
    std::vector<int64_t> rebased;
    int rc;
    for (int k = 0; k < nchunk; ++k) {
        const int slot = k % dtsdf_ctx::NSLOT;
        const int32_t f0 = k * per_chunk;
        const int32_t nf = std::min(per_chunk, n_frames - f0);
        int64_t e0 = 0, e1 = 0;
        int32_t max_px = 0;
        if (!process_mode) {
            const int32_t* hl = h_hdr + 6 * (size_t)(f0 + nf - 1);
            const int64_t nlast = ((int64_t)hl[4] - hl[2]) * ((int64_t)hl[5] - hl[3]);
            e0 = h_off[f0];
            e1 = h_off[f0 + nf - 1] + ((nlast + 3) & ~(int64_t)3);
            rebased.resize(nf);
        }
        for (int32_t i = 0; !process_mode && i < nf; ++i) {
            const int32_t* h = h_hdr + 6 * (size_t)(f0 + i);
            const int64_t n = ((int64_t)h[4] - h[2]) * ((int64_t)h[5] - h[3]);
            max_px = std::max<int32_t>(max_px, (int32_t)n);
            rebased[i] = h_off[f0 + i] - e0;
        }
        // the slot's buffers are free once its previous chunk has left them: plain mode, its D2H
        // has finished (which implies its kernel has); packed mode, the host threads are done
        // expanding it (its kernels were waited for before they started)
        const double t_slot = now();
        if (packed) c->latch[slot].wait();
        else if (k >= dtsdf_ctx::NSLOT) CK(cudaEventSynchronize(c->ev_out[slot]));
        tr_slot_wait += now() - t_slot;
        if (k == c->opt_fail_chunk) CK(cudaErrorUnknown);          // injected failure (DTSDF_OPT_TEST_FAIL_CHUNK)
        if (!process_mode) {
            if ((rc = c->d_depth[slot].reserve((size_t)(e1 - e0) * 4)) != DTSDF_OK) return rc;
            if ((rc = c->d_off[slot].reserve((size_t)nf * 8)) != DTSDF_OK) return rc;
            if ((rc = c->d_hdr[slot].reserve((size_t)nf * 24)) != DTSDF_OK) return rc;
        }
        if (aux_w && (rc = c->d_aux[slot].reserve((size_t)nf * aux_w * 4)) != DTSDF_OK) return rc;
        if ((rc = c->d_vol[slot].reserve((size_t)nf * vol_floats * 4)) != DTSDF_OK) return rc;
        if ((rc = c->d_maxl[slot].reserve((size_t)nf * 4)) != DTSDF_OK) return rc;
        if ((rc = c->d_mid[slot].reserve((size_t)nf * 12)) != DTSDF_OK) return rc;
        if (packed) {
            if ((rc = c->d_cnt[slot].reserve((size_t)nf * 4)) != DTSDF_OK) return rc;
            if ((rc = c->d_loff[slot].reserve((size_t)(nf + 1) * 8)) != DTSDF_OK) return rc;
            if ((rc = c->m_codes[slot].reserve((size_t)nf * (r3 / 4))) != DTSDF_OK) return rc;
            if ((rc = c->m_loff[slot].reserve((size_t)(nf + 1) * 8)) != DTSDF_OK) return rc;
            if ((rc = c->m_lits[slot].reserve((size_t)nf * vol_floats * 4 + 16)) != DTSDF_OK) return rc;   // worst case
            if ((rc = c->m_maxl[slot].reserve((size_t)nf * 4)) != DTSDF_OK) return rc;
            if ((rc = c->m_mid[slot].reserve((size_t)nf * 12)) != DTSDF_OK) return rc;
            if ((rc = c->d_codes[slot].reserve((size_t)nf * (r3 / 4))) != DTSDF_OK) return rc;
            if ((rc = c->d_lits[slot].reserve((size_t)nf * vol_floats * 4 + 16)) != DTSDF_OK) return rc;
        }

        if (!process_mode) {
            // H2D on s_in.  `rebased` is pageable host memory: cudaMemcpyAsync stages it before
            // returning, so reusing the vector for the next chunk is safe.
            CK(cudaMemcpyAsync(c->d_depth[slot].p, h_depth + e0, (size_t)(e1 - e0) * 4, cudaMemcpyHostToDevice, st_in));
            CK(cudaMemcpyAsync(c->d_off[slot].p, rebased.data(), (size_t)nf * 8, cudaMemcpyHostToDevice, st_in));
            CK(cudaMemcpyAsync(c->d_hdr[slot].p, h_hdr + 6 * (size_t)f0, (size_t)nf * 24, cudaMemcpyHostToDevice, st_in));
            if (aux_w)
                CK(cudaMemcpyAsync(c->d_aux[slot].p, h_aux + (size_t)aux_w * f0, (size_t)nf * aux_w * 4,
                                   cudaMemcpyHostToDevice, st_in));
        } else {
            // pre/process.py:70-84 for the frames of this chunk, in frame order, on the caller's stream
            const size_t ni = (size_t)nf * point_num;
            if ((rc = c->p_idx[slot].reserve(ni * 4)) != DTSDF_OK) return rc;
            if ((rc = c->p_ext[slot].reserve((size_t)nf * 24)) != DTSDF_OK) return rc;
            if (h_pc) {
                if ((rc = c->p_pc[slot].reserve(ni * 24)) != DTSDF_OK) return rc;
                if (packed && (rc = c->m_pc[slot].reserve(ni * 24)) != DTSDF_OK) return rc;
            }
            for (int a = 0; k == 0 && a < AHEAD && a < nchunk; ++a)
                if ((rc = scan_chunk(a)) != DTSDF_OK) return rc;
            const double t_draw = now();
            draw_latch[k % (AHEAD + 1)].wait();
            tr_draw_wait += now() - t_draw;
            CK(cudaMemcpyAsync(c->p_idx[slot].p, c->m_idx[slot].p, ni * 4, cudaMemcpyHostToDevice, st_in));
            if (c->b_pieces > 0) CK(cudaStreamWaitEvent(c->s_run, c->ev_piece[(size_t)k], 0));      // the chunk's crops
        }
        if (!single) {
            CK(cudaEventRecord(c->ev_in[slot], st_in));
            CK(cudaStreamWaitEvent(c->s_run, c->ev_in[slot], 0));
        }

        // kernel(s) on s_run
        if (!process_mode) {
            rc = dtsdf_voxelize_batch(cfg, nf, (const float*)c->d_depth[slot].p, (const int64_t*)c->d_off[slot].p,
                                      (const int32_t*)c->d_hdr[slot].p, aux_w ? (const float*)c->d_aux[slot].p : nullptr,
                                      max_px, (float*)c->d_vol[slot].p, (float*)c->d_maxl[slot].p,
                                      (float*)c->d_mid[slot].p, c->s_run);
        } else if (!h_pc) {
            // nobody wants the sampled clouds: ranks, sample bitmap and extrema inside the voxeliser launch
            rc = dtsdf_voxelize_sampled_batch(cfg, focal64, nf, (const float*)c->b_depth.p, (const int64_t*)c->b_off.p + f0,
                                              (const int32_t*)c->b_hdr.p + 6 * (size_t)f0, (const int32_t*)c->p_idx[slot].p,
                                              point_num, c->b_max_px, (float*)c->d_vol[slot].p, (float*)c->d_maxl[slot].p,
                                              (float*)c->d_mid[slot].p, c->s_run);
        } else {
            // the sampled clouds are wanted: straight from the depth map when the crops allow it (no full cloud was
            // built then, see dtsdf_process_host), else gathered from the clouds built up front
            if (c->b_n == 0)
                rc = dtsdf_sample_cloud_batch(nf, focal64, point_num, (const float*)c->b_depth.p, (const int64_t*)c->b_off.p + f0,
                                              (const int32_t*)c->b_hdr.p + 6 * (size_t)f0, (const int32_t*)c->p_idx[slot].p,
                                              c->b_max_px, (double*)c->p_pc[slot].p, (float*)c->p_ext[slot].p, nullptr,
                                              c->s_run);
            else
                rc = dtsdf_resample_extrema_batch(nf, point_num, (const double*)c->b_points.p, (const int64_t*)c->b_poff.p + f0,
                                                  (const int32_t*)c->p_idx[slot].p, h_pc ? (double*)c->p_pc[slot].p : nullptr,
                                                  (float*)c->p_ext[slot].p, c->s_run);
            if (rc == DTSDF_OK)
                rc = dtsdf_voxelize_batch(cfg, nf, (const float*)c->b_depth.p, (const int64_t*)c->b_off.p + f0,
                                          (const int32_t*)c->b_hdr.p + 6 * (size_t)f0, (const float*)c->p_ext[slot].p,
                                          c->b_max_px, (float*)c->d_vol[slot].p, (float*)c->d_maxl[slot].p,
                                          (float*)c->d_mid[slot].p, c->s_run);
        }
        if (rc != DTSDF_OK) return rc;
        // while the GPU works on this chunk: where the next chunk's draws begin, and its draws on the pool
        if (process_mode && k + AHEAD < nchunk && (rc = scan_chunk(k + AHEAD)) != DTSDF_OK) return rc;

        if (packed) {
            // pack into device memory on the kernel stream; then, on s_out so that the next chunk's
            // kernels are not held up by PCIe: the literal stream (device-known size) by a copy kernel
            // into mapped pinned memory, codes and offsets (fixed size) by the copy engine.  (With
            // everything on one stream the transfers were 60 % of that stream's time.)
            void* dv_lits = nullptr;
            CK(cudaHostGetDevicePointer(&dv_lits, c->m_lits[slot].p, 0));
            CK(dtsdf::launch_pack(nf, (int)r3, (const float*)c->d_vol[slot].p, (unsigned*)c->d_cnt[slot].p,
                                  (unsigned long long*)c->d_loff[slot].p, nullptr,
                                  (unsigned char*)c->d_codes[slot].p, (float*)c->d_lits[slot].p, c->s_run));
            CK(cudaEventRecord(c->ev_out[slot], c->s_run));          // packed mode: "chunk packed"
            CK(cudaStreamWaitEvent(c->s_out, c->ev_out[slot], 0));
            if (c->opt_push)
                CK(dtsdf::launch_pack_push(nf, (const float*)c->d_lits[slot].p,
                                           (const unsigned long long*)c->d_loff[slot].p, (float*)dv_lits, c->s_out));
            CK(cudaMemcpyAsync(c->m_codes[slot].p, c->d_codes[slot].p, (size_t)nf * (r3 / 4), cudaMemcpyDeviceToHost, c->s_out));
            CK(cudaMemcpyAsync(c->m_loff[slot].p, c->d_loff[slot].p, (size_t)(nf + 1) * 8, cudaMemcpyDeviceToHost, c->s_out));
            if (h_maxl)
                CK(cudaMemcpyAsync(c->m_maxl[slot].p, c->d_maxl[slot].p, (size_t)nf * 4, cudaMemcpyDeviceToHost, c->s_out));
            if (h_mid)
                CK(cudaMemcpyAsync(c->m_mid[slot].p, c->d_mid[slot].p, (size_t)nf * 12, cudaMemcpyDeviceToHost, c->s_out));
            if (process_mode && h_pc)
                CK(cudaMemcpyAsync(c->m_pc[slot].p, c->p_pc[slot].p, (size_t)nf * point_num * 24, cudaMemcpyDeviceToHost, c->s_out));
            // behind the copies: an event for the finalizer thread
            dtsdf_ctx::ChunkRecord& r = c->chunk[slot];
            r.ctx = c; r.slot = slot; r.jobs = std::min<int>(nf, nthreads); r.nf = nf;
            r.r3 = r3; r.vol_floats = vol_floats; r.dst = h_vol ? h_vol + (size_t)f0 * vol_floats : nullptr;
            r.dst_maxl = h_maxl ? h_maxl + f0 : nullptr;
            r.dst_mid = h_mid ? h_mid + 3 * (size_t)f0 : nullptr;
            r.keep = result ? &result->chunks[(size_t)k] : nullptr;
            if (r.keep) { r.keep->f0 = f0; r.keep->nf = nf; }
            r.pc_doubles = (size_t)point_num * 3;
            r.dst_pc = (process_mode && h_pc) ? h_pc + (size_t)f0 * r.pc_doubles : nullptr;
            r.fixed_bytes = (long long)nf * (long long)(r3 / 4) + (long long)(nf + 1) * 8 +
                            (h_maxl ? 4ll * nf : 0ll) + (h_mid ? 12ll * nf : 0ll) + (r.dst_pc ? 24ll * nf * point_num : 0ll);
            CK(cudaEventRecord(c->ev_meta[slot], c->s_out));
            c->latch[slot].add(r.jobs);
            enqueue_for_finalizer(slot);
            continue;
        }
        // D2H on s_out
        if (!single) {
            CK(cudaEventRecord(c->ev_run[slot], c->s_run));
            CK(cudaStreamWaitEvent(c->s_out, c->ev_run[slot], 0));
        }
        CK(cudaMemcpyAsync(h_vol + (size_t)f0 * vol_floats, c->d_vol[slot].p, (size_t)nf * vol_floats * 4,
                           cudaMemcpyDeviceToHost, st_out));
        if (h_maxl)
            CK(cudaMemcpyAsync(h_maxl + f0, c->d_maxl[slot].p, (size_t)nf * 4, cudaMemcpyDeviceToHost, st_out));
        if (h_mid)
            CK(cudaMemcpyAsync(h_mid + 3 * (size_t)f0, c->d_mid[slot].p, (size_t)nf * 12, cudaMemcpyDeviceToHost, st_out));
        if (process_mode && h_pc)
            CK(cudaMemcpyAsync(h_pc + (size_t)f0 * point_num * 3, c->p_pc[slot].p, (size_t)nf * point_num * 24,
                               cudaMemcpyDeviceToHost, st_out));
        if (!single) CK(cudaEventRecord(c->ev_out[slot], c->s_out));
    }
    return DTSDF_OK;
}

void HostPipeline::configure() {
    const size_t R = (size_t)cfg->voxel_res;
    r3 = R * R * R;
    vol_floats = 3 * r3;
    // chunk size: about 64 MB of output per chunk by default, at least one frame
    per_chunk = (int32_t)std::max<size_t>(1, ((size_t)c->opt_chunk_mb << 20) / (vol_floats * 4));
    if (per_chunk > n_frames) per_chunk = n_frames;
    nchunk = (n_frames + per_chunk - 1) / per_chunk;
    // packed device -> host transfer: the GPU packs each chunk into mapped pinned memory, host
    // threads expand it into h_vol while the next chunks are in flight.  Small batches go back as
    // plain copies even when the packed transfer is on: its extra kernels, stream hand-offs and
    // thread wake-ups cost more than they save below ~32 frames (one frame: 350 us packed, 110 us
    // plain; 64 frames: 26 us per frame packed, 33 us plain).
    packed = result != nullptr || (c->opt_packed != 0 && (size_t)n_frames * vol_floats * 4 >= ((size_t)16 << 20));
    if (result) {
        result->n_frames = n_frames; result->voxel_res = (int32_t)R; result->r3 = r3; result->per_chunk = per_chunk;
        result->chunks.clear();
        result->chunks.resize((size_t)nchunk);
        result->bytes = 0;
    }
    // A single plain chunk needs no pipeline: everything on one stream, one synchronisation.
    single = !packed && nchunk == 1;
    c->stat_d2h_bytes = packed ? 0ll
                                 : (long long)n_frames * (long long)(vol_floats * 4 + (h_maxl ? 4 : 0) + (h_mid ? 12 : 0));
    nthreads = 1;
    if (packed) {
        // default: three quarters of the hardware threads, at most 8: the expansion saturates the host's
        // memory system with 8 threads (16-core box: 8 and 12 threads alike; 32-core box: 286 k volumes/s
        // with 8, 271 k with 12, 234 k with 16), and the calling thread and the CUDA driver need cores too
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        nthreads = c->opt_threads > 0 ? c->opt_threads : (int)std::min(8u, std::max(1u, hw - hw / 4));
    }
}

int HostPipeline::run() {
    if (packed) {
        c->pool.ensure(nthreads);                                  // (may throw: nothing is in flight yet)
        fin = std::thread([this] { finalizer(); });
    }
    int rc;
    const double t_run = now();
    try {
        rc = submit_chunks();
    } catch (const std::bad_alloc&) {
        rc = DTSDF_ERR_ALLOC;
    } catch (...) {
        rc = DTSDF_ERR_INTERNAL;
    }
    // ---- the one exit path
    tr_submit = now();
    if (rc != DTSDF_OK) cancel.store(true);
    stop_finalizer();
    if (rc == DTSDF_OK && failed.load()) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "a device -> host transfer of the host pipeline failed");
        rc = no_memory.load() ? DTSDF_ERR_ALLOC : DTSDF_ERR_CUDA;
    }
    for (int i = 0; i < dtsdf_ctx::NSLOT; ++i) c->latch[i].wait();         // expansion jobs that did start
    for (Latch& l : draw_latch) l.wait();                                  // draw jobs read this object's spans
    if (packed) { cudaStreamSynchronize(c->s_lit); cudaGetLastError(); }
    if (process_mode) { cudaStreamSynchronize(c->s_idx); cudaGetLastError(); }
    cudaError_t e_out = single ? cudaSuccess : cudaStreamSynchronize(c->s_out);
    cudaError_t e_run = cudaStreamSynchronize(c->s_run);
    cudaError_t e_in = single ? cudaSuccess : cudaStreamSynchronize(c->s_in);
    if (getenv("DTSDF_TRACE")) {
        const double t_end = now();
        fprintf(stderr, "[dtsdf trace] frames %d chunks %d x %d: call %.2f ms = submit loop %.2f (scan %.2f of which piece wait "
                        "%.2f; draw wait %.2f; slot wait %.2f) + exit path %.2f\n", n_frames, nchunk, per_chunk,
                (t_end - t_run) * 1e3, (tr_submit - t_run) * 1e3, tr_scan * 1e3, tr_piece_wait * 1e3, tr_draw_wait * 1e3,
                tr_slot_wait * 1e3, (t_end - tr_submit) * 1e3);
    }
    if (rc != DTSDF_OK) { cudaGetLastError(); return rc; }
    if (e_out != cudaSuccess) return cuda_fail(e_out, "cudaStreamSynchronize(s_out)");
    if (e_run != cudaSuccess) return cuda_fail(e_run, "cudaStreamSynchronize(s_run)");
    if (e_in != cudaSuccess) return cuda_fail(e_in, "cudaStreamSynchronize(s_in)");
    return DTSDF_OK;
}

}  // namespace

extern "C" {

static int voxelize_host_common(dtsdf_ctx* c, const dtsdf_config* cfg, int32_t n_frames,
                                const float* h_depth, int64_t packed_len, const int64_t* h_off,
                                const int32_t* h_hdr, const float* h_aux, float* h_vol, float* h_maxl,
                                float* h_mid, dtsdf_packed_host* result) {
    if (!c) return DTSDF_ERR_INVALID_ARG;
    int rc = check_cfg(cfg);
    if (rc != DTSDF_OK) return rc;
    if (n_frames < 0 || packed_len < 0) return DTSDF_ERR_INVALID_ARG;
    if (n_frames == 0) return DTSDF_OK;
    if (!h_depth || !h_off || !h_hdr || (!h_vol && !result)) return DTSDF_ERR_INVALID_ARG;
    if (cfg->bbox_mode != DTSDF_BBOX_PIXELS && !h_aux) return DTSDF_ERR_INVALID_ARG;
    // validate headers / offsets on the host (the per-frame shims raise on these)
    for (int32_t i = 0; i < n_frames; ++i) {
        const int32_t* h = h_hdr + 6 * (size_t)i;
        if (!header_ok(h)) return DTSDF_ERR_BAD_HEADER;
        const int64_t n = ((int64_t)h[4] - h[2]) * ((int64_t)h[5] - h[3]);
        if (h_off[i] < 0 || (h_off[i] & 3) || h_off[i] + ((n + 3) & ~(int64_t)3) > packed_len)
            return DTSDF_ERR_INVALID_ARG;
        if (i > 0 && h_off[i] < h_off[i - 1]) return DTSDF_ERR_INVALID_ARG;   // chunking needs ascending offsets
    }
    USE_DEVICE(c->device);
    return guarded([&]() -> int {
        HostPipeline p;
        p.c = c; p.cfg = cfg; p.n_frames = n_frames; p.h_depth = h_depth; p.h_off = h_off; p.h_hdr = h_hdr;
        p.h_aux = h_aux; p.h_vol = h_vol; p.h_maxl = h_maxl; p.h_mid = h_mid;
        p.aux_w = cfg->bbox_mode == DTSDF_BBOX_EXTREMA ? 6 : (cfg->bbox_mode == DTSDF_BBOX_GRID ? 5 : 0);
        p.result = result;
        p.configure();
        return p.run();
    });
}

int dtsdf_voxelize_host(dtsdf_ctx* c, const dtsdf_config* cfg, int32_t n_frames,
                        const float* h_depth, int64_t packed_len, const int64_t* h_off,
                        const int32_t* h_hdr, const float* h_aux, float* h_vol, float* h_maxl,
                        float* h_mid) {
    if (n_frames > 0 && !h_vol) return DTSDF_ERR_INVALID_ARG;
    return voxelize_host_common(c, cfg, n_frames, h_depth, packed_len, h_off, h_hdr, h_aux, h_vol, h_maxl, h_mid, nullptr);
}

int dtsdf_voxelize_host_packed(dtsdf_ctx* c, const dtsdf_config* cfg, int32_t n_frames,
                               const float* h_depth, int64_t packed_len, const int64_t* h_off,
                               const int32_t* h_hdr, const float* h_aux, dtsdf_packed_host** out,
                               float* h_maxl, float* h_mid) {
    if (!out) return DTSDF_ERR_INVALID_ARG;
    *out = nullptr;
    dtsdf_packed_host* h = new (std::nothrow) dtsdf_packed_host();
    if (!h) return DTSDF_ERR_ALLOC;
    if (cfg) { h->voxel_res = cfg->voxel_res; h->r3 = (size_t)cfg->voxel_res * cfg->voxel_res * cfg->voxel_res; }
    const int rc = voxelize_host_common(c, cfg, n_frames, h_depth, packed_len, h_off, h_hdr, h_aux, nullptr, h_maxl, h_mid, h);
    if (rc != DTSDF_OK) { delete h; return rc; }
    *out = h;
    return DTSDF_OK;
}

void dtsdf_packed_host_free(dtsdf_packed_host* h) { delete h; }

int dtsdf_packed_host_info(const dtsdf_packed_host* h, int32_t* n_frames, int32_t* voxel_res, int64_t* bytes) {
    if (!h) return DTSDF_ERR_INVALID_ARG;
    if (n_frames) *n_frames = h->n_frames;
    if (voxel_res) *voxel_res = h->voxel_res;
    if (bytes) *bytes = (int64_t)h->bytes.load();
    return DTSDF_OK;
}

int dtsdf_expand_frames(const dtsdf_packed_host* h, int32_t first, int32_t count, float* h_out, int32_t n_threads) {
    if (!h || first < 0 || count < 0 || (int64_t)first + count > h->n_frames) return DTSDF_ERR_INVALID_ARG;
    if (count == 0) return DTSDF_OK;
    if (!h_out) return DTSDF_ERR_INVALID_ARG;
    const size_t r3 = h->r3, vol_floats = 3 * r3;
    auto expand = [=](int32_t a, int32_t b) {
        for (int32_t f = a; f < b; ++f) {
            const dtsdf_packed_host::Chunk& ch = h->chunks[(size_t)(f / h->per_chunk)];
            const int32_t i = f - ch.f0;
            dtsdf::decode_frame(ch.codes + (size_t)i * (r3 / 4), ch.lits + 3 * ch.loff[i],
                                (size_t)(ch.loff[i + 1] - ch.loff[i]), r3, h_out + (size_t)(f - first) * vol_floats);
        }
    };
    int nt = n_threads > 0 ? n_threads : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    nt = std::min<int>(nt, (count + 3) / 4);              // a thread is not worth fewer than four frames
    if (nt <= 1) { expand(first, first + count); return DTSDF_OK; }
    return guarded([&]() -> int {
        std::vector<std::thread> th;
        th.reserve((size_t)nt);
        int started = 0;
        try {
            for (; started < nt - 1; ++started) {
                const int32_t a = first + (int32_t)((int64_t)count * started / nt), b = first + (int32_t)((int64_t)count * (started + 1) / nt);
                th.emplace_back(expand, a, b);
            }
        } catch (...) {
            // no more threads to be had: this one does the rest
        }
        expand(first + (int32_t)((int64_t)count * started / nt), first + count);
        for (auto& t : th) t.join();
        return DTSDF_OK;
    });
}

int dtsdf_point_cloud_host(dtsdf_ctx* c, double focal, const int32_t* h_header, const float* h_depth,
                           double* h_out_points, int32_t* n_valid) {
    if (!c || !h_header || !h_depth || !n_valid || !(focal > 0.0)) return DTSDF_ERR_INVALID_ARG;
    if (!header_ok(h_header)) return DTSDF_ERR_BAD_HEADER;
    USE_DEVICE(c->device);
    const int64_t n = ((int64_t)h_header[4] - h_header[2]) * ((int64_t)h_header[5] - h_header[3]);
    const int64_t n_pad = (n + 3) & ~(int64_t)3;
    int rc;
    if ((rc = c->d_depth[0].reserve((size_t)n_pad * 4)) != DTSDF_OK) return rc;
    if ((rc = c->d_points.reserve((size_t)n * 24)) != DTSDF_OK) return rc;
    // misc: [0..7] crop offset, [8..15] point offset, [16..39] header, [40..43] n_valid
    if ((rc = c->d_misc.reserve(64)) != DTSDF_OK) return rc;
    unsigned char misc[64];
    memset(misc, 0, sizeof(misc));
    memcpy(misc + 16, h_header, 24);
    cudaStream_t s = c->s_run;
    CK(cudaMemcpyAsync(c->d_depth[0].p, h_depth, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->d_misc.p, misc, 64, cudaMemcpyHostToDevice, s));
    unsigned char* dm = (unsigned char*)c->d_misc.p;
    rc = dtsdf_point_cloud_batch(1, focal, (const float*)c->d_depth[0].p, (const int64_t*)dm,
                                 (const int32_t*)(dm + 16), (double*)c->d_points.p, (const int64_t*)(dm + 8),
                                 (int32_t*)(dm + 40), s);
    if (rc != DTSDF_OK) return rc;
    CK(cudaMemcpyAsync(n_valid, dm + 40, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    c->last_n_valid = *n_valid;
    if (h_out_points && *n_valid > 0) {
        CK(cudaMemcpyAsync(h_out_points, c->d_points.p, (size_t)(*n_valid) * 24, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    return DTSDF_OK;
}

int dtsdf_resample_extrema_host(dtsdf_ctx* c, int32_t point_num, const double* h_points, int32_t n_points,
                                const int32_t* h_indices, double* h_out_pc, float* h_out_extrema) {
    if (!c || point_num <= 0 || !h_out_extrema) return DTSDF_ERR_INVALID_ARG;
    USE_DEVICE(c->device);
    cudaStream_t s = c->s_run;
    int rc;
    int32_t avail;
    if (h_points) {
        if (n_points <= 0) return DTSDF_ERR_INVALID_ARG;
        if ((rc = c->d_points.reserve((size_t)n_points * 24)) != DTSDF_OK) return rc;
        CK(cudaMemcpyAsync(c->d_points.p, h_points, (size_t)n_points * 24, cudaMemcpyHostToDevice, s));
        c->last_n_valid = n_points;
        avail = n_points;
    } else {
        if (c->last_n_valid <= 0) return DTSDF_ERR_INVALID_ARG;   // no resident cloud
        avail = c->last_n_valid;
    }
    if (h_indices) {
        for (int32_t i = 0; i < point_num; ++i)
            if (h_indices[i] < 0 || h_indices[i] >= avail) return DTSDF_ERR_INVALID_ARG;
        if ((rc = c->d_idx.reserve((size_t)point_num * 4)) != DTSDF_OK) return rc;
        CK(cudaMemcpyAsync(c->d_idx.p, h_indices, (size_t)point_num * 4, cudaMemcpyHostToDevice, s));
    } else if (point_num > avail) {
        return DTSDF_ERR_INVALID_ARG;
    }
    if ((rc = c->d_misc.reserve(64)) != DTSDF_OK) return rc;
    if ((rc = c->d_ext.reserve(24)) != DTSDF_OK) return rc;
    if (h_out_pc && (rc = c->d_pc.reserve((size_t)point_num * 24)) != DTSDF_OK) return rc;
    CK(cudaMemsetAsync(c->d_misc.p, 0, 16, s));                   // point offset 0
    rc = dtsdf_resample_extrema_batch(1, point_num, (const double*)c->d_points.p,
                                      (const int64_t*)((unsigned char*)c->d_misc.p + 8),
                                      h_indices ? (const int32_t*)c->d_idx.p : nullptr,
                                      h_out_pc ? (double*)c->d_pc.p : nullptr, (float*)c->d_ext.p, s);
    if (rc != DTSDF_OK) return rc;
    CK(cudaMemcpyAsync(h_out_extrema, c->d_ext.p, 24, cudaMemcpyDeviceToHost, s));
    if (h_out_pc)
        CK(cudaMemcpyAsync(h_out_pc, c->d_pc.p, (size_t)point_num * 24, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return DTSDF_OK;
}

static int process_begin_impl(dtsdf_ctx* c, double focal, int32_t n_frames, const float* h_depth,
                              int64_t packed_len, const int64_t* h_off, const int32_t* h_hdr,
                              int32_t* h_n_valid, bool want_points, int32_t piece_frames = 0);

int dtsdf_process_begin_host(dtsdf_ctx* c, double focal, int32_t n_frames, const float* h_depth,
                             int64_t packed_len, const int64_t* h_off, const int32_t* h_hdr,
                             int32_t* h_n_valid) {
    return guarded([&] { return process_begin_impl(c, focal, n_frames, h_depth, packed_len, h_off, h_hdr, h_n_valid, true); });
}

// want_points = false: only the valid-pixel counts are needed (the sampled voxeliser finds the box of the
// resampled cloud from the depth map itself); no cloud is built and none of its memory reserved
// piece_frames > 0 (only without clouds): nothing is waited for - the crops go up on s_in in pieces of that many
// frames, each followed by its count kernel, the copy of its counts into pinned memory (m_nvalid) and an event
// (ev_piece[j]); the caller (HostPipeline, one piece per chunk) synchronises piece by piece.
static int process_begin_impl(dtsdf_ctx* c, double focal, int32_t n_frames, const float* h_depth,
                              int64_t packed_len, const int64_t* h_off, const int32_t* h_hdr,
                              int32_t* h_n_valid, bool want_points, int32_t piece_frames) {
    if (!c || n_frames < 0 || packed_len < 0 || !(focal > 0.0)) return DTSDF_ERR_INVALID_ARG;
    c->b_n = 0;
    c->b_pieces = 0;
    if (n_frames == 0) return DTSDF_OK;
    if (!h_depth || !h_off || !h_hdr || !h_n_valid) return DTSDF_ERR_INVALID_ARG;
    std::vector<int64_t> poff(n_frames);
    int64_t total_px = 0, prev_end = 0;
    int32_t max_px = 0;
    bool in_order = true;                   // every crop starts at or behind the end of the one before it
    for (int32_t i = 0; i < n_frames; ++i) {
        const int32_t* h = h_hdr + 6 * (size_t)i;
        if (!header_ok(h)) return DTSDF_ERR_BAD_HEADER;
        const int64_t n = ((int64_t)h[4] - h[2]) * ((int64_t)h[5] - h[3]);
        if (h_off[i] < 0 || (h_off[i] & 3) || h_off[i] + ((n + 3) & ~(int64_t)3) > packed_len)
            return DTSDF_ERR_INVALID_ARG;
        in_order = in_order && h_off[i] >= prev_end;
        prev_end = h_off[i] + ((n + 3) & ~(int64_t)3);
        poff[i] = total_px;                 // room for every pixel of the crop
        total_px += n;
        max_px = std::max<int32_t>(max_px, (int32_t)n);
    }
    USE_DEVICE(c->device);
    int rc;
    if ((rc = c->b_depth.reserve((size_t)packed_len * 4)) != DTSDF_OK) return rc;
    if ((rc = c->b_off.reserve((size_t)n_frames * 8)) != DTSDF_OK) return rc;
    if ((rc = c->b_hdr.reserve((size_t)n_frames * 24)) != DTSDF_OK) return rc;
    if ((rc = c->b_nvalid.reserve((size_t)n_frames * 4)) != DTSDF_OK) return rc;
    if (piece_frames > 0 && !want_points && in_order) {
        const int32_t np = (n_frames + piece_frames - 1) / piece_frames;
        if ((rc = c->m_nvalid.reserve((size_t)n_frames * 4)) != DTSDF_OK) return rc;
        while ((int32_t)c->ev_piece.size() < np) {
            cudaEvent_t e = nullptr;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->ev_piece.push_back(e);
        }
        c->b_nvalid_h.assign((size_t)n_frames, 0);
        cudaStream_t s = c->s_in;
        CK(cudaMemcpyAsync(c->b_off.p, h_off, (size_t)n_frames * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(c->b_hdr.p, h_hdr, (size_t)n_frames * 24, cudaMemcpyHostToDevice, s));
        for (int32_t j = 0; j < np; ++j) {
            const int32_t f0 = j * piece_frames, nf = std::min(piece_frames, n_frames - f0);
            const int32_t* hl = h_hdr + 6 * (size_t)(f0 + nf - 1);
            const int64_t nlast = ((int64_t)hl[4] - hl[2]) * ((int64_t)hl[5] - hl[3]);
            const int64_t e0 = h_off[f0], e1 = h_off[f0 + nf - 1] + ((nlast + 3) & ~(int64_t)3);
            CK(cudaMemcpyAsync((float*)c->b_depth.p + e0, h_depth + e0, (size_t)(e1 - e0) * 4, cudaMemcpyHostToDevice, s));
            rc = dtsdf_count_valid_batch(nf, (const float*)c->b_depth.p, (const int64_t*)c->b_off.p + f0,
                                         (const int32_t*)c->b_hdr.p + 6 * (size_t)f0, (int32_t*)c->b_nvalid.p + f0, s);
            if (rc != DTSDF_OK) return rc;
            CK(cudaMemcpyAsync((int32_t*)c->m_nvalid.p + f0, (const int32_t*)c->b_nvalid.p + f0, (size_t)nf * 4,
                               cudaMemcpyDeviceToHost, s));
            CK(cudaEventRecord(c->ev_piece[(size_t)j], s));
        }
        c->b_pieces = np;
        c->b_max_px = max_px;
        return DTSDF_OK;
    }
    if (want_points) {
        if ((rc = c->b_poff.reserve((size_t)n_frames * 8)) != DTSDF_OK) return rc;
        if ((rc = c->b_points.reserve((size_t)total_px * 24)) != DTSDF_OK) return rc;
    }
    cudaStream_t s = c->s_run;
    CK(cudaMemcpyAsync(c->b_depth.p, h_depth, (size_t)packed_len * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->b_off.p, h_off, (size_t)n_frames * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->b_hdr.p, h_hdr, (size_t)n_frames * 24, cudaMemcpyHostToDevice, s));
    if (want_points) {
        CK(cudaMemcpyAsync(c->b_poff.p, poff.data(), (size_t)n_frames * 8, cudaMemcpyHostToDevice, s));
        rc = dtsdf_point_cloud_batch(n_frames, focal, (const float*)c->b_depth.p, (const int64_t*)c->b_off.p,
                                     (const int32_t*)c->b_hdr.p, (double*)c->b_points.p, (const int64_t*)c->b_poff.p,
                                     (int32_t*)c->b_nvalid.p, s);
    } else {
        rc = dtsdf_count_valid_batch(n_frames, (const float*)c->b_depth.p, (const int64_t*)c->b_off.p,
                                     (const int32_t*)c->b_hdr.p, (int32_t*)c->b_nvalid.p, s);
    }
    if (rc != DTSDF_OK) return rc;
    CK(cudaMemcpyAsync(h_n_valid, c->b_nvalid.p, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));           // also covers the pageable poff staging
    c->b_nvalid_h.assign(h_n_valid, h_n_valid + n_frames);
    c->b_n = want_points ? n_frames : 0;       // dtsdf_process_finish_host needs the clouds
    c->b_max_px = max_px;
    return DTSDF_OK;
}

int dtsdf_process_finish_host(dtsdf_ctx* c, const dtsdf_config* cfg_in, int32_t point_num,
                              const int32_t* h_idx, double* h_pc, float* h_vol, float* h_maxl, float* h_mid) {
    if (!c) return DTSDF_ERR_INVALID_ARG;
    int rc = check_cfg(cfg_in);
    if (rc != DTSDF_OK) return rc;
    if (point_num <= 0) return DTSDF_ERR_INVALID_ARG;
    const int32_t n = c->b_n;
    if (n == 0) return DTSDF_OK;
    if (!h_idx || !h_vol) return DTSDF_ERR_INVALID_ARG;
    for (int32_t i = 0; i < n; ++i) {
        const int32_t nv = c->b_nvalid_h[i];
        const int32_t* idx = h_idx + (size_t)i * point_num;
        for (int32_t k = 0; k < point_num; ++k)
            if (idx[k] < 0 || idx[k] >= nv) return DTSDF_ERR_INVALID_ARG;      // also: a frame without points
    }
    dtsdf_config cfg = *cfg_in;
    cfg.bbox_mode = DTSDF_BBOX_EXTREMA;
    const size_t R = (size_t)cfg.voxel_res, vol_floats = 3 * R * R * R;
    USE_DEVICE(c->device);
    if ((rc = c->b_idx.reserve((size_t)n * point_num * 4)) != DTSDF_OK) return rc;
    if ((rc = c->b_ext.reserve((size_t)n * 24)) != DTSDF_OK) return rc;
    if (h_pc && (rc = c->b_pc.reserve((size_t)n * point_num * 24)) != DTSDF_OK) return rc;
    if ((rc = c->b_vol.reserve((size_t)n * vol_floats * 4)) != DTSDF_OK) return rc;
    if ((rc = c->b_maxl.reserve((size_t)n * 4)) != DTSDF_OK) return rc;
    if ((rc = c->b_mid.reserve((size_t)n * 12)) != DTSDF_OK) return rc;
    cudaStream_t s = c->s_run;
    CK(cudaMemcpyAsync(c->b_idx.p, h_idx, (size_t)n * point_num * 4, cudaMemcpyHostToDevice, s));
    rc = dtsdf_resample_extrema_batch(n, point_num, (const double*)c->b_points.p, (const int64_t*)c->b_poff.p,
                                      (const int32_t*)c->b_idx.p, h_pc ? (double*)c->b_pc.p : nullptr,
                                      (float*)c->b_ext.p, s);
    if (rc != DTSDF_OK) return rc;
    rc = dtsdf_voxelize_batch(&cfg, n, (const float*)c->b_depth.p, (const int64_t*)c->b_off.p,
                              (const int32_t*)c->b_hdr.p, (const float*)c->b_ext.p, c->b_max_px,
                              (float*)c->b_vol.p, (float*)c->b_maxl.p, (float*)c->b_mid.p, s);
    if (rc != DTSDF_OK) return rc;
    if (h_pc) CK(cudaMemcpyAsync(h_pc, c->b_pc.p, (size_t)n * point_num * 24, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h_vol, c->b_vol.p, (size_t)n * vol_floats * 4, cudaMemcpyDeviceToHost, s));
    if (h_maxl) CK(cudaMemcpyAsync(h_maxl, c->b_maxl.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    if (h_mid) CK(cudaMemcpyAsync(h_mid, c->b_mid.p, (size_t)n * 12, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return DTSDF_OK;
}

int dtsdf_np_randint(uint32_t* mt_key, int32_t* mt_pos, int32_t high, int32_t count, int32_t* out) {
    if (!mt_key || !mt_pos || count < 0 || (count > 0 && !out)) return DTSDF_ERR_INVALID_ARG;
    if (*mt_pos < 0 || *mt_pos > 624) return DTSDF_ERR_INVALID_ARG;
    if (high <= 0) return DTSDF_ERR_EMPTY_CLOUD;              // np.random.randint: "low >= high"
    int pos = *mt_pos;
    dtsdf::np_legacy_randint(mt_key, &pos, (uint32_t)high, count, out);
    *mt_pos = pos;
    return DTSDF_OK;
}

// The same draws in two phases for large batches: this thread finds where every frame's draws begin, block by
// block of 1024 frames (np_legacy_randint_skip: nothing written), and up to four helper threads produce the draws
// of block b from the snapshots while block b + 1 is being scanned.  Output and final stream state are those of
// resample_indices (tests/test_host_logic.py checks both against np.random itself).
static int resample_indices_two_phase(uint32_t* key, int* pos, int32_t n, const int32_t* n_valid, int32_t point_num,
                                      int32_t* out, int helpers) {
    constexpr int32_t BLOCK = 1024;          // (helper threads are started per block: ~30 us each)
    std::vector<dtsdf::DrawSpan> spans[2];
    std::vector<std::thread> th[2];
    auto join = [&](int par) { for (auto& t : th[par]) t.join(); th[par].clear(); };
    int rc = DTSDF_OK;
    try {
        for (int32_t b0 = 0, b = 0; b0 < n && rc == DTSDF_OK; b0 += BLOCK, ++b) {
            const int par = b & 1;
            const int32_t nb = std::min(BLOCK, n - b0);
            join(par);                                    // block b - 2 used these snapshots
            spans[par].resize((size_t)nb);
            int32_t ok = 0;
            for (; ok < nb; ++ok) {
                const int32_t nv = n_valid[b0 + ok];
                if (nv <= 0) { rc = DTSDF_ERR_EMPTY_CLOUD; break; }
                dtsdf::np_legacy_randint_skip(key, pos, (uint32_t)nv, nv < point_num ? point_num - nv : point_num,
                                              &spans[par][(size_t)ok]);
            }
            const int jobs = std::max(1, std::min<int>(helpers, ok));
            for (int j = 0; j < jobs; ++j) {
                const int32_t f0 = (int32_t)((int64_t)ok * j / jobs), f1 = (int32_t)((int64_t)ok * (j + 1) / jobs);
                const dtsdf::DrawSpan* sp = spans[par].data();
                int32_t* o = out + (size_t)b0 * point_num;
                th[par].emplace_back([=] {
                    for (int32_t f = f0; f < f1; ++f) {
                        int32_t* idx = o + (size_t)f * point_num;
                        const int32_t lead = point_num - sp[f].count;
                        for (int32_t k = 0; k < lead; ++k) idx[k] = k;
                        dtsdf::draw_fill(sp[f], idx + lead);
                    }
                });
            }
        }
    } catch (...) {
        join(0); join(1);
        throw;
    }
    join(0); join(1);
    return rc;
}

int dtsdf_resample_indices(uint32_t* mt_key, int32_t* mt_pos, int32_t n_frames, const int32_t* n_valid,
                           int32_t point_num, int32_t* out_idx) {
    if (!mt_key || !mt_pos || n_frames < 0 || point_num <= 0) return DTSDF_ERR_INVALID_ARG;
    if (*mt_pos < 0 || *mt_pos > 624) return DTSDF_ERR_INVALID_ARG;
    if (n_frames == 0) return DTSDF_OK;
    if (!n_valid || !out_idx) return DTSDF_ERR_INVALID_ARG;
    return guarded([&]() -> int {
        int pos = *mt_pos;
        const int helpers = (int)std::min(4u, std::max(1u, std::thread::hardware_concurrency()) - 1u);
        // (small requests, one core, DTSDF_RNG_ISA=scalar... : the plain form; same results)
        const int rc = (n_frames >= 64 && helpers >= 1)
                           ? resample_indices_two_phase(mt_key, &pos, n_frames, n_valid, point_num, out_idx, helpers)
                           : resample_indices(mt_key, &pos, n_frames, n_valid, point_num, out_idx);
        *mt_pos = pos;
        return rc;
    });
}

int dtsdf_process_host(dtsdf_ctx* c, const dtsdf_config* cfg_in, double focal, int32_t n_frames,
                       const float* h_depth, int64_t packed_len, const int64_t* h_off, const int32_t* h_hdr,
                       int32_t point_num, uint32_t* mt_key, int32_t* mt_pos, double* h_out_pc,
                       float* h_out_vol, float* h_out_max_l, float* h_out_mid_p) {
    if (!c) return DTSDF_ERR_INVALID_ARG;
    int rc = check_cfg(cfg_in);
    if (rc != DTSDF_OK) return rc;
    if (point_num <= 0 || !mt_key || !mt_pos || *mt_pos < 0 || *mt_pos > 624) return DTSDF_ERR_INVALID_ARG;
    if (n_frames < 0) return DTSDF_ERR_INVALID_ARG;
    if (n_frames == 0) return DTSDF_OK;
    if (!h_out_vol) return DTSDF_ERR_INVALID_ARG;
    return guarded([&]() -> int {
        // back-projection + compaction of every frame; crops, tables and clouds stay resident (b_*)
        USE_DEVICE(c->device);
        dtsdf_config cfg = *cfg_in;
        cfg.bbox_mode = DTSDF_BBOX_EXTREMA;
        int pos = *mt_pos;
        HostPipeline p;
        p.c = c; p.cfg = &cfg; p.n_frames = n_frames; p.h_depth = nullptr; p.h_off = nullptr; p.h_hdr = nullptr;
        p.h_aux = nullptr; p.h_vol = h_out_vol; p.h_maxl = h_out_max_l; p.h_mid = h_out_mid_p;
        p.aux_w = 6;
        // (a 64-byte aligned working copy of the generator state: its vector steps then never split a cache line)
        alignas(64) uint32_t key_al[640];
        memcpy(key_al, mt_key, 624 * sizeof(uint32_t));
        p.process_mode = true; p.point_num = point_num; p.mt_key = key_al; p.mt_pos = &pos; p.h_pc = h_out_pc;
        p.focal64 = focal;
        p.configure();
        // with clouds: back-projection + compaction of every frame up front, crops, tables and clouds resident
        // (b_*); without: the crops go up and are counted one pipeline chunk at a time behind the draws
        std::vector<int32_t> n_valid((size_t)n_frames);
        // the full clouds are an intermediate: with crops of at most DTSDF_SAMPLE_CLOUD_MAX_PX pixels the sampled
        // points come straight from the depth map (dtsdf_sample_cloud_batch) and the call takes the same route
        // as without clouds
        bool full_clouds = h_out_pc != nullptr;
        if (full_clouds) {
            int64_t biggest = 0;
            for (int32_t i = 0; i < n_frames; ++i) {
                const int32_t* hh = h_hdr + 6 * (size_t)i;
                biggest = std::max<int64_t>(biggest, ((int64_t)hh[4] - hh[2]) * ((int64_t)hh[5] - hh[3]));
            }
            full_clouds = biggest > DTSDF_SAMPLE_CLOUD_MAX_PX;
        }
        int rc2 = process_begin_impl(c, focal, n_frames, h_depth, packed_len, h_off, h_hdr, n_valid.data(),
                                     full_clouds, p.nchunk > 1 ? p.per_chunk : 0);
        if (rc2 != DTSDF_OK) {
            cudaStreamSynchronize(c->s_in);          // pieces may be in flight from the caller's buffer
            cudaGetLastError();
            c->b_pieces = 0;
            return rc2;
        }
        rc2 = p.run();
        c->b_pieces = 0;
        memcpy(mt_key, key_al, 624 * sizeof(uint32_t));
        *mt_pos = pos;
        return rc2;
    });
}

int dtsdf_normalize_joints_host(dtsdf_ctx* c, int32_t n_frames, const float* h_joints,
                                const float* h_max_l, const float* h_mid_p, int32_t clamp01,
                                float* h_out) {
    if (!c || n_frames < 0) return DTSDF_ERR_INVALID_ARG;
    if (n_frames == 0) return DTSDF_OK;
    if (!h_joints || !h_max_l || !h_mid_p || !h_out) return DTSDF_ERR_INVALID_ARG;
    USE_DEVICE(c->device);
    cudaStream_t s = c->s_run;
    int rc;
    const size_t nj = (size_t)n_frames * 63 * 4;
    if ((rc = c->d_j_in.reserve(nj)) != DTSDF_OK) return rc;
    if ((rc = c->d_j_out.reserve(nj)) != DTSDF_OK) return rc;
    if ((rc = c->d_j_ml.reserve((size_t)n_frames * 4)) != DTSDF_OK) return rc;
    if ((rc = c->d_j_mp.reserve((size_t)n_frames * 12)) != DTSDF_OK) return rc;
    CK(cudaMemcpyAsync(c->d_j_in.p, h_joints, nj, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->d_j_ml.p, h_max_l, (size_t)n_frames * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->d_j_mp.p, h_mid_p, (size_t)n_frames * 12, cudaMemcpyHostToDevice, s));
    rc = dtsdf_normalize_joints_batch(n_frames, (const float*)c->d_j_in.p, (const float*)c->d_j_ml.p,
                                      (const float*)c->d_j_mp.p, clamp01, (float*)c->d_j_out.p, s);
    if (rc != DTSDF_OK) return rc;
    CK(cudaMemcpyAsync(h_out, c->d_j_out.p, nj, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return DTSDF_OK;
}

}  // extern "C"

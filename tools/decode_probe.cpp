// Dev probe (host only, no GPU): throughput of the packed-volume expansion of dtsdf_voxelize_host.
//   g++ -O2 -std=c++17 -pthread tools/decode_probe.cpp handposeestimation-with-3d-cnns_b200/csrc/dtsdf_decode.cpp \
//       -o tools/_exp/decode_probe && tools/_exp/decode_probe tools/_exp/vol64.f32 [threads]
// Input: raw float32 volumes [k][3][32^3] (e.g. written from the C oracle).  Packs them on the CPU in the
// format of dtsdf_internal.h (launch_pack), replicates to 4096 frames and times the decoders.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>

namespace dtsdf { void decode_frame(const unsigned char*, const float*, size_t, float*); const char* decode_isa(); }

static void decode_sse(const unsigned char* codes, const float* lits, size_t r3, float* out) {
    // the round-1 decoder: 16-byte streaming stores, the three planes interleaved
    static float val[256][4];
    static unsigned char nlit[256];
    static bool init = false;
    if (!init) {
        const float v[4] = {0.0f, 1.0f, -1.0f, 0.0f};
        for (int b = 0; b < 256; ++b) { int n = 0; for (int j = 0; j < 4; ++j) { int c = (b >> (2 * j)) & 3; val[b][j] = v[c]; n += c == 3; } nlit[b] = n; }
        init = true;
    }
    float *o0 = out, *o1 = out + r3, *o2 = out + 2 * r3;
    for (size_t k = 0; k < r3 / 4; ++k) {
        const unsigned b = codes[k];
        __m128 v0, v1, v2;
        if (nlit[b] == 0) v0 = v1 = v2 = _mm_loadu_ps(val[b]);
        else { v0 = _mm_load_ps(lits); v1 = _mm_load_ps(lits + 4); v2 = _mm_load_ps(lits + 8); lits += 12; }
        _mm_stream_ps(o0 + 4 * k, v0); _mm_stream_ps(o1 + 4 * k, v1); _mm_stream_ps(o2 + 4 * k, v2);
    }
    _mm_sfence();
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: decode_probe vol.f32 [threads]\n"); return 2; }
    const size_t r3 = 32768;
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 1; }
    fseek(f, 0, SEEK_END);
    const size_t k = ftell(f) / (3 * r3 * 4);
    fseek(f, 0, SEEK_SET);
    std::vector<float> vol(k * 3 * r3);
    if (fread(vol.data(), 4, vol.size(), f) != vol.size()) return 1;
    fclose(f);
    // pack
    std::vector<unsigned char> codes(k * r3 / 4);
    std::vector<float> lits;
    std::vector<size_t> loff(k + 1);
    for (size_t i = 0; i < k; ++i) {
        loff[i] = lits.size() / 12;
        const float* v = vol.data() + i * 3 * r3;
        for (size_t g = 0; g < r3 / 4; ++g) {
            unsigned byte = 0;
            for (int j = 0; j < 4; ++j) {
                uint32_t a, b, c;
                memcpy(&a, v + 4 * g + j, 4); memcpy(&b, v + r3 + 4 * g + j, 4); memcpy(&c, v + 2 * r3 + 4 * g + j, 4);
                const bool same = a == b && b == c;
                const unsigned code = (same && a == 0u) ? 0u : (same && a == 0x3F800000u) ? 1u : (same && a == 0xBF800000u) ? 2u : 3u;
                byte |= code << (2 * j);
            }
            codes[i * r3 / 4 + g] = (unsigned char)byte;
            if (byte & (byte >> 1) & 0x55u) {
                for (int p = 0; p < 3; ++p) for (int j = 0; j < 4; ++j) lits.push_back(v[p * r3 + 4 * g + j]);
            }
        }
    }
    loff[k] = lits.size() / 12;
    float* litbuf = (float*)aligned_alloc(64, lits.size() * 4 + 64);
    memcpy(litbuf, lits.data(), lits.size() * 4);
    printf("%zu frames, %.1f literal groups per frame, packed %.1f KB per frame, decoder ISA: %s\n", k, (double)loff[k] / k,
           (r3 / 4 + 48.0 * loff[k] / k) / 1024, dtsdf::decode_isa());
    const size_t N = 4096;
    const size_t mis = getenv("PROBE_MISALIGN") ? atoi(getenv("PROBE_MISALIGN")) : 0;   // floats
    float* out = (float*)aligned_alloc(64, N * 3 * r3 * 4 + 64) + mis;
    memset(out, 0, N * 3 * r3 * 4);
    const int hw = std::thread::hardware_concurrency();
    std::vector<int> tcounts = {1, 4, 8, 12, 16, 32};
    if (argc > 2) tcounts = {atoi(argv[2])};
    for (int which = 0; which < 2; ++which) {
        auto fn = which == 0 ? decode_sse : dtsdf::decode_frame;
        // correctness first
        bool ok = true;
        for (size_t i = 0; i < k; ++i) {
            fn(codes.data() + i * r3 / 4, litbuf + 12 * loff[i], r3, out + i * 3 * r3);
            ok = ok && memcmp(out + i * 3 * r3, vol.data() + i * 3 * r3, 3 * r3 * 4) == 0;
        }
        printf("%s: round trip %s\n", which == 0 ? "sse (round 1)" : "new", ok ? "bit-exact" : "MISMATCH");
        for (int nt : tcounts) {
            if (nt > hw) continue;
            double best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                auto t0 = std::chrono::steady_clock::now();
                std::vector<std::thread> th;
                for (int t = 0; t < nt; ++t) th.emplace_back([=, &codes, &loff] {
                    for (size_t i = N * t / nt; i < N * (t + 1) / nt; ++i) {
                        const size_t s = i % k;
                        fn(codes.data() + s * r3 / 4, litbuf + 12 * loff[s], r3, out + i * 3 * r3);
                    }
                });
                for (auto& x : th) x.join();
                best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            }
            printf("  threads %2d: %.0f volumes/s  %.1f GB/s written\n", nt, N / best, N * 3 * r3 * 4 / best / 1e9);
        }
    }
    return 0;
}

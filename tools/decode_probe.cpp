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

namespace dtsdf { void decode_frame(const unsigned char*, const float*, size_t, size_t, float*); const char* decode_isa(); }

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
    // pack: per frame x[n], y[n], z[n] of its literal voxels (csrc/dtsdf_internal.h)
    std::vector<unsigned char> codes(k * r3 / 4);
    std::vector<float> lits;
    std::vector<size_t> loff(k + 1);
    for (size_t i = 0; i < k; ++i) {
        loff[i] = lits.size() / 3;
        const float* v = vol.data() + i * 3 * r3;
        std::vector<float> lx, ly, lz;
        for (size_t g = 0; g < r3 / 4; ++g) {
            unsigned byte = 0;
            for (int j = 0; j < 4; ++j) {
                uint32_t a, b, c;
                memcpy(&a, v + 4 * g + j, 4); memcpy(&b, v + r3 + 4 * g + j, 4); memcpy(&c, v + 2 * r3 + 4 * g + j, 4);
                const bool same = a == b && b == c;
                const unsigned code = (same && a == 0u) ? 0u : (same && a == 0x3F800000u) ? 1u : (same && a == 0xBF800000u) ? 2u : 3u;
                byte |= code << (2 * j);
                if (code == 3) { lx.push_back(v[4 * g + j]); ly.push_back(v[r3 + 4 * g + j]); lz.push_back(v[2 * r3 + 4 * g + j]); }
            }
            codes[i * r3 / 4 + g] = (unsigned char)byte;
        }
        lits.insert(lits.end(), lx.begin(), lx.end()); lits.insert(lits.end(), ly.begin(), ly.end()); lits.insert(lits.end(), lz.begin(), lz.end());
    }
    loff[k] = lits.size() / 3;
    float* litbuf = (float*)aligned_alloc(64, lits.size() * 4 + 64);
    memcpy(litbuf, lits.data(), lits.size() * 4);
    printf("%zu frames, %.1f literal voxels per frame, packed %.1f KB per frame, decoder ISA: %s\n", k, (double)loff[k] / k,
           (r3 / 4 + 12.0 * loff[k] / k) / 1024, dtsdf::decode_isa());
    const size_t N = 4096;
    const size_t mis = getenv("PROBE_MISALIGN") ? atoi(getenv("PROBE_MISALIGN")) : 0;   // floats
    float* out = (float*)aligned_alloc(64, N * 3 * r3 * 4 + 64) + mis;
    memset(out, 0, N * 3 * r3 * 4);
    const int hw = std::thread::hardware_concurrency();
    std::vector<int> tcounts = {1, 4, 8, 12, 16, 32};
    if (argc > 2) tcounts = {atoi(argv[2])};
    {
        bool ok = true;
        for (size_t i = 0; i < k; ++i) {
            dtsdf::decode_frame(codes.data() + i * r3 / 4, litbuf + 3 * loff[i], loff[i + 1] - loff[i], r3, out + i * 3 * r3);
            ok = ok && memcmp(out + i * 3 * r3, vol.data() + i * 3 * r3, 3 * r3 * 4) == 0;
        }
        printf("round trip %s\n", ok ? "bit-exact" : "MISMATCH");
        for (int nt : tcounts) {
            if (nt > hw) continue;
            double best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                auto t0 = std::chrono::steady_clock::now();
                std::vector<std::thread> th;
                for (int t = 0; t < nt; ++t) th.emplace_back([=, &codes, &loff] {
                    for (size_t i = N * t / nt; i < N * (t + 1) / nt; ++i) {
                        const size_t s = i % k;
                        dtsdf::decode_frame(codes.data() + s * r3 / 4, litbuf + 3 * loff[s], loff[s + 1] - loff[s], r3, out + i * 3 * r3);
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

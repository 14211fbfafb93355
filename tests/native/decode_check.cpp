// Host-side check of the volume codec's expansion (csrc/dtsdf_decode.cpp) against a plain
// restatement of the format (csrc/dtsdf_internal.h: launch_pack), for every output alignment
// a caller can hand in and both instruction sets (DTSDF_DECODE_ISA=sse2 forces the baseline).
// Build: g++ -O2 -std=c++17 tests/native/decode_check.cpp <pkg>/csrc/dtsdf_decode.cpp
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace dtsdf {
void decode_frame(const unsigned char*, const float*, size_t, size_t, float*);
const char* decode_isa();
}

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 32);
}

int main() {
    int bad = 0, cases = 0;
    for (size_t R : {32, 48, 64}) {
        const size_t r3 = R * R * R, nb = r3 / 4;
        for (int density = 0; density < 4; ++density) {          // literal density 0, sparse, medium, all
            std::vector<unsigned char> codes(nb);
            std::vector<float> want(3 * r3), lx, ly, lz;
            const float val[4] = {0.0f, 1.0f, -1.0f, 0.0f};
            for (size_t k = 0; k < nb; ++k) {
                unsigned b = 0;
                for (int j = 0; j < 4; ++j) {
                    unsigned c = rnd() % 3;
                    const unsigned roll = rnd() % 100;
                    if (density == 3 || (density == 2 && roll < 30) || (density == 1 && roll < 3)) c = 3;
                    b |= c << (2 * j);
                }
                if (rnd() % 4 == 0) b = (rnd() % 3) * 0x55u;          // runs of one value
                codes[k] = (unsigned char)b;
                for (int j = 0; j < 4; ++j) {
                    const unsigned c = (b >> (2 * j)) & 3;
                    float v[3] = {val[c], val[c], val[c]};
                    if (c == 3) {
                        for (int p = 0; p < 3; ++p) {
                            uint32_t bits = rnd();                    // any bit pattern, NaNs and -0 included
                            memcpy(&v[p], &bits, 4);
                        }
                        lx.push_back(v[0]); ly.push_back(v[1]); lz.push_back(v[2]);
                    }
                    for (int p = 0; p < 3; ++p) want[p * r3 + 4 * k + j] = v[p];
                }
            }
            const size_t nl = lx.size();
            // the literal block sits at an odd float offset: only 4-byte alignment is promised
            float* litbuf = (float*)aligned_alloc(64, (3 * nl + 32) * 4);
            float* lits = litbuf + 1;
            memcpy(lits, lx.data(), nl * 4); memcpy(lits + nl, ly.data(), nl * 4); memcpy(lits + 2 * nl, lz.data(), nl * 4);
            for (size_t mis = 0; mis < 20; ++mis) {                   // output offset in floats: 0..19
                float* raw = (float*)aligned_alloc(64, (3 * r3 + 64) * 4);
                memset(raw, 0xCD, (3 * r3 + 64) * 4);
                float* out = raw + 16 + mis;
                dtsdf::decode_frame(codes.data(), lits, nl, r3, out);
                const bool body = memcmp(out, want.data(), 3 * r3 * 4) == 0;
                bool guard = true;                                    // nothing written outside the frame
                for (size_t i = 0; i < 16 + mis; ++i) { uint32_t u; memcpy(&u, raw + i, 4); guard = guard && u == 0xCDCDCDCDu; }
                for (size_t i = 16 + mis + 3 * r3; i < 3 * r3 + 64; ++i) { uint32_t u; memcpy(&u, raw + i, 4); guard = guard && u == 0xCDCDCDCDu; }
                ++cases;
                if (!body || !guard) { ++bad; printf("FAIL R=%zu density=%d offset=%zu body=%d guard=%d\n", R, density, mis, body, guard); }
                free(raw);
            }
            free(litbuf);
        }
    }
    printf("isa=%s cases=%d mismatches=%d\n", dtsdf::decode_isa(), cases, bad);
    return bad != 0;
}

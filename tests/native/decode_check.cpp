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
void decode_frame(const unsigned char*, const float*, size_t, float*);
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
            float* lits = (float*)aligned_alloc(64, nb * 48 + 64);
            std::vector<float> want(3 * r3);
            size_t nl = 0;
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
                if (b & (b >> 1) & 0x55u) {
                    for (int p = 0; p < 3; ++p)
                        for (int j = 0; j < 4; ++j) {
                            uint32_t bits = rnd();                    // any bit pattern, NaNs and -0 included
                            float f; memcpy(&f, &bits, 4);
                            lits[12 * nl + 4 * p + j] = f;
                            want[p * r3 + 4 * k + j] = f;
                        }
                    ++nl;
                } else {
                    for (int p = 0; p < 3; ++p)
                        for (int j = 0; j < 4; ++j) want[p * r3 + 4 * k + j] = val[(b >> (2 * j)) & 3];
                }
            }
            for (size_t mis = 0; mis < 20; ++mis) {                   // output offset in floats: 0..19
                float* raw = (float*)aligned_alloc(64, (3 * r3 + 64) * 4);
                memset(raw, 0xCD, (3 * r3 + 64) * 4);
                float* out = raw + 16 + mis;
                dtsdf::decode_frame(codes.data(), lits, r3, out);
                const bool body = memcmp(out, want.data(), 3 * r3 * 4) == 0;
                bool guard = true;                                    // nothing written outside the frame
                for (size_t i = 0; i < 16 + mis; ++i) { uint32_t u; memcpy(&u, raw + i, 4); guard = guard && u == 0xCDCDCDCDu; }
                for (size_t i = 16 + mis + 3 * r3; i < 3 * r3 + 64; ++i) { uint32_t u; memcpy(&u, raw + i, 4); guard = guard && u == 0xCDCDCDCDu; }
                ++cases;
                if (!body || !guard) { ++bad; printf("FAIL R=%zu density=%d offset=%zu body=%d guard=%d\n", R, density, mis, body, guard); }
                free(raw);
            }
            free(lits);
        }
    }
    printf("isa=%s cases=%d mismatches=%d\n", dtsdf::decode_isa(), cases, bad);
    return bad != 0;
}

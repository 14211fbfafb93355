// NumPy's legacy global random stream, restated: MT19937 (Matsumoto & Nishimura) exactly as
// numpy/random/src/mt19937/mt19937.c steps it, and the bounded integers RandomState.randint draws from it
// (numpy/random/src/distributions/distributions.c: random_bounded_uint64_fill -> for a range below 2^32,
// buffered_bounded_masked_uint32: one 32-bit output per attempt, masked with the smallest 2^k - 1 >= range,
// rejected while above the range).  NumPy is a dependency of the reference, not part of it
// (pre/process.py:75-80 calls np.random.randint on the global stream); the pinned behaviour is that of the
// NumPy in this image (2.3) and of every NumPy since 1.17's legacy-seeding guarantee, and
// tests/test_host_logic.py checks it against np.random itself over 10^5 frames.
//
// Why it is here: DataProcess.set_length draws up to point_num indices per frame from the caller's global
// np.random stream, and the index vectors decide the bounding box, hence every voxel.  Drawing them in a
// Python loop (one np.random.randint call per frame) capped the batched process() path at ~11 k frames/s.
// Plain C++ (no CUDA).
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "dtsdf_internal.h"

namespace dtsdf {

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;

inline uint32_t twist(uint32_t u, uint32_t v) {
    const uint32_t y = (u & UPPER) | (v & LOWER);
    return (y >> 1) ^ ((v & 1u) ? MATRIX_A : 0u);          // mag01[y & 1]: the low bit of y is the low bit of v
}

// mt19937_gen: the next 624 words of state.  The three loops only read words that this pass has not
// written yet or wrote at least 227 iterations earlier, so the compiler may vectorise them (clones for the
// wider vector units are picked at load time).
__attribute__((target_clones("avx512f", "avx2", "default")))
void mt_generate(uint32_t* key) {
    int i = 0;
    for (; i < MT_N - MT_M; ++i) key[i] = key[i + MT_M] ^ twist(key[i], key[i + 1]);
    for (; i < MT_N - 1; ++i) key[i] = key[i + (MT_M - MT_N)] ^ twist(key[i], key[i + 1]);
    key[MT_N - 1] = key[MT_M - 1] ^ twist(key[MT_N - 1], key[0]);
}

inline uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// Tempers and filters raw words key[0..n): every masked value <= rng is a draw, in order, until `want` draws
// have been written.  Returns the number of raw words consumed (up to and including the last accepted one
// when `want` was reached); *got = draws written.
int filter_scalar(const uint32_t* key, int n, uint32_t mask, uint32_t rng, int32_t* out, int want, int* got) {
    int done = 0, k = 0;
    for (; k < n && done < want; ++k) {
        const uint32_t v = temper(key[k]) & mask;
        out[done] = (int32_t)v;
        done += (v <= rng);
    }
    *got = done;
    return k;
}

// 16 words per step: vector tempering, one compare, one compress-store
__attribute__((target("avx512f,bmi2")))
int filter_avx512(const uint32_t* key, int n, uint32_t mask, uint32_t rng, int32_t* out, int want, int* got) {
    const __m512i vmask = _mm512_set1_epi32((int)mask), vrng = _mm512_set1_epi32((int)rng);
    const __m512i c1 = _mm512_set1_epi32((int)0x9d2c5680u), c2 = _mm512_set1_epi32((int)0xefc60000u);
    int done = 0, k = 0;
    for (; k + 16 <= n && done < want; k += 16) {
        __m512i y = _mm512_loadu_si512(key + k);
        y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 11));
        y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 7), c1));
        y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 15), c2));
        y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 18));
        y = _mm512_and_si512(y, vmask);
        unsigned m = _mm512_cmple_epu32_mask(y, vrng);
        const int c = __builtin_popcount(m);
        if (done + c >= want) {
            // the step that completes the request: keep the first (want - done) accepted words only
            const int need = want - done;                       // >= 1
            const unsigned last = _pdep_u32(1u << (need - 1), m);   // the need-th set bit of m
            const int used = __builtin_ctz(last) + 1;
            m &= (last << 1) - 1u;
            _mm512_mask_compressstoreu_epi32(out + done, (__mmask16)m, y);
            *got = want;
            return k + used;
        }
        _mm512_mask_compressstoreu_epi32(out + done, (__mmask16)m, y);
        done += c;
    }
    int tail_got = 0;
    const int tail = filter_scalar(key + k, n - k, mask, rng, out + done, want - done, &tail_got);
    *got = done + tail_got;
    return k + tail;
}

bool has_avx512() {
    // DTSDF_RNG_ISA=scalar forces the portable path (tests)
    static const bool has = [] {
        const char* e = getenv("DTSDF_RNG_ISA");
        if (e != nullptr && strcmp(e, "scalar") == 0) return false;
        return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("bmi2");
    }();
    return has;
}

}  // namespace

// count draws of RandomState.randint(0, high) (int64 default dtype, high - 1 < 2^32 - 1) into out[0..count)
void np_legacy_randint(uint32_t* key, int* pos, uint32_t high, int count, int32_t* out) {
    const uint32_t rng = high - 1u;
    if (rng == 0u) {                         // random_bounded_uint64_fill: rng == 0 consumes nothing
        for (int i = 0; i < count; ++i) out[i] = 0;
        return;
    }
    uint32_t mask = rng;                     // gen_mask: smallest bit mask >= rng
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    const bool wide = has_avx512();
    int p = *pos, done = 0;
    while (done < count) {
        if (p >= MT_N) { mt_generate(key); p = 0; }
        int got = 0;
        const int used = wide ? filter_avx512(key + p, MT_N - p, mask, rng, out + done, count - done, &got)
                              : filter_scalar(key + p, MT_N - p, mask, rng, out + done, count - done, &got);
        p += used;
        done += got;
    }
    *pos = p;
}

}  // namespace dtsdf

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
__attribute__((target_clones("avx2", "default")))
void mt_generate_auto(uint32_t* key) {
    int i = 0;
    for (; i < MT_N - MT_M; ++i) key[i] = key[i + MT_M] ^ twist(key[i], key[i + 1]);
    for (; i < MT_N - 1; ++i) key[i] = key[i + (MT_M - MT_N)] ^ twist(key[i], key[i + 1]);
    key[MT_N - 1] = key[MT_M - 1] ^ twist(key[MT_N - 1], key[0]);
}

// The same pass written out for AVX-512 (16 words per step: one bit select, one test-and-move, one
// three-way xor): 0.13-0.19 ns per word against 0.3-0.6 for what the compiler makes of the loops above.
__attribute__((target("avx512f")))
inline __m512i twist16(__m512i u, __m512i v, __m512i far) {
    const __m512i up = _mm512_set1_epi32((int)UPPER), one = _mm512_set1_epi32(1), a = _mm512_set1_epi32((int)MATRIX_A);
    const __m512i y = _mm512_ternarylogic_epi32(up, u, v, 0xCA);                        // up ? u : v
    const __m512i mag = _mm512_maskz_mov_epi32(_mm512_test_epi32_mask(v, one), a);
    return _mm512_ternarylogic_epi32(_mm512_srli_epi32(y, 1), mag, far, 0x96);          // xor of the three
}
__attribute__((target("avx512f")))
void mt_generate_avx512(uint32_t* key) {
    // 39 vectors of 16 words.  The "next word" and "word 397 further on" operands are funnel shifts (valignd)
    // of whole vectors instead of loads at odd offsets: two loads per step, none of them splitting a cache line
    // when the state is 64-byte aligned (the working copies are).  397 = 24 * 16 + 13: the far operand of
    // vector j comes from vectors j + 24 and j + 25 (mod 39) - old words up to j = 13, words this pass has
    // already written from j = 14 on, and vector j + 1 is still old when step j reads it.
    __m512i cur = _mm512_loadu_si512(key);
    __m512i fcur = _mm512_loadu_si512(key + 16 * 24);
    int fj = 25;
    for (int j = 0; j < 38; ++j) {
        const __m512i nxt = _mm512_loadu_si512(key + 16 * j + 16);
        const __m512i fnxt = _mm512_loadu_si512(key + 16 * fj);
        fj = fj == 38 ? 0 : fj + 1;
        _mm512_storeu_si512(key + 16 * j, twist16(cur, _mm512_alignr_epi32(nxt, cur, 1), _mm512_alignr_epi32(fnxt, fcur, 13)));
        cur = nxt;
        fcur = fnxt;
    }
    const __m512i nxt = _mm512_set1_epi32((int)key[0]);           // word 624 of the old pass is the new word 0
    const __m512i fnxt = _mm512_loadu_si512(key + 16 * fj);
    _mm512_storeu_si512(key + 16 * 38, twist16(cur, _mm512_alignr_epi32(nxt, cur, 1), _mm512_alignr_epi32(fnxt, fcur, 13)));
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
    for (; k < n && done < want; k += 16) {
        // (the last step of a block is a masked load: a block seldom starts on a multiple of 16 words)
        const __mmask16 live = n - k >= 16 ? (__mmask16)0xffff : (__mmask16)((1u << (n - k)) - 1u);
        __m512i y = _mm512_maskz_loadu_epi32(live, key + k);
        y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 11));
        y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 7), c1));
        y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 15), c2));
        y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 18));
        y = _mm512_and_si512(y, vmask);
        unsigned m = _mm512_mask_cmple_epu32_mask(live, y, vrng);
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
    *got = done;                    // all n words looked at (or want == 0)
    return k < n ? k : n;
}

// The counting half of the filters: how many of key[0..n) are needed for `want` acceptances (up to and including
// the last accepted word), *got = acceptances found (< want: all n words were looked at).
size_t count_scalar(const uint32_t* key, size_t n, uint32_t mask, uint32_t rng, int want, int* got) {
    int done = 0;
    size_t k = 0;
    for (; k < n && done < want; ++k) done += ((temper(key[k]) & mask) <= rng);
    *got = done;
    return k;
}

__attribute__((target("avx512f,bmi2")))
size_t count_avx512(const uint32_t* key, size_t n, uint32_t mask, uint32_t rng, int want, int* got) {
    const __m512i vmask = _mm512_set1_epi32((int)mask), vrng = _mm512_set1_epi32((int)rng);
    const __m512i c1 = _mm512_set1_epi32((int)0x9d2c5680u), c2 = _mm512_set1_epi32((int)0xefc60000u);
    int done = 0;
    size_t k = 0;
    for (; k < n && done < want; k += 16) {
        const __mmask16 live = n - k >= 16 ? (__mmask16)0xffff : (__mmask16)((1u << (n - k)) - 1u);
        __m512i y = _mm512_maskz_loadu_epi32(live, key + k);
        y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 11));
        y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 7), c1));
        y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 15), c2));
        y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 18));
        const unsigned m = _mm512_mask_cmple_epu32_mask(live, _mm512_and_si512(y, vmask), vrng);
        const int c = __builtin_popcount(m);
        if (done + c >= want) {
            const unsigned last = _pdep_u32(1u << (want - done - 1), m);       // the last acceptance needed
            *got = want;
            return k + (size_t)__builtin_ctz(last) + 1;
        }
        done += c;
    }
    *got = done;
    return k < n ? k : n;
}

// Acceptances among ALL of key[0..n): what the scan uses for every block that cannot complete the request
// (fewer words left in it than draws still wanted) - no early exit, no mask leaves the vector unit.
__attribute__((target("avx512f")))
int count_all_avx512(const uint32_t* key, int n, uint32_t mask, uint32_t rng) {
    const __m512i vmask = _mm512_set1_epi32((int)mask), vrng = _mm512_set1_epi32((int)rng), one = _mm512_set1_epi32(1);
    const __m512i c1 = _mm512_set1_epi32((int)0x9d2c5680u), c2 = _mm512_set1_epi32((int)0xefc60000u);
    __m512i acc = _mm512_setzero_si512();
    for (int k = 0; k < n; k += 16) {
        const __mmask16 live = n - k >= 16 ? (__mmask16)0xffff : (__mmask16)((1u << (n - k)) - 1u);
        __m512i y = _mm512_maskz_loadu_epi32(live, key + k);
        y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 11));
        y = _mm512_ternarylogic_epi32(y, _mm512_slli_epi32(y, 7), c1, 0x78);        // y ^ (s & c)
        y = _mm512_ternarylogic_epi32(y, _mm512_slli_epi32(y, 15), c2, 0x78);
        y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 18));
        const __mmask16 m = _mm512_mask_cmple_epu32_mask(live, _mm512_and_si512(y, vmask), vrng);
        acc = _mm512_mask_add_epi32(acc, m, acc, one);
    }
    return _mm512_reduce_add_epi32(acc);
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

inline void mt_generate(uint32_t* key) {
    if (has_avx512()) mt_generate_avx512(key);
    else mt_generate_auto(key);
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

void np_legacy_randint_skip(uint32_t* key, int* pos, uint32_t high, int count, DrawSpan* span) {
    memcpy(span->key, key, sizeof(span->key));
    span->pos = *pos;
    span->high = high;
    span->count = count;
    const uint32_t rng = high - 1u;
    if (rng == 0u || count <= 0) return;      // random_bounded_uint64_fill: rng == 0 consumes nothing
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    const bool wide = has_avx512();
    int p = *pos, done = 0;
    while (done < count) {
        if (p >= MT_N) { mt_generate(key); p = 0; }
        if (wide && count - done > MT_N - p) {             // the block cannot complete the request: count all of it
            done += count_all_avx512(key + p, MT_N - p, mask, rng);
            p = MT_N;
            continue;
        }
        int got = 0;
        const size_t used = wide ? count_avx512(key + p, (size_t)(MT_N - p), mask, rng, count - done, &got)
                                 : count_scalar(key + p, (size_t)(MT_N - p), mask, rng, count - done, &got);
        p += (int)used;
        done += got;
    }
    *pos = p;
}

void draw_fill(const DrawSpan& span, int32_t* out) {
    uint32_t key[MT_N];
    memcpy(key, span.key, sizeof(key));
    int pos = span.pos;
    np_legacy_randint(key, &pos, span.high, span.count, out);
}

}  // namespace dtsdf

// Host side of the volume codec (dtsdf_internal.h, launch_pack): expansion of one packed frame into its
// three channel planes in the caller's buffer.  Plain C++ (no CUDA), compiled by the host compiler.
//
// The expansion is the end-to-end bottleneck of dtsdf_voxelize_host (3 x R^3 floats written per frame
// against ~70 KB read), so it is written for the memory system: every 64-byte line of the output is
// produced whole and leaves with non-temporal stores - one 512-bit store where the CPU has AVX-512F
// (chosen at run time), four back-to-back 128-bit stores to the same line otherwise - so the
// write-combining buffers never hold a partial line when the next plane's line is opened.
//
// Codes: voxel 4k+j of the frame sits in bits 2j..2j+1 of byte k: 0 = (+0,+0,+0), 1 = (+1,+1,+1),
// 2 = (-1,-1,-1), 3 = literal.  The frame's literal block holds, for its nlit literal voxels in voxel
// order, x[nlit], then y[nlit], then z[nlit] (floats, 4-byte aligned).
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>

namespace dtsdf {

namespace {

struct DecodeTables {
    alignas(16) float val[256][4];   // the four voxel values of a code byte (code 3 -> 0, overwritten)
    DecodeTables() {
        const float v[4] = {0.0f, 1.0f, -1.0f, 0.0f};
        for (int b = 0; b < 256; ++b)
            for (int j = 0; j < 4; ++j) val[b][j] = v[(b >> (2 * j)) & 3];
    }
};
const DecodeTables g_dec;

inline unsigned literal_mask(unsigned byte) { return byte & (byte >> 1) & 0x55u; }   // bit 2j: voxel j is a literal

// the three running literal streams of a frame
struct Lits {
    const float* x;
    const float* y;
    const float* z;
};

// one code byte (4 voxels) -> one 128-bit vector per plane
inline void expand_byte(unsigned b, Lits& l, __m128& x, __m128& y, __m128& z) {
    x = y = z = _mm_load_ps(g_dec.val[b]);
    const unsigned m = literal_mask(b);
    if (m) {
        alignas(16) float xv[4], yv[4], zv[4];
        _mm_store_ps(xv, x); _mm_store_ps(yv, x); _mm_store_ps(zv, x);
        for (int j = 0; j < 4; ++j)
            if ((m >> (2 * j)) & 1u) { xv[j] = *l.x++; yv[j] = *l.y++; zv[j] = *l.z++; }
        x = _mm_load_ps(xv); y = _mm_load_ps(yv); z = _mm_load_ps(zv);
    }
}

// ---- baseline x86-64 (SSE2), 64-byte aligned output: 16 voxels (4 code bytes) -> one line per
// plane, plane after plane
template <bool NT>
void decode_sse2(const unsigned char* codes, Lits l, size_t r3, float* out) {
    float* o0 = out;
    float* o1 = out + r3;
    float* o2 = out + 2 * r3;
    for (size_t k = 0; k < r3 / 4; k += 4) {
        __m128 x[4], y[4], z[4];
        for (int j = 0; j < 4; ++j) expand_byte(codes[k + j], l, x[j], y[j], z[j]);
        if (NT) {
            for (int j = 0; j < 4; ++j) _mm_stream_ps(o0 + 4 * (k + j), x[j]);
            for (int j = 0; j < 4; ++j) _mm_stream_ps(o1 + 4 * (k + j), y[j]);
            for (int j = 0; j < 4; ++j) _mm_stream_ps(o2 + 4 * (k + j), z[j]);
        } else {
            for (int j = 0; j < 4; ++j) _mm_storeu_ps(o0 + 4 * (k + j), x[j]);
            for (int j = 0; j < 4; ++j) _mm_storeu_ps(o1 + 4 * (k + j), y[j]);
            for (int j = 0; j < 4; ++j) _mm_storeu_ps(o2 + 4 * (k + j), z[j]);
        }
    }
    if (NT) _mm_sfence();
}

// one code byte (4 voxels) -> 16 bytes per plane; for outputs whose lines cannot be written whole
template <bool NT>
inline void decode_byte(unsigned b, Lits& l, float* o0, float* o1, float* o2) {
    __m128 x, y, z;
    expand_byte(b, l, x, y, z);
    if (NT) { _mm_stream_ps(o0, x); _mm_stream_ps(o1, y); _mm_stream_ps(o2, z); }
    else { _mm_storeu_ps(o0, x); _mm_storeu_ps(o1, y); _mm_storeu_ps(o2, z); }
}

template <bool NT>
void decode_bytes(const unsigned char* codes, Lits l, size_t r3, float* out) {
    for (size_t k = 0; k < r3 / 4; ++k)
        decode_byte<NT>(codes[k], l, out + 4 * k, out + r3 + 4 * k, out + 2 * r3 + 4 * k);
    if (NT) _mm_sfence();
}

// ---- AVX-512F: 16 two-bit codes -> 16 floats with one variable shift + one permute; the literal
// voxels of the line are filled in by a masked expand-load per plane; one 512-bit store per plane and line
template <bool NT>
__attribute__((target("avx512f")))
void decode_avx512(const unsigned char* codes, Lits l, size_t r3, float* out) {
    float* o0 = out;
    float* o1 = out + r3;
    float* o2 = out + 2 * r3;
    const __m512 table = _mm512_setr_ps(0.0f, 1.0f, -1.0f, 0.0f, 0.0f, 1.0f, -1.0f, 0.0f,
                                        0.0f, 1.0f, -1.0f, 0.0f, 0.0f, 1.0f, -1.0f, 0.0f);
    const __m512i shifts = _mm512_setr_epi32(0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30);
    const __m512i three = _mm512_set1_epi32(3);
    const size_t nb = r3 / 4;                    // code bytes: 4 voxels, 16 output bytes per plane each
    // NT: `out` is 16-byte aligned; peel code bytes until the planes reach a 64-byte boundary (all
    // three at once: a plane is a multiple of 64 bytes), the 32-bit code words are then read unaligned
    size_t k0 = NT ? (((64 - (reinterpret_cast<uintptr_t>(out) & 63)) & 63) / 16) : 0;
    for (size_t b = 0; b < k0; ++b) decode_byte<NT>(codes[b], l, o0 + 4 * b, o1 + 4 * b, o2 + 4 * b);
    const size_t nw = (nb - k0) / 4;             // whole 32-bit code words: 16 voxels = one line per plane
    codes += k0; o0 += 4 * k0; o1 += 4 * k0; o2 += 4 * k0;
    for (size_t k = 0; k < nw; ++k) {
        uint32_t w;
        memcpy(&w, codes + 4 * k, 4);
        const __m512i idx = _mm512_srlv_epi32(_mm512_set1_epi32((int)w), shifts);   // permutexvar uses the low bits
        __m512 x = _mm512_permutexvar_ps(idx, table);
        __m512 y = x, z = x;
        const uint32_t lit = w & (w >> 1) & 0x55555555u;
        if (lit) {
            const __mmask16 m = _mm512_cmpeq_epi32_mask(_mm512_and_si512(idx, three), three);
            x = _mm512_mask_expandloadu_ps(x, m, l.x);
            y = _mm512_mask_expandloadu_ps(y, m, l.y);
            z = _mm512_mask_expandloadu_ps(z, m, l.z);
            const int n = __builtin_popcount(lit);
            l.x += n; l.y += n; l.z += n;
        }
        if (NT) {
            _mm512_stream_ps(o0 + 16 * k, x); _mm512_stream_ps(o1 + 16 * k, y); _mm512_stream_ps(o2 + 16 * k, z);
        } else {
            _mm512_storeu_ps(o0 + 16 * k, x); _mm512_storeu_ps(o1 + 16 * k, y); _mm512_storeu_ps(o2 + 16 * k, z);
        }
    }
    for (size_t b = 4 * nw; b < nb - k0; ++b)     // the bytes after the last whole word (k0 != 0 only)
        decode_byte<NT>(codes[b], l, o0 + 4 * b, o1 + 4 * b, o2 + 4 * b);
    if (NT) _mm_sfence();
}

bool cpu_has_avx512f() {
    // DTSDF_DECODE_ISA=sse2 forces the baseline path (tests/native/decode_check.cpp)
    static const bool has = [] {
        const char* e = std::getenv("DTSDF_DECODE_ISA");
        if (e != nullptr && std::strcmp(e, "sse2") == 0) return false;
        return __builtin_cpu_supports("avx512f") != 0;
    }();
    return has;
}

}  // namespace

const char* decode_isa() { return cpu_has_avx512f() ? "avx512f" : "sse2"; }

// r3 = R^3 is a multiple of 16 (R is a multiple of 4... 32, 48, 64), so the planes of a frame share
// the alignment of `out`.
void decode_frame(const unsigned char* codes, const float* lits, size_t nlit, size_t r3, float* out) {
    const Lits l = {lits, lits + nlit, lits + 2 * nlit};
    const uintptr_t a = reinterpret_cast<uintptr_t>(out);
    if (cpu_has_avx512f()) {
        if ((a & 15) == 0) decode_avx512<true>(codes, l, r3, out);
        else decode_avx512<false>(codes, l, r3, out);
    } else if ((a & 63) == 0) {
        decode_sse2<true>(codes, l, r3, out);
    } else if ((a & 15) == 0) {
        decode_bytes<true>(codes, l, r3, out);
    } else {
        decode_bytes<false>(codes, l, r3, out);
    }
}

}  // namespace dtsdf

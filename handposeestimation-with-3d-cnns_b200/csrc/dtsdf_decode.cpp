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
// 2 = (-1,-1,-1), 3 = literal.  A byte that holds a 3 is a literal group: its 3 x 4 floats
// (x[4], y[4], z[4]) are the next 48 bytes of the frame's literal stream (16-byte aligned).
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

inline bool has_literal(unsigned byte) { return (byte & (byte >> 1) & 0x55u) != 0u; }

// ---- baseline x86-64 (SSE2), 64-byte aligned output: 16 voxels (4 code bytes) -> one line per
// plane, plane after plane
template <bool NT>
void decode_sse2(const unsigned char* codes, const float* lits, size_t r3, float* out) {
    float* o0 = out;
    float* o1 = out + r3;
    float* o2 = out + 2 * r3;
    for (size_t k = 0; k < r3 / 4; k += 4) {
        __m128 x[4], y[4], z[4];
        for (int j = 0; j < 4; ++j) {
            const unsigned b = codes[k + j];
            if (!has_literal(b)) {
                x[j] = y[j] = z[j] = _mm_load_ps(g_dec.val[b]);
            } else {
                x[j] = _mm_load_ps(lits); y[j] = _mm_load_ps(lits + 4); z[j] = _mm_load_ps(lits + 8);
                lits += 12;
            }
        }
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
inline const float* decode_byte(unsigned b, const float* lits, float* o0, float* o1, float* o2) {
    __m128 x, y, z;
    if (!has_literal(b)) {
        x = y = z = _mm_load_ps(g_dec.val[b]);
    } else {
        x = _mm_load_ps(lits); y = _mm_load_ps(lits + 4); z = _mm_load_ps(lits + 8);
        lits += 12;
    }
    if (NT) { _mm_stream_ps(o0, x); _mm_stream_ps(o1, y); _mm_stream_ps(o2, z); }
    else { _mm_storeu_ps(o0, x); _mm_storeu_ps(o1, y); _mm_storeu_ps(o2, z); }
    return lits;
}

template <bool NT>
void decode_bytes(const unsigned char* codes, const float* lits, size_t r3, float* out) {
    for (size_t k = 0; k < r3 / 4; ++k)
        lits = decode_byte<NT>(codes[k], lits, out + 4 * k, out + r3 + 4 * k, out + 2 * r3 + 4 * k);
    if (NT) _mm_sfence();
}

// ---- AVX-512F: 16 two-bit codes -> 16 floats with one variable shift + one permute; literal groups
// are inserted as 128-bit lanes; one 512-bit store per plane and line
template <bool NT>
__attribute__((target("avx512f")))
void decode_avx512(const unsigned char* codes, const float* lits, size_t r3, float* out) {
    float* o0 = out;
    float* o1 = out + r3;
    float* o2 = out + 2 * r3;
    const __m512 table = _mm512_setr_ps(0.0f, 1.0f, -1.0f, 0.0f, 0.0f, 1.0f, -1.0f, 0.0f,
                                        0.0f, 1.0f, -1.0f, 0.0f, 0.0f, 1.0f, -1.0f, 0.0f);
    const __m512i shifts = _mm512_setr_epi32(0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30);
    const size_t nb = r3 / 4;                    // code bytes: 4 voxels, 16 output bytes per plane each
    // NT: `out` is 16-byte aligned; peel code bytes until the planes reach a 64-byte boundary (all
    // three at once: a plane is a multiple of 64 bytes), the 32-bit code words are then read unaligned
    size_t k0 = NT ? (((64 - (reinterpret_cast<uintptr_t>(out) & 63)) & 63) / 16) : 0;
    for (size_t b = 0; b < k0; ++b)
        lits = decode_byte<NT>(codes[b], lits, o0 + 4 * b, o1 + 4 * b, o2 + 4 * b);
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
            if (lit & 0x000000FFu) {
                x = _mm512_insertf32x4(x, _mm_load_ps(lits), 0); y = _mm512_insertf32x4(y, _mm_load_ps(lits + 4), 0);
                z = _mm512_insertf32x4(z, _mm_load_ps(lits + 8), 0); lits += 12;
            }
            if (lit & 0x0000FF00u) {
                x = _mm512_insertf32x4(x, _mm_load_ps(lits), 1); y = _mm512_insertf32x4(y, _mm_load_ps(lits + 4), 1);
                z = _mm512_insertf32x4(z, _mm_load_ps(lits + 8), 1); lits += 12;
            }
            if (lit & 0x00FF0000u) {
                x = _mm512_insertf32x4(x, _mm_load_ps(lits), 2); y = _mm512_insertf32x4(y, _mm_load_ps(lits + 4), 2);
                z = _mm512_insertf32x4(z, _mm_load_ps(lits + 8), 2); lits += 12;
            }
            if (lit & 0xFF000000u) {
                x = _mm512_insertf32x4(x, _mm_load_ps(lits), 3); y = _mm512_insertf32x4(y, _mm_load_ps(lits + 4), 3);
                z = _mm512_insertf32x4(z, _mm_load_ps(lits + 8), 3); lits += 12;
            }
        }
        if (NT) {
            _mm512_stream_ps(o0 + 16 * k, x); _mm512_stream_ps(o1 + 16 * k, y); _mm512_stream_ps(o2 + 16 * k, z);
        } else {
            _mm512_storeu_ps(o0 + 16 * k, x); _mm512_storeu_ps(o1 + 16 * k, y); _mm512_storeu_ps(o2 + 16 * k, z);
        }
    }
    for (size_t b = 4 * nw; b < nb - k0; ++b)     // the bytes after the last whole word (k0 != 0 only)
        lits = decode_byte<NT>(codes[b], lits, o0 + 4 * b, o1 + 4 * b, o2 + 4 * b);
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
void decode_frame(const unsigned char* codes, const float* lits, size_t r3, float* out) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(out);
    if (cpu_has_avx512f()) {
        if ((a & 15) == 0) decode_avx512<true>(codes, lits, r3, out);
        else decode_avx512<false>(codes, lits, r3, out);
    } else if ((a & 63) == 0) {
        decode_sse2<true>(codes, lits, r3, out);
    } else if ((a & 15) == 0) {
        decode_bytes<true>(codes, lits, r3, out);
    } else {
        decode_bytes<false>(codes, lits, r3, out);
    }
}

}  // namespace dtsdf

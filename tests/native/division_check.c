/* Checks the FMA-based divisions used by the CUDA kernel (csrc/dtsdf_kernels.cu: div_fixed)
 * against IEEE division on the CPU (fmaf is a single-rounding FMA by the C standard).
 *
 *   mode "focal":  a / 241.42f with ONE residual correction, EXHAUSTIVELY over every float a
 *                  with biased exponent in [lo, hi] and both signs.  The kernel uses this form
 *                  for 1 <= |d| < 2^40 (biased exponents 127..166).
 *   mode "random": a / b with TWO residual corrections over n pseudo-random pairs, mantissas
 *                  biased toward the hard cases (all-ones / all-zeros), exponents in +-20.
 * Prints the number of mismatches; exit code 0 only if there are none.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline float asf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t asu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline uint64_t rng(uint64_t* s) { uint64_t x = *s; x ^= x << 13; x ^= x >> 7; x ^= x << 17; *s = x; return x; }

int main(int argc, char** argv) {
    if (argc >= 4 && !strcmp(argv[1], "focal")) {
        const int lo = atoi(argv[2]), hi = atoi(argv[3]);
        const float b = 241.42f, y = 1.0f / b;
        long bad = 0, n = 0;
#pragma omp parallel for reduction(+ : bad, n) schedule(static)
        for (long i = (long)lo << 23; i < ((long)hi + 1) << 23; ++i) {
            for (int sgn = 0; sgn < 2; ++sgn) {
                const float a = asf((uint32_t)i | ((uint32_t)sgn << 31));
                const float q = a / b;
                float q1 = a * y;
                q1 = fmaf(fmaf(-b, q1, a), y, q1);
                bad += asu(q1) != asu(q);
                ++n;
            }
        }
        printf("focal exhaustive exponents %d..%d: cases=%ld mismatches=%ld\n", lo, hi, n, bad);
        return bad != 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "random")) {
        const long N = atol(argv[2]);
        long bad = 0;
#pragma omp parallel reduction(+ : bad)
        {
            uint64_t s = 88172645463325252ULL;
#ifdef _OPENMP
            extern int omp_get_thread_num(void);
            s ^= 0x9E3779B97F4A7C15ULL * (uint64_t)(omp_get_thread_num() + 1);
#endif
#pragma omp for schedule(static)
            for (long i = 0; i < N; ++i) {
                const uint64_t x = rng(&s);
                uint32_t ma = x & 0x7fffff, mb = (x >> 23) & 0x7fffff;
                const uint32_t ea = 127 - 20 + ((x >> 46) % 41), eb = 127 - 20 + ((x >> 52) % 41);
                const uint32_t sel = (x >> 58) & 7;
                if (sel == 0) mb |= 0x7ff000; else if (sel == 1) mb &= 0x000fff;
                else if (sel == 2) ma &= 0xfff; else if (sel == 3) ma |= 0x7ff000;
                const float a = asf((ea << 23) | ma), b = asf((eb << 23) | mb);
                const float y = 1.0f / b, q = a / b;
                float q2 = a * y;
                q2 = fmaf(fmaf(-b, q2, a), y, q2);
                q2 = fmaf(fmaf(-b, q2, a), y, q2);
                bad += asu(q2) != asu(q);
            }
        }
        printf("random two-correction: cases=%ld mismatches=%ld\n", N, bad);
        return bad != 0;
    }
    fprintf(stderr, "usage: division_check focal LO HI | random N\n");
    return 2;
}

/* Float64 division by a fixed divisor as point_cloud_kernel does it (csrc/dtsdf_kernels.cu: ddiv_fixed):
 *     y = RN(1/b);  q = RN(a y);  q = fma(fma(-b, q, a), y, q);  q = fma(fma(-b, q, a), y, q)
 * against the host's IEEE division, bit for bit.  usage: ddivision_check COUNT
 * Half of the dividends are random (sign, exponent in the kernel's admitted range 2^-900 .. 2^900, random
 * significand, plus the products u * d the kernel really forms: small integers and halves times float32 depths);
 * the other half are HARD: a = RN(b * (q + ulp(q)/2)) +- {0,1,2} ulp for a random q, so that a / b lies within
 * a few units of 2^-53 ulp of the boundary between two doubles - the cases on which a single correction can fail.
 * Divisors: 241.42 (the default focal as a double), the float32 241.42f widened, and a few hundred random ones
 * (the divisor class the theorem excepts, a significand of all ones, is excluded by the kernel too). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s_state;
static inline uint64_t rnd(uint64_t* s) {          /* splitmix64 */
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

static inline double div_fixed(double a, double b, double y) {
    double q = a * y;
    q = fma(fma(-b, q, a), y, q);
    q = fma(fma(-b, q, a), y, q);
    return q;
}

int main(int argc, char** argv) {
    const long long count = argc > 1 ? atoll(argv[1]) : 100000000ll;
    long long bad = 0, done = 0;
    double divisors[260];
    int nd = 0;
    divisors[nd++] = 241.42;
    divisors[nd++] = (double)241.42f;
    divisors[nd++] = 3.0; divisors[nd++] = 1.0; divisors[nd++] = 0.1; divisors[nd++] = 588.03;
    s_state = 12345;
    while (nd < 260) {
        const uint64_t m = rnd(&s_state) & 0xfffffffffffffull;
        if (m == 0xfffffffffffffull) continue;
        const int e = 1023 - 90 + (int)(rnd(&s_state) % 180);
        divisors[nd++] = from_bits(((uint64_t)e << 52) | m);
    }
#pragma omp parallel for reduction(+ : bad, done) schedule(static)
    for (int di = 0; di < nd; ++di) {
        const double b = divisors[di], y = 1.0 / b;
        uint64_t st = 777 + 1000003ull * (uint64_t)di;
        const long long per = (di < 2 ? count / 4 : count / 2 / (nd - 2)) + 1;
        for (long long i = 0; i < per; ++i) {
            double a;
            const uint64_t r = rnd(&st);
            switch (r & 3) {
                case 0: {           /* random in range */
                    const int e = 1023 - 900 + (int)(rnd(&st) % 1800);
                    a = from_bits(((r >> 2) & 1 ? 0x8000000000000000ull : 0) | ((uint64_t)e << 52) | (rnd(&st) & 0xfffffffffffffull));
                    break;
                }
                case 1: {           /* what the kernel forms: (integer or half) * float32 depth */
                    const double u = (double)((int)(rnd(&st) % 1281) - 640) * 0.5;
                    uint32_t fb = (uint32_t)rnd(&st);
                    float fd; memcpy(&fd, &fb, 4);
                    a = u * (double)fd;
                    break;
                }
                default: {          /* hard: the quotient next to a rounding boundary */
                    const int e = 1023 - 800 + (int)(rnd(&st) % 1600);
                    const double q = from_bits(((uint64_t)e << 52) | (rnd(&st) & 0xfffffffffffffull));
                    const double mid = q + from_bits((uint64_t)(e - 53) << 52);          /* q + ulp/2: exact in 54 bits - rounds; use fma */
                    a = fma(b, q, b * from_bits((uint64_t)(e - 53) << 52));              /* RN(b q + b ulp/2) */
                    (void)mid;
                    const int k = (int)(rnd(&st) % 5) - 2;
                    a = from_bits(bits(a) + (uint64_t)(int64_t)k);
                    if (r & 4) a = -a;
                    break;
                }
            }
            const double m = fabs(a);
            if (!(m >= 0x1p-900 && m <= 0x1p900)) continue;              /* the kernel takes the IEEE path */
            const double want = a / b, got = div_fixed(a, b, y);
            ++done;
            if (bits(want) != bits(got)) {
                if (bad < 5) printf("MISMATCH a=%a b=%a want=%a got=%a\n", a, b, want, got);
                ++bad;
            }
        }
    }
    printf("checked=%lld mismatches=%lld\n", done, bad);
    return bad != 0;
}

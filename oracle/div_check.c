/* div_check.c -- test infrastructure (not product code): checks on the host that the 3-instruction quotient the CUDA
 * kernels use for World::XtoL ((x-x0)/dh, ch3/ver2/World.h:75-81),
 *      q = a*y;  r = fma(-dh, q, a);  q' = fma(r, y, q)      with y = RN(1/dh),
 * equals the IEEE division a/dh bit for bit (plasma-simulations-by-example_b200/csrc/espic_internal.cuh: div_by_dh).
 * usage: div_check <samples per (axis, mesh size)>; prints the number of mismatches and exits 1 if there is any. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static uint64_t s[2] = {0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull};
static inline uint64_t next(void)
{
    uint64_t a = s[0], b = s[1];
    s[0] = b;
    a ^= a << 23;
    s[1] = a ^ b ^ (a >> 17) ^ (b >> 26);
    return s[1] + b;
}

static inline double div_by_dh(double a, double dh, double rdh)
{
    if (fabs(a) < 1e-280) return a / dh;
    double q = a * rdh;
    double r = fma(-dh, q, a);
    return fma(r, rdh, q);
}

int main(int argc, char **argv)
{
    long n = argc > 1 ? atol(argv[1]) : 1000000, bad = 0, total = 0;
    const double spans[3] = {0.2, 0.2, 0.4};
    const int sizes[] = {21, 41, 128, 256, 9, 13, 1000, 127, 11, 33};
    for (int si = 0; si < 3; si++)
        for (unsigned ni = 0; ni < sizeof(sizes) / sizeof(sizes[0]); ni++) {
            const double dh = spans[si] / (sizes[ni] - 1), rdh = 1.0 / dh;
            for (long t = 0; t < n; t++, total++) {
                uint64_t r = next();
                double a = (double)(r >> 11) * (1.0 / 9007199254740992.0) * spans[si];
                if (t & 1) { /* cell boundaries +- a few ulp: where a wrong last bit would change the cell index */
                    int k = (int)(r % sizes[ni]);
                    a = k * dh;
                    int d = (int)((r >> 40) % 7) - 3;
                    for (int q = 0; q < abs(d); q++) a = nextafter(a, d > 0 ? 1e9 : -1e9);
                    if (a < 0) a = 0;
                }
                if (div_by_dh(a, dh, rdh) != a / dh) {
                    if (bad++ < 5) printf("mismatch a=%a dh=%a\n", a, dh);
                }
            }
        }
    printf("mismatches=%ld of %ld\n", bad, total);
    return bad != 0;
}

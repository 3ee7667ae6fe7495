// host check of the table-based exp used by the Gaussian epilogue: exp(x) = 2^m * T[j] * (1 + p(r)),
// k = round(x * 32/ln2), m = k >> 5, j = k & 31, r = x - k * ln2/32 (Cody-Waite, two terms), |r| <= ln2/64
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
static double T[32];
static const double INV = 46.166241308446828384;            // 32 / ln2
static double HI, LO;
static inline double as_d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t as_u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static double fexp(double x) {
    const double MAGIC = 6755399441055744.0;
    if (x > 0.0) x = 0.0;
    double tm = fma(x, INV, MAGIC);
    int32_t k = (int32_t)(uint32_t)as_u(tm);
    double kf = tm - MAGIC;
    double r = fma(kf, -HI, x);
    r = fma(kf, -LO, r);
    int j = k & 31, m = k >> 5;
    double q = 1.0 / 720.0;
    q = fma(q, r, 1.0 / 120.0);
    q = fma(q, r, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    double r2 = r * r;
    double p = fma(q, r2, r);
    double res = fma(T[j], p, T[j]);
    if (m < -1020) return 0.0;
    uint64_t u = as_u(res) + ((uint64_t)(int64_t)m << 52);
    return as_d(u);
}
int main() {
    for (int j = 0; j < 32; ++j) T[j] = exp2(j / 32.0);
    // ln2/32 split: HI has its low 16 mantissa bits cleared so that k*HI is exact for |k| < 2^16
    double l = 0.021660849392498290;     // ln2/32
    l = log(2.0) / 32.0;
    uint64_t u = as_u(l) & ~((1ull << 17) - 1);
    HI = as_d(u);
    long double lo = (0.693147180559945309417232121458176568L / 32.0L) - (long double)HI;
    LO = (double)lo;   /* the kernel uses the 60-digit value 0x1.cf79abc9e3b3ap-45 */
    printf("HI=%.20e LO=%.20e INV=%.20e\n", HI, LO, 32.0 / log(2.0));
    double maxulp = 0; double worst = 0;
    srand48(1);
    long n = 40000000;
    for (long i = 0; i < n; ++i) {
        double x = (i & 1) ? -drand48() * 700.0 : -drand48() * 30.0;
        if ((i & 7) == 7) x = -drand48() * 1e-3;
        double a = fexp(x), b = exp(x);
        long double bl = expl((long double)x);
        double ulp = fabs((double)(((long double)a - bl) / (long double)(nextafter(b, INFINITY) - b)));
        if (ulp > maxulp) { maxulp = ulp; worst = x; }
    }
    printf("max ulp error vs expl over %ld args: %.4f at x=%.17g\n", n, maxulp, worst);
    printf("exp(0)=%.17g exp(-0.0)=%.17g exp(-708)=%g exp(-745)=%g exp(-1e5)=%g exp(5)=%g\n", fexp(0.0), fexp(-0.0), fexp(-708.0), fexp(-745.0), fexp(-1e5), fexp(5.0));
    return 0;
}

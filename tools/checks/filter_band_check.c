// CPU check of the error band of the centred nearest-neighbour filter ("error band of the centred
// filter"; constants kBandCentredA / kBandCentredX of csrc/icpb_internal.h).
//
// The kernels (nn_partial_centred / nn_partial_warp / nn_grid_coop) rank targets by
//     W(a,t) = |t'|^2 - 2 a'.t',   a' = fl(a - c),  t' = fl(t - c)            (float, FMA chains, see filter_w below)
// and nn_finalize declares a target j "strictly farther than target 1 in the reference's arithmetic" when
//     W_j > W_1 + ((A * kA + max(W_1 + A, 0) * kX) + kAbs),   A = fl|a'|^2 * 1.000001f.
// A target declared farther is never looked at again, so the declaration must never be wrong:
//     W_j > lim(W_1, A)   ==>   d_ref(a, t_j) > d_ref(a, t_1)       d_ref = icp.cpp:606-620 (float differences,
//                                                                      double squares, one rounding, sqrtf)
// This program reproduces both computations operation for operation and hunts for a counterexample over seeded
// near-tie configurations: |a - t_2| = |a - t_1| (1 + delta) with |delta| from 0 to 2e-5 in every direction, neighbour
// distances 1 mm .. 2 m, centres 0.1 mm .. 1 m away from the query, world coordinates 0 .. 100 m -- several 1e8
// triples per regime.  It prints, per regime, the number of violations (must be 0) and the tightest case seen:
// the largest (W_2 - W_1) / band among pairs the reference ties or ranks the other way (d_2 <= d_1).  1.0 would be
// the edge of the band; the derivation promises about a third of slack on the A term and a fifth on the D term.
//
// build: gcc -O2 -ffp-contract=off -mfma -fopenmp -o filter_band_check filter_band_check.c -lm
// run:   ./filter_band_check [samples per regime, default 2e8] [band scale, default 1]   (-> profiles/r02_check_filter_band.txt)
//        a band scale < 1 shrinks the band: the control run that shows the hunt finds violations when there are some
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static const float kU = 5.9604644775390625e-08f; // 2^-24
static const float kA = 24.0f * 5.9604644775390625e-08f, kX = 128.0f * 5.9604644775390625e-08f, kAbs = 1.0e-30f;

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rnd(rng_t *r)
{ // splitmix64
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double uni(rng_t *r) { return (double)(rnd(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline double logu(rng_t *r, double lo, double hi) { return lo * exp(uni(r) * log(hi / lo)); }
static void unit(rng_t *r, double v[3])
{
    for (;;) {
        double x = 2 * uni(r) - 1, y = 2 * uni(r) - 1, z = 2 * uni(r) - 1, n = x * x + y * y + z * z;
        if (n > 1e-4 && n <= 1.0) { n = sqrt(n); v[0] = x / n; v[1] = y / n; v[2] = z / n; return; }
    }
}

// icp.cpp:606-620 with COLOR_WEIGHT 0 (CANON-2)
static inline float d_ref(const float a[3], const float t[3])
{
    const float x = a[0] - t[0], y = a[1] - t[1], z = a[2] - t[2];
    const float s = (float)((double)x * (double)x + (double)y * (double)y + (double)z * (double)z);
    return sqrtf(s);
}

// the filter value exactly as the kernels form it (nn.cu, nn_partial_centred_kernel): targets are stored negated,
// nt = fl(-t + c) = -t'; N = fma(z,z, fma(y,y, x*x)); W = fma(2a'_z, nt_z, fma(2a'_y, nt_y, fma(2a'_x, nt_x, N)))
static inline float filter_w(const float ap[3], const float t[3], const float c[3])
{
    const float nx = -t[0] + c[0], ny = -t[1] + c[1], nz = -t[2] + c[2];
    const float N = fmaf(nz, nz, fmaf(ny, ny, nx * nx));
    const float qx = 2.f * ap[0], qy = 2.f * ap[1], qz = 2.f * ap[2];
    return fmaf(qz, nz, fmaf(qy, ny, fmaf(qx, nx, N)));
}

typedef struct {
    const char *name;
    double world_lo, world_hi; // coordinates of the query
    double s_lo, s_hi;         // neighbour distance
    double x_lo, x_hi;         // distance query -> centre
} regime_t;

int main(int argc, char **argv)
{
    const long long per = argc > 1 ? (long long)atof(argv[1]) : 200000000ll;
    const float scale = argc > 2 ? (float)atof(argv[2]) : 1.0f;
    const regime_t regs[] = {
        {"room scale (3-8 m), per-thread centres (mm-cm)", 3, 8, 1e-3, 0.5, 1e-4, 0.05},
        {"room scale (3-8 m), per-warp centres (cm-dm)", 3, 8, 1e-3, 0.5, 0.01, 0.5},
        {"room scale, far neighbours / far centres", 3, 8, 0.05, 2.0, 0.1, 1.0},
        {"near the origin (0-0.5 m)", 0, 0.5, 1e-3, 0.3, 1e-4, 0.3},
        {"large coordinates (50-100 m)", 50, 100, 1e-3, 1.0, 1e-3, 0.5},
        {"centre much farther than the neighbour (A >> D)", 3, 8, 1e-3, 0.01, 0.1, 1.0},
        {"neighbour much farther than the centre (D >> A)", 3, 8, 0.2, 2.0, 1e-4, 1e-3},
    };
    const int nreg = (int)(sizeof regs / sizeof regs[0]);
    printf("# filter_band_check: %lld near-tie triples per regime; band = %g x (24u*A + 128u*max(W1+A,0)) + 1e-30, u = 2^-24\n", per, scale);
    printf("# %-52s %12s %12s %14s %14s\n", "regime", "declared", "violations", "must-hold", "tightest");
    long long total_viol = 0;
    for (int ri = 0; ri < nreg; ++ri) {
        const regime_t rg = regs[ri];
        long long viol = 0, declared = 0, must = 0;
        double tight = 0.0;
#pragma omp parallel reduction(+ : viol, declared, must) reduction(max : tight)
        {
            int tid = 0, nth = 1;
#ifdef _OPENMP
            extern int omp_get_thread_num(void);
            extern int omp_get_num_threads(void);
            tid = omp_get_thread_num();
            nth = omp_get_num_threads();
#endif
            rng_t r = {0x1234567ull * (uint64_t)(ri + 1) + 0x9999ull * (uint64_t)tid};
            for (long long it = tid; it < per; it += nth) {
                double u1[3], u2[3], u3[3];
                unit(&r, u1); unit(&r, u2); unit(&r, u3);
                const double s = logu(&r, rg.s_lo, rg.s_hi), x = logu(&r, rg.x_lo, rg.x_hi);
                // |delta|: 0 one time in eight, else log-uniform 1e-9 .. 2e-5, either sign
                double delta = 0.0;
                const uint64_t pick = rnd(&r);
                if (pick & 7) delta = logu(&r, 1e-9, 2e-5) * ((pick & 8) ? 1.0 : -1.0);
                float a[3], c[3], t1[3], t2[3];
                for (int k = 0; k < 3; ++k) {
                    const double ak = rg.world_lo + uni(&r) * (rg.world_hi - rg.world_lo);
                    a[k] = (float)ak;
                    c[k] = (float)((double)a[k] + x * u3[k]);
                    t1[k] = (float)((double)a[k] + s * u1[k]);
                    t2[k] = (float)((double)a[k] + s * (1.0 + delta) * u2[k]);
                }
                // one time in four the second target mirrors the first through the query: equal distances by symmetry
                if ((pick >> 4 & 3) == 0)
                    for (int k = 0; k < 3; ++k) t2[k] = (float)((double)a[k] - ((double)t1[k] - (double)a[k]));
                const float ap[3] = {a[0] - c[0], a[1] - c[1], a[2] - c[2]};
                const float Araw = (ap[0] * ap[0] + ap[1] * ap[1]) + ap[2] * ap[2]; // d.pa, separately rounded (-fmad=false)
                const float A = Araw * 1.000001f;
                float w1 = filter_w(ap, t1, c), w2 = filter_w(ap, t2, c);
                float d1 = d_ref(a, t1), d2 = d_ref(a, t2);
                if (w2 < w1) { float tw = w1; w1 = w2; w2 = tw; float td = d1; d1 = d2; d2 = td; }
                // nn_finalize: band and limit in float, exactly as written there
                const float X = fmaxf(w1 + A, 0.f);
                const float band = (A * kA + X * kX) * scale + kAbs;
                const float lim = w1 + band;
                if (w2 > lim) { // declared strictly farther: the reference must agree
                    ++declared;
                    if (!(d2 > d1)) ++viol;
                }
                if (d2 <= d1) { // the reference ties or prefers target 2: W_2 must have stayed inside the band
                    ++must;
                    const double ratio = ((double)w2 - (double)w1) / (double)band;
                    if (ratio > tight) tight = ratio;
                }
            }
        }
        total_viol += viol;
        printf("  %-52s %12lld %12lld %14lld %14.4f\n", rg.name, declared, viol, must, tight);
        fflush(stdout);
    }
    printf("# total violations: %lld (%s)\n", total_viol, total_viol == 0 ? "the band held everywhere" : "THE BAND IS TOO TIGHT");
    (void)kU;
    return total_viol == 0 ? 0 : 1;
}

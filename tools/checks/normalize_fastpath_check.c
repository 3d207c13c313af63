// Exhaustive CPU check of the fast normalisation path: for every (k1, k2) the three products must round to the same
// float as the reference sequence whenever the boundary check does not flag them.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline uint64_t dbits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline int flagged(double p, int margin)
{
    int64_t L = (int64_t)(dbits(p) & 0x1FFFFFFFull) - 0x10000000ll;
    if (L < 0) L = -L;
    return L <= margin;
}
int main(int argc, char **argv)
{
    const double pert = argc > 1 ? atof(argv[1]) : 0.0; // relative perturbation of the seed
    const int margin = argc > 2 ? atoi(argv[2]) : 64;
    const int kmax = argc > 3 ? atoi(argv[3]) : 65535;
    long long bad = 0, flags = 0, total = 0;
    double worst = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : bad, flags, total) reduction(max : worst)
    for (int k1 = 0; k1 <= kmax; ++k1) {
        for (int k2 = k1; k2 <= 65535; ++k2) {
            const float v0 = -(float)k1 / 2.0f, v1 = -(float)k2 / 2.0f, v2 = 1.0f;
            const double s = ((double)v0 * (double)v0 + (double)v1 * (double)v1) + (double)v2 * (double)v2;
            // reference
            const double nv = sqrt(s);
            const double inv = 1.0 / nv;
            // fast: seed with the stated error, two Newton steps in fma form
            double y = (double)(float)(1.0 / sqrt((double)(float)s)) * (1.0 + pert);
            const double h = 0.5 * s;
            for (int it = 0; it < 2; ++it) {
                const double hy = h * y;
                const double e = fma(-hy, y, 0.5);
                y = fma(y, e, y);
            }
            const double rel = fabs(y - inv) / inv;
            if (rel > worst) worst = rel;
            const float vv[3] = {v0, v1, v2};
            for (int c = 0; c < 3; ++c) {
                const double pr = (double)vv[c] * inv, pf = (double)vv[c] * y;
                const float fr = (float)pr, ff = (float)pf;
                ++total;
                if (flagged(pf, margin)) ++flags;
                else if (memcmp(&fr, &ff, 4) != 0) ++bad;
            }
        }
    }
    printf("pert=%g margin=%d kmax=%d: products=%lld flagged=%lld (%.3g) mismatches among unflagged=%lld worst rel err of y vs inv=%.3g (2^%.1f)\n",
           pert, margin, kmax, total, flags, (double)flags / total, bad, worst, log2(worst));
    return bad != 0;
}

// Exhaustive CPU check of the division the back-projection kernel uses (csrc/cloud.cu, div_by<true>): for a divisor c
// known on the host, rc = RN(1/c) and the five-instruction FMA sequence
//     q = a*rc;  r = fma(-c, q, a);  q = fma(r, rc, q);  r = fma(-c, q, a);  q = fma(r, rc, q)
// must return the correctly rounded a / c of pointcloud.cpp:37-39 for every value the kernel can feed it:
//     z = d / 5000                         for every depth d = 1 .. 65535
//     x = ((u - cx) * z) / fx              for every d and every column / row u = 0 .. 2047
// with the reference's intrinsics (pointcloud.hpp:7-10: CX and FX on both axes) and the Kinect v2 ones
// (SLAM.cpp:26-29).  The same sweep runs on the device in tests/test_gpu_cloud.py; this is its host twin.
// build: gcc -O2 -ffp-contract=off -mfma -fopenmp -o fast_div_check fast_div_check.c -lm ; prints "ok <checks>" or the
// first mismatches.
#include <math.h>
#include <stdio.h>
#include <string.h>

static inline float div_by(float a, float c, float rc)
{
    float q = a * rc;
    float r = fmaf(-c, q, a);
    q = fmaf(r, rc, q);
    r = fmaf(-c, q, a);
    return fmaf(r, rc, q);
}
static inline unsigned bits(float x) { unsigned u; memcpy(&u, &x, 4); return u; }

int main(void)
{
    const float scale = 5000.0f, rscale = (float)(1.0 / (double)scale);
    const float cxs[4] = {318.27f, 243.99f, 250.32f, 212.55f};
    const float fxs[4] = {468.60f, 468.61f, 363.58f, 363.53f};
    long long checks = 0, bad = 0;
    for (int d = 1; d <= 65535; ++d) {
        const float z = (float)d / scale;
        ++checks;
        if (bits(z) != bits(div_by((float)d, scale, rscale))) { if (bad++ < 5) printf("z mismatch d=%d\n", d); }
    }
    for (int k = 0; k < 4; ++k) {
        const float cx = cxs[k];
        for (int j = 0; j < 4; ++j) {
            const float fx = fxs[j], rfx = (float)(1.0 / (double)fx);
            long long b = 0, n = 0;
#pragma omp parallel for schedule(static) reduction(+ : b, n)
            for (int d = 1; d <= 65535; ++d) {
                const float z = (float)d / scale;
                for (int u = 0; u < 2048; ++u) {
                    const float num = ((float)u - cx) * z;
                    ++n;
                    if (bits(num / fx) != bits(div_by(num, fx, rfx))) ++b;
                }
            }
            checks += n; bad += b;
            if (b) printf("cx=%g fx=%g: %lld mismatches\n", cx, fx, b);
        }
    }
    if (bad) { printf("FAILED %lld of %lld\n", bad, checks); return 1; }
    printf("ok %lld\n", checks);
    return 0;
}

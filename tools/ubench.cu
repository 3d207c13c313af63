// Micro-benchmarks of the FP32 pipes on sm_100a: what instruction-issue rate can the NN inner loop hope for?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int CH = 8;
// mode 0: scalar FFMA; 1: FFMA2; 2: FADD2 only; 3: kernel mix without min; 4: kernel mix with FMNMX3;
// 5: scalar mix (FADD,FMUL,FFMA,FMNMX3 per 2 pairs); 6: half packed / half scalar mix
template <int MODE>
__global__ void k(float *out, int iters, float seed)
{
    float a[CH];
    u64 p[CH];
    float m[CH];
    for (int c = 0; c < CH; ++c) { a[c] = seed + c + threadIdx.x * 1e-3f; p[c] = pack2(a[c], a[c] + 0.5f); m[c] = 1e30f; }
    const float b = 0.999f, cc = 1e-3f;
    const u64 b2 = pack2(b, b), c2 = pack2(cc, cc);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (MODE == 0) { a[c] = fmaf(a[c], b, cc); a[c] = fmaf(a[c], b, cc); a[c] = fmaf(a[c], b, cc); a[c] = fmaf(a[c], b, cc); a[c] = fmaf(a[c], b, cc); a[c] = fmaf(a[c], b, cc); }
                if (MODE == 1) { p[c] = fma2(p[c], b2, c2); p[c] = fma2(p[c], b2, c2); p[c] = fma2(p[c], b2, c2); }
                if (MODE == 2) { p[c] = add2(p[c], c2); p[c] = add2(p[c], c2); p[c] = add2(p[c], c2); }
                if (MODE == 3 || MODE == 4) {
                    // one query against 2 targets: 3 FADD2 + FMUL2 + 2 FFMA2 (+ 1 FMNMX3)
                    u64 qx = pack2(a[c], a[c]);
                    u64 dx = add2(qx, p[c]), dy = add2(qx, p[(c + 1) % CH]), dz = add2(qx, p[(c + 2) % CH]);
                    u64 s = mul2(dx, dx); s = fma2(dy, dy, s); s = fma2(dz, dz, s);
                    float s0, s1; unpack2(s, s0, s1);
                    if (MODE == 4) m[c] = min3(m[c], s0, s1);
                    else m[c] += s0 * 0.f + s1 * 0.f > 1e30f ? 1.f : 0.f; // keep s live cheaply (never true)
                }
                if (MODE == 5) {
                    float t0, t1; unpack2(p[c], t0, t1);
                    float dx0 = a[c] + t0, dy0 = a[c] + t1, dz0 = a[c] + m[c] * 0.f;
                    float dx1 = a[c] - t0, dy1 = a[c] - t1, dz1 = a[c] - cc;
                    float s0 = dx0 * dx0; s0 = fmaf(dy0, dy0, s0); s0 = fmaf(dz0, dz0, s0);
                    float s1 = dx1 * dx1; s1 = fmaf(dy1, dy1, s1); s1 = fmaf(dz1, dz1, s1);
                    m[c] = min3(m[c], s0, s1);
                }
                if (MODE == 6) {
                    // packed pair for targets (0,1), scalar for target 2: 3 pairs per "step"
                    u64 qx = pack2(a[c], a[c]);
                    u64 dx = add2(qx, p[c]), dy = add2(qx, p[(c + 1) % CH]), dz = add2(qx, p[(c + 2) % CH]);
                    u64 s = mul2(dx, dx); s = fma2(dy, dy, s); s = fma2(dz, dz, s);
                    float s0, s1; unpack2(s, s0, s1);
                    float t0, t1; unpack2(p[(c + 3) % CH], t0, t1);
                    float ex = a[c] + t0, ey = a[c] + t1, ez = a[c] + cc;
                    float s2 = ex * ex; s2 = fmaf(ey, ey, s2); s2 = fmaf(ez, ez, s2);
                    m[c] = min3(m[c], s0, s1); m[c] = fminf(m[c], s2);
                }
            }
        }
    }
    float r = 0;
    for (int c = 0; c < CH; ++c) { float lo, hi; unpack2(p[c], lo, hi); r += a[c] + lo + hi + m[c]; }
    if (r == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, double flop_per_inner, int threads, int blocks_per_sm)
{
    int dev; cudaGetDevice(&dev);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    int blocks = prop.multiProcessorCount * blocks_per_sm;
    float *out; cudaMalloc(&out, (size_t)blocks * threads * 4);
    const int iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, iters, 1.f);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, iters, 1.f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double inner = (double)blocks * threads * iters * 4 * CH;
    printf("%-34s thr=%4d blk/SM=%d  %8.3f ms  %7.2f 'TFLOP/s' (%.1f flop/inner)\n", name, threads, blocks_per_sm, best,
           inner * flop_per_inner / (best * 1e-3) / 1e12, flop_per_inner);
    cudaFree(out);
}

int main()
{
    for (int bps : {1, 2, 4}) {
        for (int thr : {128, 256, 512}) {
            if (thr * bps > 2048) continue;
            run<0>("scalar FFMA (6/inner)", 12, thr, bps);
            run<1>("FFMA2 (3/inner)", 12, thr, bps);
            run<2>("FADD2 (3/inner, count as 2flop/lane)", 12, thr, bps);
            run<3>("mix 3FADD2+FMUL2+2FFMA2 (8flop/pair)", 16, thr, bps);
            run<4>("mix + FMNMX3 (8flop/pair)", 16, thr, bps);
            run<5>("scalar mix + FMNMX3 (8flop/pair)", 16, thr, bps);
            run<6>("packed+scalar mix, 3 pairs", 24, thr, bps);
            printf("\n");
        }
    }
    return 0;
}

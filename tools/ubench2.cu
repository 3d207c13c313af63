// Replica of nn_partial's inner loop on synthetic shared-memory data: what pipe utilisation can this
// instruction mix reach?  Variants: MIN (FMNMX3 tracking on/off), EPI (group epilogue on/off), QPT.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ void sink(u64 &v) { asm volatile("" : "+l"(v)); }

constexpr int G = 32, NG = 64; // 64 groups of 32 targets in smem (24 KB)
template <int QPT, int MIN, bool EPI, int UNR>
__global__ void __launch_bounds__(128) k(float *out, int reps)
{
    __shared__ __align__(16) float s[NG * G * 3];
    for (int i = threadIdx.x; i < NG * G * 3; i += blockDim.x) s[i] = -(float)(i % 97) * 0.01f;
    __syncthreads();
    float ax[QPT], ay[QPT], az[QPT], m1[QPT], m2[QPT];
    int g1[QPT];
    for (int q = 0; q < QPT; ++q) { ax[q] = threadIdx.x * 0.001f + q; ay[q] = ax[q] + 0.5f; az[q] = ax[q] - 0.25f; m1[q] = m2[q] = 1e30f; g1[q] = 0; }
    for (int r = 0; r < reps; ++r) {
        for (int gi = 0; gi < NG; ++gi) {
            const float4 *s4 = reinterpret_cast<const float4 *>(&s[gi * (G * 3)]);
            float gm[QPT];
#pragma unroll
            for (int q = 0; q < QPT; ++q) gm[q] = 1e30f;
#pragma unroll UNR
            for (int j = 0; j < G / 4; ++j) {
                const float4 X = s4[j], Y = s4[G / 4 + j], Z = s4[2 * (G / 4) + j];
                const u64 x01 = pack2(X.x, X.y), x23 = pack2(X.z, X.w);
                const u64 y01 = pack2(Y.x, Y.y), y23 = pack2(Y.z, Y.w);
                const u64 z01 = pack2(Z.x, Z.y), z23 = pack2(Z.z, Z.w);
#pragma unroll
                for (int q = 0; q < QPT; ++q) {
                    const u64 qx = pack2(ax[q], ax[q]), qy = pack2(ay[q], ay[q]), qz = pack2(az[q], az[q]);
                    u64 dxa = add2(qx, x01), dxb = add2(qx, x23);
                    u64 dya = add2(qy, y01), dyb = add2(qy, y23);
                    u64 dza = add2(qz, z01), dzb = add2(qz, z23);
                    u64 sa = mul2(dxa, dxa), sb = mul2(dxb, dxb);
                    sa = fma2(dya, dya, sa); sb = fma2(dyb, dyb, sb);
                    sa = fma2(dza, dza, sa); sb = fma2(dzb, dzb, sb);
                    if (MIN == 1) {
                        float s0, s1, s2, s3; unpack2(sa, s0, s1); unpack2(sb, s2, s3);
                        gm[q] = min3(gm[q], s0, s1); gm[q] = min3(gm[q], s2, s3);
                    } else if (MIN == 2) { // 2-input FMNMX x4
                        float s0, s1, s2, s3; unpack2(sa, s0, s1); unpack2(sb, s2, s3);
                        gm[q] = fminf(fminf(gm[q], s0), fminf(s1, fminf(s2, s3)));
                    } else if (MIN == 3) { // half the FMNMX3 (measurement only)
                        float s0, s1; unpack2(sa, s0, s1); sink(sb);
                        gm[q] = min3(gm[q], s0, s1);
                    } else { sink(sa); sink(sb); }
                }
            }
            if (EPI) {
#pragma unroll
                for (int q = 0; q < QPT; ++q) {
                    m2[q] = fminf(m2[q], fmaxf(m1[q], gm[q]));
                    if (gm[q] < m1[q]) { m1[q] = gm[q]; g1[q] = gi; }
                }
            } else {
#pragma unroll
                for (int q = 0; q < QPT; ++q) m1[q] = fminf(m1[q], gm[q]);
            }
        }
    }
    float rsum = 0;
    for (int q = 0; q < QPT; ++q) rsum += m1[q] + m2[q] + g1[q];
    if (rsum == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = rsum;
}

template <int QPT, int MIN, bool EPI, int UNR>
void run(const char *name, int bps)
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int blocks = prop.multiProcessorCount * bps;
    float *out; cudaMalloc(&out, (size_t)blocks * 128 * 4);
    const int reps = 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<QPT, MIN, EPI, UNR><<<blocks, 128>>>(out, reps);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<QPT, MIN, EPI, UNR><<<blocks, 128>>>(out, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double pairs = (double)blocks * 128 * QPT * reps * NG * G;
    double tf = pairs * 8 / (best * 1e-3) / 1e12;
    printf("%-40s CTAs/SM=%d  %8.3f ms  %6.2f TFLOP/s (8 flop/pair) = %.1f%% of 6-cycle/pair bound @1.965GHz\n", name, bps, best, tf,
           100.0 * tf / (148 * 128 * 2 * 1.965e9 / 1e12 * 2.0 / 3.0));
    cudaFree(out);
}

int main()
{
    for (int bps : {2, 4, 5, 6, 8}) {
        run<8, 1, true, 2>("QPT8 FMNMX3 x2/4pairs (real)", bps);
        run<8, 0, false, 2>("QPT8 no min (sink)", bps);
        run<8, 2, true, 2>("QPT8 FMNMX x4/4pairs", bps);
        run<8, 3, true, 2>("QPT8 FMNMX3 x1/4pairs", bps);
        run<4, 1, true, 2>("QPT4 FMNMX3 x2/4pairs", bps);
        run<4, 0, false, 2>("QPT4 no min (sink)", bps);
        printf("\n");
    }
}

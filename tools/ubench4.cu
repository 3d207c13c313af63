// Issue cost of the expansion-form inner step: 3 FFMA2 (broadcast query operand) per two pairs, plus one
// reduction instruction per two pairs (FMNMX3 / LOP3-OR of the sign bits / none).  Self-feeding chains.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ unsigned or3(unsigned a, unsigned b, unsigned c) { unsigned d; asm("lop3.b32 %0, %1, %2, %3, 0xFE;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

constexpr int CH = 16;
// MODE 0: 3 FFMA2 only   1: + FMNMX3   2: + LOP3.OR   3: + LOP3.OR into 4 rotating accumulators   4: + 2-input FMNMX x2
// MODE 5: + FMNMX3 every second step on half the values (cost probe)   6: + IADD3 (sum of raw bits)
template <int MODE>
__global__ void __launch_bounds__(128) k(float *out, int iters)
{
    u64 t[CH];
    float qa[CH], m[CH];
    unsigned acc[4] = {0, 0, 0, 0};
    for (int c = 0; c < CH; ++c) { qa[c] = 1.0f + c * 0.01f + threadIdx.x * 1e-4f; t[c] = pack2(-qa[c] + 1e-3f, -qa[c] - 1e-3f); m[c] = 1e30f; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            u64 qx = pack2(qa[c], qa[c]);
            u64 qy = pack2(qa[(c + 3) % CH], qa[(c + 3) % CH]);
            u64 qz = pack2(qa[(c + 5) % CH], qa[(c + 5) % CH]);
            u64 s = fma2(qx, t[c], t[(c + 7) % CH]);
            s = fma2(qy, t[(c + 1) % CH], s);
            s = fma2(qz, t[(c + 2) % CH], s);
            float s0, s1; unpack2(s, s0, s1);
            if (MODE == 1) m[c] = min3(m[c], s0, s1);
            if (MODE == 2) acc[0] = or3(acc[0], __float_as_uint(s0), __float_as_uint(s1));
            if (MODE == 3) acc[c & 3] = or3(acc[c & 3], __float_as_uint(s0), __float_as_uint(s1));
            if (MODE == 4) m[c] = fminf(m[c], fminf(s0, s1));
            if (MODE == 5 && (c & 1)) m[c] = min3(m[c], s0, s1);
            if (MODE == 6) acc[c & 3] += __float_as_uint(s0) + __float_as_uint(s1);
            t[c] = s;
        }
    }
    float r = (float)(acc[0] ^ acc[1] ^ acc[2] ^ acc[3]);
    for (int c = 0; c < CH; ++c) { float lo, hi; unpack2(t[c], lo, hi); r += lo + hi + m[c]; }
    if (r == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, int bps)
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int blocks = prop.multiProcessorCount * bps;
    float *out; cudaMalloc(&out, (size_t)blocks * 128 * 4);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 128>>>(out, iters);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<MODE><<<blocks, 128>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double steps = (double)blocks * 4 * iters * CH;
    double cyc = best * 1e-3 * 1.965e9 * prop.multiProcessorCount * 4 / steps;
    printf("%-52s CTAs/SM=%d %8.3f ms  %.2f SMSP-cycles per 2-pair step (3 FFMA2 = 6.00)\n", name, bps, best, cyc);
    cudaFree(out);
}

int main()
{
    for (int bps : {3, 4, 8}) {
        run<0>("3 FFMA2", bps);
        run<1>("3 FFMA2 + FMNMX3", bps);
        run<2>("3 FFMA2 + LOP3.OR (one accumulator)", bps);
        run<3>("3 FFMA2 + LOP3.OR (four accumulators)", bps);
        run<4>("3 FFMA2 + 2 FMNMX", bps);
        run<5>("3 FFMA2 + FMNMX3 on every second step", bps);
        run<6>("3 FFMA2 + 2 IADD (raw bits)", bps);
        printf("\n");
    }
}

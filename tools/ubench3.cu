// Isolate what limits the packed-FP32 issue rate: self-feeding chains (nothing can be hoisted or eliminated).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int CH = 16; // independent chains per thread
// MODE 0: mix, broadcast query operand, no min   1: + FMNMX3 per chain step   2: + 2x FMNMX   3: + LOP3 (xor)
// MODE 4: mix with pair (non-broadcast) query operand, no min
// MODE 5: scalar mix (FADD/FMUL/FFMA), no min       6: scalar mix + FMNMX3 per 2 steps
template <int MODE>
__global__ void __launch_bounds__(128) k(float *out, int iters)
{
    u64 t[CH];
    float qa[CH], m[CH];
    unsigned acc = 0;
    unsigned mi[CH];
    float ps0 = 0.f, ps1 = 0.f;
    for (int c = 0; c < CH; ++c) { qa[c] = 1.0f + c * 0.01f + threadIdx.x * 1e-4f; t[c] = pack2(-qa[c] + 1e-3f, -qa[c] - 1e-3f); m[c] = 1e30f; mi[c] = 0x7f000000u; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (MODE <= 4 || MODE >= 7) {
                u64 qx = (MODE == 4) ? pack2(qa[c], qa[(c + 1) % CH]) : pack2(qa[c], qa[c]);
                u64 dx = add2(qx, t[c]), dy = add2(qx, t[(c + 1) % CH]), dz = add2(qx, t[(c + 2) % CH]);
                u64 s = mul2(dx, dx); s = fma2(dy, dy, s); s = fma2(dz, dz, s);
                float s0, s1; unpack2(s, s0, s1);
                if (MODE == 1) m[c] = min3(m[c], s0, s1);
                if (MODE == 2) m[c] = fminf(m[c], fminf(s0, s1));
                if (MODE == 3) acc ^= __float_as_uint(s0) ^ __float_as_uint(s1);
                if (MODE == 7) mi[c] = __vimin3_u32(mi[c], __float_as_uint(s0), __float_as_uint(s1));
                if (MODE == 8) mi[c] = min(mi[c], min(__float_as_uint(s0), __float_as_uint(s1)));
                if (MODE == 9) mi[c] = (unsigned)__vimin3_s32((int)mi[c], __float_as_int(s0), __float_as_int(s1));
                if (MODE == 10) { if (c & 1) { m[c] = min3(m[c], s0, s1); m[c] = min3(m[c], ps0, ps1);} else { ps0 = s0; ps1 = s1; } }
                t[c] = s; // feeds the next iteration: nothing is loop invariant or dead
            } else {
                float a, b; unpack2(t[c], a, b);
                float dx = qa[c] + a, dy = qa[c] + b, dz = qa[c] + m[(c + 1) % CH];
                float s = dx * dx; s = fmaf(dy, dy, s); s = fmaf(dz, dz, s);
                float ex = qa[c] - a, ey = qa[c] - b, ez = qa[c] - m[(c + 2) % CH];
                float s2 = ex * ex; s2 = fmaf(ey, ey, s2); s2 = fmaf(ez, ez, s2);
                if (MODE == 6) m[c] = min3(m[c], s, s2);
                t[c] = pack2(s, s2);
            }
        }
    }
    float r = (float)acc;
    for (int c = 0; c < CH; ++c) r += (float)mi[c];
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
    // every chain step = 2 pairs (6 packed or 12 scalar FMA-pipe instructions = 12 pipe cycles per warp)
    double steps = (double)blocks * 4 /*warps*/ * iters * CH;
    double cyc_per_step = best * 1e-3 * 1.965e9 * prop.multiProcessorCount * 4 / steps; // SMSP cycles per warp-step
    printf("%-46s CTAs/SM=%d %8.3f ms  %.2f SMSP-cycles per 2-pair step (ideal 12.00) -> %.1f%% pipe\n", name, bps, best,
           cyc_per_step, 100.0 * 12.0 / cyc_per_step);
    cudaFree(out);
}

int main()
{
    for (int bps : {4, 8}) {
        run<0>("packed mix, broadcast query, no min", bps);
        run<4>("packed mix, pair query operand, no min", bps);
        run<1>("packed mix + FMNMX3 / 2 pairs", bps);
        run<2>("packed mix + 2 FMNMX / 2 pairs", bps);
        run<3>("packed mix + LOP3 / 2 pairs", bps);
        run<7>("packed mix + VIMNMX3.U32 / 2 pairs", bps);
        run<8>("packed mix + 2 IMNMX.U32 / 2 pairs", bps);
        run<9>("packed mix + VIMNMX3.S32 / 2 pairs", bps);
        run<10>("packed mix + 2 FMNMX3 back-to-back / 4 pairs", bps);
        run<5>("scalar mix, no min", bps);
        run<6>("scalar mix + FMNMX3 / 2 pairs", bps);
        printf("\n");
    }
}

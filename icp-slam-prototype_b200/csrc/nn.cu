// Brute-force nearest-neighbour association + rigid solve, device resident.
//
// Replaces (reference file:line, relative to the reference repository):
//   icp::distance                          icp.cpp:606-620   (N1)
//   icp::getNearestPoint                   icp.cpp:566-593   (N2)
//   icp::findGlobalNearestNeighborAssociations icp.cpp:541-563 (N3)
//   the solve / pose update / convergence test of icp::getTransformation
//                                          icp.cpp:155-258   (S1-S3)
//   icp::calculateOffset icp.cpp:314-344, icp::meanSquareError icp.cpp:622-638
//   PointCloud::rotate / translate         pointcloud.cpp:321-359 (P2, fused into the query load)
//
// Built with -fmad=false: every float / double expression below is evaluated
// with separately rounded IEEE operations, like the reference's /fp:precise
// build.  The only fused multiply-adds are the explicit fma.rn.f32x2 of the
// approximate filter, whose result never reaches an output.
#include <math_constants.h>

#include "icpb_internal.h"

namespace icpb {

// --------------------------------------------------------------------------
// packed FP32x2 helpers (sm_100a FADD2 / FMUL2 / FFMA2) and 3-input min (FMNMX3)
// --------------------------------------------------------------------------
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float min3(float a, float b, float c)
{
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// --------------------------------------------------------------------------
// mbarrier + 1-D bulk TMA (cp.async.bulk -> SASS UBLKCP)
// --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// --------------------------------------------------------------------------
// exact reference arithmetic
// --------------------------------------------------------------------------

// icp.cpp:606-620: float differences; squares and their left-to-right sum in
// double (pow(float,2) promotes); ONE rounding to float; correctly rounded
// float sqrt.  No FMA anywhere (-fmad=false).
__device__ __forceinline__ float exact_distance(float ax, float ay, float az, float bx, float by, float bz)
{
    float x = ax - bx;
    float y = ay - by;
    float z = az - bz;
    double s = ((double)x * (double)x + (double)y * (double)y) + (double)z * (double)z;
    float xyz = (float)s;
    return sqrtf(xyz);
}

// pointcloud.cpp:321-331 (rotate: OpenCV 3x3 by 3xN float gemm = float, no
// FMA, left to right) followed by :349-359 (translate: float add).
__device__ __forceinline__ float4 apply_rt(float4 p, const float *R, const float *t)
{
    float4 o;
    o.x = ((R[0] * p.x + R[1] * p.y) + R[2] * p.z) + t[0];
    o.y = ((R[3] * p.x + R[4] * p.y) + R[5] * p.z) + t[1];
    o.z = ((R[6] * p.x + R[7] * p.y) + R[8] * p.z) + t[2];
    o.w = p.w;
    return o;
}

// --------------------------------------------------------------------------
// target preparation: AoS -> negated, group-tiled SoA, padded to whole groups
// --------------------------------------------------------------------------
__global__ void target_prep_kernel(const float4 *__restrict__ tgt, int m, float *__restrict__ soa, int ngroups)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ngroups * kGroup) return;
    float x = -kPadCoord, y = -kPadCoord, z = -kPadCoord;
    if (t < m) {
        float4 p = tgt[t];
        x = -p.x; y = -p.y; z = -p.z;
    }
    int g = t / kGroup, l = t % kGroup;
    float *base = soa + (size_t)g * (3 * kGroup);
    base[l] = x;
    base[kGroup + l] = y;
    base[2 * kGroup + l] = z;
}

void launch_target_prep(const float4 *tgt, int m, float *soa, int ngroups, cudaStream_t s)
{
    int total = ngroups * kGroup;
    target_prep_kernel<<<(total + 255) / 256, 256, 0, s>>>(tgt, m, soa, ngroups);
}

// --------------------------------------------------------------------------
// nn_partial: the N x M scan (approximate FP32 filter, group minima)
// --------------------------------------------------------------------------
//
// Each thread keeps QPT queries in registers and streams the targets of its
// split through shared memory (bulk-TMA ring).  Per (query, target) pair the
// FMA pipe sees 3 FADD + 1 FMUL + 2 FFMA (packed two targets at a time), and
// the ALU pipe half an FMNMX3.  Per group of 32 targets it keeps the best and
// second-best GROUP minimum and the best group's id; nn_finalize re-evaluates
// the best group in the reference's exact arithmetic.
template <int QPT>
__global__ void __launch_bounds__(kNnThreads) nn_partial_kernel(const RegDesc *__restrict__ descs, IcpState *states,
                                                                int splits, int pass)
{
    pdl_enter(); // icpb_internal.h: the grid before this one is complete from here on
    // the loop state sits at states[blockIdx.z] (== d.st): addressed from the kernel argument, its load does not wait
    // for the descriptor's
    IcpState *st = states + blockIdx.z;
    const RegDesc d = descs[blockIdx.z]; // by value: no pointer reloads after stores
    if (st->done) return;
    const int n = d.n;
    const int q0 = blockIdx.x * (kNnThreads * QPT);
    if (q0 >= n) return;
    const int split = blockIdx.y;
    const int tid = threadIdx.x;

    const int gps = (d.ngroups + splits - 1) / splits;
    const int g_begin = split * gps;
    const int g_end = min(d.ngroups, g_begin + gps);
    const int n_tiles = (g_end > g_begin) ? (g_end - g_begin + kStageGroups - 1) / kStageGroups : 0;

    constexpr int kStageFloats = kStageGroups * kGroup * 3;
    __shared__ __align__(128) float s_tile[kStages][kStageFloats];
    __shared__ __align__(8) uint64_t s_full[kStages];

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&s_full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const float *soa = d.tgt_soa;
    auto issue = [&](int tile) {
        int gb = g_begin + tile * kStageGroups;
        int ng = min(kStageGroups, g_end - gb);
        uint32_t bytes = (uint32_t)ng * (kGroup * 3 * sizeof(float));
        int s = tile % kStages;
        mbar_expect_tx(&s_full[s], bytes);
        tma_bulk_g2s(&s_tile[s][0], soa + (size_t)gb * (kGroup * 3), bytes, &s_full[s]);
    };
    if (tid == 0) {
        for (int t = 0; t < kStages && t < n_tiles; ++t) issue(t);
    }

    // ---- queries: load, apply the pending rigid motion, hand on to the next buffer
    const float4 *src = d.D[pass & 1];
    float4 *dst = d.D[(pass + 1) & 1];
    const int apply = st->apply;
    float R[9], T[3];
    if (apply) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = st->Rf[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) T[k] = st->tf[k];
    }
    float ax[QPT], ay[QPT], az[QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        int i = q0 + q * kNnThreads + tid;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n) {
            p = src[i];
            if (apply) p = apply_rt(p, R, T);
            if (split == 0) dst[i] = p;
        }
        ax[q] = p.x; ay[q] = p.y; az[q] = p.z;
    }

    float m1[QPT], m2[QPT];
    int g1[QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        m1[q] = CUDART_INF_F; m2[q] = CUDART_INF_F; g1[q] = g_begin < d.ngroups ? g_begin : 0;
    }

    for (int tile = 0; tile < n_tiles; ++tile) {
        const int s = tile % kStages;
        mbar_wait(&s_full[s], (uint32_t)((tile / kStages) & 1));
        const int gb = g_begin + tile * kStageGroups;
        const int ng = min(kStageGroups, g_end - gb);
        for (int gi = 0; gi < ng; ++gi) {
            const float4 *s4 = reinterpret_cast<const float4 *>(&s_tile[s][gi * (kGroup * 3)]);
            float gm[QPT];
#pragma unroll
            for (int q = 0; q < QPT; ++q) gm[q] = CUDART_INF_F;
#pragma unroll kUnrollJ
            for (int j = 0; j < kGroup / 4; ++j) {
                const float4 X = s4[j];
                const float4 Y = s4[kGroup / 4 + j];
                const float4 Z = s4[2 * (kGroup / 4) + j];
                const u64 x01 = pack2(X.x, X.y), x23 = pack2(X.z, X.w);
                const u64 y01 = pack2(Y.x, Y.y), y23 = pack2(Y.z, Y.w);
                const u64 z01 = pack2(Z.x, Z.y), z23 = pack2(Z.z, Z.w);
#pragma unroll
                for (int q = 0; q < QPT; ++q) {
                    const u64 qx = pack2(ax[q], ax[q]);
                    const u64 qy = pack2(ay[q], ay[q]);
                    const u64 qz = pack2(az[q], az[q]);
                    u64 dxa = add2(qx, x01), dxb = add2(qx, x23);
                    u64 dya = add2(qy, y01), dyb = add2(qy, y23);
                    u64 dza = add2(qz, z01), dzb = add2(qz, z23);
                    u64 sa = mul2(dxa, dxa), sb = mul2(dxb, dxb);
                    sa = fma2(dya, dya, sa); sb = fma2(dyb, dyb, sb);
                    sa = fma2(dza, dza, sa); sb = fma2(dzb, dzb, sb);
                    float s0, s1, s2, s3;
                    unpack2(sa, s0, s1);
                    unpack2(sb, s2, s3);
                    gm[q] = min3(gm[q], s0, s1);
                    gm[q] = min3(gm[q], s2, s3);
                }
            }
            const int g = gb + gi;
#pragma unroll
            for (int q = 0; q < QPT; ++q) {
                m2[q] = fminf(m2[q], fmaxf(m1[q], gm[q]));
                if (gm[q] < m1[q]) { m1[q] = gm[q]; g1[q] = g; }
            }
        }
        __syncthreads(); // every warp is done with stage s
        if (tid == 0 && tile + kStages < n_tiles) issue(tile + kStages);
    }

    const size_t row = (size_t)split * d.n_stride;
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        int i = q0 + q * kNnThreads + tid;
        if (i < n) {
            d.pm1[row + i] = m1[q];
            d.pm2[row + i] = m2[q];
            d.pg[row + i] = g1[q];
        }
    }
}

// --------------------------------------------------------------------------
// nn_partial_centred: the N x M scan with the expanded, locally centred filter
// --------------------------------------------------------------------------
//
// Same contract as nn_partial_kernel (best / second-best GROUP minimum and the
// best group's id per query and split), but the filter value is
//     W(a,t) = |t'|^2 - 2 a'.t'  =  |a - t|^2 - |a'|^2,   a' = a - c,  t' = t - c
// with c the centre of the QPT *consecutive* queries a thread owns.  Per
// (query, target) pair the FMA pipe sees 3 FFMA (packed two targets at a time)
// instead of 3 FADD + FMUL + 2 FFMA; centring the targets costs 6 packed-lane
// ops per target per thread, amortised over the thread's QPT queries.
// Centring is what keeps the expansion usable: its rounding error scales with
// (|a'| + |t'|)^2, i.e. with the spread of one thread's queries (centimetres
// for a raster-ordered cloud) instead of the 5-8 m world coordinates.
//
// Error bound and the band nn_finalize applies (kBandCentredA / kBandCentredX, icpb_internal.h): ONE derivation, in
// DESIGN.md section 4, "error band of the centred filter".  A = |a'|^2 is written to d.pa by split 0.
template <int QPT>
__global__ void __launch_bounds__(kNnThreads, (QPT >= 16 ? 2 : QPT >= 12 ? 3 : 4)) nn_partial_centred_kernel(const RegDesc *__restrict__ descs,
                                                                        IcpState *states, int splits, int pass)
{
    pdl_enter(); // icpb_internal.h: the grid before this one is complete from here on
    // the loop state sits at states[blockIdx.z] (== d.st): addressed from the kernel argument, its load does not wait
    // for the descriptor's
    IcpState *st = states + blockIdx.z;
    const RegDesc d = descs[blockIdx.z]; // by value: no pointer reloads after stores
    if (st->done) return;
    const int n = d.n;
    const int q0 = blockIdx.x * (kNnThreads * QPT);
    if (q0 >= n) return;
    const int split = blockIdx.y;
    const int tid = threadIdx.x;

    const int gps = (d.ngroups + splits - 1) / splits;
    const int g_begin = split * gps;
    const int g_end = min(d.ngroups, g_begin + gps);
    const int n_tiles = (g_end > g_begin) ? (g_end - g_begin + kStageGroups - 1) / kStageGroups : 0;

    constexpr int kStageFloats = kStageGroups * kGroup * 3;
    __shared__ __align__(128) float s_tile[kStages][kStageFloats];
    __shared__ __align__(8) uint64_t s_full[kStages];

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&s_full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const float *soa = d.tgt_soa;
    auto issue = [&](int tile) {
        int gb = g_begin + tile * kStageGroups;
        int ng = min(kStageGroups, g_end - gb);
        uint32_t bytes = (uint32_t)ng * (kGroup * 3 * sizeof(float));
        int s = tile % kStages;
        mbar_expect_tx(&s_full[s], bytes);
        tma_bulk_g2s(&s_tile[s][0], soa + (size_t)gb * (kGroup * 3), bytes, &s_full[s]);
    };
    if (tid == 0) {
        for (int t = 0; t < kStages && t < n_tiles; ++t) issue(t);
    }

    // ---- queries: QPT consecutive points per thread; apply the pending rigid motion, hand on to the next buffer
    const float4 *src = d.D[pass & 1];
    float4 *dst = d.D[(pass + 1) & 1];
    const int apply = st->apply;
    float R[9], T[3];
    if (apply) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = st->Rf[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) T[k] = st->tf[k];
    }
    const int i0 = q0 + tid * QPT;
    float ax[QPT], ay[QPT], az[QPT];
    float lox = CUDART_INF_F, loy = CUDART_INF_F, loz = CUDART_INF_F;
    float hix = -CUDART_INF_F, hiy = -CUDART_INF_F, hiz = -CUDART_INF_F;
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        const int i = i0 + q;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n) {
            p = src[i];
            if (apply) p = apply_rt(p, R, T);
            if (split == 0) dst[i] = p;
            lox = fminf(lox, p.x); loy = fminf(loy, p.y); loz = fminf(loz, p.z);
            hix = fmaxf(hix, p.x); hiy = fmaxf(hiy, p.y); hiz = fmaxf(hiz, p.z);
        }
        ax[q] = p.x; ay[q] = p.y; az[q] = p.z;
    }
    // centre of the thread's own queries (any value is valid; it only sets the size of the error bound)
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (i0 < n) { cx = 0.5f * lox + 0.5f * hix; cy = 0.5f * loy + 0.5f * hiy; cz = 0.5f * loz + 0.5f * hiz; }
    // a' = a - c; the scan uses 2a' against the negated centred targets -(t - c)
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        const bool v = (i0 + q) < n;
        const float x = v ? ax[q] - cx : 0.f, y = v ? ay[q] - cy : 0.f, z = v ? az[q] - cz : 0.f;
        if (v && split == 0) d.pa[i0 + q] = (x * x + y * y) + z * z;
        ax[q] = 2.f * x; ay[q] = 2.f * y; az[q] = 2.f * z;
    }
    const u64 c2x = pack2(cx, cx), c2y = pack2(cy, cy), c2z = pack2(cz, cz);

    // three best GROUP minima per query (ids for the best two): nn_finalize re-evaluates one or two groups exactly
    float m1[QPT], m2[QPT], m3[QPT];
    int g1[QPT], g2[QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        m1[q] = CUDART_INF_F; m2[q] = CUDART_INF_F; m3[q] = CUDART_INF_F;
        g1[q] = g_begin < d.ngroups ? g_begin : 0; g2[q] = -1;
    }

    for (int tile = 0; tile < n_tiles; ++tile) {
        const int s = tile % kStages;
        mbar_wait(&s_full[s], (uint32_t)((tile / kStages) & 1));
        const int gb = g_begin + tile * kStageGroups;
        const int ng = min(kStageGroups, g_end - gb);
        for (int gi = 0; gi < ng; ++gi) {
            const float4 *s4 = reinterpret_cast<const float4 *>(&s_tile[s][gi * (kGroup * 3)]);
            float gm[QPT];
#pragma unroll
            for (int q = 0; q < QPT; ++q) gm[q] = CUDART_INF_F;
#pragma unroll kUnrollJ
            for (int j = 0; j < kGroup / 4; ++j) {
                const float4 X = s4[j];
                const float4 Y = s4[kGroup / 4 + j];
                const float4 Z = s4[2 * (kGroup / 4) + j];
                // -(t - c) for four targets, and |t - c|^2
                const u64 x01 = add2(pack2(X.x, X.y), c2x), x23 = add2(pack2(X.z, X.w), c2x);
                const u64 y01 = add2(pack2(Y.x, Y.y), c2y), y23 = add2(pack2(Y.z, Y.w), c2y);
                const u64 z01 = add2(pack2(Z.x, Z.y), c2z), z23 = add2(pack2(Z.z, Z.w), c2z);
                const u64 n01 = fma2(z01, z01, fma2(y01, y01, mul2(x01, x01)));
                const u64 n23 = fma2(z23, z23, fma2(y23, y23, mul2(x23, x23)));
#pragma unroll
                for (int q = 0; q < QPT; ++q) {
                    const u64 qx = pack2(ax[q], ax[q]);
                    const u64 qy = pack2(ay[q], ay[q]);
                    const u64 qz = pack2(az[q], az[q]);
                    u64 sa = fma2(qx, x01, n01), sb = fma2(qx, x23, n23);
                    sa = fma2(qy, y01, sa); sb = fma2(qy, y23, sb);
                    sa = fma2(qz, z01, sa); sb = fma2(qz, z23, sb);
                    float s0, s1, s2, s3;
                    unpack2(sa, s0, s1);
                    unpack2(sb, s2, s3);
                    gm[q] = min3(gm[q], s0, s1);
                    gm[q] = min3(gm[q], s2, s3);
                }
            }
            const int g = gb + gi;
            // most groups beat none of the thread's third-best minima: one compare per query, then skip
            bool hit = false;
#pragma unroll
            for (int q = 0; q < QPT; ++q) hit |= gm[q] < m3[q];
            if (hit) {
#pragma unroll
                for (int q = 0; q < QPT; ++q) {
                    const float v = gm[q];
                    const bool lt1 = v < m1[q], lt2 = v < m2[q];
                    m3[q] = lt2 ? m2[q] : fminf(m3[q], v);
                    g2[q] = lt1 ? g1[q] : (lt2 ? g : g2[q]);
                    m2[q] = lt1 ? m1[q] : (lt2 ? v : m2[q]);
                    g1[q] = lt1 ? g : g1[q];
                    m1[q] = lt1 ? v : m1[q];
                }
            }
        }
        __syncthreads(); // every warp is done with stage s
        if (tid == 0 && tile + kStages < n_tiles) issue(tile + kStages);
    }

    const size_t row = (size_t)split * d.n_stride;
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        const int i = i0 + q;
        if (i < n) {
            d.pm1[row + i] = m1[q];
            d.pm2[row + i] = m2[q];
            d.pm3[row + i] = m3[q];
            d.pg[row + i] = g1[q];
            d.pg2[row + i] = g2[q];
        }
    }
}

// --------------------------------------------------------------------------
// nn_partial_warp: the centred filter about ONE centre per warp
// --------------------------------------------------------------------------
//
// Same filter value and the same error bound as nn_partial_centred_kernel, with c the centre of the WARP's queries:
// the 32 targets of a group are centred once per warp -- lane l lifts target l into a warp-private shared-memory
// buffer -- instead of once per thread, which removes the 12 packed operations per 4-target step that the
// per-thread centring costs (7 - 9 % of the scan).  A warp's queries must then be spatial neighbours: threads take
// their queries through d.perm, the Morton order built once per registration (grid.cu); a rigid motion keeps
// neighbours neighbours, so the order is not rebuilt between passes.  A = |a - c|^2 is larger than with per-thread
// centres (decimetres instead of centimetres), so a few more queries need the second group or the full rescan.
template <int QPT>
__global__ void __launch_bounds__(kNnThreads, (QPT >= 16 ? 2 : QPT >= 12 ? 3 : 4)) nn_partial_warp_kernel(
    const RegDesc *__restrict__ descs, IcpState *states, int splits, int pass)
{
    pdl_enter(); // icpb_internal.h: the grid before this one is complete from here on
    // the loop state sits at states[blockIdx.z] (== d.st): addressed from the kernel argument, its load does not wait
    // for the descriptor's
    IcpState *st = states + blockIdx.z;
    const RegDesc d = descs[blockIdx.z]; // by value: no pointer reloads after stores
    if (st->done) return;
    const int n = d.n;
    const int q0 = blockIdx.x * (kNnThreads * QPT);
    if (q0 >= n) return;
    const int split = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    const int gps = (d.ngroups + splits - 1) / splits;
    const int g_begin = split * gps;
    const int g_end = min(d.ngroups, g_begin + gps);
    const int n_tiles = (g_end > g_begin) ? (g_end - g_begin + kStageGroups - 1) / kStageGroups : 0;

    constexpr int kStageFloats = kStageGroups * kGroup * 3;
    __shared__ __align__(128) float s_tile[kStages][kStageFloats];
    __shared__ __align__(16) float s_wb[kNnThreads / 32][2][4 * kGroup]; // per warp, double buffered: x32 y32 z32 n32
    __shared__ __align__(8) uint64_t s_full[kStages];

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&s_full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const float *soa = d.tgt_soa;
    auto issue = [&](int tile) {
        int gb = g_begin + tile * kStageGroups;
        int ng = min(kStageGroups, g_end - gb);
        uint32_t bytes = (uint32_t)ng * (kGroup * 3 * sizeof(float));
        int s = tile % kStages;
        mbar_expect_tx(&s_full[s], bytes);
        tma_bulk_g2s(&s_tile[s][0], soa + (size_t)gb * (kGroup * 3), bytes, &s_full[s]);
    };
    if (tid == 0) {
        for (int t = 0; t < kStages && t < n_tiles; ++t) issue(t);
    }

    // ---- queries: QPT consecutive SORTED slots per thread; apply the pending rigid motion, hand on to the next buffer
    const float4 *src = d.D[pass & 1];
    float4 *dst = d.D[(pass + 1) & 1];
    const int apply = st->apply;
    float R[9], T[3];
    if (apply) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = st->Rf[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) T[k] = st->tf[k];
    }
    const int k0 = q0 + tid * QPT;
    int qi[QPT];
    float ax[QPT], ay[QPT], az[QPT];
    float lox = CUDART_INF_F, loy = CUDART_INF_F, loz = CUDART_INF_F;
    float hix = -CUDART_INF_F, hiy = -CUDART_INF_F, hiz = -CUDART_INF_F;
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        const int k = k0 + q;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        qi[q] = -1;
        if (k < n) {
            const int i = d.perm[k];
            qi[q] = i;
            p = src[i];
            if (apply) p = apply_rt(p, R, T);
            if (split == 0) dst[i] = p;
            lox = fminf(lox, p.x); loy = fminf(loy, p.y); loz = fminf(loz, p.z);
            hix = fmaxf(hix, p.x); hiy = fmaxf(hiy, p.y); hiz = fmaxf(hiz, p.z);
        }
        ax[q] = p.x; ay[q] = p.y; az[q] = p.z;
    }
    // centre of the warp's queries (any value is valid; it only sets the size of the error bound)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, off)); hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, off));
        loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, off)); hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, off));
        loz = fminf(loz, __shfl_xor_sync(0xffffffffu, loz, off)); hiz = fmaxf(hiz, __shfl_xor_sync(0xffffffffu, hiz, off));
    }
    const bool any_query = lox <= hix; // false only for a warp past the end of the cloud
    const float cx = any_query ? 0.5f * lox + 0.5f * hix : 0.f, cy = any_query ? 0.5f * loy + 0.5f * hiy : 0.f,
                cz = any_query ? 0.5f * loz + 0.5f * hiz : 0.f;
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        const bool v = qi[q] >= 0;
        const float x = v ? ax[q] - cx : 0.f, y = v ? ay[q] - cy : 0.f, z = v ? az[q] - cz : 0.f;
        if (v && split == 0) d.pa[qi[q]] = (x * x + y * y) + z * z;
        ax[q] = 2.f * x; ay[q] = 2.f * y; az[q] = 2.f * z;
    }

    float m1[QPT], m2[QPT], m3[QPT];
    int g1[QPT], g2[QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        m1[q] = CUDART_INF_F; m2[q] = CUDART_INF_F; m3[q] = CUDART_INF_F;
        g1[q] = g_begin < d.ngroups ? g_begin : 0; g2[q] = -1;
    }

    float *wb0 = &s_wb[wid][0][0], *wb1 = &s_wb[wid][1][0];
    // lane l centres target l of a group: -(t - c) and |t - c|^2 into the warp's buffer
    auto centre_group = [&](const float *grp, float *wb) {
        const float tx = grp[lane] + cx, ty = grp[kGroup + lane] + cy, tz = grp[2 * kGroup + lane] + cz;
        wb[lane] = tx; wb[kGroup + lane] = ty; wb[2 * kGroup + lane] = tz;
        wb[3 * kGroup + lane] = __fmaf_rn(tz, tz, __fmaf_rn(ty, ty, tx * tx));
    };

    for (int tile = 0; tile < n_tiles; ++tile) {
        const int s = tile % kStages;
        mbar_wait(&s_full[s], (uint32_t)((tile / kStages) & 1));
        const int gb = g_begin + tile * kStageGroups;
        const int ng = min(kStageGroups, g_end - gb);
        centre_group(&s_tile[s][0], wb0);
        for (int gi = 0; gi < ng; ++gi) {
            float *cur = (gi & 1) ? wb1 : wb0;
            if (gi + 1 < ng) centre_group(&s_tile[s][(gi + 1) * (kGroup * 3)], (gi & 1) ? wb0 : wb1);
            __syncwarp();
            const float4 *w4 = reinterpret_cast<const float4 *>(cur);
            float gm[QPT];
#pragma unroll
            for (int q = 0; q < QPT; ++q) gm[q] = CUDART_INF_F;
#pragma unroll kUnrollJ
            for (int j = 0; j < kGroup / 4; ++j) {
                const float4 X = w4[j];
                const float4 Y = w4[kGroup / 4 + j];
                const float4 Z = w4[2 * (kGroup / 4) + j];
                const float4 N = w4[3 * (kGroup / 4) + j];
                const u64 x01 = pack2(X.x, X.y), x23 = pack2(X.z, X.w);
                const u64 y01 = pack2(Y.x, Y.y), y23 = pack2(Y.z, Y.w);
                const u64 z01 = pack2(Z.x, Z.y), z23 = pack2(Z.z, Z.w);
                const u64 n01 = pack2(N.x, N.y), n23 = pack2(N.z, N.w);
#pragma unroll
                for (int q = 0; q < QPT; ++q) {
                    const u64 qx = pack2(ax[q], ax[q]);
                    const u64 qy = pack2(ay[q], ay[q]);
                    const u64 qz = pack2(az[q], az[q]);
                    u64 sa = fma2(qx, x01, n01), sb = fma2(qx, x23, n23);
                    sa = fma2(qy, y01, sa); sb = fma2(qy, y23, sb);
                    sa = fma2(qz, z01, sa); sb = fma2(qz, z23, sb);
                    float s0, s1, s2, s3;
                    unpack2(sa, s0, s1);
                    unpack2(sb, s2, s3);
                    gm[q] = min3(gm[q], s0, s1);
                    gm[q] = min3(gm[q], s2, s3);
                }
            }
            const int g = gb + gi;
            bool hit = false;
#pragma unroll
            for (int q = 0; q < QPT; ++q) hit |= gm[q] < m3[q];
            if (hit) {
#pragma unroll
                for (int q = 0; q < QPT; ++q) {
                    const float v = gm[q];
                    const bool lt1 = v < m1[q], lt2 = v < m2[q];
                    m3[q] = lt2 ? m2[q] : fminf(m3[q], v);
                    g2[q] = lt1 ? g1[q] : (lt2 ? g : g2[q]);
                    m2[q] = lt1 ? m1[q] : (lt2 ? v : m2[q]);
                    g1[q] = lt1 ? g : g1[q];
                    m1[q] = lt1 ? v : m1[q];
                }
            }
            __syncwarp(); // every lane is done with `cur` before it is refilled two groups later
        }
        __syncthreads(); // every warp is done with stage s
        if (tid == 0 && tile + kStages < n_tiles) issue(tile + kStages);
    }

    const size_t row = (size_t)split * d.n_stride;
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        const int i = qi[q];
        if (i >= 0) {
            d.pm1[row + i] = m1[q];
            d.pm2[row + i] = m2[q];
            d.pm3[row + i] = m3[q];
            d.pg[row + i] = g1[q];
            d.pg2[row + i] = g2[q];
        }
    }
}

template <int QPT>
static void launch_warp(dim3 grid, const RegDesc *descs, IcpState *states, int splits, int pass, cudaStream_t s)
{
    launch_pdl<true>(nn_partial_warp_kernel<QPT>, grid, dim3(kNnThreads), 0, s, descs, states, splits, pass);
}

template <int QPT>
static void launch_centred(dim3 grid, const RegDesc *descs, IcpState *states, int splits, int pass, cudaStream_t s)
{
    launch_pdl<true>(nn_partial_centred_kernel<QPT>, grid, dim3(kNnThreads), 0, s, descs, states, splits, pass);
}

void launch_nn_partial(const RegDesc *descs, IcpState *states, int batch, int max_n, int qpt, int splits, int pass,
                       int filter, cudaStream_t s)
{
    dim3 block(kNnThreads);
    dim3 grid((max_n + kNnThreads * qpt - 1) / (kNnThreads * qpt), splits, batch);
    if (filter == kFilterWarp) {
        switch (qpt) {
        case 16: launch_warp<16>(grid, descs, states, splits, pass, s); break;
        case 12: launch_warp<12>(grid, descs, states, splits, pass, s); break;
        case 8: launch_warp<8>(grid, descs, states, splits, pass, s); break;
        case 4: launch_warp<4>(grid, descs, states, splits, pass, s); break;
        default: launch_warp<2>(grid, descs, states, splits, pass, s); break;
        }
    } else if (filter == kFilterCentred) {
        switch (qpt) {
        case 16: launch_centred<16>(grid, descs, states, splits, pass, s); break;
        case 12: launch_centred<12>(grid, descs, states, splits, pass, s); break;
        case 8: launch_centred<8>(grid, descs, states, splits, pass, s); break;
        case 4: launch_centred<4>(grid, descs, states, splits, pass, s); break;
        default: launch_centred<2>(grid, descs, states, splits, pass, s); break;
        }
    } else {
        switch (qpt) {
        case 8: launch_pdl<true>(nn_partial_kernel<8>, grid, block, 0, s, descs, states, splits, pass); break;
        case 4: launch_pdl<true>(nn_partial_kernel<4>, grid, block, 0, s, descs, states, splits, pass); break;
        default: launch_pdl<true>(nn_partial_kernel<2>, grid, block, 0, s, descs, states, splits, pass); break;
        }
    }
}

// --------------------------------------------------------------------------
// canonical FP64 block-ordered reduction pieces (CANON-3)
// --------------------------------------------------------------------------
__device__ __forceinline__ double warp_fold(double v)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_down_sync(0xffffffffu, v, off);
    return v;
}

// The value warp_fold leaves in lane 0, computed by ONE thread from the 32 inputs x[0..31].  The shuffle steps give
// v16[l] = x[l] + x[l+16], v8[l] = v16[l] + v16[l+8], ..., v1[0] = v2[0] + v2[1]; with F<l, 32> = x[l] and
// F<l, s> = F<l, 2s> + F<l+s, 2s> that is F<0, 1>: the same additions in the same association, evaluated depth first
// (five live partial sums).
template <int L, int S>
__device__ __forceinline__ double fold_tree(const double *x)
{
    if constexpr (S == 32) return x[L];
    else return fold_tree<L, 2 * S>(x) + fold_tree<L + S, 2 * S>(x);
}

// --------------------------------------------------------------------------
// 3x3 SVD (one-sided Jacobi, double) -- stands in for cv::SVD, icp.cpp:215.
// Same operation sequence as the host-side oracle so results are bit equal.
// --------------------------------------------------------------------------
// One Jacobi rotation on the column pair (P, Q); compile-time indices keep A and V in registers.
template <int P, int Q>
__device__ __forceinline__ int jacobi_pair(double (&A)[9], double (&V)[9])
{
    const double eps = 2.220446049250313e-16;
    double alpha = (A[P] * A[P] + A[3 + P] * A[3 + P]) + A[6 + P] * A[6 + P];
    double beta = (A[Q] * A[Q] + A[3 + Q] * A[3 + Q]) + A[6 + Q] * A[6 + Q];
    double gamma = (A[P] * A[Q] + A[3 + P] * A[3 + Q]) + A[6 + P] * A[6 + Q];
    if (fabs(gamma) <= eps * sqrt(alpha * beta)) return 0;
    double zeta = (beta - alpha) / (2.0 * gamma);
    double az = fabs(zeta);
    double t = 1.0 / (az + sqrt(1.0 + zeta * zeta));
    if (zeta < 0.0) t = -t;
    double c = 1.0 / sqrt(1.0 + t * t);
    double sn = c * t;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double a = A[3 * r + P], b = A[3 * r + Q];
        A[3 * r + P] = c * a - sn * b;
        A[3 * r + Q] = sn * a + c * b;
        double va = V[3 * r + P], vb = V[3 * r + Q];
        V[3 * r + P] = c * va - sn * vb;
        V[3 * r + Q] = sn * va + c * vb;
    }
    return 1;
}

template <int I, int J>
__device__ __forceinline__ void swap_cols(double (&A)[9], double (&V)[9], double (&w)[3])
{
    double tw = w[I]; w[I] = w[J]; w[J] = tw;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double ta = A[3 * r + I]; A[3 * r + I] = A[3 * r + J]; A[3 * r + J] = ta;
        double tv = V[3 * r + I]; V[3 * r + I] = V[3 * r + J]; V[3 * r + J] = tv;
    }
}

#ifdef ICPB_SOLVE_CLOCKS
__device__ long long g_svd_clocks;
#endif
__device__ void svd3(const double *Ain, double (&U)[9], double (&w)[3], double (&Vt)[9])
{
    double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
    for (int k = 0; k < 9; ++k) A[k] = Ain[k];
    for (int sweep = 0; sweep < 30; ++sweep) {
        int changed = jacobi_pair<0, 1>(A, V);
        changed |= jacobi_pair<0, 2>(A, V);
        changed |= jacobi_pair<1, 2>(A, V);
        if (!changed) break;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) w[k] = sqrt((A[k] * A[k] + A[3 + k] * A[3 + k]) + A[6 + k] * A[6 + k]);
    // the oracle's selection sort (descending, first maximum wins), spelled out with constant indices
    {
        int best = 0;
        if (w[1] > w[best]) best = 1;
        if (w[2] > w[best]) best = 2;
        if (best == 1) swap_cols<0, 1>(A, V, w);
        else if (best == 2) swap_cols<0, 2>(A, V, w);
        if (w[2] > w[1]) swap_cols<1, 2>(A, V, w);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (w[k] > 0.0) {
#pragma unroll
            for (int r = 0; r < 3; ++r) U[3 * r + k] = A[3 * r + k] / w[k];
        } else {
#pragma unroll
            for (int r = 0; r < 3; ++r) U[3 * r + k] = 0.0;
        }
    }
    if (!(w[2] > 0.0) && w[1] > 0.0) {
        U[2] = U[3] * U[7] - U[6] * U[4];
        U[5] = U[6] * U[1] - U[0] * U[7];
        U[8] = U[0] * U[4] - U[3] * U[1];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Vt[3 * i + j] = V[3 * j + i];
}

// cv::Mat operator* on 3x3 CV_32F (icp.cpp:218,231,237): float, no FMA.
__device__ void gemm33f(const float *A, const float *B, float *C)
{
    float T[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[3 * i + j] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j]) + A[3 * i + 2] * B[6 + j];
    for (int k = 0; k < 9; ++k) C[k] = T[k];
}
// cv::determinant on 3x3 CV_32F (icp.cpp:220): evaluated in double.
__device__ double det33f(const float *m)
{
    double m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7],
           m22 = m[8];
    return m00 * (m11 * m22 - m12 * m21) - m01 * (m10 * m22 - m12 * m20) + m02 * (m10 * m21 - m11 * m20);
}
// Mat::inv() on 3x3 CV_32F (icp.cpp:235): adjugate in double times 1/det.
__device__ void inv33f(const float *Sf, float *D)
{
    double d = det33f(Sf);
    if (d == 0.0) {
        for (int k = 0; k < 9; ++k) D[k] = 0.f;
        return;
    }
    d = 1.0 / d;
    double S[9];
    for (int i = 0; i < 9; ++i) S[i] = Sf[i];
    float T[9];
    T[0] = (float)((S[4] * S[8] - S[5] * S[7]) * d);
    T[1] = (float)((S[2] * S[7] - S[1] * S[8]) * d);
    T[2] = (float)((S[1] * S[5] - S[2] * S[4]) * d);
    T[3] = (float)((S[5] * S[6] - S[3] * S[8]) * d);
    T[4] = (float)((S[0] * S[8] - S[2] * S[6]) * d);
    T[5] = (float)((S[2] * S[3] - S[0] * S[5]) * d);
    T[6] = (float)((S[3] * S[7] - S[4] * S[6]) * d);
    T[7] = (float)((S[1] * S[6] - S[0] * S[7]) * d);
    T[8] = (float)((S[0] * S[4] - S[1] * S[3]) * d);
    for (int k = 0; k < 9; ++k) D[k] = T[k];
}

// meanSquareError, icp.cpp:622-638: (sum(errors)/n)^2.
__device__ float mse_from_sums(const double *sums)
{
    if (!(sums[19] > 0.0)) return 0.f;
    float e = (float)(sums[15] / sums[19]);
    return (float)((double)e * (double)e);
}

// The body of the while loop of icp.cpp:155-258 for one iteration, minus the
// association itself.  Runs in one thread.
__device__ void solve_step_local(IcpState *st, const IcpParamsDev *prm, const double *sums, int pass, float *mlog);

// Works on a register / local copy of the state: one global read and one global write of the block.
__device__ __noinline__ void solve_step(IcpState *gst, const IcpParamsDev *gprm, const double *sums, int pass, float *mlog)
{
    IcpState st = *gst;
    IcpParamsDev prm = *gprm;
    double lsums[kTerms];
#pragma unroll
    for (int k = 0; k < kTerms; ++k) lsums[k] = sums[k];
    solve_step_local(&st, &prm, lsums, pass, mlog);
    *gst = st;
}

__device__ __forceinline__ void solve_step_local(IcpState *st, const IcpParamsDev *prm, const double *sums, int pass,
                                                 float *mlog)
{
    st->passes = pass + 1;
    st->last_buf = (pass + 1) & 1;
    st->apply = 0;
    const float mse = mse_from_sums(sums);
    const int n_assoc = (int)sums[19];
    st->mse = mse;
    st->n_assoc = n_assoc;
    const int i = st->iterations;
    if (!(mse > prm->threshold && i < prm->max_iterations)) { // icp.cpp:155
        st->done = 1;
        return;
    }
    if (n_assoc < 3) { // icp.cpp:163-182
        st->iterations = prm->max_iterations;
        for (int k = 0; k < 3; ++k) {
            st->offset[k] = -prm->last_translation[k];
            st->tf[k] = prm->last_translation[k];
            st->Pt[k] += (double)prm->last_translation[k];
        }
        st->pending_translate = 1;
        st->small_exit = 1;
        st->done = 1;
        return;
    }
    float Rf[9], tf[3];
    double U[9], w[3], Vt[9], Rd[9];
    if (prm->solve_mode == ICPB_SOLVE_REFERENCE) {
#ifdef ICPB_SOLVE_CLOCKS
        const long long cs = clock64();
#endif
        svd3(&sums[6], U, w, Vt); // icp.cpp:212-215, M = sum b a^T (uncentred)
#ifdef ICPB_SOLVE_CLOCKS
        g_svd_clocks = clock64() - cs;
#endif
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                Rd[3 * r + c] = (Vt[r] * U[3 * c] + Vt[3 + r] * U[3 * c + 1]) + Vt[6 + r] * U[3 * c + 2]; // :218
        float R[9];
        for (int k = 0; k < 9; ++k) R[k] = (float)Rd[k];
        if (det33f(R) < 0) { R[2] *= -1; R[5] *= -1; R[8] *= -1; } // :220-223
        if (i == 0) { for (int k = 0; k < 9; ++k) st->rigid[k] = R[k]; } // :227-229
        else gemm33f(R, st->rigid, st->rigid);                           // :230-233
        inv33f(R, Rf);                                                   // :235
        gemm33f(st->camR, Rf, st->camR);                                 // :237
        for (int k = 0; k < 3; ++k) {
            st->offset[k] = (float)(sums[16 + k] / sums[19]);            // :240, :314-344
            tf[k] = -st->offset[k];                                      // :245
            st->camP[k] -= st->offset[k];                                // :246
        }
    } else {
        // rigid_transform_3D.py:14-37 with A = data, B = matches
        double cnt = sums[19];
        double cA[3] = {sums[0] / cnt, sums[1] / cnt, sums[2] / cnt};
        double cB[3] = {sums[3] / cnt, sums[4] / cnt, sums[5] / cnt};
        double H[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) H[3 * r + c] = sums[6 + 3 * c + r] - (cnt * cA[r]) * cB[c];
        svd3(H, U, w, Vt);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    Rd[3 * r + c] = (Vt[r] * U[3 * c] + Vt[3 + r] * U[3 * c + 1]) + Vt[6 + r] * U[3 * c + 2];
            double det = Rd[0] * (Rd[4] * Rd[8] - Rd[5] * Rd[7]) - Rd[1] * (Rd[3] * Rd[8] - Rd[5] * Rd[6]) +
                         Rd[2] * (Rd[3] * Rd[7] - Rd[4] * Rd[6]);
            if (p == 0 && det < 0) { Vt[6] *= -1; Vt[7] *= -1; Vt[8] *= -1; }
            else break;
        }
        for (int k = 0; k < 9; ++k) Rf[k] = (float)Rd[k];
        for (int r = 0; r < 3; ++r) {
            tf[r] = (float)(cB[r] - ((Rd[3 * r] * cA[0] + Rd[3 * r + 1] * cA[1]) + Rd[3 * r + 2] * cA[2]));
            st->offset[r] = -tf[r];
        }
    }
    // composed pose, in double, from the float motion actually applied
    double NR[9], Nt[3];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
            NR[3 * r + c] = ((double)Rf[3 * r] * st->PR[c] + (double)Rf[3 * r + 1] * st->PR[3 + c]) +
                            (double)Rf[3 * r + 2] * st->PR[6 + c];
        Nt[r] = (((double)Rf[3 * r] * st->Pt[0] + (double)Rf[3 * r + 1] * st->Pt[1]) + (double)Rf[3 * r + 2] * st->Pt[2]) +
                (double)tf[r];
    }
    for (int k = 0; k < 9; ++k) { st->PR[k] = NR[k]; st->Rf[k] = Rf[k]; }
    for (int k = 0; k < 3; ++k) { st->Pt[k] = Nt[k]; st->tf[k] = tf[k]; }
    if (mlog) { // key-point variant: the carried cloud replays every motion after the loop
        for (int k = 0; k < 9; ++k) mlog[12 * i + k] = Rf[k];
        for (int k = 0; k < 3; ++k) mlog[12 * i + 9 + k] = tf[k];
        st->n_log = i + 1;
    }
    st->apply = 1;
    st->iterations = i + 1;
}

struct FinalizeShared;
__device__ __noinline__ void finalize_last_cta(const double *chunk_sums, float *mlog, IcpState *st,
                                               const IcpParamsDev *__restrict__ prm, int pass, int nchunks, int tid,
                                               FinalizeShared &sh);

// Everything after the association of a chunk is known: per-query outputs, CANON-3 sums, and in the last CTA of the
// registration the second summation level and the solve.  Shared by nn_finalize_kernel and nn_finalize_coop_kernel.
struct FinalizeShared {
    double w[kChunk / 32][kTerms];
    double tot[kTerms];
    int last;
};

template <bool kTicket = true>
__device__ __forceinline__ void finalize_tail(const RegDesc &d, IcpState *st, const IcpParamsDev *__restrict__ prm, int pass,
                                              int chunk, int nchunks, int tid, bool valid, int i, int n, const float4 a,
                                              int best_i, float best_d, const float4 best_b, bool have_b, int n_amb,
                                              FinalizeShared &sh, double *fold_buf = nullptr)
{
    double (&s_w)[kChunk / 32][kTerms] = sh.w;
    int &s_last = sh.last;
    if (valid) {
        d.idx[i] = best_i;
        d.dist[i] = best_d;
        if (d.idx_trace) d.idx_trace[(size_t)pass * n + i] = best_i;
        if (d.dist_trace) d.dist_trace[(size_t)pass * n + i] = best_d;
        if (d.rej_flag) { // icp.cpp:507-509: the rejects of every pass accumulate
            const bool rejected = !(best_d < prm->max_nn_distance);
            d.rej_flag[(size_t)pass * n + i] = rejected ? 1 : 0;
            if (rejected) d.rej_pts[(size_t)pass * n + i] = a;
        }
    }

    // ---- association sums (CANON-3 level 1)
    // one term at a time (value -> fold -> shared memory): twenty live doubles would cost forty registers per thread
    const bool accepted = valid && (best_d < prm->max_nn_distance); // icp.cpp:553
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (accepted) b = have_b ? best_b : d.tgt[best_i];
    auto term = [&](int k) -> double {
        if (!accepted) return 0.0;
        switch (k) {
        case 0: return a.x; case 1: return a.y; case 2: return a.z;
        case 3: return b.x; case 4: return b.y; case 5: return b.z;
        case 6: return (double)b.x * (double)a.x; case 7: return (double)b.x * (double)a.y; case 8: return (double)b.x * (double)a.z;
        case 9: return (double)b.y * (double)a.x; case 10: return (double)b.y * (double)a.y; case 11: return (double)b.y * (double)a.z;
        case 12: return (double)b.z * (double)a.x; case 13: return (double)b.z * (double)a.y; case 14: return (double)b.z * (double)a.z;
        case 15: return best_d;
        case 16: return (double)(a.x - b.x); case 17: return (double)(a.y - b.y); case 18: return (double)(a.z - b.z);
        default: return 1.0;
        }
    };
    if (fold_buf) {
        // The same fold tree (v[l] += v[l + off], off = 16, 8, 4, 2, 1) evaluated from shared memory: every thread
        // stores its twenty terms, then one thread per (group of 32, term) adds the 32 values depth first in exactly the
        // shuffle tree's association -- 320 warp-level shared-memory operations per chunk instead of 1,600 shuffles
        // (the shuffle version kept the LSU pipe 60 % busy and cost 46 us per pass at full resolution).
        // layout [term][group][33]: stores run along the lanes, loads see a two-way bank conflict at worst
#pragma unroll
        for (int k = 0; k < kTerms; ++k) fold_buf[(k * (kChunk / 32) + (tid >> 5)) * 33 + (tid & 31)] = term(k);
        __syncthreads();
        if (tid < kTerms * (kChunk / 32)) {
            const int g = tid % (kChunk / 32), k = tid / (kChunk / 32);
            const double *x = fold_buf + (k * (kChunk / 32) + g) * 33;
            s_w[g][k] = fold_tree<0, 1>(x);
        }
    } else {
#pragma unroll
        for (int k = 0; k < kTerms; ++k) {
            const double v = warp_fold(term(k));
            if ((tid & 31) == 0) s_w[tid >> 5][k] = v;
        }
    }
    __syncthreads();
    if (tid < kTerms) {
        double s = s_w[0][tid];
        for (int wv = 1; wv < kChunk / 32; ++wv) s = s + s_w[wv][tid];
        d.chunk_sums[(size_t)tid * nchunks + chunk] = s; // term-major: level 2 reads every term coalesced
    }
    if (!kTicket) return; // the second level and the solve run as their own one-CTA kernel (nn_solve_kernel)
    // the barrier orders the chunk sums written by threads 0..19 before thread 0's fence, and the fence (cumulative)
    // before its ticket: one thread waits for the memory system instead of 256
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (n_amb) atomicAdd(&st->rescans, n_amb);
        unsigned int ticket = atomicAdd(&st->block_counter, 1u);
        s_last = (ticket == (unsigned int)(nchunks - 1));
    }
    __syncthreads();
    if (!s_last) return;

    // ---- last CTA of this registration: CANON-3 level 2, then the solve (out of line: its forty accumulator registers
    //      and the solve's must not set the register count of the 1,100 CTAs that never get here)
    finalize_last_cta(d.chunk_sums, d.mlog, st, prm, pass, nchunks, tid, sh);
}

__device__ __forceinline__ void finalize_last_cta_body(const double *chunk_sums, float *mlog, IcpState *st,
                                                       const IcpParamsDev *__restrict__ prm, int pass, int nchunks, int tid,
                                                       FinalizeShared &sh, bool rearm_ticket)
{
    double (&s_w)[kChunk / 32][kTerms] = sh.w;
    double (&s_tot)[kTerms] = sh.tot;
#ifdef ICPB_SOLVE_CLOCKS
    const long long sh_clock0 = clock64();
#endif
    __threadfence();
    {
        double acc[kTerms];
#pragma unroll
        for (int k = 0; k < kTerms; ++k) acc[k] = 0.0;
        for (int c = tid; c < nchunks; c += kChunk) {
            double v[kTerms];
#pragma unroll
            for (int k = 0; k < kTerms; ++k) v[k] = __ldcg(&chunk_sums[(size_t)k * nchunks + c]); // issued together, coalesced
#pragma unroll
            for (int k = 0; k < kTerms; ++k) acc[k] = acc[k] + v[k];
        }
#pragma unroll
        for (int k = 0; k < kTerms; ++k) {
            double v = warp_fold(acc[k]);
            if ((tid & 31) == 0) s_w[tid >> 5][k] = v;
        }
    }
    __syncthreads();
    if (tid < kTerms) {
        double s = s_w[0][tid];
        for (int wv = 1; wv < kChunk / 32; ++wv) s = s + s_w[wv][tid];
        s_tot[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
        if (rearm_ticket) st->block_counter = 0; // re-armed before solve_step copies the state
#ifdef ICPB_SOLVE_CLOCKS
        const long long c1 = clock64();
#endif
        solve_step(st, prm, s_tot, pass, mlog);
#ifdef ICPB_SOLVE_CLOCKS
        const long long c2 = clock64();
        if (pass == 5) printf("solve clocks: level2 %lld  solve %lld of which svd3 %lld\n", c1 - sh_clock0, c2 - c1, g_svd_clocks);
#endif
    }
}

__device__ __noinline__ void finalize_last_cta(const double *chunk_sums, float *mlog, IcpState *st,
                                               const IcpParamsDev *__restrict__ prm, int pass, int nchunks, int tid,
                                               FinalizeShared &sh)
{
    finalize_last_cta_body(chunk_sums, mlog, st, prm, pass, nchunks, tid, sh, true);
}

// --------------------------------------------------------------------------
// nn_finalize: exact resolution, association sums, solve
// --------------------------------------------------------------------------
// Launched with kChunk threads (grid mode) or 2 * kChunk (brute-force modes): in the latter the group selection and
// the exact re-evaluation of a query are shared by a PAIR of adjacent threads (each takes every second split and one
// half of each candidate group; shuffles merge the halves), which halves the stage's serial length on the few CTAs a
// small cloud gives (40 for 10k points).  From the sums on, threads 0 .. kChunk-1 own one query each, as before.
__global__ void __launch_bounds__(2 * kChunk) nn_finalize_kernel(const RegDesc *__restrict__ descs, IcpState *states,
                                                                 const IcpParamsDev *__restrict__ prm, int splits,
                                                                 int pass, int filter)
{
    pdl_enter(); // icpb_internal.h: the grid before this one is complete from here on
    IcpState *st = states + blockIdx.z;  // == d.st; see nn_partial*
    const RegDesc d = descs[blockIdx.z]; // by value: no pointer reloads after stores
    if (st->done) return;
    const int n = d.n, m = d.m;
    const int chunk = blockIdx.x;
    if (chunk * kChunk >= n) return;
    const int nchunks = (n + kChunk - 1) / kChunk;
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const bool paired = nthr == 2 * kChunk;
    // exact stage: query slot and half; later stages: slot = tid for the first kChunk threads
    const int slot = paired ? (tid >> 1) : tid;
    const int half = paired ? (tid & 1) : 0;
    int i = chunk * kChunk + slot;
    bool valid = i < n;

    __shared__ float4 s_pts[kChunk];
    __shared__ float4 s_bb[kChunk];
    __shared__ float s_bd[kChunk];
    __shared__ int s_bi[kChunk];
    __shared__ int s_list[kChunk];
    __shared__ int s_cnt;
    __shared__ float s_rd[2 * kChunk / 32];
    __shared__ int s_ri[2 * kChunk / 32];
    __shared__ FinalizeShared s_fin;

    if (tid == 0) s_cnt = 0;
    const float4 *cur = d.D[(pass + 1) & 1];
    float4 a = valid ? cur[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (half == 0) s_pts[slot] = a;
    __syncthreads();

    int best_i = 0;
    float best_d = 0.f;
    float4 best_b = make_float4(0.f, 0.f, 0.f, 0.f);
    bool have_b = false;
    int n_amb = 0;
    if (splits < 0) {
        // ICPB_NN_GRID, cooperative search: the neighbour's coordinates and index arrive in one coalesced 16-byte
        // record per query; the distance is re-evaluated from them (same inputs, same arithmetic: same bits)
        if (valid) {
            const float4 nb = __ldcg(&d.gnb[i]);
            best_i = __float_as_int(nb.w);
            best_d = CUDART_INF_F;
            if (best_i >= 0) {
                best_b = nb; have_b = true;
                best_d = exact_distance(a.x, a.y, a.z, nb.x, nb.y, nb.z);
            }
        }
    } else if (splits == 0) {
        // ICPB_NN_GRID: nn_grid_kernel already resolved (idx, dist) exactly
        if (valid) { best_i = d.idx[i]; best_d = d.dist[i]; }
    } else {
        // ---- combine the per-split records into the three best group minima; id -1 = "some group other than
        //      the ones named" (a bound without an address).  Only the three splits with the smallest best value
        //      can contribute: the best split its three records, the second its first two, the third its first.
        float m1 = CUDART_INF_F, m2 = CUDART_INF_F, m3 = CUDART_INF_F;
        int g = 0, g2 = -1;
        auto insert = [&](float v, int id) {
            if (v < m1) { m3 = m2; m2 = m1; g2 = g; m1 = v; g = id; }
            else if (v < m2) { m3 = m2; m2 = v; g2 = id; }
            else if (v < m3) m3 = v;
        };
        if (valid) {
            float a1 = CUDART_INF_F, a2 = CUDART_INF_F, a3 = CUDART_INF_F;
            int s1 = -1, s2 = -1, s3 = -1;
            constexpr int kB = 16; // 30 splits at 10k points: one round of independent loads per half
            const int sstep = paired ? 2 : 1;
            for (int s0 = half; s0 < splits; s0 += kB * sstep) {
                float p1[kB];
    #pragma unroll
                for (int k = 0; k < kB; ++k) // independent loads, issued together
                    p1[k] = __ldcg(&d.pm1[(size_t)min(s0 + k * sstep, splits - 1) * d.n_stride + i]);
    #pragma unroll
                for (int k = 0; k < kB; ++k) {
                    const float v = p1[k];
                    const int sp = s0 + k * sstep;
                    if (sp < splits) {
                        if (v < a1 || s1 < 0) { a3 = a2; s3 = s2; a2 = a1; s2 = s1; a1 = v; s1 = sp; }
                        else if (v < a2 || s2 < 0) { a3 = a2; s3 = s2; a2 = v; s2 = sp; }
                        else if (v < a3 || s3 < 0) { a3 = v; s3 = sp; }
                    }
                }
            }
            if (paired) {
                // the partner scanned the other splits: fold its three records in, then both halves adopt half 0's
                // list so that the pair takes identical decisions (ties between splits may be ordered either way; the
                // exact stage settles them)
                const unsigned pm = __activemask();
                const float b1 = __shfl_xor_sync(pm, a1, 1), b2 = __shfl_xor_sync(pm, a2, 1), b3 = __shfl_xor_sync(pm, a3, 1);
                const int t1 = __shfl_xor_sync(pm, s1, 1), t2 = __shfl_xor_sync(pm, s2, 1), t3 = __shfl_xor_sync(pm, s3, 1);
                auto take = [&](float v, int sp) {
                    if (sp < 0) return;
                    if (v < a1 || s1 < 0) { a3 = a2; s3 = s2; a2 = a1; s2 = s1; a1 = v; s1 = sp; }
                    else if (v < a2 || s2 < 0) { a3 = a2; s3 = s2; a2 = v; s2 = sp; }
                    else if (v < a3 || s3 < 0) { a3 = v; s3 = sp; }
                };
                take(b1, t1); take(b2, t2); take(b3, t3);
                const int src = (threadIdx.x & 31) & ~1;
                a1 = __shfl_sync(pm, a1, src); a2 = __shfl_sync(pm, a2, src); a3 = __shfl_sync(pm, a3, src);
                s1 = __shfl_sync(pm, s1, src); s2 = __shfl_sync(pm, s2, src); s3 = __shfl_sync(pm, s3, src);
            }
            const bool top3 = filter != kFilterDirect;
            const size_t o1 = (size_t)max(s1, 0) * d.n_stride + i;
            const size_t o2 = (size_t)max(s2, 0) * d.n_stride + i;
            const size_t o3 = (size_t)max(s3, 0) * d.n_stride + i;
            const int ga = __ldcg(&d.pg[o1]);
            const float p2a = __ldcg(&d.pm2[o1]);
            const int g2a = top3 ? __ldcg(&d.pg2[o1]) : -1;
            const float p3a = top3 ? __ldcg(&d.pm3[o1]) : CUDART_INF_F;
            const int gb = __ldcg(&d.pg[o2]);
            const float p2b = __ldcg(&d.pm2[o2]);
            const int g2b = top3 ? __ldcg(&d.pg2[o2]) : -1;
            const int gc = __ldcg(&d.pg[o3]);
            insert(a1, ga);
            insert(p2a, g2a);
            insert(p3a, -1);
            if (s2 >= 0) { insert(a2, gb); insert(p2b, g2b); }
            if (s3 >= 0) insert(a3, gc);
            if (m1 == CUDART_INF_F) g = ga; // nothing finite anywhere (NaN / overflow inputs): keep a valid address
        }
        // is every target outside the best group (or the best two) provably farther, in the reference's
        // arithmetic, than the best target?
        float band;
        if (filter != kFilterDirect) {
            // m1..m3 are W = |a-t|^2 - A; bound derived at nn_partial_centred_kernel
            const float A = valid ? __ldcg(&d.pa[i]) * 1.000001f : 0.f;
            const float X = fmaxf(m1 + A, 0.f);
            band = (A * kBandCentredA + X * kBandCentredX) + kBandAbs;
        } else {
            band = m1 * (kBandRel - 1.0f) + kBandAbs;
        }
        const float lim = m1 + band;
        const bool one = m2 > lim;                                  // the best group alone decides
        const bool two = !one && g2 >= 0 && g >= 0 && m3 > lim;     // the best two groups decide
        const bool ambiguous = valid && !(one || two);
        if (valid && !ambiguous) {
            // exact re-evaluation (reference arithmetic, ascending index, strict <) of the deciding group(s); a pair
            // splits every group into its lower and upper half and merges on (distance, index)
            const int ga = two ? min(g, g2) : g;
            const int gbb = two ? max(g, g2) : -1;
            const int span = paired ? kGroup / 2 : kGroup;
            float bd = CUDART_INF_F;
            int bi = 0x7fffffff;
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int grp = 0; grp < 2; ++grp) {
                const int gsel = grp == 0 ? ga : gbb;
                if (gsel < 0) break;
                const int t0 = gsel * kGroup + half * span;
    #pragma unroll 16
                for (int k = 0; k < span; ++k) {
                    const int t = t0 + k;
                    const float4 b = d.tgt[min(t, m - 1)];
                    const float dd = exact_distance(a.x, a.y, a.z, b.x, b.y, b.z);
                    if (t < m && dd < bd) { bd = dd; bi = t; bb = b; }
                }
            }
            if (paired) {
                const unsigned pm = __activemask();
                const float od = __shfl_xor_sync(pm, bd, 1);
                const int oi = __shfl_xor_sync(pm, bi, 1);
                const float ox = __shfl_xor_sync(pm, bb.x, 1), oy = __shfl_xor_sync(pm, bb.y, 1);
                const float oz = __shfl_xor_sync(pm, bb.z, 1), ow = __shfl_xor_sync(pm, bb.w, 1);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; bb = make_float4(ox, oy, oz, ow); }
            }
            if (bi == 0x7fffffff) { // no finite distance in the group (NaN inputs): keep its first target, like a strict-< scan
                bi = ga * kGroup;
                bb = d.tgt[min(bi, m - 1)];
                bd = exact_distance(a.x, a.y, a.z, bb.x, bb.y, bb.z);
            }
            best_d = bd; best_i = bi; best_b = bb; have_b = true; // the winner's coordinates feed the sums: no second gather
        }
        if (ambiguous && half == 0) {
            const int e = atomicAdd(&s_cnt, 1);
            s_list[e] = slot;
        }
        if (half == 0) { s_bd[slot] = best_d; s_bi[slot] = have_b ? best_i : -1 - best_i; s_bb[slot] = best_b; }
        __syncthreads();

        // ---- near-tie queries: CTA-cooperative full scan in exact arithmetic (all threads of the CTA)
        n_amb = s_cnt;
        for (int e = 0; e < n_amb; ++e) {
            const int owner = s_list[e];
            const float4 q = s_pts[owner];
            float bd = CUDART_INF_F;
            int bi = 0x7fffffff;
            for (int t = tid; t < m; t += nthr) {
                float4 b = d.tgt[t];
                float dd = exact_distance(q.x, q.y, q.z, b.x, b.y, b.z);
                if (dd < bd) { bd = dd; bi = t; }
            }
    #pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                float od = __shfl_xor_sync(0xffffffffu, bd, off);
                int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
            if ((tid & 31) == 0) { s_rd[tid >> 5] = bd; s_ri[tid >> 5] = bi; }
            __syncthreads();
            if (tid == 0) {
                bd = s_rd[0]; bi = s_ri[0];
                for (int wv = 1; wv < nthr / 32; ++wv) {
                    float od = s_rd[wv];
                    int oi = s_ri[wv];
                    if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
                }
                if (bi == 0x7fffffff) { // every distance was NaN: the reference keeps target[0]
                    float4 b = d.tgt[0];
                    bi = 0;
                    bd = exact_distance(q.x, q.y, q.z, b.x, b.y, b.z);
                }
                s_bd[owner] = bd;
                s_bi[owner] = -1 - bi; // negative: coordinates not staged, gathered later
            }
            __syncthreads();
        }
        // ---- from here on one thread per query
        if (tid >= kChunk) return;
        i = chunk * kChunk + tid;
        valid = i < n;
        a = s_pts[tid];
        best_d = s_bd[tid];
        {
            const int enc = s_bi[tid];
            have_b = enc >= 0;
            best_i = have_b ? enc : -1 - enc;
            best_b = s_bb[tid];
        }
    }

    finalize_tail(d, st, prm, pass, chunk, nchunks, tid, valid, i, n, a, best_i, best_d, best_b, have_b, n_amb, s_fin);
}

// ICPB_NN_GRID with the cooperative search: the association arrives as one 16-byte record per query (the neighbour's
// coordinates and index, grid.cu); nothing to resolve, so the kernel is the tail alone -- 256 threads, few registers,
// eight CTAs per SM instead of the two of the general kernel (whose 59 us per pass were a third of a registration).
// q_cur / q_nb / q_n (count == 1: known to the host): the kernel's first loads go out at once instead of behind two
// dependent reads of the descriptor -- the kernel is a chain of memory round trips, not work.
__global__ void __launch_bounds__(kChunk, 5) nn_finalize_coop_kernel(const RegDesc *__restrict__ descs, IcpState *states,
                                                                  const IcpParamsDev *__restrict__ prm, int pass,
                                                                  const float4 *__restrict__ q_cur,
                                                                  const float4 *__restrict__ q_nb, int q_n)
{
    pdl_enter(); // icpb_internal.h: the grid before this one is complete from here on
    IcpState *st = states + blockIdx.z;
    const int chunk = blockIdx.x;
    float4 a_early = make_float4(0.f, 0.f, 0.f, 0.f), nb_early = a_early;
    if (q_cur && chunk * kChunk + (int)threadIdx.x < q_n) {
        a_early = q_cur[chunk * kChunk + threadIdx.x];
        nb_early = __ldcg(&q_nb[chunk * kChunk + threadIdx.x]);
    }
    if (st->done) return;
    const RegDesc &d = descs[blockIdx.z];
    const int n = q_cur ? q_n : d.n;
    if (chunk * kChunk >= n) return;
    if (d.gord) {
        // work order of the next pass's search (grid.cu): this pass's cost classes, heaviest first, as one flat list --
        // one entry per thread of the first CTAs, its two round trips under the loads above
        const int e = chunk * kChunk + (int)threadIdx.x;
        if (e < ((n + 31) >> 5)) {
            const int stride = d.n_stride >> 5;
            const int *cnt = d.gord_count + pass * kOrderBins;
            const int *lists = d.gord + (pass & 1) * kOrderBins * stride;
            int before = 0, src = -1;
#pragma unroll
            for (int b = 0; b < kOrderBins; ++b) {
                const int c = __ldcg(&cnt[b]);
                if (e >= before && e < before + c) src = b * stride + (e - before);
                before += c;
            }
            if (src >= 0) d.gord_flat[e] = __ldcg(&lists[src]);
        }
    }
    const int nchunks = (n + kChunk - 1) / kChunk;
    const int tid = threadIdx.x;
    const int i = chunk * kChunk + tid;
    const bool valid = i < n;
    __shared__ FinalizeShared s_fin;
    __shared__ double s_fold[kTerms * (kChunk / 32) * 33];
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), best_b = a;
    int best_i = -1;
    float best_d = CUDART_INF_F;
    bool have_b = false;
    if (valid) {
        a = q_cur ? a_early : d.D[(pass + 1) & 1][i];
        const float4 nb = q_cur ? nb_early : __ldcg(&d.gnb[i]);
        best_i = __float_as_int(nb.w);
        if (best_i >= 0) {
            best_b = nb; have_b = true;
            best_d = exact_distance(a.x, a.y, a.z, nb.x, nb.y, nb.z); // same inputs, same arithmetic: the search's own bits
        }
    }
    finalize_tail<false>(d, st, prm, pass, chunk, nchunks, tid, valid, i, n, a, best_i, best_d, best_b, have_b, 0, s_fin, s_fold);
}

// CANON-3 level 2 + the solve for the cooperative search, one CTA per registration with all the registers it wants.
// Inside nn_finalize_coop_kernel (48 registers for occupancy) the same code ran out of spilled local memory: 43 us of
// a 49 us kernel were this tail on one SM while 147 idled (profiles/r02_ncu_nn_finalize_coop_tail.txt).
__global__ void __launch_bounds__(kChunk, 1) nn_solve_kernel(const RegDesc *__restrict__ descs, IcpState *states,
                                                             const IcpParamsDev *__restrict__ prm, int pass)
{
    pdl_enter(); // icpb_internal.h: the grid before this one is complete from here on
    IcpState *st = states + blockIdx.z;
    if (st->done) return;
    const RegDesc &d = descs[blockIdx.z];
    __shared__ FinalizeShared s_fin;
    const int nchunks = (d.n + kChunk - 1) / kChunk;
    finalize_last_cta_body(d.chunk_sums, d.mlog, st, prm, pass, nchunks, threadIdx.x, s_fin, false);
}

void launch_nn_finalize(const RegDesc *descs, IcpState *states, const IcpParamsDev *prm, int batch, int max_n, int splits,
                        int pass, int filter, cudaStream_t s, const float4 *q_cur, const float4 *q_nb)
{
    dim3 grid((max_n + kChunk - 1) / kChunk, 1, batch);
    if (splits < 0) { // cooperative cell-grid search
        launch_pdl(nn_finalize_coop_kernel, grid, dim3(kChunk), 0, s, descs, states, prm, pass, batch == 1 ? q_cur : nullptr, q_nb, max_n);
        launch_pdl(nn_solve_kernel, dim3(1, 1, batch), dim3(kChunk), 0, s, descs, states, prm, pass);
        return;
    }
    // brute-force modes on few CTAs (a small cloud; latency-bound: 10k points are 40 CTAs on 148 SMs): a pair of
    // threads per query for the group selection and the exact stage.  Many CTAs (full resolution, batches) are
    // throughput-bound and keep one thread per query (measured: pairs -6 % on the 1024-registration batch).
    const bool paired = splits > 0 && (long long)grid.x * batch <= 2 * 148;
    launch_pdl<true>(nn_finalize_kernel, grid, dim3(paired ? 2 * kChunk : kChunk), 0, s, descs, states, prm, splits, pass, filter);
}

// --------------------------------------------------------------------------
// stand-alone P2 transform, and the deferred translation of the <3 rule
// --------------------------------------------------------------------------
struct RtArgs {
    float R[9];
    float t[3];
};

// R and t travel as kernel arguments: no staging buffer, no host synchronisation
__global__ void transform_kernel(float4 *pts, int n, RtArgs a, int have_R, int have_t, const int *__restrict__ n_dev)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev) n = max(0, min(n, *n_dev)); // -1 in a band header: the lift failed, no points
    if (i >= n) return;
    float4 p = pts[i];
    if (have_R) {
        float x = (a.R[0] * p.x + a.R[1] * p.y) + a.R[2] * p.z;
        float y = (a.R[3] * p.x + a.R[4] * p.y) + a.R[5] * p.z;
        float z = (a.R[6] * p.x + a.R[7] * p.y) + a.R[8] * p.z;
        p.x = x; p.y = y; p.z = z;
    }
    if (have_t) { p.x += a.t[0]; p.y += a.t[1]; p.z += a.t[2]; }
    pts[i] = p;
}

void launch_transform(float4 *pts, int n, const float *R, const float *t, int have_R, int have_t, cudaStream_t s,
                      const int *n_dev)
{
    if (n <= 0) return;
    RtArgs a;
    for (int k = 0; k < 9; ++k) a.R[k] = have_R ? R[k] : 0.f;
    for (int k = 0; k < 3; ++k) a.t[k] = have_t ? t[k] : 0.f;
    transform_kernel<<<(n + 255) / 256, 256, 0, s>>>(pts, n, a, have_R, have_t, n_dev);
}

// z-slab exchange helpers: a band = one 16-byte header row (point count) followed by band_capacity point rows
__global__ void pack_band_kernel(const float4 *__restrict__ pts, int n, float4 *dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) dst[0] = make_float4(__int_as_float(n), 0.f, 0.f, 0.f);
    if (i < n) dst[1 + i] = pts[i];
}

void launch_pack_band(const float4 *pts, int n, float4 *dst, cudaStream_t s)
{
    pack_band_kernel<<<(max(n, 1) + 255) / 256, 256, 0, s>>>(pts, n, dst);
}

// blockIdx.y = band; output offset = sum of the counts of the bands before it (rank order == raster order)
__global__ void assemble_bands_kernel(const float4 *__restrict__ bands, int world, int band_capacity, float4 *out,
                                      int out_capacity, int *total)
{
    const int r = blockIdx.y;
    const float4 *band = bands + (size_t)r * (band_capacity + 1);
    int base = 0, all = 0;
    for (int q = 0; q < world; ++q) {
        int c = __float_as_int(bands[(size_t)q * (band_capacity + 1)].x);
        c = min(max(c, 0), band_capacity);
        if (q < r) base += c;
        all += c;
    }
    const int cnt = min(max(__float_as_int(band[0].x), 0), band_capacity);
    if (r == 0 && blockIdx.x == 0 && threadIdx.x == 0) *total = all;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x)
        if (base + i < out_capacity) out[base + i] = band[1 + i];
}

void launch_assemble_bands(const float4 *bands, int world, int band_capacity, float4 *out, int out_capacity, int *total,
                           cudaStream_t s)
{
    dim3 grid((band_capacity + 1023) / 1024, world);
    assemble_bands_kernel<<<grid, 256, 0, s>>>(bands, world, band_capacity, out, out_capacity, total);
}

__global__ void pending_translate_kernel(const RegDesc *__restrict__ descs)
{
    pdl_enter(); // icpb_internal.h: the grid before this one is complete from here on
    const RegDesc &d = descs[blockIdx.z];
    const IcpState *st = d.st;
    if (!st->pending_translate) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    float4 *buf = d.D[st->last_buf];
    float4 p = buf[i];
    p.x += st->tf[0]; p.y += st->tf[1]; p.z += st->tf[2];
    buf[i] = p;
}

// Key-point variant, after the loop.  blockIdx.y == 0 .. : the carried cloud replays the recorded motions, each as
// rotate-then-translate in float exactly like the associated key-points experienced them (pointcloud.cpp:321-359),
// then the deferred translation of the <3 rule.  The LAST y-slice (one CTA) compacts the reject flags of all executed
// passes, pass-major and in query order, into the non-association list (icp.cpp:96,508).
__global__ void __launch_bounds__(256) keypoint_epilogue_kernel(const RegDesc *__restrict__ descs)
{
    pdl_enter(); // icpb_internal.h: the grid before this one is complete from here on
    const RegDesc &d = descs[blockIdx.z];
    const IcpState *st = d.st;
    if (blockIdx.y + 1 < gridDim.y) {
        const int i = (blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
        if (!d.carry || i >= d.n_carry) return;
        float4 p = d.carry[i];
        const int nl = st->n_log;
        for (int k = 0; k < nl; ++k) p = apply_rt(p, d.mlog + 12 * k, d.mlog + 12 * k + 9);
        if (st->pending_translate) { p.x += st->tf[0]; p.y += st->tf[1]; p.z += st->tf[2]; }
        d.carry[i] = p;
        return;
    }
    if (blockIdx.x != 0 || !d.rej_flag) return;
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    const long long total = (long long)st->passes * d.n;
    for (long long c0 = 0; c0 < total; c0 += 256) {
        const long long e = c0 + tid;
        const int f = (e < total) ? d.rej_flag[e] : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < wid; ++w) before += s_warp[w];
        const int pos = before + __popc(bal & ((1u << lane) - 1u));
        if (f && pos < d.nonassoc_capacity) d.nonassoc[pos] = d.rej_pts[e];
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
    if (tid == 0) d.st->n_nonassoc = s_base;
}

void launch_keypoint_epilogue(const RegDesc *descs, int batch, int max_carry, cudaStream_t s)
{
    const int blocks = (max_carry + 255) / 256;
    dim3 grid(blocks > 0 ? blocks : 1, 2, batch); // y = 0: carried points; y = 1: reject compaction
    launch_pdl(keypoint_epilogue_kernel, grid, dim3(256), 0, s, descs);
}

// The associated cloud back into the caller's buffer when the loop ended with it in the alternate one.
__global__ void copy_back_kernel(const RegDesc *__restrict__ descs)
{
    pdl_enter(); // icpb_internal.h
    const RegDesc &d = descs[blockIdx.z];
    if (d.st->last_buf != 1) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d.n) d.D[0][i] = d.D[1][i];
}

void launch_copy_back(const RegDesc *descs, int batch, int max_n, cudaStream_t s)
{
    dim3 grid((max_n + 255) / 256, 1, batch);
    launch_pdl(copy_back_kernel, grid, dim3(256), 0, s, descs);
}

void launch_pending_translate(const RegDesc *descs, int batch, int max_n, cudaStream_t s)
{
    dim3 grid((max_n + 255) / 256, 1, batch);
    launch_pdl(pending_translate_kernel, grid, dim3(256), 0, s, descs);
}

// --------------------------------------------------------------------------
// PointCloud::center (pointcloud.cpp:43-45,100-102) in the canonical FP64 order
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(kChunk) center_kernel(const float4 *__restrict__ pts, int n, double *chunk_sums,
                                                        double *out3, unsigned int *counter)
{
    const int tid = threadIdx.x;
    const int i = blockIdx.x * kChunk + tid;
    const int nchunks = (n + kChunk - 1) / kChunk;
    __shared__ double s_w[kChunk / 32][3];
    __shared__ int s_last;
    double t[3] = {0.0, 0.0, 0.0};
    if (i < n) {
        float4 p = pts[i];
        t[0] = p.x; t[1] = p.y; t[2] = p.z;
    }
    for (int k = 0; k < 3; ++k) {
        double v = warp_fold(t[k]);
        if ((tid & 31) == 0) s_w[tid >> 5][k] = v;
    }
    __syncthreads();
    if (tid < 3) {
        double s = s_w[0][tid];
        for (int wv = 1; wv < kChunk / 32; ++wv) s = s + s_w[wv][tid];
        chunk_sums[(size_t)blockIdx.x * 3 + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(counter, 1u) == (unsigned int)(nchunks - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int k = 0; k < 3; ++k) {
        double acc = 0.0;
        for (int c = tid; c < nchunks; c += kChunk) acc = acc + __ldcg(&chunk_sums[(size_t)c * 3 + k]);
        double v = warp_fold(acc);
        if ((tid & 31) == 0) s_w[tid >> 5][k] = v;
    }
    __syncthreads();
    if (tid < 3) {
        double s = s_w[0][tid];
        for (int wv = 1; wv < kChunk / 32; ++wv) s = s + s_w[wv][tid];
        out3[tid] = s / (double)n;
    }
    if (tid == 0) *counter = 0;
}

void launch_center(const float4 *pts, int n, double *chunk_sums, double *out3, unsigned int *counter, cudaStream_t s)
{
    center_kernel<<<(n + kChunk - 1) / kChunk, kChunk, 0, s>>>(pts, n, chunk_sums, out3, counter);
}

// --------------------------------------------------------------------------
// FP32 FMA peak micro-benchmark (roofline denominator of nn_partial)
// --------------------------------------------------------------------------
__global__ void fp32_peak_kernel(float *out, int iters)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float b = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = __fmaf_rn(a0, b, c); a1 = __fmaf_rn(a1, b, c); a2 = __fmaf_rn(a2, b, c); a3 = __fmaf_rn(a3, b, c);
            a4 = __fmaf_rn(a4, b, c); a5 = __fmaf_rn(a5, b, c); a6 = __fmaf_rn(a6, b, c); a7 = __fmaf_rn(a7, b, c);
        }
    }
    float r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

void launch_fp32_peak(float *out, int blocks, int threads, int iters, cudaStream_t s)
{
    fp32_peak_kernel<<<blocks, threads, 0, s>>>(out, iters);
}

} // namespace icpb

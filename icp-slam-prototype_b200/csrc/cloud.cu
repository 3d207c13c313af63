// Image-space stages: depth -> XYZ back-projection with order-preserving
// compaction, normals, depth pre-filter.
//
// Replaces (reference file:line):
//   PointCloud::PointCloud(cv::Mat&, cv::Mat[, keypoints])  pointcloud.cpp:11-58, 109-165  (P1)
//   getNormalMap                                            SLAM.cpp:412-430               (P3)
//   filterDepthImage                                        SLAM.cpp:553-573               (8f-1)
//
// Built with -fmad=false; float divisions are IEEE (nvcc default -prec-div=true).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "icpb_internal.h"

namespace icpb {

constexpr int kBpThreads = 256;
constexpr int kBpPix = 8;                         // pixels per thread: one 16-byte load of u16
constexpr int kBpTile = kBpThreads * kBpPix;      // 2048 pixels: one staging buffer
constexpr int kBpSubMax = 4;                      // most tiles a CTA of the chained kernel walks through its buffer

int backproject_tiles(int w, int h) { return (w * h + kBpTile - 1) / kBpTile; }

// murmur3 finaliser; the ICPB_SUB_HASH stand-in for rand() (pointcloud.cpp:28)
__device__ __forceinline__ uint32_t hash32(uint32_t seed, uint32_t pixel)
{
    uint32_t h = seed ^ (pixel * 0x9E3779B9u);
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

// ---- decoupled look-back over 64-bit status words ---------------------------
// word = (epoch*4 + flag) << 32 | value ; flag 1 = tile aggregate, 2 = inclusive prefix.
// Words left over from earlier launches carry an older epoch and read as "not ready".
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by one full warp; returns the exclusive prefix of `tile`.  The spin is
// bounded: a predecessor that never publishes (an internal error) makes the
// launch report failure instead of hanging the GPU.
__device__ uint32_t lookback(unsigned long long *state, int tile, uint32_t epoch, int *failed)
{
    const int lane = threadIdx.x & 31;
    uint32_t exclusive = 0;
    int base = tile - 1;
    unsigned int spins = 0;
    while (base >= 0) {
        if (++spins > (1u << 24)) { *failed = 1; break; }
        int j = base - lane;
        uint32_t flag = 2, val = 0;
        if (j >= 0) {
            unsigned long long v = ld_volatile_u64(&state[j]);
            uint32_t hi = (uint32_t)(v >> 32);
            flag = ((hi >> 2) == epoch) ? (hi & 3u) : 0u;
            val = (uint32_t)v;
        }
        if (__any_sync(0xffffffffu, flag == 0)) continue; // a predecessor has not published yet
        uint32_t incl_mask = __ballot_sync(0xffffffffu, flag == 2);
        int first = incl_mask ? (__ffs(incl_mask) - 1) : 31;
        uint32_t c = (lane <= first) ? val : 0;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        exclusive += c;
        if (incl_mask) break;
        base -= 32;
    }
    return exclusive;
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *s_warp, uint32_t &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += o;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < kBpThreads / 32; ++k) {
        uint32_t c = s_warp[k];
        if (k < wid) wbase += c;
        tot += c;
    }
    __syncthreads();
    total = tot;
    return wbase + inc - v;
}

// The staged tile leaves with fully coalesced 16-byte stores: all shared-memory reads first, then all the stores, so
// the eight stores of a thread are in flight together instead of each waiting for its own LDS.
__device__ __forceinline__ void store_tile(const float4 *s_pts, uint32_t tile_keep, uint32_t tile_base, float4 *out,
                                           int capacity, int tid)
{
    float4 v[kBpPix];
#pragma unroll
    for (int k = 0; k < kBpPix; ++k) {
        const uint32_t j = (uint32_t)tid + (uint32_t)k * kBpThreads;
        if (j < tile_keep) v[k] = s_pts[j];
    }
#pragma unroll
    for (int k = 0; k < kBpPix; ++k) {
        const uint32_t j = (uint32_t)tid + (uint32_t)k * kBpThreads;
        const uint32_t o = tile_base + j;
        if (j < tile_keep && (int)o < capacity) out[o] = v[k];
    }
}

// Correctly rounded a / c for a divisor known on the host: rc = RN(1/c) (computed in double), then the classical
// FMA sequence -- quotient estimate, exact residual, correction, twice.  The first correction leaves a faithful
// quotient, the second is then the correctly rounded one (Markstein's theorem; needs rc correctly rounded, which the
// host guarantees, and no over/underflow, which the launcher checks).  Same result as the `/` of
// pointcloud.cpp:37-39, five instructions instead of the ~ten of a general IEEE division.
struct DivConst {
    float c, rc;
};
template <bool FAST>
__device__ __forceinline__ float div_by(float a, DivConst d)
{
    if (!FAST) return a / d.c;
    float q = __fmul_rn(a, d.rc);
    float r = __fmaf_rn(-d.c, q, a);
    q = __fmaf_rn(r, d.rc, q);
    r = __fmaf_rn(-d.c, q, a);
    return __fmaf_rn(r, d.rc, q);
}

struct BackprojectDiv {
    DivConst scale, fx_u, fx_v;
    unsigned long long row_magic; // floor(2^40 / w) + 1, or 0 when the image is too large for the multiply-shift form
};

// p / w for 0 <= p < 2^24 and p * w < 2^40: (p * magic) >> 40 is the exact floor (the estimate is high by less than
// p / 2^40 < 1 / w) -- three instructions instead of the ~25 of a general integer division
__device__ __forceinline__ int row_of(int p, int w, unsigned long long magic)
{
    return magic ? (int)(((unsigned long long)(uint32_t)p * magic) >> 40) : p / w;
}

__device__ __forceinline__ void load_depth8(const BackprojectArgs &a, int p0, int npx, uint32_t (&dw)[4])
{
    dw[0] = dw[1] = dw[2] = dw[3] = 0;
    if (p0 + kBpPix <= npx) {
        const uint4 raw = *reinterpret_cast<const uint4 *>(a.depth + p0);
        dw[0] = raw.x; dw[1] = raw.y; dw[2] = raw.z; dw[3] = raw.w;
    } else {
#pragma unroll
        for (int k = 0; k < kBpPix; ++k)
            if (p0 + k < npx) dw[k >> 1] |= (uint32_t)a.depth[p0 + k] << (16 * (k & 1));
    }
}

// Lift a thread's 8 pixels and store the kept ones from `dst` on (pointcloud.cpp:37-39 / 134-136: all float, left to
// right, true division).
template <bool HAS_BGR, bool FASTDIV>
__device__ __forceinline__ void lift8(const BackprojectArgs &a, const BackprojectDiv &dv, int p0, int npx,
                                      const uint32_t (&dw)[4], uint32_t keep_mask, float4 *dst)
{
    if (!keep_mask) return;
    uint32_t cw[6] = {0, 0, 0, 0, 0, 0}; // the thread's 8 BGR triples: 24 bytes, three 8-byte loads when aligned
    if (HAS_BGR) {
        const uint8_t *c = a.bgr + (size_t)p0 * 3;
        if (p0 + kBpPix <= npx && ((reinterpret_cast<uintptr_t>(c) & 7) == 0)) {
            const uint2 *c2 = reinterpret_cast<const uint2 *>(c);
            const uint2 w0 = c2[0], w1 = c2[1], w2 = c2[2];
            cw[0] = w0.x; cw[1] = w0.y; cw[2] = w1.x; cw[3] = w1.y; cw[4] = w2.x; cw[5] = w2.y;
        } else {
            for (int b = 0; b < 3 * kBpPix; ++b)
                if (p0 * 3 + b < npx * 3) cw[b >> 2] |= (uint32_t)c[b] << (8 * (b & 3));
        }
    }
    const int v0 = row_of(p0, a.w, dv.row_magic);
    const int u0 = p0 - v0 * a.w;
    // Image widths are multiples of 8 in practice (640, 512): the thread's 8 pixels then share a row and the
    // coordinates are u0 + k, exactly representable float sums; otherwise every pixel finds its own (u, v).
    const bool same_row = (a.w % kBpPix) == 0;
    const float uf0 = (float)u0, vf0 = (float)(v0 + a.v_offset);
    // branch-free: all 8 pixels are lifted (a zero depth just gives z = 0), only the kept ones are stored
#pragma unroll
    for (int k = 0; k < kBpPix; ++k) {
        float uf = uf0 + (float)k, vf = vf0;
        if (!same_row) {
            const int pk = p0 + k, vv = pk / a.w;
            uf = (float)(pk - vv * a.w);
            vf = (float)(vv + a.v_offset);
        }
        const uint32_t d16 = (dw[k >> 1] >> (16 * (k & 1))) & 0xffffu;
        // (float)d for d < 2^23 without a conversion instruction: 0x4B000000 | d is 2^23 + d
        const float df = __fsub_rn(__uint_as_float(0x4B000000u | d16), 8388608.0f);
        const float pz = div_by<FASTDIV>(df, dv.scale);
        const float px = div_by<FASTDIV>(__fmul_rn(__fsub_rn(uf, a.K.cx_u), pz), dv.fx_u);
        const float py = div_by<FASTDIV>(__fmul_rn(__fsub_rn(vf, a.K.cx_v), pz), dv.fx_v);
        uint32_t cbits = 0;
        if (HAS_BGR) { // bytes 3k .. 3k+2 of the 24-byte run (:47): a funnel shift across two words
            const int w = (3 * k) >> 2, sh = 8 * ((3 * k) & 3);
            cbits = __funnelshift_r(cw[w], w + 1 < 6 ? cw[w + 1] : 0u, sh) & 0x00ffffffu;
        }
        if (keep_mask & (1u << k)) {
            *dst = make_float4(px, py, pz, __uint_as_float(cbits));
            ++dst;
        }
    }
}

// P1.  One CTA per run of kBpSub consecutive 2048-pixel tiles (run = blockIdx.x: predecessors are dispatched first).
// Two chained scans: the ordinal among NON-ZERO pixels (consumed by the subsample rule exactly where the reference
// consumes one rand(), pointcloud.cpp:22-28; only the STRIDE / STREAM rules need it) and the output position among KEPT
// pixels (raster order == push_back order, :54).  The look-back for the output position runs in warp 0 AFTER it has
// lifted its first pixels, while the other warps lift theirs: its latency hides behind the arithmetic.
// Why several tiles per CTA: under a saturated memory system every dependent round trip costs a microsecond or two, and
// the staging buffer (what bounds the CTAs per SM) is held for the whole chain depth-load -> scan -> look-back -> copy.
// A CTA loads the depth of all its tiles at once and looks back once, then walks the tiles through the same 32 KB
// buffer: one load round trip and one look-back per kBpSub tiles instead of per tile.  Measured on 256 resident frames:
// 4,227 GB/s with one tile per CTA, 4,978 with two, 5,113 with four at 59 registers (4 CTAs per SM) and 5,450 with
// four at 40 registers (6 CTAs per SM); launches with few tiles (a single frame is 150) take two so that more SMs
// take part.
// Specialised on the subsample rule, on the presence of a colour image and on the division path.
template <int RULE, bool HAS_BGR, bool FASTDIV, int kBpSub>
// (six CTAs per SM is what the 32 KB staging buffer allows: hold the four-tile variant to 40 registers for it -- a
//  dozen bytes of spill -- 5,113 -> 5,450 GB/s)
__global__ void __launch_bounds__(kBpThreads, 6)
backproject_kernel(BackprojectArgs a, BackprojectDiv dv, uint32_t epoch)
{
    constexpr int kBpSuper = kBpTile * kBpSub;
    // batched launch: frame = blockIdx.y, every per-frame pointer advances by its stride
    {
        const long long f = blockIdx.y;
        a.depth += f * a.depth_stride;
        if (HAS_BGR) a.bgr += f * a.bgr_stride;
        a.out += f * a.out_stride;
        a.tile_state += f * a.state_stride;
        a.out_count = (int *)((unsigned long long *)a.out_count + f * a.state_stride);
    }
    __shared__ uint32_t s_warp[kBpThreads / 32];
    __shared__ uint32_t s_bcast[2];
    __shared__ float4 s_pts[kBpTile];
    const int tid = threadIdx.x;
    const int tile = blockIdx.x; // index of the run: the unit of both chained scans
    int failed = 0;
    unsigned long long *stateV = a.tile_state;
    unsigned long long *stateK = a.tile_state + a.n_tiles;
    const int npx = a.w * a.h;
    const int p_first = tile * kBpSuper + tid * kBpPix;

    uint32_t dw[kBpSub][4]; // 8 u16 depths per tile of the run
    uint32_t valid_mask[kBpSub], keep_mask[kBpSub];
#pragma unroll
    for (int j = 0; j < kBpSub; ++j) load_depth8(a, p_first + j * kBpTile, npx, dw[j]);
#pragma unroll
    for (int j = 0; j < kBpSub; ++j) {
        valid_mask[j] = 0;
#pragma unroll
        for (int k = 0; k < kBpPix; ++k)
            if ((dw[j][k >> 1] >> (16 * (k & 1))) & 0xffffu) valid_mask[j] |= 1u << k;
        keep_mask[j] = valid_mask[j];
    }

    constexpr bool need_ordinal = (RULE == ICPB_SUB_STRIDE) || (RULE == ICPB_SUB_STREAM);
    const unsigned long long tag = (unsigned long long)(epoch << 2) << 32;
    if (RULE != ICPB_SUB_NONE) {
        uint32_t v_off[kBpSub];
#pragma unroll
        for (int j = 0; j < kBpSub; ++j) v_off[j] = 0;
        uint32_t v_base = 0;
        if (need_ordinal) {
            uint32_t run_valid = 0;
#pragma unroll
            for (int j = 0; j < kBpSub; ++j) {
                uint32_t t;
                v_off[j] = run_valid + block_exclusive_scan(__popc(valid_mask[j]), s_warp, t);
                run_valid += t;
            }
            if (tid == 0) st_volatile_u64(&stateV[tile], tag | ((tile == 0 ? 2ull : 1ull) << 32) | run_valid);
            if (tid < 32) {
                const uint32_t ex = (tile == 0) ? 0u : lookback(stateV, tile, epoch, &failed);
                if (tid == 0) {
                    if (tile != 0) st_volatile_u64(&stateV[tile], tag | (2ull << 32) | (ex + run_valid));
                    s_bcast[0] = ex;
                }
            }
            __syncthreads();
            v_base = s_bcast[0];
        }
        const uint32_t rule_arg = a.rule_arg ? a.rule_arg : 1u;
#pragma unroll
        for (int j = 0; j < kBpSub; ++j) {
            keep_mask[j] = 0;
            uint32_t ord = v_base + v_off[j];
#pragma unroll
            for (int k = 0; k < kBpPix; ++k) {
                if (valid_mask[j] & (1u << k)) {
                    bool keep = true;
                    if (RULE == ICPB_SUB_STRIDE) keep = (ord % rule_arg) == 0;
                    else if (RULE == ICPB_SUB_HASH) keep = (hash32(a.seed, (uint32_t)(p_first + j * kBpTile + k)) % rule_arg) == 0;
                    else if (RULE == ICPB_SUB_STREAM) keep = (ord < (uint32_t)a.keep_stream_len) && a.keep_stream[ord] != 0;
                    if (keep) keep_mask[j] |= 1u << k;
                    ++ord;
                }
            }
        }
    }
    uint32_t k_off[kBpSub], sub_keep[kBpSub], run_keep = 0;
#pragma unroll
    for (int j = 0; j < kBpSub; ++j) {
        k_off[j] = block_exclusive_scan(__popc(keep_mask[j]), s_warp, sub_keep[j]);
        run_keep += sub_keep[j];
    }
    if (tid == 0) st_volatile_u64(&stateK[tile], tag | ((tile == 0 ? 2ull : 1ull) << 32) | run_keep);

    // ---- the first tile of the run goes into the staging buffer at its tile-local positions
    lift8<HAS_BGR, FASTDIV>(a, dv, p_first, npx, dw[0], keep_mask[0], &s_pts[k_off[0]]);
    // ---- warp 0: exclusive prefix of this run among the kept pixels (its own lifting is already done)
    if (tid < 32) {
        const uint32_t ex = (tile == 0) ? 0u : lookback(stateK, tile, epoch, &failed);
        if (tid == 0) {
            if (tile != 0) st_volatile_u64(&stateK[tile], tag | (2ull << 32) | (ex + run_keep));
            s_bcast[1] = ex;
            if (failed) *a.out_count = -1;
            else if (tile == (int)gridDim.x - 1) *a.out_count = (int)(ex + run_keep);
        }
    }
    // ---- every tile leaves as ONE bulk copy shared -> global (1-D TMA, UBLKCP): the staged points are contiguous in
    //      shared memory and in the output, 16-byte aligned at both ends, so no thread has to read them back and store
    //      them (that loop was a quarter of the kernel's instructions and half of its stall samples).  The writers
    //      make their generic-proxy stores visible to the async proxy, the barrier orders them before the copy, and
    //      the issuing thread holds the buffer (and at the end the CTA's shared memory) until the copy has read it.
    uint32_t done = 0; // points of the run already copied out (thread 0)
#pragma unroll
    for (int j = 0; j < kBpSub; ++j) {
        if (j > 0) lift8<HAS_BGR, FASTDIV>(a, dv, p_first + j * kBpTile, npx, dw[j], keep_mask[j], &s_pts[k_off[j]]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const uint32_t base = s_bcast[1] + done;
            const long long room = (long long)a.capacity - (long long)base;
            const uint32_t n_out = (uint32_t)max(0ll, min((long long)sub_keep[j], room));
            if (n_out) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.out + base),
                             "r"((uint32_t)__cvta_generic_to_shared(s_pts)), "r"(n_out * 16u)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            done += sub_keep[j];
        }
        if (j + 1 < kBpSub) __syncthreads(); // the buffer is free for the next tile
    }
}

// ---- two-pass variant for the rules whose keep decision is local to the pixel (NONE, HASH) -------------------
// Pass 1 counts the kept pixels of every tile; pass 2 lifts and writes, each CTA summing the counts of the tiles
// before it in its frame (<= a few hundred L2-resident words, fetched before anything else).  No CTA ever waits for
// another one, which is what bounded the chained-scan kernel (30 % of its stall samples sat behind the look-back).
// Costs a second read of the depth image (2 of 18 bytes per pixel, mostly from L2).
template <int RULE>
__device__ __forceinline__ uint32_t local_keep_mask(const uint32_t (&dw)[4], const BackprojectArgs &a, int p0)
{
    uint32_t m = 0;
    const uint32_t rule_arg = a.rule_arg ? a.rule_arg : 1u;
#pragma unroll
    for (int k = 0; k < kBpPix; ++k) {
        bool keep = ((dw[k >> 1] >> (16 * (k & 1))) & 0xffffu) != 0;
        if (RULE == ICPB_SUB_HASH) keep = keep && (hash32(a.seed, (uint32_t)(p0 + k)) % rule_arg) == 0;
        if (keep) m |= 1u << k;
    }
    return m;
}

template <int RULE>
__global__ void __launch_bounds__(kBpThreads) backproject_count_kernel(BackprojectArgs a)
{
    a.depth += (long long)blockIdx.y * a.depth_stride;
    int *counts = reinterpret_cast<int *>(a.tile_state + (long long)blockIdx.y * a.state_stride + 2 * a.n_tiles);
    __shared__ uint32_t s_warp[kBpThreads / 32];
    const int tid = threadIdx.x, tile = blockIdx.x;
    const int p0 = tile * kBpTile + tid * kBpPix;
    uint32_t dw[4];
    load_depth8(a, p0, a.w * a.h, dw);
    uint32_t c = __popc(local_keep_mask<RULE>(dw, a, p0));
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((tid & 31) == 0) s_warp[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int k = 0; k < kBpThreads / 32; ++k) t += s_warp[k];
        counts[tile] = (int)t;
    }
}

template <int RULE, bool HAS_BGR, bool FASTDIV>
__global__ void __launch_bounds__(kBpThreads) backproject_write_kernel(BackprojectArgs a, BackprojectDiv dv)
{
    {
        const long long f = blockIdx.y;
        a.depth += f * a.depth_stride;
        if (HAS_BGR) a.bgr += f * a.bgr_stride;
        a.out += f * a.out_stride;
        a.tile_state += f * a.state_stride;
        a.out_count = (int *)((unsigned long long *)a.out_count + f * a.state_stride);
    }
    __shared__ uint32_t s_warp[kBpThreads / 32];
    __shared__ uint32_t s_before[kBpThreads / 32];
    __shared__ float4 s_pts[kBpTile];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile = blockIdx.x;
    const int npx = a.w * a.h;
    const int p0 = tile * kBpTile + tid * kBpPix;

    // points kept by the tiles before this one: independent loads, issued first
    const int *counts = reinterpret_cast<const int *>(a.tile_state + 2 * a.n_tiles);
    uint32_t before = 0;
    for (int j = tid; j < tile; j += kBpThreads) before += (uint32_t)__ldcg(&counts[j]);

    uint32_t dw[4];
    load_depth8(a, p0, npx, dw);
    const uint32_t keep_mask = local_keep_mask<RULE>(dw, a, p0);
    const uint32_t nkeep = __popc(keep_mask);
    uint32_t inc = nkeep;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += o;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) before += __shfl_xor_sync(0xffffffffu, before, off);
    if (lane == 31) s_warp[wid] = inc;
    if (lane == 0) s_before[wid] = before;
    __syncthreads();
    uint32_t wbase = 0, tile_keep = 0, tile_base = 0;
#pragma unroll
    for (int k = 0; k < kBpThreads / 32; ++k) {
        const uint32_t c = s_warp[k];
        if (k < wid) wbase += c;
        tile_keep += c;
        tile_base += s_before[k];
    }
    const uint32_t k_off = wbase + inc - nkeep;
    if (tid == 0 && tile == a.n_tiles - 1) *a.out_count = (int)(tile_base + tile_keep);

    if (keep_mask) {
        uint32_t cw[6] = {0, 0, 0, 0, 0, 0}; // the thread's 8 BGR triples: 24 bytes, three 8-byte loads when aligned
        if (HAS_BGR) {
            const uint8_t *c = a.bgr + (size_t)p0 * 3;
            if (p0 + kBpPix <= npx && ((reinterpret_cast<uintptr_t>(c) & 7) == 0)) {
                const uint2 *c2 = reinterpret_cast<const uint2 *>(c);
                const uint2 w0 = c2[0], w1 = c2[1], w2 = c2[2];
                cw[0] = w0.x; cw[1] = w0.y; cw[2] = w1.x; cw[3] = w1.y; cw[4] = w2.x; cw[5] = w2.y;
            } else {
                for (int b = 0; b < 3 * kBpPix; ++b)
                    if (p0 * 3 + b < npx * 3) cw[b >> 2] |= (uint32_t)c[b] << (8 * (b & 3));
            }
        }
        const int v0 = row_of(p0, a.w, dv.row_magic);
        const int u0 = p0 - v0 * a.w;
        const bool same_row = (a.w % kBpPix) == 0; // see backproject_kernel
        const float uf0 = (float)u0, vf0 = (float)(v0 + a.v_offset);
        float4 *dst = &s_pts[k_off];
#pragma unroll
        for (int k = 0; k < kBpPix; ++k) {
            float uf = uf0 + (float)k, vf = vf0;
            if (!same_row) {
                const int pk = p0 + k, vv = pk / a.w;
                uf = (float)(pk - vv * a.w);
                vf = (float)(vv + a.v_offset);
            }
            const uint32_t d16 = (dw[k >> 1] >> (16 * (k & 1))) & 0xffffu;
            const float df = __fsub_rn(__uint_as_float(0x4B000000u | d16), 8388608.0f);
            const float pz = div_by<FASTDIV>(df, dv.scale);                                             // pointcloud.cpp:37
            const float px = div_by<FASTDIV>(__fmul_rn(__fsub_rn(uf, a.K.cx_u), pz), dv.fx_u);          // :38
            const float py = div_by<FASTDIV>(__fmul_rn(__fsub_rn(vf, a.K.cx_v), pz), dv.fx_v);          // :39
            uint32_t cbits = 0;
            if (HAS_BGR) {
                const int w = (3 * k) >> 2, sh = 8 * ((3 * k) & 3);
                cbits = __funnelshift_r(cw[w], w + 1 < 6 ? cw[w + 1] : 0u, sh) & 0x00ffffffu;        // :47
            }
            if (keep_mask & (1u << k)) {
                *dst = make_float4(px, py, pz, __uint_as_float(cbits));
                ++dst;
            }
        }
    }
    __syncthreads();
    store_tile(s_pts, tile_keep, tile_base, a.out, a.capacity, tid);
}

static bool fast_div_ok(float c)
{
    // the FMA sequence needs a normal divisor well inside the exponent range; an all-ones significand is kept on the
    // plain division path as well (the reciprocal's rounding is least favourable there)
    uint32_t b;
    memcpy(&b, &c, 4);
    const uint32_t e = (b >> 23) & 0xffu, m = b & 0x7fffffu;
    return e >= 127 - 60 && e <= 127 + 60 && m != 0x7fffffu;
}

// The two-pass kernels keep plain per-tile counts behind the status words.  A later chained launch over a LARGER image
// reads those bytes as epoch-tagged status words, and a small count can spell a small epoch: zero them once the write
// kernel is done (epoch 0 is never used, so a zero word reads as "not ready").
static void clear_two_pass_counts(const BackprojectArgs &a, cudaStream_t s)
{
    unsigned long long *counts = a.tile_state + 2 * (long long)a.n_tiles;
    const size_t bytes = sizeof(unsigned long long) * (size_t)a.n_tiles;
    if (a.frames > 1 && a.state_stride > 0)
        cudaMemset2DAsync(counts, sizeof(unsigned long long) * (size_t)a.state_stride, 0, bytes, (size_t)a.frames, s);
    else cudaMemsetAsync(counts, 0, bytes, s);
}

void launch_backproject(const BackprojectArgs &a, cudaStream_t s, bool force_two_pass)
{
    // per-process launch counter (atomic: contexts may be driven from different host threads); stale tile words never
    // match it.  1 .. 2^30 - 1, never 0.
    static std::atomic<uint32_t> launches{0};
    const uint32_t epoch = launches.fetch_add(1, std::memory_order_relaxed) % 0x3fffffffu + 1u;
    dim3 grid(a.n_tiles, a.frames > 0 ? a.frames : 1);
    // chained kernel: one CTA per run of 2 or 4 tiles (4 once the launch has tiles for several waves of CTAs)
    bool sub4 = (long long)a.n_tiles * grid.y >= 8192;
    if (const char *e = getenv("ICPB_BP_SUB")) sub4 = atoi(e) == 4; // tests pin either run length
    const int sub = sub4 ? 4 : 2;
    dim3 cgrid((a.n_tiles + sub - 1) / sub, grid.y);
    BackprojectDiv dv;
    dv.scale = {a.K.depth_scale, (float)(1.0 / (double)a.K.depth_scale)};
    dv.fx_u = {a.K.fx_u, (float)(1.0 / (double)a.K.fx_u)};
    dv.fx_v = {a.K.fx_v, (float)(1.0 / (double)a.K.fx_v)};
    {
        const unsigned long long reach = (unsigned long long)a.w * a.h + kBpTile * kBpSubMax; // one past the largest p0
        dv.row_magic = (a.w >= 2 && reach < (1ull << 24) && reach * (unsigned long long)a.w < (1ull << 40))
                           ? (1ull << 40) / (unsigned long long)a.w + 1ull : 0ull;
    }
    // depths are 1..65535 and image coordinates a few thousand at most: with divisors inside 2^+-60 nothing can
    // overflow or underflow in the FMA sequence
    const bool fast = fast_div_ok(a.K.depth_scale) && fast_div_ok(a.K.fx_u) && fast_div_ok(a.K.fx_v) &&
                      fabsf(a.K.cx_u) < 1.0e6f && fabsf(a.K.cx_v) < 1.0e6f;
#define ICPB_BP_LAUNCH_SUB(R, S)                                                                             \
    do {                                                                                                    \
        if (a.bgr && fast) backproject_kernel<R, true, true, S><<<cgrid, kBpThreads, 0, s>>>(a, dv, epoch); \
        else if (a.bgr) backproject_kernel<R, true, false, S><<<cgrid, kBpThreads, 0, s>>>(a, dv, epoch);   \
        else if (fast) backproject_kernel<R, false, true, S><<<cgrid, kBpThreads, 0, s>>>(a, dv, epoch);    \
        else backproject_kernel<R, false, false, S><<<cgrid, kBpThreads, 0, s>>>(a, dv, epoch);             \
    } while (0)
#define ICPB_BP_LAUNCH(R)                                                                                   \
    do {                                                                                                    \
        if (sub4) ICPB_BP_LAUNCH_SUB(R, 4);                                                                 \
        else ICPB_BP_LAUNCH_SUB(R, 2);                                                                      \
    } while (0)
#define ICPB_BP_TWO_PASS(R)                                                                                 \
    do {                                                                                                    \
        backproject_count_kernel<R><<<grid, kBpThreads, 0, s>>>(a);                                         \
        if (a.bgr && fast) backproject_write_kernel<R, true, true><<<grid, kBpThreads, 0, s>>>(a, dv);      \
        else if (a.bgr) backproject_write_kernel<R, true, false><<<grid, kBpThreads, 0, s>>>(a, dv);        \
        else if (fast) backproject_write_kernel<R, false, true><<<grid, kBpThreads, 0, s>>>(a, dv);         \
        else backproject_write_kernel<R, false, false><<<grid, kBpThreads, 0, s>>>(a, dv);                  \
        clear_two_pass_counts(a, s);                                                                        \
    } while (0)
    // Default: the single-launch chained-scan kernel (0.349 vs 0.359 ms for 256 frames).  The two-pass kernels never
    // wait on another CTA: they are the retry path when a chained launch reports that a predecessor never published
    // (out_count < 0), and can be forced with ICPB_BP_TWOPASS=1.
    static const bool env_two_pass = getenv("ICPB_BP_TWOPASS") && atoi(getenv("ICPB_BP_TWOPASS")) != 0;
    const bool two_pass = env_two_pass || force_two_pass;
    switch (a.rule) {
    case ICPB_SUB_STRIDE: ICPB_BP_LAUNCH(ICPB_SUB_STRIDE); break; // ordinal-dependent decisions: chained scans
    case ICPB_SUB_STREAM: ICPB_BP_LAUNCH(ICPB_SUB_STREAM); break;
    case ICPB_SUB_HASH: if (two_pass) ICPB_BP_TWO_PASS(ICPB_SUB_HASH); else ICPB_BP_LAUNCH(ICPB_SUB_HASH); break;
    default: if (two_pass) ICPB_BP_TWO_PASS(ICPB_SUB_NONE); else ICPB_BP_LAUNCH(ICPB_SUB_NONE); break;
    }
#undef ICPB_BP_LAUNCH
#undef ICPB_BP_LAUNCH_SUB
#undef ICPB_BP_TWO_PASS
}

// P3, SLAM.cpp:412-430.  Central differences on raw depth units; normalize as
// cv::normalize(Vec3f): v * (1.0 / sqrt(sum v^2)) in double, rounded to float.
// Border rows / cols (never written, or read out of bounds, by the reference)
// are defined as zeros.
__global__ void normals_kernel(const uint16_t *__restrict__ depth, int w, int h, float *__restrict__ normals)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int r = blockIdx.y;
    if (c >= w) return;
    float n0 = 0.f, n1 = 0.f, n2 = 0.f;
    if (r >= 1 && r < h - 1 && c >= 1 && c < w - 1) {
        float up = (float)depth[(size_t)(r - 1) * w + c];
        float dn = (float)depth[(size_t)(r + 1) * w + c];
        float lf = (float)depth[(size_t)r * w + c - 1];
        float rt = (float)depth[(size_t)r * w + c + 1];
        float dzdx = (dn - up) / 2.0f;
        float dzdy = (rt - lf) / 2.0f;
        float v0 = -dzdx, v1 = -dzdy, v2 = 1.0f;
        double nv = sqrt(((double)v0 * (double)v0 + (double)v1 * (double)v1) + (double)v2 * (double)v2);
        double inv = 1.0 / nv;
        n0 = (float)((double)v0 * inv);
        n1 = (float)((double)v1 * inv);
        n2 = (float)((double)v2 * inv);
    }
    float *o = normals + ((size_t)r * w + c) * 3;
    o[0] = n0; o[1] = n1; o[2] = n2;
}

// The same stencil, four pixels of a row per thread (image widths that are multiples of 4): the rows above and below
// arrive as one 8-byte load each, the centre row as one 8-byte load plus its two outer neighbours, and the 12 floats of
// the four normals leave as three 16-byte stores.  blockIdx.z = frame of a batch.
__device__ __forceinline__ void normal_of(float up, float dn, float lf, float rt, bool inside, float &n0, float &n1,
                                          float &n2)
{
    n0 = n1 = n2 = 0.f;
    if (!inside) return;
    const float dzdx = (dn - up) / 2.0f;
    const float dzdy = (rt - lf) / 2.0f;
    const float v0 = -dzdx, v1 = -dzdy, v2 = 1.0f;
    const double nv = sqrt(((double)v0 * (double)v0 + (double)v1 * (double)v1) + (double)v2 * (double)v2);
    const double inv = 1.0 / nv;
    n0 = (float)((double)v0 * inv);
    n1 = (float)((double)v1 * inv);
    n2 = (float)((double)v2 * inv);
}

// kNormRows rows per thread: the kNormRows + 2 depth rows a thread needs are all requested before the first normal is
// computed (one round trip for the lot; the rows shared between a thread's normals are fetched once), and a CTA spans a
// whole image row when the width allows, so a launch is tens of thousands of CTAs rather than a quarter of a million
// 128-thread ones.
constexpr int kNormRows = 2; // measured on 256 resident frames: 1 row 4,322 GB/s, 2 rows 4,776, 3 rows 4,319, 4 rows 4,268

__global__ void __launch_bounds__(256) normals4_kernel(const uint16_t *__restrict__ depth, int w, int h,
                                                       float *__restrict__ normals)
{
    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int r0 = blockIdx.y * kNormRows;
    if (c0 >= w) return;
    depth += (size_t)blockIdx.z * w * h;
    normals += (size_t)blockIdx.z * w * h * 3;
    // rows r0 - 1 .. r0 + kNormRows (those inside the image): four pixels each, plus the two outer neighbours
    uint2 row[kNormRows + 2];
    uint32_t lf[kNormRows + 2], rt[kNormRows + 2];
#pragma unroll
    for (int j = 0; j < kNormRows + 2; ++j) {
        const int r = r0 - 1 + j;
        row[j] = make_uint2(0u, 0u);
        lf[j] = rt[j] = 0u;
        if (r >= 0 && r < h) {
            const uint16_t *p = depth + (size_t)r * w + c0;
            row[j] = *reinterpret_cast<const uint2 *>(p);
            if (j >= 1 && j <= kNormRows) { // only centre rows look sideways
                if (c0 > 0) lf[j] = p[-1];
                if (c0 + 4 < w) rt[j] = p[4];
            }
        }
    }
#pragma unroll
    for (int j = 1; j <= kNormRows; ++j) {
        const int r = r0 - 1 + j;
        if (r >= h) break;
        float out[12];
        if (r >= 1 && r < h - 1) {
            const uint2 upw = row[j - 1], dnw = row[j + 1], ctw = row[j];
            const float up[4] = {(float)(upw.x & 0xffffu), (float)(upw.x >> 16), (float)(upw.y & 0xffffu), (float)(upw.y >> 16)};
            const float dn[4] = {(float)(dnw.x & 0xffffu), (float)(dnw.x >> 16), (float)(dnw.y & 0xffffu), (float)(dnw.y >> 16)};
            float ct[6];
            ct[0] = (float)lf[j];
            ct[1] = (float)(ctw.x & 0xffffu); ct[2] = (float)(ctw.x >> 16);
            ct[3] = (float)(ctw.y & 0xffffu); ct[4] = (float)(ctw.y >> 16);
            ct[5] = (float)rt[j];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = c0 + k;
                normal_of(up[k], dn[k], ct[k], ct[k + 2], c >= 1 && c < w - 1, out[3 * k], out[3 * k + 1], out[3 * k + 2]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 12; ++k) out[k] = 0.f;
        }
        float4 *o = reinterpret_cast<float4 *>(normals + ((size_t)r * w + c0) * 3);
        o[0] = make_float4(out[0], out[1], out[2], out[3]);
        o[1] = make_float4(out[4], out[5], out[6], out[7]);
        o[2] = make_float4(out[8], out[9], out[10], out[11]);
    }
}

void launch_normals(const uint16_t *depth, int w, int h, float *normals, cudaStream_t s, int frames)
{
    if ((w & 3) == 0 && ((uintptr_t)depth & 7) == 0 && ((uintptr_t)normals & 15) == 0) {
        const int quads = w / 4;
        const int threads = std::min(256, (quads + 31) / 32 * 32); // 640 -> 160, 512 -> 128: one CTA per row strip
        dim3 grid((quads + threads - 1) / threads, (h + kNormRows - 1) / kNormRows, frames);
        normals4_kernel<<<grid, threads, 0, s>>>(depth, w, h, normals);
        return;
    }
    for (int f = 0; f < frames; ++f) { // odd widths: one pixel per thread
        dim3 grid((w + 127) / 128, h);
        normals_kernel<<<grid, 128, 0, s>>>(depth + (size_t)f * w * h, w, h, normals + (size_t)f * w * h * 3);
    }
}

// 8f-1, SLAM.cpp:553-573: range threshold (:559-565), then 5x5 rect dilate and
// erode anchored at (3,3) (:567-573): window offsets [-3,+1] on both axes,
// out-of-image pixels ignored.
template <bool kDilate, bool kThreshold>
__global__ void morph5_kernel(const uint16_t *__restrict__ in, uint16_t *__restrict__ out, int w, int h, int min_d,
                              int max_d)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    int best = kDilate ? 0 : 65535;
    for (int dy = -3; dy <= 1; ++dy) {
        int yy = y + dy;
        if (yy < 0 || yy >= h) continue;
        for (int dx = -3; dx <= 1; ++dx) {
            int xx = x + dx;
            if (xx < 0 || xx >= w) continue;
            int v = in[(size_t)yy * w + xx];
            if (kThreshold) v = (v > max_d || v < min_d) ? 0 : v;
            best = kDilate ? max(best, v) : min(best, v);
        }
    }
    out[(size_t)y * w + x] = (uint16_t)best;
}

void launch_depth_filter(const uint16_t *in, uint16_t *tmp_a, uint16_t *tmp_b, uint16_t *out, int w, int h,
                         int min_d, int max_d, cudaStream_t s)
{
    (void)tmp_b;
    dim3 grid((w + 127) / 128, h);
    morph5_kernel<true, true><<<grid, 128, 0, s>>>(in, tmp_a, w, h, min_d, max_d);
    morph5_kernel<false, false><<<grid, 128, 0, s>>>(tmp_a, out, w, h, min_d, max_d);
}

} // namespace icpb

// Device-resident uint8 certainty grid.
//
// Replaces (reference file:line):
//   Map::getVoxelCoordinates   map.cpp:55-85                      (M2)
//   Map::update overloads      map.cpp:88-119, 122-151, 220-269   (M3, certainty part)
//   Map::rayTrace              map.cpp:272-439                    (M4, semantics in DESIGN.md)
//
// Both update rules are pure functions of the voxel byte, so k hits on one
// voxel give f^k(c) in any order: the kernels apply them with 32-bit CAS on
// the word holding four packed voxels and the grid is bit-exact regardless
// of scheduling.
#include <cstdlib>

#include "icpb_internal.h"

namespace icpb {

// map.cpp:60-62 + :65-82: float true division, truncation toward zero, clamp.
__device__ __forceinline__ int voxel_axis(float p, float cell, int dim)
{
    int q = (int)(p / cell);
    if (q < 0) q = 0;
    if (q >= dim) q = dim - 1;
    return q;
}

__device__ __forceinline__ uint32_t rule_apply(uint32_t c, int rule, int delta, int max_conf)
{
    if (rule == ICPB_RULE_A) return (c > (uint32_t)(255 - delta)) ? 255u : c + delta;            // map.cpp:249-253
    return ((int)c >= max_conf - delta) ? 255u : ((c + delta) & 0xffu);                           // map.cpp:139-149
}

// CAS on the 32-bit word holding voxel `lin`; fn maps the old byte to the new byte.
template <typename F>
__device__ __forceinline__ void byte_rmw(uint8_t *grid, size_t lin, F fn)
{
    uint32_t *word = reinterpret_cast<uint32_t *>(grid) + (lin >> 2);
    const uint32_t shift = (uint32_t)(lin & 3) * 8;
    uint32_t old = *word;
    while (true) {
        uint32_t c = (old >> shift) & 0xffu;
        uint32_t nc = fn(c) & 0xffu;
        if (nc == c) return;
        uint32_t want = (old & ~(0xffu << shift)) | (nc << shift);
        uint32_t seen = atomicCAS(word, old, want);
        if (seen == old) return;
        old = seen;
    }
}

// ---- brick occupancy (icpb_internal.h): bit (bx*nby + by)*nbz + bz covers voxels [8bx, 8bx+8) x [8by, ..) x [8bz, ..)
__device__ __forceinline__ unsigned brick_bit(const MapDev &m, int x, int y, int zr)
{
    return (unsigned)(((x >> kBrickLog) * m.nby + (y >> kBrickLog)) * m.nbz + (zr >> kBrickLog));
}

__device__ __forceinline__ unsigned brick2_bit(const MapDev &m, int x, int y, int zr)
{
    return (unsigned)(((x >> kBrick2Log) * m.nby2 + (y >> kBrick2Log)) * m.nbz2 + (zr >> kBrick2Log));
}

// a voxel of the brick became (or stays) non-zero; the read keeps the common case free of atomics
__device__ __forceinline__ void brick_mark(const MapDev &m, int x, int y, int zr)
{
    const unsigned b = brick_bit(m, x, y, zr);
    uint32_t *w = m.bricks + (b >> 5);
    const uint32_t mask = 1u << (b & 31);
    if (!(*reinterpret_cast<volatile uint32_t *>(w) & mask)) {
        atomicOr(w, mask);
        const unsigned b2 = brick2_bit(m, x, y, zr);
        uint32_t *w2 = m.bricks2 + (b2 >> 5);
        const uint32_t mask2 = 1u << (b2 & 31);
        if (!(*reinterpret_cast<volatile uint32_t *>(w2) & mask2)) atomicOr(w2, mask2);
    }
}

// icpb_map_upload: recompute every bit from the grid.  One thread per brick column segment would be unbalanced for
// thin slabs; one thread per brick, 8x8 rows of up to 8 bytes each, is plenty for a call that follows a host copy.
__global__ void map_rebuild_bricks_kernel(MapDev m, int nbx)
{
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nb = (long long)nbx * m.nby * m.nbz;
    if (b >= nb) return;
    const int bz = (int)(b % m.nbz), by = (int)((b / m.nbz) % m.nby), bx = (int)(b / ((long long)m.nbz * m.nby));
    bool any = false;
    for (int x = bx * kBrick; x < min((bx + 1) * kBrick, m.dims[0]) && !any; ++x)
        for (int y = by * kBrick; y < min((by + 1) * kBrick, m.dims[1]) && !any; ++y) {
            const uint8_t *row = m.grid + ((long long)x * m.dims[1] + y) * m.zs;
            for (int z = bz * kBrick; z < min((bz + 1) * kBrick, m.zs); ++z) any |= row[z] != 0;
        }
    if (any) {
        atomicOr(m.bricks + (b >> 5), 1u << (b & 31));
        const unsigned b2 = brick2_bit(m, bx * kBrick, by * kBrick, bz * kBrick);
        atomicOr(m.bricks2 + (b2 >> 5), 1u << (b2 & 31));
    }
}

void launch_map_rebuild_bricks(const MapDev &m, long long brick_words, cudaStream_t s)
{
    cudaMemsetAsync(m.bricks, 0, sizeof(uint32_t) * (size_t)brick_words, s);
    const int nbx = (m.dims[0] + kBrick - 1) / kBrick;
    const long long nb = (long long)nbx * m.nby * m.nbz;
    map_rebuild_bricks_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, s>>>(m, nbx);
}

// ---- PointSrc (icpb_internal.h): per-thread view of where point i lives --------------------------------------------
struct SrcView {
    int total;
};

__device__ __forceinline__ int band_count(const PointSrc &s, int b)
{
    return max(0, min(s.band_cap, __ldg(reinterpret_cast<const int *>(s.pts + (long long)b * s.band_stride))));
}

__device__ __forceinline__ void src_open(const PointSrc &s, SrcView &v)
{
    if (s.bands == 0) {
        // a negative live count (-1: the lift's look-back timed out) means "no points"
        v.total = s.n_dev ? max(0, min(s.n, *s.n_dev)) : s.n;
        return;
    }
    int acc = 0;
    for (int b = 0; b < s.bands; ++b) acc += band_count(s, b);
    v.total = acc;
}

// banded: the headers are a few L1-resident words; walking them per fetched point keeps them out of the registers of
// the long-running ray kernel (a prefix array held for the kernel's lifetime cost it two resident CTAs per SM)
__device__ __forceinline__ float4 src_point(const PointSrc &s, const SrcView &, int i)
{
    if (s.bands == 0) return s.pts[i];
    int b = 0;
    for (; b < s.bands - 1; ++b) {
        const int c = band_count(s, b);
        if (i < c) break;
        i -= c;
    }
    return s.pts[(long long)b * s.band_stride + 1 + i];
}

__global__ void map_endpoints_kernel(MapDev m, PointSrc src, int rule, int delta, int max_conf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    SrcView sv;
    src_open(src, sv);
    const int n = sv.total;
    const int lane = threadIdx.x & 31;
    long long lin = -1;
    int vx = 0, vy = 0, vz = 0;
    if (i < n) {
        float4 p = src_point(src, sv, i);
        vx = voxel_axis(p.x, m.cell, m.dims[0]);
        vy = voxel_axis(p.y, m.cell, m.dims[1]);
        vz = voxel_axis(p.z, m.cell, m.dims[2]);
        if (vz >= m.z_lo && vz < m.z_hi) lin = ((long long)vx * m.dims[1] + vy) * m.zs + (vz - m.z_lo);
    }
    // neighbouring pixels land in the same voxel: one CAS per distinct voxel per warp
    const unsigned peers = __match_any_sync(0xffffffffu, lin);
    if (lin < 0) return;
    if ((__ffs(peers) - 1) != lane) return;
    const int k = __popc(peers);
    // a positive delta leaves the voxel non-zero whatever it held: the brick is occupied from now on.  Marked BEFORE the
    // voxel write; the ray walk of the same frame has finished (phase 1 precedes phase 2 on the stream).
    if (delta > 0) brick_mark(m, vx, vy, vz - m.z_lo);
    byte_rmw(m.grid, (size_t)lin, [&](uint32_t c) {
        for (int r = 0; r < k && c != 255u; ++r) c = rule_apply(c, rule, delta, max_conf);
        return c;
    });
}

void launch_map_endpoints(const MapDev &m, const PointSrc &src, int rule, int delta, int max_conf, cudaStream_t s)
{
    if (src.n <= 0) return;
    map_endpoints_kernel<<<(src.n + 255) / 256, 256, 0, s>>>(m, src, rule, delta, max_conf);
}

// M3 with the reference's lookup-table / mapCloud bookkeeping (map.cpp:101-113, 136-149, 246-259).  One CTA,
// n <= 65536 points (key-points in the reference).  For point i: rank = hits of lower index on the same voxel,
// k = all hits on it.  The hit of rank 0 writes f^k(c0) (one writer per voxel, no atomics); the hit whose rank
// equals r* = the first rank satisfying the variant's condition inserts, if the voxel has no table entry yet.
// Inserted points are appended to dst in point order (block scans over chunks of 1024).
constexpr int kTrackThreads = 1024;

__device__ __forceinline__ uint32_t track_apply(uint32_t c, int variant, int delta, int max_conf)
{
    if (variant == ICPB_TRACK_NONASSOC) return ((int)c >= max_conf - delta) ? 255u : ((c + delta) & 0xffu);
    return (c > (uint32_t)(255 - delta)) ? 255u : c + delta;
}

__global__ void __launch_bounds__(kTrackThreads) map_tracked_kernel(MapDev m, int *table, const float4 *__restrict__ pts,
                                                                    int n, int variant, int delta, int max_conf,
                                                                    float4 *dst, int dst_n, int dst_capacity,
                                                                    int *d_appended, long long *vox_scratch, int table_base)
{
    __shared__ long long s_vox[kTrackThreads];
    __shared__ int s_warp[33];
    __shared__ int s_base;
    const int tid = threadIdx.x;
    // voxel of every point (scratch in global memory: n <= 65536)
    for (int i = tid; i < n; i += kTrackThreads) {
        float4 p = pts[i];
        int vx = voxel_axis(p.x, m.cell, m.dims[0]), vy = voxel_axis(p.y, m.cell, m.dims[1]);
        int vz = voxel_axis(p.z, m.cell, m.dims[2]);
        vox_scratch[i] = ((long long)vx * m.dims[1] + vy) * m.zs + vz;
    }
    if (tid == 0) s_base = dst_n;
    __syncthreads();
    // ---- phase A: decide everything from the PRE-update grid / table; nothing is written to them yet
    int *flags = reinterpret_cast<int *>(vox_scratch + 65536); // bit 0 insert, bit 1 writer, bits 8..15 new value
    for (int c0 = 0; c0 < n; c0 += kTrackThreads) {
        const int i = c0 + tid;
        const long long v = (i < n) ? vox_scratch[i] : -1;
        int rank = 0, k = 0;
        for (int t0 = 0; t0 < n; t0 += kTrackThreads) {
            __syncthreads();
            s_vox[tid] = (t0 + tid < n) ? vox_scratch[t0 + tid] : -2;
            __syncthreads();
            const int lim = min(kTrackThreads, n - t0);
            for (int j = 0; j < lim; ++j) {
                const bool same = (s_vox[j] == v);
                k += same;
                rank += same && (t0 + j < i);
            }
        }
        if (i < n) {
            // values before hit r: c_r = f^r(c0); r* = first hit rank whose condition holds
            uint32_t c = m.grid[v];
            int rstar = -1;
            for (int r = 0; r < k; ++r) {
                const uint32_t after = track_apply(c, variant, delta, max_conf);
                bool cond;
                if (variant == ICPB_TRACK_INIT) cond = (int)after >= max_conf;
                else if (variant == ICPB_TRACK_ASSOC) cond = c > (uint32_t)(255 - delta);
                else cond = (int)c >= max_conf - delta;
                if (cond && rstar < 0) rstar = r;
                c = after;
            }
            const int insert = (rank == rstar) && (table[v] < 0);
            flags[i] = insert | ((rank == 0) ? 2 : 0) | ((int)c << 8);
        }
    }
    __syncthreads();
    // ---- phase B: one writer per voxel stores f^k(c0); inserted points are appended in point order
    for (int c0 = 0; c0 < n; c0 += kTrackThreads) {
        const int i = c0 + tid;
        const int fl = (i < n) ? flags[i] : 0;
        const int insert = fl & 1;
        if (fl & 2) {
            const long long v = vox_scratch[i];
            m.grid[v] = (uint8_t)(fl >> 8);
            if ((fl >> 8) & 0xff) brick_mark(m, (int)(v / ((long long)m.zs * m.dims[1])), (int)((v / m.zs) % m.dims[1]), (int)(v % m.zs));
        }
        const int lane = tid & 31, wid = tid >> 5;
        int inc = insert;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { int o = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += o; }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int w = s_warp[lane], winc = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { int o = __shfl_up_sync(0xffffffffu, winc, off); if (lane >= off) winc += o; }
            s_warp[lane] = winc - w;
            if (lane == 31) s_warp[32] = winc;
        }
        __syncthreads();
        const int total = s_warp[32];
        const int pos = s_base + s_warp[wid] + inc - insert;
        if (insert && pos < dst_capacity) { // past the capacity nothing is recorded: the call fails with ICPB_ERR_CAPACITY
            dst[pos] = pts[i];
            table[vox_scratch[i]] = table_base + (pos - dst_n); // the caller's numbering of the stored points
        }
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
    }
    if (tid == 0) *d_appended = s_base - dst_n;
}

void launch_map_tracked(const MapDev &m, int *table, const float4 *pts, int n, int variant, int delta, int max_conf,
                        float4 *dst, int dst_n, int dst_capacity, int *d_appended, int table_base, cudaStream_t s)
{
    long long *scratch = reinterpret_cast<long long *>(d_appended + 2);
    map_tracked_kernel<<<1, kTrackThreads, 0, s>>>(m, table, pts, n, variant, delta, max_conf, dst, dst_n, dst_capacity,
                                                   d_appended, scratch, table_base);
}

// M4 phase 1: exact integer Amanatides-Woo walk from the origin voxel centre to the endpoint voxel centre.
// Axis k crosses its i-th wall at t = (2i+1)/(2 n_k); scaled by 2 P (P = product of max(n_k,1)) the wall times
// are the integers e_k = (2i+1) P / n_k; the smallest goes first, ties x < y < z.  An axis that has crossed all
// its walls carries e_k = (2 n_k + 1) P / n_k > 2 P, larger than every real wall time, and an axis with n_k = 0
// starts at 3 P: neither is ever selected, so the loop needs no per-axis counters.  Every voxel entered except
// the endpoint voxel is decremented with clamp at 0.  I = int when 3 * dims product < 2^31, else long long.
//
// Scheduling: persistent warps, one ray per LANE, rays handed out from a global counter.  Walk lengths differ by
// 2-4x between near and far surfaces; with one thread per ray a warp ran at 21.8 of 32 lanes and the last wave left
// half the SMs idle.  Here a lane that finishes its ray waits only until 8 lanes of its warp are idle, then the
// warp draws new rays for all of them at once (the set-up, ~100 instructions, is amortised over >= 8 lanes).  The
// certainty updates commute (g^k), so the order in which rays are walked does not matter.
#ifndef ICPB_RAY_REFILL
#define ICPB_RAY_REFILL 8
#endif
#ifndef ICPB_RAY_BURST
#define ICPB_RAY_BURST 16
#endif
#ifndef ICPB_RAY_GROUP
#define ICPB_RAY_GROUP 4
#endif
constexpr int kRayRefill = ICPB_RAY_REFILL; // idle lanes that trigger a refill
constexpr int kRayBurst = ICPB_RAY_BURST;   // walk steps between two refill checks
constexpr int kRayGroup = ICPB_RAY_GROUP;   // steps whose voxel reads are issued together

template <typename I>
struct RayState {
    I ex, ey, ez, dxs, dys, dzs;
    long long lin, stx, sty;
    int sz, zrel, rem; // rem: walk steps still to take
};

// Set-up of ray i (also performs the visit of the slab-entry voxel when the walk is entered by a jump).
template <typename I>
__device__ __forceinline__ void ray_setup(const MapDev &m, const float4 p, int ox, int oy, int oz, int delta_dec,
                                          RayState<I> &r, unsigned long long &visits)
{
    const int ex_ = voxel_axis(p.x, m.cell, m.dims[0]);
    const int ey_ = voxel_axis(p.y, m.cell, m.dims[1]);
    const int ez_ = voxel_axis(p.z, m.cell, m.dims[2]);
    const int nx = abs(ex_ - ox), ny = abs(ey_ - oy), nz = abs(ez_ - oz);
    const int sx = (ex_ > ox) - (ex_ < ox), sy = (ey_ > oy) - (ey_ < oy), sz = (ez_ > oz) - (ez_ < oz);
    const I mx = nx ? nx : 1, my = ny ? ny : 1, mz = nz ? nz : 1;
    const I P3 = 3 * mx * my * mz;
    r.ex = nx ? my * mz : P3; r.ey = ny ? mx * mz : P3; r.ez = nz ? mx * my : P3;
    r.dxs = 2 * my * mz; r.dys = 2 * mx * mz; r.dzs = 2 * mx * my;
    const int steps = nx + ny + nz;
    visits += steps > 0 ? (unsigned long long)(steps - 1) : 0ull;
    int z = oz;
    int done_steps = 0;
    bool live = true, entered_now = false;
    long long lin0 = ((long long)ox * m.dims[1] + oy) * m.zs + (oz - m.z_lo);
    // Slab clipping: z moves monotonically, so the part of the walk inside [z_lo, z_hi) is one contiguous
    // range of steps.  If the origin is outside the slab, jump to the state right after the z-step that
    // enters it: that is z-step number k; by then every x / y wall with time <= (2k-1)/(2 nz) has been
    // crossed (x and y go first on ties), i.e. floor(((2k-1) n + nz) / (2 nz)) of them.
    if (oz < m.z_lo || oz >= m.z_hi) {
        long long k = 0;
        if (sz > 0 && oz < m.z_lo) k = (long long)m.z_lo - oz;
        else if (sz < 0 && oz >= m.z_hi) k = (long long)oz - (m.z_hi - 1);
        if (k <= 0 || k > nz) live = false; // the ray never enters this slab
        else {
            const long long cz = k;
            const long long cx = nx ? min((long long)nx, ((2 * k - 1) * nx + nz) / (2LL * nz)) : 0;
            const long long cy = ny ? min((long long)ny, ((2 * k - 1) * ny + nz) / (2LL * nz)) : 0;
            z = oz + sz * (int)cz;
            lin0 = ((long long)(ox + sx * (int)cx) * m.dims[1] + (oy + sy * (int)cy)) * m.zs + (z - m.z_lo);
            if (nx) r.ex = (I)(2 * cx + 1) * my * mz;
            if (ny) r.ey = (I)(2 * cy + 1) * mx * mz;
            r.ez = (I)(2 * cz + 1) * mx * my;
            done_steps = (int)(cx + cy + cz);
            entered_now = true;
        }
    }
    r.lin = lin0;
    r.stx = (long long)sx * m.dims[1] * m.zs; r.sty = (long long)sy * m.zs;
    r.sz = sz;
    r.zrel = z - m.z_lo;
    r.rem = live ? max(steps - 1 - done_steps, 0) : 0;
    // the voxel just entered by the jump is itself a visit unless it is the endpoint
    if (live && entered_now && done_steps <= steps - 1 && m.grid[lin0] != 0)
        byte_rmw(m.grid, (size_t)lin0, [&](uint32_t c) { return c > (uint32_t)delta_dec ? c - delta_dec : 0u; });
}

template <typename I>
__global__ void __launch_bounds__(128) map_rays_kernel(MapDev m, PointSrc src, int ox, int oy,
                                                      int oz, int delta_dec, unsigned long long *visited,
                                                      unsigned int *next_ray)
{
    SrcView sv;
    src_open(src, sv);
    const int n = sv.total;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    RayState<I> r;
    r.ex = r.ey = r.ez = r.dxs = r.dys = r.dzs = 0;
    r.lin = r.stx = r.sty = 0;
    r.sz = r.zrel = r.rem = 0;
    unsigned long long my_visits = 0;
    bool drained = false; // warp-uniform: the counter has passed n
    uint8_t *g = m.grid;
    const unsigned zs = (unsigned)m.zs;
    while (true) {
        const unsigned idle = __ballot_sync(0xffffffffu, r.rem <= 0);
        if (!drained && __popc(idle) >= kRayRefill) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(next_ray, (unsigned)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (r.rem <= 0) {
                const unsigned i = base + (unsigned)__popc(idle & lt);
                if (i < (unsigned)n) ray_setup<I>(m, src_point(src, sv, (int)i), ox, oy, oz, delta_dec, r, my_visits);
            }
            drained = base + (unsigned)__popc(idle) >= (unsigned)n;
        }
        if (__ballot_sync(0xffffffffu, r.rem > 0) == 0u) {
            if (drained) break;
            continue; // every ray drawn this round was empty (never enters the slab): draw again
        }
        // kRayGroup steps at a time: first all the DDA advances (pure integer state), then all the voxel reads --
        // independent loads, in flight together -- then the rare decrements.  One step per load made the walk wait a
        // full L2 / DRAM latency per voxel (65 % of the stall samples sat on the first use of the loaded byte).
#pragma unroll 1
        for (int k = 0; k < kRayBurst / kRayGroup; ++k) {
            long long at[kRayGroup];
            bool in[kRayGroup];
#pragma unroll
            for (int u = 0; u < kRayGroup; ++u) {
                in[u] = r.rem > 0;
                if (in[u]) {
                    const bool px = (r.ex <= r.ey) && (r.ex <= r.ez);
                    const bool py = !px && (r.ey <= r.ez);
                    const bool pz = !px && !py;
                    r.lin += px ? r.stx : (py ? r.sty : (long long)r.sz);
                    r.ex += px ? r.dxs : (I)0;
                    r.ey += py ? r.dys : (I)0;
                    r.ez += pz ? r.dzs : (I)0;
                    r.zrel += pz ? r.sz : 0;
                    --r.rem;
                    if ((unsigned)r.zrel >= zs) { r.rem = 0; in[u] = false; } // left the slab for good
                }
                at[u] = r.lin;
            }
            uint8_t v[kRayGroup];
#pragma unroll
            for (int u = 0; u < kRayGroup; ++u) v[u] = in[u] ? g[at[u]] : (uint8_t)0;
            // phase 1 only lowers values, so a cached non-zero byte is at worst stale-high: the CAS re-reads it
#pragma unroll
            for (int u = 0; u < kRayGroup; ++u)
                if (v[u] != 0)
                    byte_rmw(g, (size_t)at[u], [&](uint32_t c) { return c > (uint32_t)delta_dec ? c - delta_dec : 0u; });
        }
    }
    if (visited) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) my_visits += __shfl_xor_sync(0xffffffffu, my_visits, off);
        if (lane == 0 && my_visits) atomicAdd(visited, my_visits);
    }
}

// ---- M4 phase 1 with brick skipping --------------------------------------------------------------------------
// The same walk, but a brick whose occupancy bit is clear holds only zero voxels, and decrementing a zero voxel is a
// no-op (map.cpp:423): the walk crosses such a brick in ONE jump instead of one dependent byte load per voxel (the
// byte-at-a-time kernel above spent 9.6 of every 10 issue slots waiting on those loads, profiles/r01_ncu_map_rays_*).
// The jump is computed from the walk's own integer state, so the voxels entered afterwards are exactly the ones the
// step-by-step walk enters: with k_a = steps along axis a that leave the brick (capped by the walls the ray still has
// to cross), the exit wall is the earliest of the three T_a = e_a + (k_a - 1) d_a in the walk's order (ties x < y < z);
// before it the other axes cross c_a = #{ j : e_a + j d_a precedes T } walls (<= for an axis that goes first on ties,
// < otherwise) -- two integer divisions.  The jump stops on the LAST voxel inside the brick; the ordinary step that
// follows enters the next brick and looks its bit up (only steps that cross a brick boundary do).  In occupied bricks
// the walk steps voxel by voxel as before, the byte reads of up to kRayGroup steps in flight together.
template <typename I>
struct RayBrick {
    I ex, ey, ez, dxs, dys, dzs;
    float rdx, rdy, rdz; // 1 / dxs ... : quotient estimates of the jump
    int x, y, zr;     // current voxel, z relative to the slab
    int rx, ry, rz;   // walls still to cross per axis
    int sgn;          // step signs + 1, two bits per axis
    int rem;          // visits still to make
    bool empty;       // the current voxel's brick (8^3) holds only zeros
    bool empty2;      // so does its coarse brick (32^3)
};

__device__ __forceinline__ bool brick_is_empty(const MapDev &m, int x, int y, int zr)
{
    const unsigned b = brick_bit(m, x, y, zr);
    return ((__ldg(m.bricks + (b >> 5)) >> (b & 31)) & 1u) == 0u;
}

__device__ __forceinline__ bool brick2_is_empty(const MapDev &m, int x, int y, int zr)
{
    const unsigned b = brick2_bit(m, x, y, zr);
    return ((__ldg(m.bricks2 + (b >> 5)) >> (b & 31)) & 1u) == 0u;
}

template <typename I>
__device__ __forceinline__ void ray_lookup(const MapDev &m, RayBrick<I> &r, bool coarse_too)
{
    if (coarse_too) r.empty2 = brick2_is_empty(m, r.x, r.y, r.zr);
    r.empty = r.empty2 || brick_is_empty(m, r.x, r.y, r.zr); // a clear coarse bit covers all its fine bricks
}

template <typename I>
__device__ __forceinline__ void ray_setup_brick(const MapDev &m, const float4 p, int ox, int oy, int oz, int delta_dec,
                                                RayBrick<I> &r, unsigned long long &visits)
{
    const int ex_ = voxel_axis(p.x, m.cell, m.dims[0]);
    const int ey_ = voxel_axis(p.y, m.cell, m.dims[1]);
    const int ez_ = voxel_axis(p.z, m.cell, m.dims[2]);
    const int nx = abs(ex_ - ox), ny = abs(ey_ - oy), nz = abs(ez_ - oz);
    const int sx = (ex_ > ox) - (ex_ < ox), sy = (ey_ > oy) - (ey_ < oy), sz = (ez_ > oz) - (ez_ < oz);
    const I mx = nx ? nx : 1, my = ny ? ny : 1, mz = nz ? nz : 1;
    const I P3 = 3 * mx * my * mz;
    r.ex = nx ? my * mz : P3; r.ey = ny ? mx * mz : P3; r.ez = nz ? mx * my : P3;
    r.dxs = 2 * my * mz; r.dys = 2 * mx * mz; r.dzs = 2 * mx * my;
    r.rdx = 1.0f / (float)r.dxs; r.rdy = 1.0f / (float)r.dys; r.rdz = 1.0f / (float)r.dzs;
    const int steps = nx + ny + nz;
    visits += steps > 0 ? (unsigned long long)(steps - 1) : 0ull;
    int cx = 0, cy = 0, cz = 0;
    bool live = true, entered_now = false;
    // slab clipping: see ray_setup above
    if (oz < m.z_lo || oz >= m.z_hi) {
        int k = 0;
        if (sz > 0 && oz < m.z_lo) k = m.z_lo - oz;
        else if (sz < 0 && oz >= m.z_hi) k = oz - (m.z_hi - 1);
        if (k <= 0 || k > nz) live = false;
        else {
            cz = k;
            // (2k - 1) n + nz < 2 dims^2: 32-bit for every grid the int walk serves, 64-bit otherwise
            if (sizeof(I) == 4) {
                cx = nx ? min(nx, (int)(((unsigned)(2 * k - 1) * (unsigned)nx + (unsigned)nz) / (2u * (unsigned)nz))) : 0;
                cy = ny ? min(ny, (int)(((unsigned)(2 * k - 1) * (unsigned)ny + (unsigned)nz) / (2u * (unsigned)nz))) : 0;
            } else {
                cx = nx ? (int)min((long long)nx, ((2LL * k - 1) * nx + nz) / (2LL * nz)) : 0;
                cy = ny ? (int)min((long long)ny, ((2LL * k - 1) * ny + nz) / (2LL * nz)) : 0;
            }
            if (nx) r.ex = (I)(2 * cx + 1) * my * mz;
            if (ny) r.ey = (I)(2 * cy + 1) * mx * mz;
            r.ez = (I)(2 * cz + 1) * mx * my;
            entered_now = true;
        }
    }
    r.x = ox + sx * cx; r.y = oy + sy * cy; r.zr = oz + sz * cz - m.z_lo;
    r.rx = nx - cx; r.ry = ny - cy; r.rz = nz - cz;
    r.sgn = (sx + 1) | ((sy + 1) << 2) | ((sz + 1) << 4);
    const int done_steps = cx + cy + cz;
    r.rem = live ? max(steps - 1 - done_steps, 0) : 0;
    r.empty = r.empty2 = false;
    if (live) {
        ray_lookup<I>(m, r, true);
        // the voxel just entered by the jump is itself a visit unless it is the endpoint
        if (entered_now && done_steps <= steps - 1 && !r.empty) {
            const long long lin0 = ((long long)r.x * m.dims[1] + r.y) * m.zs + r.zr;
            if (m.grid[lin0] != 0)
                byte_rmw(m.grid, (size_t)lin0, [&](uint32_t c) { return c > (uint32_t)delta_dec ? c - delta_dec : 0u; });
        }
    }
}

// walls j in [0, r) of an axis with e + j d before time T: "<= T" when the axis goes first on ties, "< T" otherwise.
// floor(lim / d) from a float estimate settled by the exact remainder: the quotient is below the axis' wall count
// (<= a grid dimension), far inside the range where the estimate is within one of it.
template <typename I>
__device__ __forceinline__ int walls_before(I e, I d, float rcp_d, int r, I T, bool first_on_ties)
{
    const I lim = T - e - (first_on_ties ? (I)0 : (I)1);
    if (lim < 0) return 0;
    int c = (int)((float)lim * rcp_d);
    const I rmd = lim - (I)c * d;
    c += (rmd >= d) - (rmd < 0);
    return min(c + 1, r);
}

// LOG = log2 of the brick edge: kBrickLog for the fine level, kBrick2Log for the coarse one
template <typename I, int LOG>
__device__ __forceinline__ void ray_jump(const MapDev &m, RayBrick<I> &r)
{
    constexpr int B = 1 << LOG;
    const int sx = (r.sgn & 3) - 1, sy = ((r.sgn >> 2) & 3) - 1, sz = ((r.sgn >> 4) & 3) - 1;
    const int lx = r.x & (B - 1), ly = r.y & (B - 1), lz = r.zr & (B - 1);
    // steps along each axis that leave the brick (the slab's upper face also ends the z-run)
    const int kx = sx > 0 ? B - lx : lx + 1;
    const int ky = sy > 0 ? B - ly : ly + 1;
    const int kz = sz > 0 ? min(B - lz, m.zs - r.zr) : lz + 1;
    const bool vx = kx <= r.rx, vy = ky <= r.ry, vz = kz <= r.rz; // r_a = 0 for an axis that does not move
    const I big = sizeof(I) == 4 ? (I)0x7fffffff : (I)0x7fffffffffffffffLL;
    const I Tx = vx ? r.ex + (I)(kx - 1) * r.dxs : big;
    const I Ty = vy ? r.ey + (I)(ky - 1) * r.dys : big;
    const I Tz = vz ? r.ez + (I)(kz - 1) * r.dzs : big;
    if (!(vx || vy || vz)) { r.rem = 0; return; } // the ray ends inside this (empty) brick
    const bool bx = (Tx <= Ty) && (Tx <= Tz);
    const bool by = !bx && (Ty <= Tz);
    const I T = bx ? Tx : (by ? Ty : Tz);
    // walls crossed inside the brick: all but the last of the exit axis, and whatever the others cross before T
    const int cx = bx ? kx - 1 : walls_before<I>(r.ex, r.dxs, r.rdx, r.rx, T, true);
    const int cy = by ? ky - 1 : walls_before<I>(r.ey, r.dys, r.rdy, r.ry, T, !bx);
    const int cz = (!bx && !by) ? kz - 1 : walls_before<I>(r.ez, r.dzs, r.rdz, r.rz, T, false);
    const int skip = cx + cy + cz;
    if (skip >= r.rem) { r.rem = 0; return; } // every remaining visit lies in the empty brick
    r.rem -= skip;
    r.x += sx * cx; r.y += sy * cy; r.zr += sz * cz;
    r.rx -= cx; r.ry -= cy; r.rz -= cz;
    r.ex += (I)cx * r.dxs; r.ey += (I)cy * r.dys; r.ez += (I)cz * r.dzs;
}

// layer_work (nullable, dims[2] words): calibration of the z-slab boundaries -- every step adds 1 and every jump 4 (their
// relative instruction cost) to the z-layer it happens in, so that equal shares of the histogram are equal shares of
// this kernel's work (icpb_map_integrate_rays_profiled).
template <typename I>
__global__ void __launch_bounds__(128) map_rays_brick_kernel(MapDev m, PointSrc src, int ox, int oy,
                                                            int oz, int delta_dec, unsigned long long *visited,
                                                            unsigned int *next_ray, unsigned int *layer_work)
{
    SrcView sv;
    src_open(src, sv);
    const int n = sv.total;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    RayBrick<I> r;
    r.ex = r.ey = r.ez = r.dxs = r.dys = r.dzs = 0;
    r.rdx = r.rdy = r.rdz = 0.f;
    r.x = r.y = r.zr = r.rx = r.ry = r.rz = 0;
    r.sgn = 0x15;
    r.rem = 0;
    r.empty = r.empty2 = false;
    unsigned long long my_visits = 0;
    bool drained = false; // warp-uniform: the counter has passed n
    uint8_t *g = m.grid;
    const unsigned zs = (unsigned)m.zs;
    const long long ystride = m.zs, xstride = (long long)m.dims[1] * m.zs;
    while (true) {
        const unsigned idle = __ballot_sync(0xffffffffu, r.rem <= 0);
        if (!drained && __popc(idle) >= kRayRefill) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(next_ray, (unsigned)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (r.rem <= 0) {
                const unsigned i = base + (unsigned)__popc(idle & lt);
                if (i < (unsigned)n) {
                    ray_setup_brick<I>(m, src_point(src, sv, (int)i), ox, oy, oz, delta_dec, r, my_visits);
                    if (layer_work && r.rem > 0) atomicAdd(&layer_work[r.zr + m.z_lo], 6u);
                }
            }
            drained = base + (unsigned)__popc(idle) >= (unsigned)n;
        }
        if (__ballot_sync(0xffffffffu, r.rem > 0) == 0u) {
            if (drained) break;
            continue;
        }
#pragma unroll 1
        for (int k = 0; k < kRayBurst / kRayGroup; ++k) {
            if (r.rem > 0 && r.empty) {
                if (layer_work) atomicAdd(&layer_work[r.zr + m.z_lo], 4u);
                if (r.empty2) ray_jump<I, kBrick2Log>(m, r);
                else ray_jump<I, kBrickLog>(m, r);
            }
            const int sx = (r.sgn & 3) - 1, sy = ((r.sgn >> 2) & 3) - 1, sz = ((r.sgn >> 4) & 3) - 1;
            long long at[kRayGroup];
            bool in[kRayGroup];
            bool halt = false; // the walk has just entered an empty brick: the next round jumps across it
#pragma unroll
            for (int u = 0; u < kRayGroup; ++u) {
                in[u] = false;
                at[u] = 0;
                if (r.rem > 0 && !halt) {
                    const bool px = (r.ex <= r.ey) && (r.ex <= r.ez);
                    const bool py = !px && (r.ey <= r.ez);
                    const bool pz = !px && !py;
                    r.x += px ? sx : 0; r.y += py ? sy : 0; r.zr += pz ? sz : 0;
                    r.rx -= px; r.ry -= py; r.rz -= pz;
                    r.ex += px ? r.dxs : (I)0;
                    r.ey += py ? r.dys : (I)0;
                    r.ez += pz ? r.dzs : (I)0;
                    --r.rem;
                    if ((unsigned)r.zr >= zs) { r.rem = 0; halt = true; } // left the slab for good
                    else {
                        if (layer_work) atomicAdd(&layer_work[r.zr + m.z_lo], 1u);
                        // the coordinate that moved tells whether a brick boundary (fine, coarse) was crossed
                        const int c = px ? r.x : (py ? r.y : r.zr);
                        const int s = px ? sx : (py ? sy : sz);
                        if ((c & (kBrick - 1)) == (s > 0 ? 0 : kBrick - 1))
                            ray_lookup<I>(m, r, (c & (kBrick2 - 1)) == (s > 0 ? 0 : kBrick2 - 1));
                        in[u] = !r.empty;
                        halt = r.empty;
                        at[u] = (long long)r.x * xstride + (long long)r.y * ystride + r.zr;
                    }
                }
            }
            uint8_t v[kRayGroup];
#pragma unroll
            for (int u = 0; u < kRayGroup; ++u) v[u] = in[u] ? g[at[u]] : (uint8_t)0;
            // phase 1 only lowers values, so a cached non-zero byte is at worst stale-high: the CAS re-reads it
#pragma unroll
            for (int u = 0; u < kRayGroup; ++u)
                if (v[u] != 0)
                    byte_rmw(g, (size_t)at[u], [&](uint32_t c) { return c > (uint32_t)delta_dec ? c - delta_dec : 0u; });
        }
    }
    if (visited) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) my_visits += __shfl_xor_sync(0xffffffffu, my_visits, off);
        if (lane == 0 && my_visits) atomicAdd(visited, my_visits);
    }
}

// layer_work == kCounterIsZero: no histogram, and next_ray already is zero (a fresh counter per frame of a sequence:
// one memset per sequence instead of one per frame)
unsigned int *const kCounterIsZero = reinterpret_cast<unsigned int *>(1);

void launch_map_rays(const MapDev &m, const PointSrc &src, const float origin[3], int delta_dec,
                     unsigned long long *visited, unsigned int *next_ray, int sm_count, cudaStream_t s, unsigned int *layer_work)
{
    const int n = src.n;
    if (n <= 0) return;
    if (!layer_work || layer_work != kCounterIsZero) cudaMemsetAsync(next_ray, 0, sizeof(unsigned int), s);
    if (layer_work == kCounterIsZero) layer_work = nullptr;
    // persistent grid: every resident warp slot is filled once (40 registers -> 12 CTAs of 128 threads per SM)
    const int blocks = min((n + 127) / 128, sm_count * 12);
    // origin voxel: same quantisation as any point (map.cpp:226 uses getVoxelCoordinates too)
    int o[3];
    for (int k = 0; k < 3; ++k) {
        int q = (int)(origin[k] / m.cell);
        if (q < 0) q = 0;
        if (q >= m.dims[k]) q = m.dims[k] - 1;
        o[k] = q;
    }
    const double prod = 3.0 * m.dims[0] * m.dims[1] * m.dims[2];
    static const bool bricks = []() { const char *e = getenv("ICPB_RAY_BRICKS"); return !(e && *e == '0'); }();
    if (bricks) { // default: cross empty bricks in one jump
        if (prod < 2147483647.0)
            map_rays_brick_kernel<int><<<blocks, 128, 0, s>>>(m, src, o[0], o[1], o[2], delta_dec, visited, next_ray, layer_work);
        else
            map_rays_brick_kernel<long long><<<blocks, 128, 0, s>>>(m, src, o[0], o[1], o[2], delta_dec, visited, next_ray, layer_work);
        return;
    }
    // ICPB_RAY_BRICKS=0: the byte-at-a-time walk (kept for A/B measurements)
    if (prod < 2147483647.0)
        map_rays_kernel<int><<<blocks, 128, 0, s>>>(m, src, o[0], o[1], o[2], delta_dec, visited, next_ray);
    else
        map_rays_kernel<long long><<<blocks, 128, 0, s>>>(m, src, o[0], o[1], o[2], delta_dec, visited, next_ray);
}

} // namespace icpb

// Exact nearest-neighbour search over a uniform cell grid of the target cloud (ICPB_NN_GRID).
//
// The reference's intended fix for the N x M cost is its voxel-indexed search
// (findMappedNearestNeighborAssociations / getNearestMappedPoint / processVoxel, icp.cpp:347-486), which is
// dead and reads out of bounds there.  An exact indexed search must return what the brute-force scan of
// icp.cpp:541-620 returns; this one does, for every query whose nearest neighbour is closer than
// MAX_NN_COLOR_DISTANCE (the only ones the registration consumes, icp.cpp:553):
//   * targets are bucketed into cubic cells (counting sort, x fastest, so a row of cells is one contiguous run);
//   * a query visits cells in growing Chebyshev shells around its own cell; a candidate is evaluated in the
//     reference's exact arithmetic whenever the FP32 filter cannot rule it out (same bound as nn.cu);
//   * the best (distance, original index) pair is final once it is strictly below a lower bound on the
//     distance to everything outside the visited cube -- nothing out there can win or tie;
//   * a query with nothing closer than the acceptance radius reports idx -1 / dist +inf (it is rejected by
//     icp.cpp:553 whatever its true neighbour is).
#include <math_constants.h>

#include "icpb_internal.h"

namespace icpb {

__device__ __forceinline__ unsigned int f2ord(float f)
{
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// bbox[0..2] = ordered-uint min, bbox[3..5] = ordered-uint max (initialised by the host to ~0u / 0u).  One set of six
// atomics per CTA (the first version issued them per warp: 49,000 atomics on six words, 35 us for 292k points).
__global__ void __launch_bounds__(256) grid_bbox_kernel(const float4 *__restrict__ pts, int m, unsigned int *bbox)
{
    __shared__ unsigned int s_lo[3][8], s_hi[3][8];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
    for (; i < m; i += gridDim.x * blockDim.x) {
        float4 p = pts[i];
        unsigned int v[3] = {f2ord(p.x), f2ord(p.y), f2ord(p.z)};
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = min(lo[k], v[k]); hi[k] = max(hi[k], v[k]); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = __reduce_min_sync(0xffffffffu, lo[k]);
        hi[k] = __reduce_max_sync(0xffffffffu, hi[k]);
        if ((threadIdx.x & 31) == 0) { s_lo[k][threadIdx.x >> 5] = lo[k]; s_hi[k][threadIdx.x >> 5] = hi[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int k = threadIdx.x;
        unsigned int a = s_lo[k][0], b = s_hi[k][0];
        for (int w = 1; w < 8; ++w) { a = min(a, s_lo[k][w]); b = max(b, s_hi[k][w]); }
        atomicMin(&bbox[k], a);
        atomicMax(&bbox[3 + k], b);
    }
}

__device__ __forceinline__ int cell_axis(float p, float mn, float h, int n)
{
    int c = (int)floorf((p - mn) / h);
    return min(max(c, 0), n - 1);
}

__device__ __forceinline__ int cell_of(const GridMeta &g, float x, float y, float z)
{
    int cx = cell_axis(x, g.mn[0], g.h, g.dim[0]);
    int cy = cell_axis(y, g.mn[1], g.h, g.dim[1]);
    int cz = cell_axis(z, g.mn[2], g.h, g.dim[2]);
    const int coarse = (cz * g.dim[1] + cy) * g.dim[0] + cx; // x fastest
    if (g.sub == 1) return coarse;
    // 2 x 2 x 2 children: which half of the cell along every axis (any consistent rule will do: the search prunes by the
    // tight boxes of the points actually assigned, not by the children's nominal cubes)
    const int sx = (x - (g.mn[0] + cx * g.h)) >= 0.5f * g.h, sy = (y - (g.mn[1] + cy * g.h)) >= 0.5f * g.h;
    const int sz = (z - (g.mn[2] + cz * g.h)) >= 0.5f * g.h;
    return coarse * 8 + (sz * 4 + sy * 2 + sx);
}

__global__ void grid_count_kernel(const float4 *__restrict__ pts, int m, GridMeta g, int *counts)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    float4 p = pts[i];
    atomicAdd(&counts[cell_of(g, p.x, p.y, p.z)], 1);
}

__global__ void grid_scatter_kernel(const float4 *__restrict__ pts, int m, GridMeta g, int *cursor, float4 *sorted)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    float4 p = pts[i];
    int pos = atomicAdd(&cursor[cell_of(g, p.x, p.y, p.z)], 1);
    p.w = __int_as_float(i); // original index; order inside a cell is irrelevant (lexicographic (d, index) minimum)
    sorted[pos] = p;
}

// ---- exclusive scan over the cell counts (three small kernels; plumbing) -------------------------------------
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ int block_scan_excl(int v, int *s_warp, int &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += o;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = s_warp[lane];
        int winc = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, winc, off);
            if (lane >= off) winc += o;
        }
        s_warp[lane] = winc - w; // exclusive over warps
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    total = s_warp[32];
    int r = s_warp[wid] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const int *in, int n, int *block_sums)
{
    __shared__ int s_warp[33];
    int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) v += (base + k < n) ? in[base + k] : 0;
    int total;
    block_scan_excl(v, s_warp, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(int *block_sums, int nblocks)
{
    __shared__ int s_warp[33];
    int carry = 0;
    for (int b0 = 0; b0 < nblocks; b0 += kScanThreads) {
        int i = b0 + threadIdx.x;
        int v = (i < nblocks) ? block_sums[i] : 0;
        int total;
        int ex = block_scan_excl(v, s_warp, total);
        if (i < nblocks) block_sums[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(int *data, int n, const int *block_sums)
{
    __shared__ int s_warp[33];
    int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int x[kScanItems], v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { x[k] = (base + k < n) ? data[base + k] : 0; v += x[k]; }
    int total;
    int ex = block_scan_excl(v, s_warp, total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) data[base + k] = ex;
        ex += x[k];
    }
}

void launch_grid_bbox(const float4 *tgt, int m, unsigned int *bbox, cudaStream_t s)
{
    int blocks = min((m + 255) / 256, 296);
    grid_bbox_kernel<<<blocks, 256, 0, s>>>(tgt, m, bbox);
}

// counts: ncells+1 ints, zeroed by the caller; on return counts[c] = start of cell c, counts[ncells] = m.
// Tight bounding box of the targets of every cell (min / max of their coordinates, exact): the cooperative search
// prunes a cell by the distance to this box, not to the cell's cube -- a surface patch fills a sliver of its cell.
// One warp per group of 32 cells, lane = cell; empty cells get an inverted box.
__global__ void grid_boxes_kernel(const float4 *__restrict__ sorted, const int *__restrict__ start, int ncells, float4 *boxes)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int t0 = start[c], t1 = start[c + 1];
    float lx = CUDART_INF_F, ly = CUDART_INF_F, lz = CUDART_INF_F, hx = -CUDART_INF_F, hy = -CUDART_INF_F, hz = -CUDART_INF_F;
    // eight loads in flight: a dense cell holds a hundred points and a thread walks them alone
    for (int t = t0; t < t1; t += 8) {
        float4 q[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) q[k] = __ldg(&sorted[min(t + k, t1 - 1)]);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            lx = fminf(lx, q[k].x); ly = fminf(ly, q[k].y); lz = fminf(lz, q[k].z);
            hx = fmaxf(hx, q[k].x); hy = fmaxf(hy, q[k].y); hz = fmaxf(hz, q[k].z);
        }
    }
    boxes[2 * (size_t)c] = make_float4(lx, ly, lz, 0.f);
    boxes[2 * (size_t)c + 1] = make_float4(hx, hy, hz, 0.f);
}

// coarse box = union of the eight children's
__global__ void grid_coarse_boxes_kernel(const float4 *__restrict__ boxes, int ncells, float4 *coarse)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    float4 lo = boxes[16 * (size_t)c], hi = boxes[16 * (size_t)c + 1];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        const float4 l = boxes[16 * (size_t)c + 2 * k], h = boxes[16 * (size_t)c + 2 * k + 1];
        lo.x = fminf(lo.x, l.x); lo.y = fminf(lo.y, l.y); lo.z = fminf(lo.z, l.z);
        hi.x = fmaxf(hi.x, h.x); hi.y = fmaxf(hi.y, h.y); hi.z = fmaxf(hi.z, h.z);
    }
    coarse[2 * (size_t)c] = lo;
    coarse[2 * (size_t)c + 1] = hi;
}

void launch_grid_build(const float4 *tgt, int m, const GridMeta &g, int *counts, int *cursor, int *block_sums,
                       float4 *sorted, float4 *boxes, float4 *coarse_boxes, cudaStream_t s)
{
    const int n = g.ncells * g.sub + 1;
    grid_count_kernel<<<(m + 255) / 256, 256, 0, s>>>(tgt, m, g, counts);
    const int nblocks = (n + kScanTile - 1) / kScanTile;
    scan_reduce_kernel<<<nblocks, kScanThreads, 0, s>>>(counts, n, block_sums);
    scan_sums_kernel<<<1, kScanThreads, 0, s>>>(block_sums, nblocks);
    scan_apply_kernel<<<nblocks, kScanThreads, 0, s>>>(counts, n, block_sums);
    cudaMemcpyAsync(cursor, counts, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, s);
    grid_scatter_kernel<<<(m + 255) / 256, 256, 0, s>>>(tgt, m, g, cursor, sorted);
    if (boxes) grid_boxes_kernel<<<(g.ncells * g.sub + 255) / 256, 256, 0, s>>>(sorted, counts, g.ncells * g.sub, boxes);
    if (boxes && coarse_boxes && g.sub == 8)
        grid_coarse_boxes_kernel<<<(g.ncells + 255) / 256, 256, 0, s>>>(boxes, g.ncells, coarse_boxes);
}

// ---- batched cell build (ICPB_NN_GRID for icpb_icp_register_batch) -----------------------------------------------------
__global__ void __launch_bounds__(256) grid_bbox_batch_kernel(const TgtRef *__restrict__ refs, unsigned int *bbox)
{
    __shared__ unsigned int s_lo[3][8], s_hi[3][8];
    const TgtRef t = refs[blockIdx.y];
    unsigned int *bb = bbox + 6 * blockIdx.y;
    unsigned int lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < t.m; i += gridDim.x * blockDim.x) {
        const float4 p = t.pts[i];
        const unsigned int v[3] = {f2ord(p.x), f2ord(p.y), f2ord(p.z)};
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = min(lo[k], v[k]); hi[k] = max(hi[k], v[k]); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = __reduce_min_sync(0xffffffffu, lo[k]);
        hi[k] = __reduce_max_sync(0xffffffffu, hi[k]);
        if ((threadIdx.x & 31) == 0) { s_lo[k][threadIdx.x >> 5] = lo[k]; s_hi[k][threadIdx.x >> 5] = hi[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int k = threadIdx.x;
        unsigned int a = s_lo[k][0], b = s_hi[k][0];
        for (int w = 1; w < 8; ++w) { a = min(a, s_lo[k][w]); b = max(b, s_hi[k][w]); }
        atomicMin(&bb[k], a);
        atomicMax(&bb[3 + k], b);
    }
}

void launch_grid_bbox_batch(const TgtRef *refs, int batch, int max_m, unsigned int *bbox, cudaStream_t s)
{
    dim3 grid(max(1, min((max_m + 255) / 256, 296 / max(1, min(batch, 296)) + 1)), batch);
    grid_bbox_batch_kernel<<<grid, 256, 0, s>>>(refs, bbox);
}

__global__ void grid_count_batch_kernel(const RegDesc *__restrict__ descs)
{
    const RegDesc &d = descs[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.m) return;
    const float4 p = d.tgt[i];
    atomicAdd(const_cast<int *>(d.gstart) + cell_of(*d.grid, p.x, p.y, p.z), 1);
}

__global__ void grid_scatter_batch_kernel(const RegDesc *__restrict__ descs)
{
    const RegDesc &d = descs[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.m) return;
    float4 p = d.tgt[i];
    const int pos = atomicAdd(&d.gcursor[cell_of(*d.grid, p.x, p.y, p.z)], 1);
    p.w = __int_as_float(i);
    const_cast<float4 *>(d.gsorted)[pos] = p;
}

__global__ void grid_coarse_boxes_batch_kernel(const RegDesc *__restrict__ descs)
{
    const RegDesc &d = descs[blockIdx.y];
    const GridMeta &g = *d.grid;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (g.sub != 8 || c >= g.ncells) return;
    const float4 *boxes = d.gbox;
    float4 lo = boxes[16 * (size_t)c], hi = boxes[16 * (size_t)c + 1];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        const float4 l = boxes[16 * (size_t)c + 2 * k], h = boxes[16 * (size_t)c + 2 * k + 1];
        lo.x = fminf(lo.x, l.x); lo.y = fminf(lo.y, l.y); lo.z = fminf(lo.z, l.z);
        hi.x = fmaxf(hi.x, h.x); hi.y = fmaxf(hi.y, h.y); hi.z = fmaxf(hi.z, h.z);
    }
    const_cast<float4 *>(d.gboxc)[2 * (size_t)c] = lo;
    const_cast<float4 *>(d.gboxc)[2 * (size_t)c + 1] = hi;
}

// counts: `total_entries` ints (every registration's cells + 1, back to back), zeroed by the caller.  After the call
// counts[] holds the start of every cell in the shared sorted array; boxes[] the tight box of every entry.
void launch_grid_build_batch(const RegDesc *descs, int batch, int max_m, int max_ncells, int *counts, int *cursor,
                             long long total_entries, int *block_sums, const float4 *sorted, float4 *boxes, bool any_children,
                             cudaStream_t s)
{
    const int n = (int)total_entries;
    dim3 pgrid((max_m + 255) / 256, batch);
    grid_count_batch_kernel<<<pgrid, 256, 0, s>>>(descs);
    const int nblocks = (n + kScanTile - 1) / kScanTile;
    scan_reduce_kernel<<<nblocks, kScanThreads, 0, s>>>(counts, n, block_sums);
    scan_sums_kernel<<<1, kScanThreads, 0, s>>>(block_sums, nblocks);
    scan_apply_kernel<<<nblocks, kScanThreads, 0, s>>>(counts, n, block_sums);
    cudaMemcpyAsync(cursor, counts, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, s);
    grid_scatter_batch_kernel<<<pgrid, 256, 0, s>>>(descs);
    // every entry but the last has a successor: its box is that of the points between the two starts (the entry that
    // closes a registration holds no points: start == next start, an inverted box)
    grid_boxes_kernel<<<(n - 1 + 255) / 256, 256, 0, s>>>(sorted, counts, n - 1, boxes);
    if (any_children) grid_coarse_boxes_batch_kernel<<<dim3((max_ncells + 255) / 256, batch), 256, 0, s>>>(descs);
}

// ---- spatial order of the QUERIES (warp-centred brute-force filter of nn.cu, cooperative grid search) ---------------
// The order is a performance matter only -- it decides which queries share a warp, never what any query's result is.
// d.perm[slot] = original index.  blockIdx.y = registration.

// batched forms of the three scan kernels: blockIdx.y = registration, arrays `stride` / `sums_stride` apart
__global__ void __launch_bounds__(kScanThreads) scan_reduce_batch_kernel(const int *in, int n, int stride, int *block_sums,
                                                                        int sums_stride)
{
    __shared__ int s_warp[33];
    in += (size_t)blockIdx.y * stride;
    int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) v += (base + k < n) ? in[base + k] : 0;
    int total;
    block_scan_excl(v, s_warp, total);
    if (threadIdx.x == 0) block_sums[(size_t)blockIdx.y * sums_stride + blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_sums_batch_kernel(int *block_sums, int nblocks, int sums_stride)
{
    __shared__ int s_warp[33];
    block_sums += (size_t)blockIdx.x * sums_stride;
    int carry = 0;
    for (int b0 = 0; b0 < nblocks; b0 += kScanThreads) {
        int i = b0 + threadIdx.x;
        int v = (i < nblocks) ? block_sums[i] : 0;
        int total;
        int ex = block_scan_excl(v, s_warp, total);
        if (i < nblocks) block_sums[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_batch_kernel(int *data, int n, int stride, const int *block_sums,
                                                                       int sums_stride)
{
    __shared__ int s_warp[33];
    data += (size_t)blockIdx.y * stride;
    int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int x[kScanItems], v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { x[k] = (base + k < n) ? data[base + k] : 0; v += x[k]; }
    int total;
    int ex = block_scan_excl(v, s_warp, total) + block_sums[(size_t)blockIdx.y * sums_stride + blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) data[base + k] = ex;
        ex += x[k];
    }
}

// ---- stable LSD radix sort of the queries by a 24-bit Morton key ----------------------------------------------------
// Three passes of 8 bits.  Every pass is a per-block digit histogram (block = 256 consecutive elements), an exclusive
// scan over (digit, block) and a scatter that ranks equal digits inside the block in element order (__match_any +
// per-warp counts): stable, hence DETERMINISTIC -- slot k of d.perm holds the same query in every run (the counting
// sort with an atomic cursor that this replaces ordered the queries of a cell by arrival).  8 bits per axis over 8 m
// around the cloud's first point: 3.1 cm cells, ties in original (raster) order, so the eight queries of a search
// group and the 32 of a warp are close neighbours.
constexpr int kRsBlock = 256;

__device__ __forceinline__ unsigned int spread3_8(unsigned int v) // 8 bits -> every third bit
{
    v &= 0xffu;
    v = (v | (v << 8)) & 0x0000f00fu;
    v = (v | (v << 4)) & 0x000c30c3u;
    v = (v | (v << 2)) & 0x00249249u;
    return v;
}

__device__ __forceinline__ unsigned int morton24(const float4 p, const float4 ref)
{
    const float cell = 8.0f / 256.0f;
    const int cx = min(max((int)floorf((p.x - ref.x) / cell) + 128, 0), 255);
    const int cy = min(max((int)floorf((p.y - ref.y) / cell) + 128, 0), 255);
    const int cz = min(max((int)floorf((p.z - ref.z) / cell) + 128, 0), 255);
    return spread3_8((unsigned)cx) | (spread3_8((unsigned)cy) << 1) | (spread3_8((unsigned)cz) << 2);
}

// keys + identity values, and the digit histogram of pass 0
__global__ void __launch_bounds__(kRsBlock) rsort_keys_kernel(const RegDesc *__restrict__ descs, unsigned int *keys, int *vals,
                                                              int stride, int *hist, int nblk)
{
    __shared__ int s_h[256];
    const RegDesc &d = descs[blockIdx.y];
    const int i = blockIdx.x * kRsBlock + threadIdx.x;
    s_h[threadIdx.x] = 0;
    __syncthreads();
    if (i < d.n) {
        const float4 *pts = d.D[0];
        const unsigned int key = morton24(pts[i], pts[0]);
        keys[(size_t)blockIdx.y * stride + i] = key;
        vals[(size_t)blockIdx.y * stride + i] = i;
        atomicAdd(&s_h[key & 255u], 1);
    }
    __syncthreads();
    hist[((size_t)blockIdx.y * 256 + threadIdx.x) * nblk + blockIdx.x] = s_h[threadIdx.x];
}

__global__ void __launch_bounds__(kRsBlock) rsort_hist_kernel(const RegDesc *__restrict__ descs, const unsigned int *keys, int stride,
                                                              int shift, int *hist, int nblk)
{
    __shared__ int s_h[256];
    const int n = descs[blockIdx.y].n;
    const int i = blockIdx.x * kRsBlock + threadIdx.x;
    s_h[threadIdx.x] = 0;
    __syncthreads();
    if (i < n) atomicAdd(&s_h[(keys[(size_t)blockIdx.y * stride + i] >> shift) & 255u], 1);
    __syncthreads();
    hist[((size_t)blockIdx.y * 256 + threadIdx.x) * nblk + blockIdx.x] = s_h[threadIdx.x];
}

// hist holds the exclusive scan over (digit, block).  out_vals == nullptr on the last pass: the values go to d.perm.
__global__ void __launch_bounds__(kRsBlock) rsort_scatter_kernel(const RegDesc *__restrict__ descs, const unsigned int *keys,
                                                                 const int *vals, unsigned int *out_keys, int *out_vals,
                                                                 int stride, int shift, const int *hist, int nblk)
{
    __shared__ int s_cnt[kRsBlock / 32][256];
    const RegDesc &d = descs[blockIdx.y];
    const int n = d.n;
    const int i = blockIdx.x * kRsBlock + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kRsBlock / 32; ++k) s_cnt[k][threadIdx.x] = 0;
    __syncthreads();
    const bool valid = i < n;
    unsigned int key = 0;
    int val = 0;
    if (valid) { key = keys[(size_t)blockIdx.y * stride + i]; val = vals[(size_t)blockIdx.y * stride + i]; }
    const unsigned int digit = valid ? ((key >> shift) & 255u) : 256u; // invalid lanes match only each other
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) s_cnt[w][digit] = __popc(peers);
    __syncthreads();
    { // exclusive prefix over the block's warps, one thread per digit
        int run = 0;
#pragma unroll
        for (int k = 0; k < kRsBlock / 32; ++k) { const int c = s_cnt[k][threadIdx.x]; s_cnt[k][threadIdx.x] = run; run += c; }
    }
    __syncthreads();
    if (!valid) return;
    const int pos = hist[((size_t)blockIdx.y * 256 + digit) * nblk + blockIdx.x] + s_cnt[w][digit] + rank;
    if (out_vals) {
        out_keys[(size_t)blockIdx.y * stride + pos] = key;
        out_vals[(size_t)blockIdx.y * stride + pos] = val;
    } else {
        const_cast<int *>(d.perm)[pos] = val;
    }
}

// work: batch * (4 * stride + 256 * nblk + sums) ints, laid out by spatial_sort_layout()
int spatial_sort_stride(int max_n) { return (max_n + kRsBlock - 1) / kRsBlock * kRsBlock; }
int spatial_sort_blocks(int max_n) { return (max_n + kRsBlock - 1) / kRsBlock; }
int spatial_sort_sum_slots(int max_n) { return (256 * spatial_sort_blocks(max_n) + kScanTile - 1) / kScanTile + 1; }
size_t spatial_sort_work_ints(int max_n, int batch)
{
    return (size_t)batch * (4 * (size_t)spatial_sort_stride(max_n) + 256 * (size_t)spatial_sort_blocks(max_n) +
                            (size_t)spatial_sort_sum_slots(max_n));
}

void launch_spatial_sort(const RegDesc *descs, int batch, int max_n, int *work, cudaStream_t s)
{
    const int stride = spatial_sort_stride(max_n), nblk = spatial_sort_blocks(max_n), ss = spatial_sort_sum_slots(max_n);
    unsigned int *keysA = (unsigned int *)work, *keysB = keysA + (size_t)batch * stride;
    int *valsA = (int *)(keysB + (size_t)batch * stride), *valsB = valsA + (size_t)batch * stride;
    int *hist = valsB + (size_t)batch * stride, *sums = hist + (size_t)batch * 256 * nblk;
    const int nscan = 256 * nblk;
    const int sblocks = (nscan + kScanTile - 1) / kScanTile;
    dim3 grid(nblk, batch);
    auto scan = [&]() {
        scan_reduce_batch_kernel<<<dim3(sblocks, batch), kScanThreads, 0, s>>>(hist, nscan, nscan, sums, ss);
        scan_sums_batch_kernel<<<batch, kScanThreads, 0, s>>>(sums, sblocks, ss);
        scan_apply_batch_kernel<<<dim3(sblocks, batch), kScanThreads, 0, s>>>(hist, nscan, nscan, sums, ss);
    };
    rsort_keys_kernel<<<grid, kRsBlock, 0, s>>>(descs, keysA, valsA, stride, hist, nblk);
    scan();
    rsort_scatter_kernel<<<grid, kRsBlock, 0, s>>>(descs, keysA, valsA, keysB, valsB, stride, 0, hist, nblk);
    rsort_hist_kernel<<<grid, kRsBlock, 0, s>>>(descs, keysB, stride, 8, hist, nblk);
    scan();
    rsort_scatter_kernel<<<grid, kRsBlock, 0, s>>>(descs, keysB, valsB, keysA, valsA, stride, 8, hist, nblk);
    rsort_hist_kernel<<<grid, kRsBlock, 0, s>>>(descs, keysA, stride, 16, hist, nblk);
    scan();
    rsort_scatter_kernel<<<grid, kRsBlock, 0, s>>>(descs, keysA, valsA, nullptr, nullptr, stride, 16, hist, nblk);
}

// ---- the search ------------------------------------------------------------------------------------------------

__device__ __forceinline__ float exact_distance_xyz(float ax, float ay, float az, float bx, float by, float bz, float &xyz)
{
    // icp.cpp:606-620 (see nn.cu): float differences, double squares summed left to right, one rounding, float sqrt
    float x = ax - bx, y = ay - by, z = az - bz;
    double s = ((double)x * (double)x + (double)y * (double)y) + (double)z * (double)z;
    xyz = (float)s;
    return sqrtf(xyz);
}

// Shared pieces of the two search kernels ----------------------------------------------------------------------

struct Best {
    float d, thr;
    int i;
};

// Visit shell r (cells at Chebyshev distance exactly r from c0).  LANES threads share the work: lane `lane`
// takes candidates lane, lane+LANES, ... of every contiguous run.
// Gap between coordinate q and the cell interval [lo, lo + h) along one axis, shrunk by the same slack the cube
// bound uses for the float rounding of the cell assignment: a LOWER bound on |q - t| for every target t of the cell.
__device__ __forceinline__ float axis_gap(float q, float lo, float h)
{
    return fmaxf(fmaxf(lo - q, q - (lo + h)), 0.f) - 1e-3f * h;
}

template <int LANES>
__device__ __forceinline__ void visit_shell(const GridMeta &g, const float4 *__restrict__ sorted, const int *__restrict__ start,
                                            const float4 &p, int c0x, int c0y, int c0z, int r, int lane, Best &best)
{
    const int nx = g.dim[0], ny = g.dim[1], nz = g.dim[2];
    const int xlo = max(c0x - r, 0), xhi = min(c0x + r, nx - 1);
    // LANES > 1: the lanes split every run by position, so they must all prune the same rows and cells -- the reach
    // is frozen at the (warp-uniform) best the shell starts with.  One thread per query uses its live best.
    const float frozen_reach2 = best.d * best.d * 1.00001f + 1e-30f;
    for (int dz = -r; dz <= r; ++dz) {
        const int z = c0z + dz;
        if (z < 0 || z >= nz) continue;
        const float gz = fmaxf(axis_gap(p.z, g.mn[2] + z * g.h, g.h), 0.f);
        for (int dy = -r; dy <= r; ++dy) {
            const int y = c0y + dy;
            if (y < 0 || y >= ny) continue;
            // Ball pruning: a row of cells whose box is farther from the query than the current best distance
            // cannot hold a better or an equal candidate.  best.d only shrinks, so the test stays valid.
            const float gy = fmaxf(axis_gap(p.y, g.mn[1] + y * g.h, g.h), 0.f);
            const float gyz2 = gy * gy + gz * gz;
            const float reach2 = (LANES > 1) ? frozen_reach2 : best.d * best.d * 1.00001f + 1e-30f; // +inf: nothing found yet
            if (gyz2 > reach2) continue;
            const int rowbase = (z * ny + y) * nx;
            const bool edge = (dz == -r) || (dz == r) || (dy == -r) || (dy == r);
            const int nseg = edge ? 1 : 2; // edge rows: the whole x-run; inner rows: the two end cells
            for (int sgm = 0; sgm < nseg; ++sgm) {
                int a, b;
                if (edge) {
                    a = xlo; b = xhi;
                    if (reach2 < CUDART_INF_F) { // cells of the run the ball can reach
                        const float rx = sqrtf(fmaxf(reach2 - gyz2, 0.f)) * 1.00001f + 2e-3f * g.h;
                        a = max(a, cell_axis(p.x - rx, g.mn[0], g.h, nx));
                        b = min(b, cell_axis(p.x + rx, g.mn[0], g.h, nx));
                        if (a > b) continue;
                    }
                } else {
                    const int x = (sgm == 0) ? c0x - r : c0x + r;
                    if (x < 0 || x >= nx) continue;
                    const float gx = fmaxf(axis_gap(p.x, g.mn[0] + x * g.h, g.h), 0.f);
                    if (gx * gx + gyz2 > reach2) continue;
                    a = b = x;
                }
                const int t0 = __ldg(&start[(rowbase + a) * g.sub]);
                const int t1 = __ldg(&start[(rowbase + b + 1) * g.sub]);
                for (int t = t0 + lane; t < t1; t += LANES) {
                    const float4 q = __ldg(&sorted[t]);
                    const float ex = p.x - q.x, ey = p.y - q.y, ez = p.z - q.z;
                    const float sf = __fmaf_rn(ez, ez, __fmaf_rn(ey, ey, ex * ex));
                    if (sf > best.thr) continue; // provably farther than the current best (nn.cu bound)
                    float xyz;
                    const float dd = exact_distance_xyz(p.x, p.y, p.z, q.x, q.y, q.z, xyz);
                    const int oi = __float_as_int(q.w);
                    if (dd < best.d || (dd == best.d && oi < best.i)) {
                        best.d = dd; best.i = oi;
                        best.thr = xyz * kBandRel + kBandAbs;
                    }
                }
            }
        }
    }
}

// Lower bound on the distance from p to anything outside the cube of radius r around its cell (+inf when the
// cube covers the grid), minus a slack for the float rounding of the cell assignment.
__device__ __forceinline__ float shell_lower_bound(const GridMeta &g, const float4 &p, int c0x, int c0y, int c0z, int r)
{
    float lb = CUDART_INF_F;
    const float q[3] = {p.x, p.y, p.z};
    const int c0[3] = {c0x, c0y, c0z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int lo_c = c0[k] - r, hi_c = c0[k] + r;
        if (lo_c > 0) lb = fminf(lb, q[k] - (g.mn[k] + lo_c * g.h));
        if (hi_c < g.dim[k] - 1) lb = fminf(lb, (g.mn[k] + (hi_c + 1) * g.h) - q[k]);
    }
    return (lb == CUDART_INF_F) ? lb : lb - 1e-3f * g.h;
}

// true when the query is farther than the acceptance radius from the target's bounding box
__device__ __forceinline__ bool beyond_reach(const GridMeta &g, const float4 &p)
{
    const float q[3] = {p.x, p.y, p.z};
    float out2 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float lo = g.mn[k], hi = g.mn[k] + g.dim[k] * g.h;
        float e = fmaxf(fmaxf(lo - q[k], q[k] - hi), 0.f);
        out2 += e * e;
    }
    const float reach = g.max_nn + 2.f * g.h;
    return out2 > reach * reach;
}


// Phase 1: one thread per query, the first shells.  Resolved queries write (idx, dist); the others leave their
// partial best in (idx, dist) and append themselves to the heavy list of this pass.
__global__ void __launch_bounds__(128) nn_grid_kernel(const RegDesc *__restrict__ descs, int pass)
{
    pdl_enter(); // icpb_internal.h: the grid before this one is complete from here on
    const RegDesc d = descs[blockIdx.z];
    IcpState *st = d.st;
    if (st->done) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= d.n) return;
    // queries in Morton order when the permutation is there: the lanes of a warp then walk (almost) the same cells,
    // so their candidate loads hit the same lines and they leave the shell loops together
    const int i = d.perm ? d.perm[k] : k;

    // P2 fused into the query load, as in nn_partial (pointcloud.cpp:321-359)
    float4 p = d.D[pass & 1][i];
    if (st->apply) {
        const float *R = st->Rf, *T = st->tf;
        float x = ((R[0] * p.x + R[1] * p.y) + R[2] * p.z) + T[0];
        float y = ((R[3] * p.x + R[4] * p.y) + R[5] * p.z) + T[1];
        float z = ((R[6] * p.x + R[7] * p.y) + R[8] * p.z) + T[2];
        p.x = x; p.y = y; p.z = z;
    }
    d.D[(pass + 1) & 1][i] = p;

    const GridMeta g = *d.grid;
    Best best = {CUDART_INF_F, CUDART_INF_F, 0x7fffffff};
    bool open = false;
    if (!beyond_reach(g, p)) {
        const int c0x = cell_axis(p.x, g.mn[0], g.h, g.dim[0]);
        const int c0y = cell_axis(p.y, g.mn[1], g.h, g.dim[1]);
        const int c0z = cell_axis(p.z, g.mn[2], g.h, g.dim[2]);
        open = true;
        for (int r = 0; r <= g.light_r && r <= g.max_r; ++r) {
            visit_shell<1>(g, d.gsorted, d.gstart, p, c0x, c0y, c0z, r, 0, best);
            const float lb = shell_lower_bound(g, p, c0x, c0y, c0z, r);
            if (lb == CUDART_INF_F || best.d < lb || lb > g.max_nn || r == g.max_r) { open = false; break; }
        }
    }
    if (open) {
        const int slot = atomicAdd(&d.gheavy_count[pass], 1);
        d.gheavy[slot] = i;
    } else if (!(best.d < g.max_nn)) {
        best.i = -1; best.d = CUDART_INF_F;
    }
    d.idx[i] = best.i;
    d.dist[i] = best.d;
}

// Phase 2: one warp per heavy query, continuing from shell light_r+1 with the candidates of every run split
// over the 32 lanes; the warp's lexicographic (distance, index) minimum is exchanged after every shell.
__global__ void __launch_bounds__(128) nn_grid_heavy_kernel(const RegDesc *__restrict__ descs, int pass, int first_shell)
{
    pdl_enter(); // icpb_internal.h: the grid before this one is complete from here on
    const RegDesc d = descs[blockIdx.z];
    IcpState *st = d.st;
    if (st->done) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int count = d.gheavy_count[pass];
    const GridMeta g = *d.grid;
    for (int e = warp; e < count; e += nwarps) {
        const int i = d.gheavy[e];
        const float4 p = d.D[(pass + 1) & 1][i];
        const int c0x = cell_axis(p.x, g.mn[0], g.h, g.dim[0]);
        const int c0y = cell_axis(p.y, g.mn[1], g.h, g.dim[1]);
        const int c0z = cell_axis(p.z, g.mn[2], g.h, g.dim[2]);
        Best best;
        best.d = d.dist[i]; best.i = d.idx[i];
        {
            // rebuild the filter threshold of the partial best: thr bounds s-tilde, and xyz <= d^2 (1 + 2^-22)
            const float dd = best.d;
            best.thr = (dd == CUDART_INF_F) ? CUDART_INF_F : (dd * dd) * (kBandRel + 4.8e-7f) + kBandAbs;
        }
        // first_shell < 0: continue after the per-thread shells; 0: a query handed over by the cooperative kernel, whose
        // partial best (a real candidate, or none) says nothing about which shells were covered
        for (int r = first_shell < 0 ? g.light_r + 1 : first_shell; r <= g.max_r; ++r) {
            visit_shell<32>(g, d.gsorted, d.gstart, p, c0x, c0y, c0z, r, lane, best);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, best.d, off);
                const int oi = __shfl_xor_sync(0xffffffffu, best.i, off);
                const float ot = __shfl_xor_sync(0xffffffffu, best.thr, off);
                if (od < best.d || (od == best.d && oi < best.i)) { best.d = od; best.i = oi; best.thr = ot; }
            }
            const float lb = shell_lower_bound(g, p, c0x, c0y, c0z, r);
            if (lb == CUDART_INF_F || best.d < lb || lb > g.max_nn) break;
        }
        if (lane == 0) {
            if (!(best.d < g.max_nn)) { best.i = -1; best.d = CUDART_INF_F; }
            d.idx[i] = best.i;
            d.dist[i] = best.d;
            if (first_shell >= 0) { // cooperative mode: nn_finalize reads the neighbour's coordinates from gnb
                float4 rec = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
                if (best.i >= 0) { const float4 t = d.tgt[best.i]; rec = make_float4(t.x, t.y, t.z, __int_as_float(best.i)); }
                d.gnb[i] = rec;
            }
        }
    }
}

// ---- warp-cooperative search -------------------------------------------------------------------------------------
// The per-thread walk above spends its time on private, divergent 16-byte candidate loads (13.9 of 32 lanes active,
// long_scoreboard the top stall, profiles/r01_ncu_nn_grid_fullres_v2.txt).  Here the 32 Morton-neighbour queries of a
// warp search TOGETHER, the way the brute-force scan of nn.cu does, but over a few hundred candidates instead of all M:
//
//   1. every query owns a search ball that is known to hold its nearest neighbour: radius = the reference-arithmetic
//      distance to the neighbour the PREVIOUS pass found (a rigid motion of a few millimetres later that target is
//      still close: temporal coherence of the ICP loop, icp.cpp:155-258); on the first pass a guess of one cell edge
//      that is corrected by a second round;
//   2. lane r takes row r (fixed y, z; a contiguous run of cells along x) of the box of cells around the warp's balls
//      and intersects it with all 32 balls: the union of the needed x-ranges is ONE contiguous run of sorted targets;
//   3. the runs are copied -- coalesced, once per warp -- into a warp-private shared-memory batch, centred on the
//      warp's centre (x', y', z', |t'|^2 in SoA);
//   4. every lane evaluates every staged candidate with the centred expansion filter of nn_partial_warp_kernel
//      (3 packed FFMA per pair from four broadcast LDS.128 per four candidates); only a candidate whose filter value is
//      within the proven error band of the lane's best EXACT squared distance is evaluated in the reference's
//      arithmetic (icp.cpp:606-620) -- a handful per query;
//   5. a lane is finished when its best distance is inside the ball it searched: every target at that distance or
//      closer was staged, so the lexicographic (distance, index) minimum over the staged set is the brute-force
//      scan's answer.  Otherwise its ball becomes its best distance and the warp goes round again.
//
// Lanes whose ball is wider than coop_r (no overlap with the target: the rim of the frame) are handed to the
// warp-per-query kernel with their partial best, like the open queries of the per-thread kernel.
typedef unsigned long long u64;
__device__ __forceinline__ u64 gpack2(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void gunpack2(u64 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 gfma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float gmin3(float a, float b, float c)
{
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

#ifndef ICPB_COOP_CAP
#define ICPB_COOP_CAP 512
#endif
constexpr int kCoopCap = ICPB_COOP_CAP;   // candidates per warp batch (20 B each in shared memory)
#ifndef ICPB_COOP_WARPS
#define ICPB_COOP_WARPS 4
#endif
constexpr int kCoopWarps = ICPB_COOP_WARPS;
#ifndef ICPB_COOP_GROUP
#define ICPB_COOP_GROUP 4
#endif
constexpr int kCoopGroup = ICPB_COOP_GROUP; // lanes per bounding sphere of the region test (1 = every query's own ball)
#ifndef ICPB_COOP_MINB
#define ICPB_COOP_MINB 5
#endif

#ifdef ICPB_COOP_CLOCKS
__device__ int g_coop_dbg[16384 * 8];
__device__ __forceinline__ unsigned __smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#endif
struct CoopBuf {
    float xs[kCoopCap], ys[kCoopCap], zs[kCoopCap], ns[kCoopCap]; // centred candidates, SoA
    int gp[kCoopCap];                                              // their slot in the sorted target array
};

// Phase A, all lanes: the centred filter value of every staged candidate of the batch [0, fill) (fill a multiple of 4),
// reduced to the best and second-best minimum over GROUPS of four consecutive slots and the best group's number.
// No branches: a lane's running minimum improves at a different slot than its neighbours', and any per-improvement
// work would be executed by one lane at a time (the first version of this kernel spent 60 % of its instructions so).
__device__ __forceinline__ void coop_scan(const CoopBuf &b, int fill, float qx, float qy, float qz, float &m1, float &m2, int &g1)
{
    const u64 q2x = gpack2(qx, qx), q2y = gpack2(qy, qy), q2z = gpack2(qz, qz);
    const float4 *X4 = reinterpret_cast<const float4 *>(b.xs);
    const float4 *Y4 = reinterpret_cast<const float4 *>(b.ys);
    const float4 *Z4 = reinterpret_cast<const float4 *>(b.zs);
    const float4 *N4 = reinterpret_cast<const float4 *>(b.ns);
#pragma unroll 4
    for (int j = 0; j < fill / 4; ++j) {
        const float4 X = X4[j], Y = Y4[j], Z = Z4[j], N = N4[j];
        u64 sa = gfma2(q2x, gpack2(X.x, X.y), gpack2(N.x, N.y)), sb = gfma2(q2x, gpack2(X.z, X.w), gpack2(N.z, N.w));
        sa = gfma2(q2y, gpack2(Y.x, Y.y), sa); sb = gfma2(q2y, gpack2(Y.z, Y.w), sb);
        sa = gfma2(q2z, gpack2(Z.x, Z.y), sa); sb = gfma2(q2z, gpack2(Z.z, Z.w), sb);
        float w0, w1, w2, w3;
        gunpack2(sa, w0, w1);
        gunpack2(sb, w2, w3);
        const float m = fminf(gmin3(w0, w1, w2), w3);
        const bool lt = m < m1;
        m2 = lt ? m1 : fminf(m2, m);
        g1 = lt ? j : g1;
        m1 = fminf(m1, m);
    }
}

struct BestB {
    float x, y, z; // coordinates of the best target so far
};

__device__ __forceinline__ void coop_take(const float4 &p, const float4 &t, Best &best, BestB &bb)
{
    float xyz;
    const float dd = exact_distance_xyz(p.x, p.y, p.z, t.x, t.y, t.z, xyz);
    const int oi = __float_as_int(t.w);
    if (dd < best.d || (dd == best.d && oi < best.i)) {
        best.d = dd; best.i = oi;
        bb.x = t.x; bb.y = t.y; bb.z = t.z;
    }
}

// Fallback for lanes whose second-best group is within the error band of the best (near-ties across groups): every
// candidate of the batch whose filter value is within the band is evaluated in the reference's arithmetic.
__device__ __forceinline__ void coop_rescan(const CoopBuf &b, int fill, const float4 *__restrict__ sorted, const float4 &p,
                                            float qx, float qy, float qz, float lim, Best &best, BestB &bb)
{
    const u64 q2x = gpack2(qx, qx), q2y = gpack2(qy, qy), q2z = gpack2(qz, qz);
    const float4 *X4 = reinterpret_cast<const float4 *>(b.xs);
    const float4 *Y4 = reinterpret_cast<const float4 *>(b.ys);
    const float4 *Z4 = reinterpret_cast<const float4 *>(b.zs);
    const float4 *N4 = reinterpret_cast<const float4 *>(b.ns);
    for (int j = 0; j < fill / 4; ++j) {
        const float4 X = X4[j], Y = Y4[j], Z = Z4[j], N = N4[j];
        u64 sa = gfma2(q2x, gpack2(X.x, X.y), gpack2(N.x, N.y)), sb = gfma2(q2x, gpack2(X.z, X.w), gpack2(N.z, N.w));
        sa = gfma2(q2y, gpack2(Y.x, Y.y), sa); sb = gfma2(q2y, gpack2(Y.z, Y.w), sb);
        sa = gfma2(q2z, gpack2(Z.x, Z.y), sa); sb = gfma2(q2z, gpack2(Z.z, Z.w), sb);
        float w[4];
        gunpack2(sa, w[0], w[1]);
        gunpack2(sb, w[2], w[3]);
        if (fminf(gmin3(w[0], w[1], w[2]), w[3]) <= lim) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (w[k] <= lim) coop_take(p, __ldg(&sorted[b.gp[4 * j + k]]), best, bb);
        }
    }
}

// One batch: phase A for all lanes, then the exact stage.  A lane whose second-best group minimum exceeds the best by
// more than the filter's error band (the bound of nn_partial_centred_kernel, as applied by nn_finalize) needs only the
// four candidates of its best group: everything outside is strictly farther, in the reference's arithmetic, than the
// candidate that produced the minimum.  Other lanes take the fallback.
__device__ __forceinline__ void coop_batch(const CoopBuf &b, int fill, const float4 *__restrict__ sorted, const float4 &p,
                                           float qx, float qy, float qz, float A, bool live, Best &best, BestB &bb)
{
    float m1 = CUDART_INF_F, m2 = CUDART_INF_F;
    int g1 = 0;
    coop_scan(b, fill, qx, qy, qz, m1, m2, g1);
    const bool have = live && m1 < CUDART_INF_F;
    const float lim = m1 + ((A * kBandCentredA + fmaxf(m1 + A, 0.f) * kBandCentredX) + kBandAbs);
    if (have) {
#pragma unroll
        for (int k = 0; k < 4; ++k) coop_take(p, __ldg(&sorted[b.gp[4 * g1 + k]]), best, bb);
    }
    const bool amb = have && !(m2 > lim);
    if (__any_sync(0xffffffffu, amb)) coop_rescan(b, fill, sorted, p, qx, qy, qz, amb ? lim : -CUDART_INF_F, best, bb);
}

// One warp's share of a pass: the 32 queries of work item `wpos`.
struct CoopMotion { // the head of IcpState (one cache line), read together with the done flag
    float R[9], T[3];
    int apply;
};

__device__ __forceinline__ void coop_warp(const RegDesc &d, const GridMeta &g, const CoopMotion &mo, CoopBuf &buf, const int pass,
                                          const float coop_r, const int wsel, const int lane)
{
    const int n = d.n;
    // ---- which 32 sorted slots this warp takes.  From the second pass on the warps are handed out heaviest first:
    //      a warp's cost (the candidates it stages, the cells it tests: up to five times the mean) changes little from
    //      pass to pass, and a heavy warp that starts late runs on while the rest of the GPU idles (measured with
    //      %globaltimer per warp: in natural order the last 28 % of the kernel's duration had < 4 of 20 warps per SM
    //      resident; tools/coop_clocks.py).  Classes by the cycles the warp took in the previous pass.
    const int ord_stride = d.n_stride >> 5;
    const long long ord_clk0 = clock64();
    const int k = wsel * 32 + lane;
    const bool valid = k < n;
    const int *perm = d.perm;
    const int i = valid ? (perm ? perm[k] : k) : 0;
    const float4 *__restrict__ sorted = d.gsorted;
    const int *__restrict__ gstart = d.gstart;
    const float4 *__restrict__ gbox = d.gbox;
    const float4 *__restrict__ gboxc = d.gboxc;

    // P2 fused into the query load (pointcloud.cpp:321-359), as in nn_grid_kernel
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
        // after the first pass the query comes from gq, the copy of the cloud in SORTED slot order this kernel keeps
        // up to date: a coalesced read that does not wait for perm[k]
        p = (d.gq && pass > 0) ? d.gq[k] : d.D[pass & 1][i];
        if (mo.apply) {
            const float *R = mo.R, *T = mo.T;
            const float x = ((R[0] * p.x + R[1] * p.y) + R[2] * p.z) + T[0];
            const float y = ((R[3] * p.x + R[4] * p.y) + R[5] * p.z) + T[1];
            const float z = ((R[6] * p.x + R[7] * p.y) + R[8] * p.z) + T[2];
            p.x = x; p.y = y; p.z = z;
        }
        d.D[(pass + 1) & 1][i] = p;
        if (d.gq) d.gq[k] = p;
    }

    // ---- the lane's ball: seeded by the previous pass's neighbour (kept per SORTED slot: a coalesced read, and the
    //      coordinates come with it), else one cell edge
    Best best = {CUDART_INF_F, CUDART_INF_F, 0x7fffffff};
    BestB bb = {0.f, 0.f, 0.f};
    bool done = !valid || beyond_reach(g, p);
    bool deferred = false;
    bool loose_seed = true; // the ball's radius is not last pass's neighbour distance
    float rad = g.h;
    if (!done && pass > 0) {
        const float4 sd = d.gseed[k];
        const int j = __float_as_int(sd.w);
        if (j >= 0) {
            float xyz;
            best.d = exact_distance_xyz(p.x, p.y, p.z, sd.x, sd.y, sd.z, xyz);
            best.i = j;
            bb.x = sd.x; bb.y = sd.y; bb.z = sd.z;
            rad = best.d;
            loose_seed = false;
        } else if (j == -2) { // nothing inside the acceptance radius last time: not worth widening the warp's search for
            deferred = true; done = true;
        }
    }
    const int nx = g.dim[0], ny = g.dim[1], nz = g.dim[2];
    if (!done && pass == 0) {
        // no previous pass to seed from: the first target of the query's own cell (or of a face neighbour) is a real
        // candidate a cell edge or so away -- a ball that is searched ONCE, instead of a guess of one cell edge that is
        // corrected by a second round for every query whose neighbour is farther than that
        const int c0x = cell_axis(p.x, g.mn[0], g.h, nx), c0y = cell_axis(p.y, g.mn[1], g.h, ny), c0z = cell_axis(p.z, g.mn[2], g.h, nz);
        const int ox[7] = {0, -1, 1, 0, 0, 0, 0}, oy[7] = {0, 0, 0, -1, 1, 0, 0}, oz[7] = {0, 0, 0, 0, 0, -1, 1};
#pragma unroll 1
        for (int k = 0; k < 7 && !(best.d < CUDART_INF_F); ++k) {
            const int x = c0x + ox[k], y = c0y + oy[k], z = c0z + oz[k];
            if (x < 0 || y < 0 || z < 0 || x >= nx || y >= ny || z >= nz) continue;
            const int cell = (z * ny + y) * nx + x;
            const int t0 = __ldg(&gstart[cell * g.sub]);
            if (__ldg(&gstart[(cell + 1) * g.sub]) > t0) coop_take(p, __ldg(&sorted[t0]), best, bb);
        }
        if (best.d < CUDART_INF_F) rad = best.d;
    }
    const float max_reach = g.max_nn * 1.00002f + 1e-6f; // nothing beyond the acceptance radius is ever needed (icp.cpp:553)

    // ---- warp centre of the centred filter
    const unsigned full = 0xffffffffu;
    float lox = valid ? p.x : CUDART_INF_F, hix = valid ? p.x : -CUDART_INF_F;
    float loy = valid ? p.y : CUDART_INF_F, hiy = valid ? p.y : -CUDART_INF_F;
    float loz = valid ? p.z : CUDART_INF_F, hiz = valid ? p.z : -CUDART_INF_F;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        lox = fminf(lox, __shfl_xor_sync(full, lox, off)); hix = fmaxf(hix, __shfl_xor_sync(full, hix, off));
        loy = fminf(loy, __shfl_xor_sync(full, loy, off)); hiy = fmaxf(hiy, __shfl_xor_sync(full, hiy, off));
        loz = fminf(loz, __shfl_xor_sync(full, loz, off)); hiz = fmaxf(hiz, __shfl_xor_sync(full, hiz, off));
    }
    const float cx = 0.5f * lox + 0.5f * hix, cy = 0.5f * loy + 0.5f * hiy, cz = 0.5f * loz + 0.5f * hiz;
    const float ax = valid ? p.x - cx : 0.f, ay = valid ? p.y - cy : 0.f, az = valid ? p.z - cz : 0.f;
    const float A = ((ax * ax + ay * ay) + az * az) * 1.000001f;
    const float qx = -2.f * ax, qy = -2.f * ay, qz = -2.f * az; // W = |t'|^2 - 2 a'.t'

    unsigned long long staged = 0; // candidates this warp put through the filter (x 32 lanes = pairs), profiling only
#ifdef ICPB_COOP_CLOCKS
    const long long clk0 = clock64();
    unsigned long long gt0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
    int dbg_rounds = 0, dbg_rows = 0, dbg_cells = 0, dbg_box = 0;
#endif
    // ---- probe rounds.  A ball of radius r around a query d away from the surface cuts a disc of area pi (r^2 - d^2)
    //      out of it and every target in the disc is a candidate: harmless for the usual query a few centimetres off
    //      the surface, ruinous for one 30 cm off whose ball was GUESSED (no seed, nothing found yet, radius
    //      quadrupled) -- such warps staged 12 - 20 thousand candidates and ran 0.5 ms, alone, at the end of the first
    //      pass.  When an active ball is wider than probe_r the round stages ONE target per child cell the balls
    //      reach: samples 5 cm apart, the nearest of which lies within millimetres of the true distance
    //      (sqrt(d^2 + s^2) - d ~ s^2 / 2d); it becomes the lane's seed and the real round that follows searches a ball
    //      that hugs the surface.  Exactness is untouched: the last round always covers the whole final ball.
    const float probe_r = g.probe_r;
    bool probed = false; // the previous round was a probe: this one is real
    for (int round = 0; round < 12; ++round) {
        // lanes whose ball outgrew the cooperative phase go to the warp-per-query kernel with their partial best
        const float rr = fminf(rad, max_reach);
        if (!done && rr > coop_r) { deferred = true; done = true; }
        const unsigned active = __ballot_sync(full, !done);
        if (active == 0u) break;
        // a ball whose radius is a guess (no seed, a first-cell or borrowed candidate) from probe_r on; any ball from
        // probe_r2 on (a temporal seed that far away still leaves a wide disc after a few centimetres of drift)
        const bool probe = !probed && __any_sync(full, !done && ((loose_seed && rr > probe_r) || rr > g.probe_r2));
        // ---- the union of the active balls.  A group of kCoopGroup consecutive lanes (Morton neighbours) is normally
        //      covered by ONE sphere: centre = middle of the group's queries, radius = the farthest reach of a member.
        //      A group that straddles a jump of the Morton curve would get a sphere metres wide: it is "loose" and its
        //      members are tested ball by ball instead.
        float glx = done ? CUDART_INF_F : p.x, ghx = done ? -CUDART_INF_F : p.x;
        float gly = done ? CUDART_INF_F : p.y, ghy = done ? -CUDART_INF_F : p.y;
        float glz = done ? CUDART_INF_F : p.z, ghz = done ? -CUDART_INF_F : p.z;
#pragma unroll
        for (int off = 1; off < kCoopGroup; off <<= 1) {
            glx = fminf(glx, __shfl_xor_sync(full, glx, off)); ghx = fmaxf(ghx, __shfl_xor_sync(full, ghx, off));
            gly = fminf(gly, __shfl_xor_sync(full, gly, off)); ghy = fmaxf(ghy, __shfl_xor_sync(full, ghy, off));
            glz = fminf(glz, __shfl_xor_sync(full, glz, off)); ghz = fmaxf(ghz, __shfl_xor_sync(full, ghz, off));
        }
        const float scx = 0.5f * glx + 0.5f * ghx, scy = 0.5f * gly + 0.5f * ghy, scz = 0.5f * glz + 0.5f * ghz;
        float sr = -1.f, far = 0.f; // no active member: the sphere needs nothing
        if (!done) {
            const float ex = p.x - scx, ey = p.y - scy, ez = p.z - scz;
            far = sqrtf((ex * ex + ey * ey) + ez * ez) * 1.00001f;
            sr = far + rr;
        }
#pragma unroll
        for (int off = 1; off < kCoopGroup; off <<= 1) {
            sr = fmaxf(sr, __shfl_xor_sync(full, sr, off));
            far = fmaxf(far, __shfl_xor_sync(full, far, off));
        }
        // loose: the sphere would cover much more than its members' balls do (uniform within the group)
        float rmax = done ? 0.f : rr;
#pragma unroll
        for (int off = 1; off < kCoopGroup; off <<= 1) rmax = fmaxf(rmax, __shfl_xor_sync(full, rmax, off));
        const bool loose = far > 0.5f * rmax + 0.01f;
        const float sr2 = (sr < 0.f || loose) ? -1.f : (sr * sr) * 1.0001f + 1e-12f;
        const float br2 = (!done && loose) ? (rr * rr) * 1.0001f + 1e-12f : -1.f; // the lane's own ball, loose groups only
        const unsigned loose_lanes = __ballot_sync(full, br2 >= 0.f);
        // needed(box): does the box [lo, hi] reach into one of the group spheres or one of the loose lanes' balls?
        auto reaches = [&](float lx, float ly, float lz, float hx, float hy, float hz) -> bool {
            bool hit = false;
#pragma unroll
            for (int s0 = 0; s0 < 32; s0 += kCoopGroup) {
                const float sx = __shfl_sync(full, scx, s0), sy = __shfl_sync(full, scy, s0), sz = __shfl_sync(full, scz, s0);
                const float s2 = __shfl_sync(full, sr2, s0);
                const float dx = fmaxf(fmaxf(lx - sx, sx - hx), 0.f), dy = fmaxf(fmaxf(ly - sy, sy - hy), 0.f);
                const float dz = fmaxf(fmaxf(lz - sz, sz - hz), 0.f);
                hit |= ((dx * dx + dy * dy) + dz * dz) <= s2;
            }
            for (unsigned rest = loose_lanes; rest; rest &= rest - 1) {
                const int j = __ffs(rest) - 1;
                const float sx = __shfl_sync(full, p.x, j), sy = __shfl_sync(full, p.y, j), sz = __shfl_sync(full, p.z, j);
                const float s2 = __shfl_sync(full, br2, j);
                const float dx = fmaxf(fmaxf(lx - sx, sx - hx), 0.f), dy = fmaxf(fmaxf(ly - sy, sy - hy), 0.f);
                const float dz = fmaxf(fmaxf(lz - sz, sz - hz), 0.f);
                hit |= ((dx * dx + dy * dy) + dz * dz) <= s2;
            }
            return hit;
        };
        // box of cells around the active balls
        const float ext = rr * 1.00001f + 2e-3f * g.h;
        int bx0 = done ? 0x7fffffff : cell_axis(p.x - ext, g.mn[0], g.h, nx), bx1 = done ? -1 : cell_axis(p.x + ext, g.mn[0], g.h, nx);
        int by0 = done ? 0x7fffffff : cell_axis(p.y - ext, g.mn[1], g.h, ny), by1 = done ? -1 : cell_axis(p.y + ext, g.mn[1], g.h, ny);
        int bz0 = done ? 0x7fffffff : cell_axis(p.z - ext, g.mn[2], g.h, nz), bz1 = done ? -1 : cell_axis(p.z + ext, g.mn[2], g.h, nz);
        bx0 = __reduce_min_sync(full, bx0); bx1 = __reduce_max_sync(full, bx1);
        by0 = __reduce_min_sync(full, by0); by1 = __reduce_max_sync(full, by1);
        bz0 = __reduce_min_sync(full, bz0); bz1 = __reduce_max_sync(full, bz1);
        const int nyb = by1 - by0 + 1, nzb = bz1 - bz0 + 1;
        const int nrows = nyb * nzb;
#ifdef ICPB_COOP_CLOCKS
        ++dbg_rounds; dbg_rows += nrows; dbg_box = max(dbg_box, nrows * (bx1 - bx0 + 1));
#endif
        const float xlo = g.mn[0] + bx0 * g.h, xhi = g.mn[0] + (bx1 + 1) * g.h;
        const float slack = 1e-3f * g.h; // float rounding of the cell assignment (as in axis_gap)
        const bool live = !done; // a lane that is done must not pick anything up any more
        int fill = 0;
        for (int rbase = 0; rbase < nrows; rbase += 32) {
            // ---- step 1, lane <-> row of cells (fixed y, z), arithmetic only: rows whose slab of space [x-range of the
            //      warp's box] x [the row's y, z interval] no sphere reaches are dropped before anything is loaded (a
            //      warp that straddles a Morton jump has a box of thousands of mostly irrelevant rows)
            const int row = rbase + lane;
            const bool rv = row < nrows;
            const int yy = by0 + (rv ? row % nyb : 0), zz = bz0 + (rv ? row / nyb : 0);
            const float ylo = g.mn[1] + yy * g.h, zlo = g.mn[2] + zz * g.h;
            const bool row_need = reaches(xlo - slack, ylo - slack, zlo - slack, xhi + slack, ylo + g.h + slack, zlo + g.h + slack) && rv;
            const int cnt = row_need ? bx1 - bx0 + 1 : 0;
            // ---- step 2: the cells of the surviving rows, flattened over the lanes: lane <-> cell, tight-box test
            int incl = cnt;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int o = __shfl_up_sync(full, incl, off);
                if (lane >= off) incl += o;
            }
            const int total = __shfl_sync(full, incl, 31);
            const int excl = incl - cnt;
#ifdef ICPB_COOP_CLOCKS
            dbg_cells += total;
#endif
            const int rowbase_l = (zz * ny + yy) * nx + bx0;
            for (int cb = 0; cb < total; cb += 32) {
                const int ci = cb + lane;
                int r = 0; // the last lane whose first flattened cell is <= ci
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const int e = __shfl_sync(full, excl, (r + step) & 31);
                    if (r + step < 32 && e <= ci) r += step;
                }
                const int r_excl = __shfl_sync(full, excl, r), r_base = __shfl_sync(full, rowbase_l, r);
                int cell = -1, t0 = 0, len = 0;
                float4 blo = make_float4(0.f, 0.f, 0.f, 0.f), bhi = blo;
                if (ci < total) {
                    cell = r_base + (ci - r_excl);
                    t0 = __ldg(&gstart[cell * g.sub]);
                    len = __ldg(&gstart[(cell + 1) * g.sub]) - t0;
                    // unconditionally: an empty cell's box is inverted and reaches nothing, and the loads leave together
                    // with the two above instead of one round trip later
                    const float4 *bx = g.sub == 8 ? gboxc : gbox;
                    blo = __ldg(&bx[2 * (size_t)cell]); bhi = __ldg(&bx[2 * (size_t)cell + 1]);
                }
                const bool need = reaches(blo.x, blo.y, blo.z, bhi.x, bhi.y, bhi.z) && len > 0;
                // stage: the runs of the lanes that hold a needed cell (my_pos, my_len; 0 = none), FLATTENED over the warp
                // -- element j of their concatenation goes to lane j mod 32, found by a five-step search over the runs'
                // prefix sums -- so that every load instruction carries 32 targets and consecutive ones are independent
                // (copying run after run left most lanes idle on the short runs of child cells and exposed one L2
                // latency per run).  Centred into the batch; a full batch is evaluated at once.
                auto stage_runs = [&](int my_pos, int my_len) {
                    int incl2 = my_len;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const int o = __shfl_up_sync(full, incl2, off);
                        if (lane >= off) incl2 += o;
                    }
                    const int all = __shfl_sync(full, incl2, 31);
                    const int excl2 = incl2 - my_len;
                    for (int base = 0; base < all;) {
                        const int take = min(all - base, kCoopCap - fill);
#pragma unroll 2
                        for (int e0 = 0; e0 < take; e0 += 32) { // warp-uniform trip count: the search shuffles need every lane
                            const int e = e0 + lane;
                            const int j = base + e;
                            int r = 0; // the last lane whose run starts at or before element j
#pragma unroll
                            for (int step = 16; step >= 1; step >>= 1) {
                                const int x2 = __shfl_sync(full, excl2, (r + step) & 31);
                                if (r + step < 32 && x2 <= j) r += step;
                            }
                            const int src = __shfl_sync(full, my_pos, r) + (j - __shfl_sync(full, excl2, r));
                            if (e < take) {
                                const float4 t = __ldg(&sorted[src]);
                                const float tx = t.x - cx, ty = t.y - cy, tz = t.z - cz;
                                buf.xs[fill + e] = tx; buf.ys[fill + e] = ty; buf.zs[fill + e] = tz;
                                buf.ns[fill + e] = __fmaf_rn(tz, tz, __fmaf_rn(ty, ty, tx * tx));
                                buf.gp[fill + e] = src;
                            }
                        }
                        fill += take; base += take;
                        if (fill == kCoopCap) {
                            __syncwarp();
                            coop_batch(buf, fill, sorted, p, qx, qy, qz, A, live, best, bb);
                            __syncwarp();
                            staged += fill;
                            fill = 0;
                        }
                    }
                };
                const unsigned needed = __ballot_sync(full, need);
                if (g.sub == 1) {
                    if (needed) stage_runs(t0, need ? (probe ? 1 : len) : 0);
                } else {
                    // ---- step 3: the eight children of the needed cells, flattened over the lanes again: their own tight
                    //      boxes decide (half the edge: a quarter of the points a cell would bring along for its rim)
                    const int kids = 8 * __popc(needed);
                    for (int kb = 0; kb < kids; kb += 32) {
                        const int ki = kb + lane;
                        int k0 = 0, klen = 0;
                        float4 klo = make_float4(0.f, 0.f, 0.f, 0.f), khi = klo;
                        const int src = __fns(needed, 0, (ki >> 3) + 1); // lane holding the (ki / 8)-th needed cell
                        const int pcell = __shfl_sync(full, cell, src & 31);
                        if (ki < kids) {
                            const int fine = pcell * 8 + (ki & 7);
                            k0 = __ldg(&gstart[fine]);
                            klen = __ldg(&gstart[fine + 1]) - k0;
                            klo = __ldg(&gbox[2 * (size_t)fine]); khi = __ldg(&gbox[2 * (size_t)fine + 1]);
                        }
                        const bool kneed = reaches(klo.x, klo.y, klo.z, khi.x, khi.y, khi.z) && klen > 0;
                        if (__any_sync(full, kneed)) stage_runs(k0, kneed ? (probe ? 1 : klen) : 0);
                    }
                }
            }
        }
        if (fill > 0) {
            // pad to a multiple of four with candidates that can never be the filter's minimum.  The exact stage evaluates
            // all four members of the best group, pads included, so a pad must still name a REAL target of THIS
            // registration: the first one of its cells (in a batch position 0 of the shared array is another
            // registration's point)
            if (lane < 4 && (fill & 3) != 0 && fill + lane < ((fill + 3) & ~3)) {
                buf.xs[fill + lane] = 0.f; buf.ys[fill + lane] = 0.f; buf.zs[fill + lane] = 0.f;
                buf.ns[fill + lane] = CUDART_INF_F; buf.gp[fill + lane] = __ldg(&gstart[0]);
            }
            __syncwarp();
            coop_batch(buf, (fill + 3) & ~3, sorted, p, qx, qy, qz, A, live, best, bb);
            __syncwarp();
            staged += fill;
        }
        // a lane that found nothing borrows its neighbours' finds: their best targets are real candidates a few
        // centimetres further away, which bounds the lane's next ball far better than quadrupling the radius
        const bool empty_handed = !done && !(best.d < CUDART_INF_F);
        if (__any_sync(full, empty_handed)) {
            for (unsigned src = __ballot_sync(full, best.d < CUDART_INF_F); src; src &= src - 1) {
                const int j = __ffs(src) - 1;
                const float4 t = make_float4(__shfl_sync(full, bb.x, j), __shfl_sync(full, bb.y, j), __shfl_sync(full, bb.z, j),
                                             __int_as_float(__shfl_sync(full, best.i, j)));
                if (empty_handed) coop_take(p, t, best, bb);
            }
        }
        if (probe) {
            // only samples were looked at: nothing is finished.  The nearest sample is a real target and bounds the
            // ball; a lane that saw none grows its ball as after an empty real round and the warp probes again (unless
            // the lane is at the cap already: then the real round comes next)
            bool again = false;
            if (!done) {
                if (best.d < CUDART_INF_F) { rad = fminf(rad, best.d); loose_seed = false; }
                else if (rr < max_reach) { rad = 4.f * rad; again = true; }
            }
            probed = !__any_sync(full, again); // warp-uniform
            continue;
        }
        probed = false;
        if (!done) {
            // finished: the best lies inside the searched ball (or the whole acceptance ball was searched)
            if (best.d <= rr || rr >= max_reach) done = true;
            else rad = (best.d < CUDART_INF_F) ? best.d : 4.f * rad;
        }
    }
#ifdef ICPB_COOP_CLOCKS
    {
        const long long dt = clock64() - clk0;
        if (lane == 0 && pass == ICPB_COOP_CLOCKS && k / 32 < 16384) {
            int *o = g_coop_dbg + 8 * (k / 32);
            o[0] = (int)dt; o[1] = dbg_rounds; o[2] = dbg_rows; o[3] = dbg_cells; o[4] = dbg_box; o[5] = (int)staged;
            unsigned long long gt1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
            o[6] = (int)(gt0 & 0x7fffffff); o[7] = (int)__smid() | ((int)((gt1 - gt0) & 0xfffff) << 8);
        }
    }
#endif
    if (d.gord && lane == 0) {
        // this warp's class for the next pass (0 = heaviest), by the time it took: candidates staged AND cells visited
        // (a warp across a jump of the Morton curve stages little and tests thousands of rows)
        const long long c = clock64() - ord_clk0;
        const int b = c >= 131072 ? 0 : c >= 98304 ? 1 : c >= 73728 ? 2 : c >= 57344 ? 3 : c >= 45056 ? 4 : c >= 34816 ? 5 : c >= 24576 ? 6 : 7;
        const int slot = atomicAdd(&d.gord_count[pass * kOrderBins + b], 1);
        d.gord[((pass & 1) * kOrderBins + b) * ord_stride + slot] = wsel;
    }
    if (d.gpairs && lane == 0 && staged) atomicAdd(d.gpairs, staged * 32ull);
    if (!valid) return;
    if (!done) deferred = true; // round limit (not reachable with radii that quadruple up to the acceptance radius)
    if (deferred) {
        // the warp-per-query kernel finishes it (and writes gnb); its partial best, a real candidate or none, goes along
        const int slot = atomicAdd(&d.gheavy_count[pass], 1);
        d.gheavy[slot] = i;
        d.idx[i] = best.i;
        d.dist[i] = best.d;
        d.gseed[k] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)); // no seed next pass: the search starts over
        return;
    }
    const bool accepted = best.d < g.max_nn;
    const float4 rec = make_float4(bb.x, bb.y, bb.z, __int_as_float(accepted ? best.i : -1));
    d.gnb[i] = rec;                                                                      // nn_finalize: sums, idx, dist
    d.gseed[k] = accepted ? rec : make_float4(0.f, 0.f, 0.f, __int_as_float(beyond_reach(g, p) ? -1 : -2));
}

// kSingle: one registration -- its descriptor and cell geometry arrive BY VALUE as kernel parameters (constant bank) and
// the loop state is addressed from an argument, so a warp's first useful load (its work item, its queries) does not wait
// behind a chain of pointer loads (descriptor -> state, descriptor -> geometry): two round trips of ~8 per work item.
template <bool kSingle>
__global__ void __launch_bounds__(32 * kCoopWarps, ICPB_COOP_MINB) nn_grid_coop_kernel(const RegDesc *__restrict__ descs, const IcpState *states,
                                                                                     const GridMeta *__restrict__ gmetas, int pass, float coop_r,
                                                                                     const RegDesc d1, const GridMeta g1)
{
    pdl_enter(); // icpb_internal.h: the grid before this one is complete from here on
    const IcpState *st = states + blockIdx.z; // == d.st
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // one work item (32 sorted query slots) per warp.  Resident warps drawing items from a counter instead measured
    // slower (154 against 141 us per pass at full resolution): the relaunch of CTAs is not what the SMs wait for, and
    // the counter adds one more round trip to every item's chain of dependent loads.
    const int wpos = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    // everything the first round trip can fetch goes out together: the motion and the flags (one line of the state)
    // and, where the descriptor is a parameter, the work item
    int wsel = wpos;
    if (kSingle && d1.gord && pass > 0 && wpos < ((d1.n + 31) >> 5)) wsel = __ldcg(&d1.gord_flat[wpos]);
    CoopMotion mo;
#pragma unroll
    for (int k = 0; k < 9; ++k) mo.R[k] = st->Rf[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) mo.T[k] = st->tf[k];
    mo.apply = st->apply;
    if (st->done) return;
    __shared__ __align__(16) CoopBuf s_buf[kCoopWarps];
    if (kSingle) {
        if (wpos < ((d1.n + 31) >> 5)) coop_warp(d1, g1, mo, s_buf[wid], pass, coop_r, wsel, lane);
    } else {
        const RegDesc &d = descs[blockIdx.z];
        if (wpos < ((d.n + 31) >> 5)) {
            if (d.gord && pass > 0) wsel = __ldcg(&d.gord_flat[wpos]); // written by nn_finalize_coop_kernel from last pass's classes
            const GridMeta g = gmetas[blockIdx.z]; // == *d.grid, addressed from the argument
            coop_warp(d, g, mo, s_buf[wid], pass, coop_r, wsel, lane);
        }
    }
}

void launch_nn_grid(const RegDesc *descs, const IcpState *states, const GridMeta *gmetas, int batch, int max_n, int pass,
                    int sm_count, cudaStream_t s, float coop_r, const RegDesc *h_desc0, const GridMeta *h_grid0)
{
    dim3 hgrid(sm_count * 8, 1, batch);
    if (coop_r > 0.f) {
        hgrid.x = sm_count * 2; // a handful of queries at most reach the fall-back (it strides over its list) // warp-cooperative search (default), open queries finished from shell 0
        dim3 grid((max_n + 32 * kCoopWarps - 1) / (32 * kCoopWarps), 1, batch);
        if (batch == 1 && h_desc0 && h_grid0)
            launch_pdl(nn_grid_coop_kernel<true>, grid, dim3(32 * kCoopWarps), 0, s, descs, states, gmetas, pass, coop_r, *h_desc0, *h_grid0);
        else
            launch_pdl(nn_grid_coop_kernel<false>, grid, dim3(32 * kCoopWarps), 0, s, descs, states, gmetas, pass, coop_r, RegDesc{}, GridMeta{});
        launch_pdl<true>(nn_grid_heavy_kernel, hgrid, dim3(128), 0, s, descs, pass, 0);
        return;
    }
    dim3 grid((max_n + 127) / 128, 1, batch);
    launch_pdl<true>(nn_grid_kernel, grid, dim3(128), 0, s, descs, pass);
    launch_pdl<true>(nn_grid_heavy_kernel, hgrid, dim3(128), 0, s, descs, pass, -1);
}

} // namespace icpb

#ifdef ICPB_COOP_CLOCKS
extern "C" int icpb_debug_coop_clocks(int *out, int n_ints)
{
    return (int)cudaMemcpyFromSymbol(out, icpb::g_coop_dbg, sizeof(int) * (size_t)n_ints);
}
#endif

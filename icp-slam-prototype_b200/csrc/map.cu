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

__global__ void map_endpoints_kernel(MapDev m, const float4 *__restrict__ pts, int n, int rule, int delta,
                                     int max_conf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    long long lin = -1;
    if (i < n) {
        float4 p = pts[i];
        int vx = voxel_axis(p.x, m.cell, m.dims[0]);
        int vy = voxel_axis(p.y, m.cell, m.dims[1]);
        int vz = voxel_axis(p.z, m.cell, m.dims[2]);
        if (vz >= m.z_lo && vz < m.z_hi) lin = ((long long)vx * m.dims[1] + vy) * m.zs + (vz - m.z_lo);
    }
    // neighbouring pixels land in the same voxel: one CAS per distinct voxel per warp
    const unsigned peers = __match_any_sync(0xffffffffu, lin);
    if (lin < 0) return;
    if ((__ffs(peers) - 1) != lane) return;
    const int k = __popc(peers);
    byte_rmw(m.grid, (size_t)lin, [&](uint32_t c) {
        for (int r = 0; r < k && c != 255u; ++r) c = rule_apply(c, rule, delta, max_conf);
        return c;
    });
}

void launch_map_endpoints(const MapDev &m, const float4 *pts, int n, int rule, int delta, int max_conf,
                          cudaStream_t s)
{
    if (n <= 0) return;
    map_endpoints_kernel<<<(n + 255) / 256, 256, 0, s>>>(m, pts, n, rule, delta, max_conf);
}

// M4 phase 1: exact integer Amanatides-Woo walk from the origin voxel centre to
// the endpoint voxel centre.  Axis k crosses its i-th wall at t = (2i+1)/(2 n_k);
// walls are ordered by the cross-multiplied integers e_k, ties x < y < z.
// Every voxel entered except the endpoint voxel is decremented with clamp at 0.
__global__ void map_rays_kernel(MapDev m, const float4 *__restrict__ pts, int n, int ox, int oy, int oz,
                                int delta_dec, unsigned long long *visited)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long my_visits = 0;
    if (i < n) {
        float4 p = pts[i];
        const int ex_ = voxel_axis(p.x, m.cell, m.dims[0]);
        const int ey_ = voxel_axis(p.y, m.cell, m.dims[1]);
        const int ez_ = voxel_axis(p.z, m.cell, m.dims[2]);
        const long long nx = abs(ex_ - ox), ny = abs(ey_ - oy), nz = abs(ez_ - oz);
        const int sx = (ex_ > ox) - (ex_ < ox), sy = (ey_ > oy) - (ey_ < oy), sz = (ez_ > oz) - (ez_ < oz);
        const long long mx = nx ? nx : 1, my = ny ? ny : 1, mz = nz ? nz : 1;
        const long long INF = (long long)1 << 62;
        long long ex = nx ? my * mz : INF, ey = ny ? mx * mz : INF, ez = nz ? mx * my : INF;
        const long long dxs = 2 * my * mz, dys = 2 * mx * mz, dzs = 2 * mx * my;
        const long long steps = nx + ny + nz;
        int x = ox, y = oy, z = oz;
        long long cx = 0, cy = 0, cz = 0;
        const int dimY = m.dims[1];
        for (long long s = 0; s + 1 < steps; ++s) {
            if (ex <= ey && ex <= ez) { x += sx; ++cx; ex = (cx < nx) ? ex + dxs : INF; }
            else if (ey <= ez) { y += sy; ++cy; ey = (cy < ny) ? ey + dys : INF; }
            else { z += sz; ++cz; ez = (cz < nz) ? ez + dzs : INF; }
            if (z < m.z_lo || z >= m.z_hi) continue;
            const size_t lin = ((size_t)x * dimY + y) * m.zs + (z - m.z_lo);
            // phase 1 only lowers values, so a cached non-zero byte is at worst stale-high: the CAS re-reads it
            if (m.grid[lin] != 0)
                byte_rmw(m.grid, lin, [&](uint32_t c) { return c > (uint32_t)delta_dec ? c - delta_dec : 0u; });
        }
        my_visits = steps > 0 ? (unsigned long long)(steps - 1) : 0ull;
    }
    if (visited) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) my_visits += __shfl_xor_sync(0xffffffffu, my_visits, off);
        if ((threadIdx.x & 31) == 0 && my_visits) atomicAdd(visited, my_visits);
    }
}

void launch_map_rays(const MapDev &m, const float4 *pts, int n, const float origin[3], int delta_dec,
                     unsigned long long *visited, cudaStream_t s)
{
    if (n <= 0) return;
    // origin voxel: same quantisation as any point (map.cpp:226 uses getVoxelCoordinates too)
    int o[3];
    for (int k = 0; k < 3; ++k) {
        int q = (int)(origin[k] / m.cell);
        if (q < 0) q = 0;
        if (q >= m.dims[k]) q = m.dims[k] - 1;
        o[k] = q;
    }
    map_rays_kernel<<<(n + 127) / 128, 128, 0, s>>>(m, pts, n, o[0], o[1], o[2], delta_dec, visited);
}

} // namespace icpb

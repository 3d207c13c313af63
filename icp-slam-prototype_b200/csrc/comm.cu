// Multi-GPU entry points of the C-ABI (include/icpb200.h, "multi-GPU"): an NCCL communicator per rank and the z-slab
// certainty map.  SURVEY.md 8e: the map of map.hpp:20-37 shards by z-slab with ONE exchange per frame group -- the
// all-gather of the ranks' lifted row bands; batches of registrations shard with no data-path collective at all.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2"; in a process that already carries torch's NCCL the loader hands
// back that copy): libicpb200.so has no link-time dependency on it, and single-GPU hosts never touch it.
#include <dlfcn.h>
#include <nccl.h> // types and prototypes only; the functions are looked up with dlsym

#include <cstring>
#include <new>
#include <vector>

#include "icpb_internal.h"

using namespace icpb;

namespace {

struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};

NcclApi &nccl()
{
    static NcclApi api = []() {
        NcclApi a;
        a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) return a;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.lib, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.lib, "ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
        a.AllGather = (decltype(a.AllGather))dlsym(a.lib, "ncclAllGather");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.GetErrorString;
        return a;
    }();
    return api;
}

int nccl_fail(icpb_ctx *ctx, const char *what, ncclResult_t r)
{
    std::string msg = what;
    if (nccl().ok) { msg += ": "; msg += nccl().GetErrorString(r); }
    return api_fail(ctx, ICPB_ERR_NCCL, msg.c_str());
}

#define CUC(ctx, call)                                                                  \
    do {                                                                                \
        cudaError_t ce__ = (call);                                                      \
        if (ce__ != cudaSuccess) return api_fail((ctx), ICPB_ERR_CUDA, #call, ce__);     \
    } while (0)

} // namespace

struct icpb_comm {
    icpb_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    cudaStream_t stream = nullptr; // collectives run here, ordered against the compute streams by events
    void *d_stage = nullptr;       // device staging of icpb_comm_allgather_host
    size_t stage_bytes = 0;
};

constexpr int kSeqCounters = 4096; // ray counters of a sequence call: one per frame, zeroed by one memset

struct icpb_slabmap {
    icpb_ctx *ctx = nullptr;
    icpb_comm *comm = nullptr;
    icpb_map *map = nullptr;
    int w = 0, h = 0, row0 = 0, row1 = 0, band_cap = 0;
    int world = 1, rank = 0;
    int group = 0;                 // frames the buffers hold per exchange
    float4 *send[2] = {nullptr, nullptr}, *recv[2] = {nullptr, nullptr};
    void *tile_state = nullptr;
    unsigned int *next_ray = nullptr;
    cudaStream_t s_lift = nullptr;
    cudaEvent_t ev_lifted[2] = {nullptr, nullptr}, ev_gathered[2] = {nullptr, nullptr}, ev_walked[2] = {nullptr, nullptr};
    cudaEvent_t ev_join = nullptr;
};

extern "C" {

int icpb_comm_unique_id(uint8_t id[ICPB_COMM_ID_BYTES])
{
    if (!id) return ICPB_ERR_INVALID;
    if (!nccl().ok) return api_fail(nullptr, ICPB_ERR_NCCL, "libnccl.so.2 could not be loaded");
    static_assert(sizeof(ncclUniqueId) == ICPB_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    ncclResult_t r = nccl().GetUniqueId(&u);
    if (r != ncclSuccess) return nccl_fail(nullptr, "ncclGetUniqueId", r);
    memcpy(id, &u, ICPB_COMM_ID_BYTES);
    return ICPB_OK;
}

int icpb_comm_create(icpb_ctx *ctx, int world, int rank, const uint8_t id[ICPB_COMM_ID_BYTES], icpb_comm **out)
{
    if (!ctx || !out || !id || world < 1 || rank < 0 || rank >= world) return ICPB_ERR_INVALID;
    *out = nullptr;
    if (world > kMaxBands) return api_fail(ctx, ICPB_ERR_INVALID, "icpb_comm_create: more ranks than kMaxBands (16)");
    if (!nccl().ok) return api_fail(ctx, ICPB_ERR_NCCL, "libnccl.so.2 could not be loaded");
    CUC(ctx, cudaSetDevice(ctx->device));
    icpb_comm *c = new (std::nothrow) icpb_comm();
    if (!c) return api_fail(ctx, ICPB_ERR_INVALID, "out of host memory");
    c->ctx = ctx; c->world = world; c->rank = rank;
    ncclUniqueId u;
    memcpy(&u, id, ICPB_COMM_ID_BYTES);
    ncclResult_t r = nccl().CommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) { delete c; return nccl_fail(ctx, "ncclCommInitRank", r); }
    cudaError_t ce = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) { nccl().CommDestroy(c->comm); delete c; return api_fail(ctx, ICPB_ERR_CUDA, "cudaStreamCreate", ce); }
    *out = c;
    return ICPB_OK;
}

int icpb_comm_destroy(icpb_comm *comm)
{
    if (!comm) return ICPB_OK;
    cudaSetDevice(comm->ctx->device);
    cudaStreamSynchronize(comm->stream);
    if (comm->comm) nccl().CommDestroy(comm->comm);
    if (comm->d_stage) cudaFree(comm->d_stage);
    cudaStreamDestroy(comm->stream);
    delete comm;
    return ICPB_OK;
}

int icpb_comm_rank(const icpb_comm *comm, int *world, int *rank)
{
    if (!comm) return ICPB_ERR_INVALID;
    if (world) *world = comm->world;
    if (rank) *rank = comm->rank;
    return ICPB_OK;
}

int icpb_comm_shard_range(const icpb_comm *comm, long long n_items, long long *lo, long long *hi)
{
    if (!lo || !hi || n_items < 0) return ICPB_ERR_INVALID;
    const long long world = comm ? comm->world : 1, rank = comm ? comm->rank : 0;
    const long long base = n_items / world, extra = n_items % world;
    *lo = rank * base + (rank < extra ? rank : extra);
    *hi = *lo + base + (rank < extra ? 1 : 0);
    return ICPB_OK;
}

int icpb_comm_allgather_host(icpb_comm *comm, const void *send, void *recv, long long bytes)
{
    if (!comm || !send || !recv || bytes <= 0) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = comm->ctx;
    CUC(ctx, cudaSetDevice(ctx->device));
    const size_t need = (size_t)bytes * (size_t)(comm->world + 1);
    if (comm->stage_bytes < need) {
        CUC(ctx, cudaStreamSynchronize(comm->stream));
        if (comm->d_stage) CUC(ctx, cudaFree(comm->d_stage));
        comm->d_stage = nullptr; comm->stage_bytes = 0;
        CUC(ctx, cudaMalloc(&comm->d_stage, need));
        comm->stage_bytes = need;
    }
    char *d_send = (char *)comm->d_stage, *d_recv = d_send + bytes;
    CUC(ctx, cudaMemcpyAsync(d_send, send, (size_t)bytes, cudaMemcpyHostToDevice, comm->stream));
    ncclResult_t r = nccl().AllGather(d_send, d_recv, (size_t)bytes, ncclChar, comm->comm, comm->stream);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllGather", r);
    CUC(ctx, cudaMemcpyAsync(recv, d_recv, (size_t)bytes * comm->world, cudaMemcpyDeviceToHost, comm->stream));
    CUC(ctx, cudaStreamSynchronize(comm->stream));
    return ICPB_OK;
}

// ---- z-slab map ----------------------------------------------------------------------------------------------------

int icpb_slabmap_create(icpb_ctx *ctx, icpb_comm *comm, const int dims[3], float cell, const int *bounds, int w, int h,
                        icpb_slabmap **out)
{
    if (!ctx || !dims || !out || w <= 0 || h <= 0) return ICPB_ERR_INVALID;
    *out = nullptr;
    if (comm && comm->ctx != ctx) return api_fail(ctx, ICPB_ERR_INVALID, "communicator belongs to another context");
    const int world = comm ? comm->world : 1, rank = comm ? comm->rank : 0;
    int z_lo, z_hi;
    if (bounds) {
        for (int g = 0; g < world; ++g)
            if (bounds[g] >= bounds[g + 1]) return api_fail(ctx, ICPB_ERR_INVALID, "slab bounds must increase");
        if (bounds[0] != 0 || bounds[world] != dims[2]) return api_fail(ctx, ICPB_ERR_INVALID, "slab bounds must span [0, dims[2]]");
        z_lo = bounds[rank]; z_hi = bounds[rank + 1];
    } else {
        const int base = dims[2] / world, extra = dims[2] % world;
        z_lo = rank * base + (rank < extra ? rank : extra);
        z_hi = z_lo + base + (rank < extra ? 1 : 0);
    }
    CUC(ctx, cudaSetDevice(ctx->device));
    icpb_slabmap *sm = new (std::nothrow) icpb_slabmap();
    if (!sm) return api_fail(ctx, ICPB_ERR_INVALID, "out of host memory");
    sm->ctx = ctx; sm->comm = comm; sm->world = world; sm->rank = rank; sm->w = w; sm->h = h;
    int rc = icpb_map_create(ctx, dims, cell, z_lo, z_hi, &sm->map);
    if (rc) { delete sm; return rc; }
    // rank r lifts image rows [row0, row1): contiguous blocks, so that rank order is raster order
    const int rows = (h + world - 1) / world;
    sm->row0 = std::min(h, rank * rows);
    sm->row1 = std::min(h, sm->row0 + rows);
    sm->band_cap = rows * w;
    cudaError_t ce = cudaStreamCreateWithFlags(&sm->s_lift, cudaStreamNonBlocking);
    for (int b = 0; b < 2 && ce == cudaSuccess; ++b) {
        ce = cudaEventCreateWithFlags(&sm->ev_lifted[b], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&sm->ev_gathered[b], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&sm->ev_walked[b], cudaEventDisableTiming);
    }
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&sm->ev_join, cudaEventDisableTiming);
    const int tiles = backproject_tiles(w, std::max(1, sm->row1 - sm->row0));
    const size_t ts_bytes = sizeof(unsigned long long) * (3 * (size_t)tiles + 2);
    if (ce == cudaSuccess) ce = cudaMalloc(&sm->tile_state, ts_bytes);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(sm->tile_state, 0, ts_bytes, ctx->stream);
    if (ce == cudaSuccess) ce = cudaMalloc((void **)&sm->next_ray, sizeof(unsigned int) * kSeqCounters);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
    if (ce != cudaSuccess) { icpb_slabmap_destroy(sm); return api_fail(ctx, ICPB_ERR_CUDA, "icpb_slabmap_create", ce); }
    *out = sm;
    return ICPB_OK;
}

int icpb_slabmap_destroy(icpb_slabmap *sm)
{
    if (!sm) return ICPB_OK;
    cudaSetDevice(sm->ctx->device);
    if (sm->s_lift) cudaStreamSynchronize(sm->s_lift);
    if (sm->comm) cudaStreamSynchronize(sm->comm->stream);
    cudaStreamSynchronize(sm->ctx->stream);
    for (int b = 0; b < 2; ++b) {
        if (sm->send[b]) cudaFree(sm->send[b]);
        if (sm->recv[b]) cudaFree(sm->recv[b]);
        if (sm->ev_lifted[b]) cudaEventDestroy(sm->ev_lifted[b]);
        if (sm->ev_gathered[b]) cudaEventDestroy(sm->ev_gathered[b]);
        if (sm->ev_walked[b]) cudaEventDestroy(sm->ev_walked[b]);
    }
    if (sm->ev_join) cudaEventDestroy(sm->ev_join);
    if (sm->tile_state) cudaFree(sm->tile_state);
    if (sm->next_ray) cudaFree(sm->next_ray);
    if (sm->s_lift) cudaStreamDestroy(sm->s_lift);
    if (sm->map) icpb_map_destroy(sm->map);
    delete sm;
    return ICPB_OK;
}

int icpb_slabmap_local(icpb_slabmap *sm, icpb_map **map, int *z_lo, int *z_hi)
{
    if (!sm) return ICPB_ERR_INVALID;
    if (map) *map = sm->map;
    if (z_lo) *z_lo = sm->map->dev.z_lo;
    if (z_hi) *z_hi = sm->map->dev.z_hi;
    return ICPB_OK;
}

int icpb_slabmap_integrate_sequence_device(icpb_slabmap *sm, const void *d_depths, int frames, const icpb_intrinsics *K,
                                           const float *R, const float *t, int delta_dec, int delta_inc,
                                           int frames_per_exchange)
{
    if (!sm || !d_depths || !K || !R || !t || frames <= 0) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = sm->ctx;
    if (delta_dec < 0 || delta_dec > 255 || delta_inc < 0 || delta_inc > 255)
        return api_fail(ctx, ICPB_ERR_INVALID, "delta out of [0,255]");
    CUC(ctx, cudaSetDevice(ctx->device));
    const int k = std::max(1, std::min(frames_per_exchange, frames));
    const long long band_rows = (long long)sm->band_cap + 1; // header row + points
    if (k > sm->group) { // (re)allocate the double buffers for k frames per exchange
        CUC(ctx, cudaStreamSynchronize(ctx->stream));
        CUC(ctx, cudaStreamSynchronize(sm->s_lift));
        if (sm->comm) CUC(ctx, cudaStreamSynchronize(sm->comm->stream));
        for (int b = 0; b < 2; ++b) {
            if (sm->send[b]) CUC(ctx, cudaFree(sm->send[b]));
            if (sm->recv[b]) CUC(ctx, cudaFree(sm->recv[b]));
            sm->send[b] = sm->recv[b] = nullptr;
            CUC(ctx, cudaMalloc((void **)&sm->send[b], sizeof(float4) * (size_t)(band_rows * k)));
            if (sm->world > 1) CUC(ctx, cudaMalloc((void **)&sm->recv[b], sizeof(float4) * (size_t)(band_rows * k * sm->world)));
        }
        sm->group = k;
    }
    cudaStream_t s_main = ctx->stream, s_lift = sm->s_lift, s_comm = sm->comm ? sm->comm->stream : nullptr;
    // everything enqueued on the context stream before this call is visible to the other two streams
    CUC(ctx, cudaEventRecord(sm->ev_join, s_main));
    CUC(ctx, cudaStreamWaitEvent(s_lift, sm->ev_join, 0));
    if (s_comm) CUC(ctx, cudaStreamWaitEvent(s_comm, sm->ev_join, 0));
    const size_t frame_px = (size_t)sm->w * sm->h;
    const int groups = (frames + k - 1) / k;
    const bool fresh_counters = frames <= kSeqCounters;
    if (fresh_counters) CUC(ctx, cudaMemsetAsync(sm->next_ray, 0, sizeof(unsigned int) * (size_t)frames, s_main));
    for (int g = 0; g < groups; ++g) {
        const int b = g & 1;
        const int f0 = g * k, kk = std::min(k, frames - f0);
        // ---- lift stream: this rank's row band of every frame of the group, into send[b]
        if (g >= 2) CUC(ctx, cudaStreamWaitEvent(s_lift, sm->world > 1 ? sm->ev_gathered[b] : sm->ev_walked[b], 0)); // send[b] free
        for (int j = 0; j < kk; ++j) {
            const int f = f0 + j;
            float4 *band = sm->send[b] + band_rows * j;
            if (sm->row1 > sm->row0) {
                int rc = api_lift_band(ctx, s_lift, sm->tile_state, (const uint16_t *)d_depths + frame_px * f, sm->w, sm->h,
                                       sm->row0, sm->row1, K, R + 9 * f, t + 3 * f, band, sm->band_cap);
                if (rc) return rc;
            } else {
                CUC(ctx, cudaMemsetAsync(band, 0, sizeof(float4), s_lift)); // more ranks than image rows: an empty band
            }
        }
        CUC(ctx, cudaEventRecord(sm->ev_lifted[b], s_lift));
        // ---- exchange stream: one all-gather per group
        const float4 *src_base = sm->send[b];
        long long stride = band_rows; // rows between the bands of consecutive RANKS for one frame
        if (sm->world > 1) {
            CUC(ctx, cudaStreamWaitEvent(s_comm, sm->ev_lifted[b], 0));
            if (g >= 2) CUC(ctx, cudaStreamWaitEvent(s_comm, sm->ev_walked[b], 0)); // recv[b] free
            ncclResult_t r = nccl().AllGather(sm->send[b], sm->recv[b], (size_t)(band_rows * kk) * 4, ncclFloat,
                                              sm->comm->comm, s_comm);
            if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllGather", r);
            CUC(ctx, cudaEventRecord(sm->ev_gathered[b], s_comm));
            CUC(ctx, cudaStreamWaitEvent(s_main, sm->ev_gathered[b], 0));
            src_base = sm->recv[b];
            stride = band_rows * kk; // rank r's kk bands lie back to back
        } else {
            CUC(ctx, cudaStreamWaitEvent(s_main, sm->ev_lifted[b], 0));
        }
        // ---- context stream: walk + endpoints of every frame of the group, bands read in place
        for (int j = 0; j < kk; ++j) {
            const int f = f0 + j;
            const PointSrc src = band_src(src_base + band_rows * j, sm->world, sm->band_cap, stride);
            int sp = api_span_begin(ctx, ICPB_PROF_MAP_RAYS);
            launch_map_rays(sm->map->dev, src, t + 3 * f, delta_dec, nullptr, fresh_counters ? sm->next_ray + f : sm->next_ray,
                            ctx->sm_count, s_main, fresh_counters ? kCounterIsZero : nullptr);
            api_span_end(ctx, sp);
            sp = api_span_begin(ctx, ICPB_PROF_MAP_ENDPOINTS);
            launch_map_endpoints(sm->map->dev, src, ICPB_RULE_A, delta_inc, 0, s_main);
            api_span_end(ctx, sp);
            ctx->launches += 2;
        }
        CUC(ctx, cudaEventRecord(sm->ev_walked[b], s_main));
        CUC(ctx, cudaGetLastError());
    }
    return ICPB_OK;
}

} // extern "C"

// Internal declarations shared by the translation units of libicpb200.so.
// Nothing here is part of the C-ABI (include/icpb200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <string>
#include <vector>

#include "icpb200.h"

namespace icpb {

// ---- tunables of the brute-force NN kernel (DESIGN.md "N1-N3") -----------
constexpr int kGroup = 32;          // targets per min-tracking group (exactly re-scanned per query)
constexpr int kNnThreads = 128;     // threads per CTA of nn_partial
#ifndef ICPB_STAGE_GROUPS
#define ICPB_STAGE_GROUPS 16
#endif
#ifndef ICPB_STAGES
#define ICPB_STAGES 4
#endif
#ifndef ICPB_UNROLL_J
#define ICPB_UNROLL_J 2
#endif
constexpr int kStageGroups = ICPB_STAGE_GROUPS; // groups per shared-memory stage (16*32*12 B = 6 KB)
constexpr int kStages = ICPB_STAGES;            // TMA bulk-copy ring depth
constexpr int kUnrollJ = ICPB_UNROLL_J;        // unroll of the 4-target inner step
constexpr int kChunk = 256;         // canonical reduction chunk (CANON-3)
constexpr int kTerms = 20;          // a(3) b(3) b a^T(9) d(1) a-b(3) count(1)
constexpr float kPadCoord = 1.0e18f; // padding targets: (a-b)^2 ~ 3e36, finite, never the minimum
// second-best group minimum within this relative band of the best => exact FP64 rescan
constexpr float kBandRel = 1.0f + 1.9073486328125e-06f; // 1 + 2^-19
constexpr float kBandAbs = 1.0e-30f;
// the two filters of the brute-force scan (nn.cu)
constexpr int kFilterCentred = ICPB_FILTER_CENTRED; // |t'|^2 - 2a'.t' about a per-thread centre: 3 per pair + 6 per target per thread
constexpr int kFilterDirect = ICPB_FILTER_DIRECT;   // (a-t)^2 in packed FP32: 6 FMA-pipe lane-ops per pair
constexpr int kFilterWarp = ICPB_FILTER_WARP;       // the centred form about a per-WARP centre over spatially sorted queries
// centred filters (nn_partial_centred / nn_partial_warp / nn_grid_coop): a target whose filter value W exceeds the best
// W by more than kBandCentredA * A + kBandCentredX * max(W_best + A, 0) is strictly farther, in the reference's own
// arithmetic, than the best target -- it can neither win nor tie.  The single derivation is DESIGN.md section 4
// ("error band of the centred filter"): it needs 18 u A + 106 u D (u = 2^-24); the constants carry a third more.
constexpr float kBandCentredA = 24.0f * 5.9604644775390625e-08f;
constexpr float kBandCentredX = 128.0f * 5.9604644775390625e-08f;

// Per-registration device state: iteration control and pose, kept on the
// device for the whole loop (icp.cpp:22-25 globals + locals of :28-285).
struct IcpState {
    float Rf[9];
    float tf[3];
    int apply;          // 1: next pass applies (Rf, tf) to the data cloud while loading it
    int done;           // loop finished; later passes are no-ops
    int iterations;     // i of icp.cpp:152
    int passes;         // association passes executed
    int last_buf;       // ping-pong buffer holding the cloud the last pass associated
    int n_assoc;
    float mse;
    int small_exit;
    int pending_translate; // <3 associations rule: apply tf as a plain translation after the loop
    int rescans;
    int n_log;          // motions recorded in RegDesc::mlog (key-point variant: replayed on the carried cloud)
    int n_nonassoc;     // key-point variant: length of the accumulated reject list
    unsigned int block_counter;
    float rigid[9];
    float camR[9];
    float camP[3];
    float offset[3];
    double PR[9];
    double Pt[3];
};

struct IcpParamsDev {
    int max_iterations;
    float threshold;
    float max_nn_distance;
    int solve_mode;
    float last_translation[3];
};

// Uniform cell grid over the target cloud (ICPB_NN_GRID, grid.cu).
struct GridMeta {
    float mn[3];   // lower corner of the target's bounding box
    float h;       // cell edge
    int dim[3];    // cells per axis (x fastest in the cell index)
    int ncells;    // coarse cells
    int sub;       // sorted-array slots per coarse cell: 1, or 8 = every cell split into 2x2x2 children (child index in the low 3 bits)
    int max_r;     // shells needed to cover the acceptance radius
    int light_r;   // shells every query walks on its own thread before it is handed to a whole warp
    float probe_r; // cooperative search: guessed balls wider than this are probed (one sample per child cell) before they are searched
    float probe_r2; // ... and any ball wider than this
    float max_nn;  // acceptance radius (MAX_NN_COLOR_DISTANCE, icp.hpp:8)
};

// One registration problem as the kernels see it (indexed by blockIdx.z).
struct RegDesc {
    float4 *D[2];            // ping-pong data cloud buffers (16-B points)
    const float4 *tgt;       // target, AoS (exact rescans, gathers)
    const float *tgt_soa;    // target, negated group-tiled SoA: [group][x[32] y[32] z[32]]
    int n, m, ngroups;
    int n_stride;            // row stride of the per-split partial arrays
    float *pm1, *pm2;        // [S][n_stride] best / second-best group minimum (approximate squared distance)
    float *pm3;              // [S][n_stride] centred filter only: third-best group minimum
    int *pg2;                // [S][n_stride] centred filter only: group holding pm2 (-1: none)
    const int *perm;         // warp-centred filter only: [n] original index of the query in sorted slot k
    float *pa;               // [n_stride] centred filters only: |a - c|^2 of every query about its thread's centre
    int *pg;                 // [S][n_stride] group holding pm1
    int *idx;                // [n] nearest target index
    float *dist;             // [n] nearest distance (reference arithmetic)
    double *chunk_sums;      // [chunks][kTerms]
    IcpState *st;
    int *idx_trace;          // nullable [(max_it+1)][n]
    float *dist_trace;
    // key-point variant (icp.cpp:98,255; 8f-2); all null / 0 otherwise
    float4 *carry;           // points that only follow the motion (dataCloud.points while the key-points are associated)
    int n_carry;
    float *mlog;             // [max_iterations][12] the (R, t) applied by every iteration, in order
    int *rej_flag;           // [passes][n] 1: the query was rejected by the acceptance test of that pass (icp.cpp:507-509)
    float4 *rej_pts;         // [passes][n] its coordinates at that pass
    float4 *nonassoc;        // out: the rejects of all passes, pass-major, query order
    int nonassoc_capacity;
    const GridMeta *grid;    // ICPB_NN_GRID only
    const float4 *gsorted;   // targets sorted by cell, w = original index
    const int *gstart;       // [ncells*sub+1] first sorted slot of every cell (positions in gsorted)
    int *gcursor;            // build scratch: running insert position per cell
    const float4 *gbox;      // [2 * ncells * sub] tight bounding box of every (child) cell's targets: (lo.xyz, -), (hi.xyz, -)
    const float4 *gboxc;     // [2 * ncells] the same per coarse cell (sub == 8)
    float4 *gnb;             // [n] cooperative search: the nearest target's coordinates, w = its distance (original query order)
    float4 *gseed;           // [n] the same in SORTED slot order, w = its index bits: the next pass's search ball
    float4 *gq;              // [n] the data cloud as of the last pass, in SORTED slot order (cooperative search only)
    unsigned long long *gpairs; // profiling mode: (query, candidate) pairs the cooperative search put through its filter
    int *gheavy;             // [n] queries still open after the per-thread shells
    int *gheavy_count;       // [passes] length of that list per pass (zeroed once per registration)
    // cooperative search, work order: the warps of pass p are launched heaviest first, by the number of candidates they
    // staged in pass p-1 (kOrderBins classes).  gord: [2][kOrderBins][n_stride/32] warp numbers (ping-pong by pass),
    // gord_count: [passes][kOrderBins] class sizes (zeroed once per registration).  Null: natural order.
    int *gord;
    int *gord_count;
    int *gord_flat;          // [n_stride/32] the classes of the last pass concatenated, heaviest first (nn_finalize_coop_kernel)
};
constexpr int kOrderBins = 8;

// ---- launchers (defined in the .cu files) ---------------------------------
void launch_target_prep(const float4 *tgt, int m, float *soa, int ngroups, cudaStream_t s);
void launch_nn_partial(const RegDesc *descs, IcpState *states, int batch, int max_n, int qpt, int splits, int pass,
                       int filter, cudaStream_t s);
void launch_nn_finalize(const RegDesc *descs, IcpState *states, const IcpParamsDev *prm, int batch, int max_n,
                        int splits, int pass, int filter, cudaStream_t s, const float4 *q_cur = nullptr,
                        const float4 *q_nb = nullptr);
// n_dev (nullable): the point count lives on the device (sync-free frame path); n is then the capacity
void launch_transform(float4 *pts, int n, const float *R, const float *t, int have_R, int have_t,
                      cudaStream_t s, const int *n_dev = nullptr);
void launch_pack_band(const float4 *pts, int n, float4 *dst, cudaStream_t s);
void launch_assemble_bands(const float4 *bands, int world, int band_capacity, float4 *out, int out_capacity, int *total,
                           cudaStream_t s);
void launch_pending_translate(const RegDesc *descs, int batch, int max_n, cudaStream_t s);
void launch_copy_back(const RegDesc *descs, int batch, int max_n, cudaStream_t s);
void launch_keypoint_epilogue(const RegDesc *descs, int batch, int max_carry, cudaStream_t s);
void launch_center(const float4 *pts, int n, double *chunk_sums, double *out3, unsigned int *counter,
                   cudaStream_t s);
void launch_fp32_peak(float *out, int blocks, int threads, int iters, cudaStream_t s);
void launch_grid_bbox(const float4 *tgt, int m, unsigned int *bbox, cudaStream_t s);
// batched forms for `batch` registrations (blockIdx.y): bounding boxes of the targets (bbox: batch x 6 ordered uints,
// initialised by the caller), and the cell build of all registrations at once -- counts / starts of all registrations
// form ONE array (desc.gstart points into it) scanned as a whole, so that start values are positions in the shared
// sorted array desc.gsorted
struct TgtRef {
    const float4 *pts;
    int m;
};
void launch_grid_bbox_batch(const TgtRef *refs, int batch, int max_m, unsigned int *bbox, cudaStream_t s);
void launch_grid_build_batch(const RegDesc *descs, int batch, int max_m, int max_ncells, int *counts, int *cursor,
                             long long total_entries, int *block_sums, const float4 *sorted, float4 *boxes, bool any_children,
                             cudaStream_t s);
void launch_grid_build(const float4 *tgt, int m, const GridMeta &g, int *counts, int *cursor, int *block_sums,
                       float4 *sorted, float4 *boxes, float4 *coarse_boxes, cudaStream_t s);
// coop_r > 0: warp-cooperative search for balls up to coop_r metres (grid.cu); 0: the per-thread shell walk
void launch_nn_grid(const RegDesc *descs, const IcpState *states, const GridMeta *gmetas, int batch, int max_n, int pass,
                    int sm_count, cudaStream_t s, float coop_r, const RegDesc *h_desc0 = nullptr, const GridMeta *h_grid0 = nullptr);
size_t spatial_sort_work_ints(int max_n, int batch); // ints of scratch launch_spatial_sort needs
void launch_spatial_sort(const RegDesc *descs, int batch, int max_n, int *work, cudaStream_t s); // 15 launches

struct BackprojectArgs {
    const uint16_t *depth;
    const uint8_t *bgr;     // nullable
    int w, h;
    icpb_intrinsics K;
    int rule;
    uint32_t rule_arg, seed;
    const uint8_t *keep_stream; // nullable
    int keep_stream_len;
    float4 *out;
    int capacity;
    int *out_count;          // device: number of points written
    unsigned long long *tile_state; // device scratch, zeroed by the launcher
    int n_tiles;
    // batched launch (blockIdx.y = frame): strides between consecutive frames, in elements
    int v_offset;            // image row of the first depth row handed in (a row band of a larger frame)
    int frames;
    long long depth_stride, bgr_stride, out_stride, state_stride; // pixels, bytes, points, u64 words
};
int backproject_tiles(int w, int h);
void launch_backproject(const BackprojectArgs &a, cudaStream_t s, bool force_two_pass = false);
void launch_normals(const uint16_t *depth, int w, int h, float *normals, cudaStream_t s, int frames = 1);
void launch_depth_filter(const uint16_t *in, uint16_t *tmp_a, uint16_t *tmp_b, uint16_t *out, int w, int h,
                         int min_d, int max_d, cudaStream_t s);

// Coarse occupancy of the certainty grid: one bit per brick of kBrick^3 voxels (slab-relative z), set whenever a voxel
// of the brick is written non-zero and only cleared by icpb_map_clear / rebuilt by icpb_map_upload.  A clear bit proves
// that every voxel of the brick is zero, so the ray walk (whose decrement of a zero voxel is a no-op, map.cpp:423) may
// cross the brick without reading it.  600x600x500 at 1 cm: 75x75x63 bits = 44 KB, L1-resident.
// A second, coarser level (kBrick2^3 voxels, 722 bytes for the same grid) lets the walk cross the empty interior of a room
// in a dozen jumps.
constexpr int kBrickLog = 3;
constexpr int kBrick = 1 << kBrickLog;
constexpr int kBrick2Log = 5;
constexpr int kBrick2 = 1 << kBrick2Log;

struct MapDev {
    uint8_t *grid;    // slab storage: [(x*dimY + y)*zs + (z - z_lo)], zs = z_hi - z_lo; the allocation is padded to 4 bytes
    int dims[3];
    int z_lo, z_hi, zs;
    float cell;
    uint32_t *bricks; // occupancy bits, index (bx*nby + by)*nbz + bz
    int nby, nbz;     // bricks along y and along the slab's z
    uint32_t *bricks2; // the coarse level, same layout with nby2 / nbz2 (stored behind the fine level)
    int nby2, nbz2;
};
void launch_map_rebuild_bricks(const MapDev &m, long long brick_words, cudaStream_t s);
// Where a map kernel finds its points.  Flat: `n` points at pts (n_dev, nullable: the live count in device memory, n
// then being the capacity).  Banded (bands > 0): `bands` bands `band_stride` rows apart, each one 16-byte header row
// (its point count in the first word; negative = none) followed by up to band_cap points -- the receive buffer of an
// all-gather as it is, no assembly pass.
constexpr int kMaxBands = 16;
struct PointSrc {
    const float4 *pts;
    const int *n_dev;
    int n;
    int bands, band_cap;
    long long band_stride;
};
inline PointSrc flat_src(const float4 *pts, int n, const int *n_dev = nullptr) { return PointSrc{pts, n_dev, n, 0, 0, 0}; }
inline PointSrc band_src(const float4 *first_header, int bands, int band_cap, long long band_stride)
{
    return PointSrc{first_header, nullptr, bands * band_cap, bands, band_cap, band_stride};
}
void launch_map_endpoints(const MapDev &m, const PointSrc &src, int rule, int delta, int max_conf, cudaStream_t s);
void launch_map_tracked(const MapDev &m, int *table, const float4 *pts, int n, int variant, int delta, int max_conf,
                        float4 *dst, int dst_n, int dst_capacity, int *d_appended, int table_base, cudaStream_t s);
void launch_map_rays(const MapDev &m, const PointSrc &src, const float origin[3], int delta_dec,
                     unsigned long long *visited, unsigned int *next_ray, int sm_count, cudaStream_t s,
                     unsigned int *layer_work = nullptr);
extern unsigned int *const kCounterIsZero; // pass as layer_work: next_ray is already zero, skip the memset

// ---- programmatic dependent launch (the kernels of the pass loop) -------------------------------------------
// A pass is a chain of short kernels (search / scan -> sums -> solve -> next pass) whose launch latencies and drain
// times add up to a tenth of a small registration.  Every kernel of the chain is launched with the programmatic
// stream-serialisation attribute and starts with pdl_enter(): wait until the grid before it has completed and its
// writes are visible (griddepcontrol.wait -- before ANY read of data the chain produces, the done flag included), then
// let the grid after it be scheduled (griddepcontrol.launch_dependents: it fires once the LAST CTA of this grid is
// resident, so the successor's CTAs fill the SMs as this grid's tail drains and sit in their own wait).  Because the
// trigger comes after the wait, a grid never becomes resident before its grandparent has finished.
// Decided per call by run_registrations (api.cu): all launches for small jobs, the ones behind a small grid for large
// jobs; ICPB_PDL=0 / 1 forces none / all.  Launched
// without the attribute the same kernels are fully serialised and the two instructions are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif
int pdl_level();          // 0: no launch takes part, 1: those whose predecessor grid is small, 2: all
void pdl_set(int level);
// after_big: the grid launched before this one fills the GPU several waves deep (level 1 leaves such launches serialised)
template <bool after_big = false, typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (pdl_level() >= (after_big ? 2 : 1)) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- helpers of api.cu used by comm.cu ------------------------------------------------------------------------
int api_fail(icpb_ctx *ctx, int status, const char *what, cudaError_t ce = cudaSuccess);
int api_ws_get(icpb_ctx *ctx, int id, size_t bytes, void **out, bool zero_new = false);
int api_span_begin(icpb_ctx *ctx, int kernel); // profiling spans on the context stream (icpb_ctx_profile_read)
void api_span_end(icpb_ctx *ctx, int id);
int api_lift_band(icpb_ctx *ctx, cudaStream_t stream, void *tile_state, const void *d_depth, int w, int h, int row0,
                  int row1, const icpb_intrinsics *K, const float *R, const float *t, void *d_band, int band_capacity);

} // namespace icpb

// ---- handle definitions -----------------------------------------------------
struct icpb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evt0 = nullptr, evt1 = nullptr;
    std::string err;
    long long launches = 0;
    int sm_count = 148;
    // grow-only device workspace
    struct Buf {
        void *p = nullptr;
        size_t bytes = 0;
    };
    std::vector<Buf> ws; // indexed by the WS_* ids in api.cu
    void *pinned = nullptr;
    size_t pinned_bytes = 0;
    // the registration loop's own staging block (descriptors, states, results read-back): not shared with the upload
    // paths, so it stays valid while a registration started by icpb_icp_register_async is in flight
    void *pinned_reg = nullptr;
    size_t pinned_reg_bytes = 0;
    icpb_pending *inflight = nullptr; // at most one per context; drained by the next registration / wait / destroy
    cudaEvent_t ev_reg_done = nullptr;
    cudaStream_t stream2 = nullptr; // side stream of the registration set-up (cell build beside the query sort)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;
    // profiling mode: CUDA-event spans around individual kernels other than nn_partial (icpb_ctx_profile_read)
    struct Span {
        cudaEvent_t a, b;
        int kernel;
    };
    std::vector<Span> spans;          // recorded, not yet read
    std::vector<cudaEvent_t> ev_pool; // recycled events
};

// A registration (or batch) whose kernels are enqueued and whose results have not been read yet.
struct icpb_pending {
    icpb_ctx *ctx = nullptr;
    int count = 0;
    std::function<int(icpb_icp_result *)> finish; // waits for the stream work, fills `count` results
    std::vector<icpb_icp_result> results;          // filled when the context drains it before the owner waits
    bool finished = false;
    int status = 0;
};

struct icpb_cloud {
    icpb_ctx *ctx = nullptr;
    float4 *d_pts = nullptr;
    int capacity = 0;
    int n = 0;
};

struct icpb_map {
    icpb_ctx *ctx = nullptr;
    icpb::MapDev dev{};
    long long bytes = 0;
    int *table = nullptr; // lazily allocated lookup table: index into the map cloud, -1 = empty (map.hpp:24)
    long long brick_words = 0; // 32-bit words of dev.bricks
};

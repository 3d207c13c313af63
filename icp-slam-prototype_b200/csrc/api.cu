// C-ABI of libicpb200.so (include/icpb200.h): contexts, handles, host-side
// orchestration of the kernels in nn.cu / cloud.cu / map.cu.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "icpb_internal.h"

using namespace icpb;

namespace {

thread_local std::string g_create_error;

enum WsId {
    WS_DESCS = 0, WS_STATES, WS_PARAMS, WS_TGT_SOA, WS_PM1, WS_PM2, WS_PG, WS_IDX, WS_DIST, WS_CHUNKS, WS_ALT,
    WS_IDX_TRACE, WS_DIST_TRACE, WS_MISC, WS_DEPTH, WS_BGR, WS_KEEP, WS_TILESTATE, WS_IMG_A, WS_IMG_B, WS_NORMALS,
    WS_RT, WS_GRID_META, WS_GRID_COUNTS, WS_GRID_CURSOR, WS_GRID_SUMS, WS_GRID_SORTED, WS_GRID_BBOX, WS_GRID_HEAVY, WS_BATCHSTATE, WS_TRACK, WS_PA, WS_PM3, WS_PG2, WS_MLOG, WS_REJ_FLAG, WS_REJ_PTS, WS_FRAME, WS_PERM, WS_SORT_COUNTS, WS_SORT_SUMS, WS_GRID_NB, WS_GRID_SEED, WS_GRID_BOX, WS_GRID_BOXC, WS_GRID_ORD, WS_GRID_ORDCNT, WS_GRID_Q, WS_COUNT
};

int fail(icpb_ctx *ctx, int status, const char *what, cudaError_t ce = cudaSuccess)
{
    std::string msg = what;
    if (ce != cudaSuccess) {
        msg += ": ";
        msg += cudaGetErrorString(ce);
    }
    if (ctx) ctx->err = msg;
    else g_create_error = msg;
    return status;
}

#define CU(ctx, call)                                                         \
    do {                                                                      \
        cudaError_t ce__ = (call);                                            \
        if (ce__ != cudaSuccess) return fail((ctx), ICPB_ERR_CUDA, #call, ce__); \
    } while (0)

// grow-only workspace buffer
int ws_get(icpb_ctx *ctx, int id, size_t bytes, void **out, bool zero_new = false)
{
    icpb_ctx::Buf &b = ctx->ws[id];
    if (b.bytes < bytes) {
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (b.p) CU(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.bytes = 0;
        size_t want = bytes + bytes / 4 + 256;
        CU(ctx, cudaMalloc(&b.p, want));
        b.bytes = want;
        if (zero_new) CU(ctx, cudaMemsetAsync(b.p, 0, want, ctx->stream));
    }
    *out = b.p;
    return ICPB_OK;
}

int pinned_get(icpb_ctx *ctx, size_t bytes, void **out)
{
    if (ctx->pinned_bytes < bytes) {
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->pinned) CU(ctx, cudaFreeHost(ctx->pinned));
        ctx->pinned = nullptr;
        ctx->pinned_bytes = 0;
        size_t want = bytes + bytes / 4 + 256;
        CU(ctx, cudaMallocHost(&ctx->pinned, want));
        ctx->pinned_bytes = want;
    }
    *out = ctx->pinned;
    return ICPB_OK;
}

// the registration loop's own pinned block (see icpb_ctx::pinned_reg)
int pinned_reg_get(icpb_ctx *ctx, size_t bytes, void **out)
{
    if (ctx->pinned_reg_bytes < bytes) {
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->pinned_reg) CU(ctx, cudaFreeHost(ctx->pinned_reg));
        ctx->pinned_reg = nullptr;
        ctx->pinned_reg_bytes = 0;
        size_t want = bytes + bytes / 4 + 256;
        CU(ctx, cudaMallocHost(&ctx->pinned_reg, want));
        ctx->pinned_reg_bytes = want;
    }
    *out = ctx->pinned_reg;
    return ICPB_OK;
}

// Completes the registration icpb_icp_register_async left in flight on this context (its results wait in the handle).
void drain_inflight(icpb_ctx *ctx)
{
    icpb_pending *p = ctx->inflight;
    if (!p) return;
    ctx->inflight = nullptr;
    p->results.resize((size_t)p->count);
    p->status = p->finish(p->results.data());
    p->finish = nullptr;
    p->finished = true;
}

// profiling spans (icpb_ctx_profile_read): no-ops unless the context is in profiling mode
int span_begin(icpb_ctx *ctx, int kernel)
{
    if (!ctx->profiling) return -1;
    icpb_ctx::Span sp;
    cudaEvent_t *ev[2] = {&sp.a, &sp.b};
    for (auto e : ev) {
        if (!ctx->ev_pool.empty()) { *e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); }
        else if (cudaEventCreate(e) != cudaSuccess) return -1;
    }
    sp.kernel = kernel;
    cudaEventRecord(sp.a, ctx->stream);
    ctx->spans.push_back(sp);
    return (int)ctx->spans.size() - 1;
}

void span_end(icpb_ctx *ctx, int id)
{
    if (id >= 0) cudaEventRecord(ctx->spans[id].b, ctx->stream);
}

int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    if (!v || !*v) return dflt;
    return atoi(v);
}

// Work decomposition of nn_partial: QPT queries per thread, `splits` target ranges.
void choose_nn_config(const icpb_ctx *ctx, int max_n, int max_m, int batch, int filter, int *qpt, int *splits)
{
    // Centred filter: the per-thread centring of the targets is amortised over the thread's queries, so large
    // scans take 12 queries per thread (168 registers, 3 CTAs per SM); direct filter: 8.
    int q = (filter != kFilterDirect && max_n >= 50000) ? 12 : 8;
    const long long slots = (long long)ctx->sm_count * 4; // ~4 resident 128-thread CTAs per SM
    auto tiles_of = [&](int qq) { return (long long)((max_n + kNnThreads * qq - 1) / (kNnThreads * qq)) * batch; };
    // small problems: smaller query tiles give the grid more CTAs before the targets must be split
    if (tiles_of(8) < slots / 8) q = 4;
    q = env_int("ICPB_QPT", q);
    if (filter != kFilterDirect) { if (q != 2 && q != 4 && q != 8 && q != 12 && q != 16) q = 8; }
    else if (q != 2 && q != 4 && q != 8) q = 8;
    const int tiles = (max_n + kNnThreads * q - 1) / (kNnThreads * q);
    const int ngroups = (max_m + kGroup - 1) / kGroup;
    long long s = (slots + (long long)tiles * batch - 1) / ((long long)tiles * batch);
    s = std::max<long long>(1, std::min<long long>(s, std::max(1, ngroups / 4)));
    // large scans: ~12 CTAs per SM so the tail (SMs running their last CTA alone) stays short,
    // as long as every split still streams >= 64 groups (2048 targets; batches of 10k-point registrations: 3 splits,
    // +11 % on bench.py --workload batch10k)
    const long long fine = ((long long)ctx->sm_count * 12 + (long long)tiles * batch - 1) / ((long long)tiles * batch);
    if (fine > s && ngroups / fine >= 64) s = fine;
    s = env_int("ICPB_SPLITS", (int)s);
    s = std::max<long long>(1, std::min<long long>(s, ngroups));
    *qpt = q;
    *splits = (int)s;
}

struct RegHost {
    icpb_cloud *data;
    const icpb_cloud *target;
    icpb_cloud *carry = nullptr;    // key-point variant: points that follow the motion
    icpb_cloud *nonassoc = nullptr; // key-point variant: receives the accumulated rejects
};

// The device-resident registration loop for `count` independent problems.
int run_registrations(icpb_ctx *ctx, const RegHost *regs, int count, const icpb_icp_params *prm,
                      icpb_icp_result *results, bool keep_transformed, icpb_pending **async_out = nullptr)
{
    if (count <= 0) return fail(ctx, ICPB_ERR_INVALID, "count <= 0");
    drain_inflight(ctx); // the staging block and the workspace below are this call's from here on
    int max_n = 0, max_m = 0;
    size_t tot_n = 0, tot_groups = 0, tot_chunks = 0;
    for (int b = 0; b < count; ++b) {
        if (!regs[b].data || !regs[b].target) return fail(ctx, ICPB_ERR_INVALID, "null cloud");
        if (regs[b].data->ctx != ctx || regs[b].target->ctx != ctx)
            return fail(ctx, ICPB_ERR_INVALID, "cloud belongs to another context");
        if (regs[b].data->n <= 0 || regs[b].target->n <= 0)
            return fail(ctx, ICPB_ERR_EMPTY, "empty cloud (the reference dereferences begin(), icp.cpp:572)");
        max_n = std::max(max_n, regs[b].data->n);
        max_m = std::max(max_m, regs[b].target->n);
        tot_n += ((size_t)regs[b].data->n + 31) / 32 * 32;
        tot_groups += ((size_t)regs[b].target->n + kGroup - 1) / kGroup;
        tot_chunks += ((size_t)regs[b].data->n + kChunk - 1) / kChunk;
    }
    if (prm->max_iterations < 0) return fail(ctx, ICPB_ERR_INVALID, "max_iterations < 0");
    const int passes = prm->max_iterations + 1;
    const bool trace = (prm->idx_trace != nullptr) || (prm->dist_trace != nullptr);
    // in a batch the traces record ONE registration: number ICPB_TRACE_REG (default 0), shape (max_iterations+1) x its n
    const int trace_reg = trace ? std::min(std::max(env_int("ICPB_TRACE_REG", 0), 0), count - 1) : 0;

    int qpt, splits;
    int filter = env_int("ICPB_NN_FILTER", prm->nn_filter);
    if (filter != kFilterDirect && filter != kFilterWarp && filter != kFilterCentred) // ICPB_FILTER_AUTO
        filter = ((count == 1 && max_n >= 50000) || (count > 1 && max_n >= 4096)) ? kFilterWarp : kFilterCentred;
    choose_nn_config(ctx, max_n, max_m, count, filter, &qpt, &splits);

    RegDesc *d_descs;
    IcpState *d_states;
    IcpParamsDev *d_prm;
    float *d_soa, *d_pm1, *d_pm2, *d_pm3, *d_pa, *d_dist;
    int *d_pg, *d_pg2, *d_idx;
    double *d_chunks;
    float4 *d_alt;
    int rc;
    if ((rc = ws_get(ctx, WS_DESCS, sizeof(RegDesc) * count, (void **)&d_descs))) return rc;
    if ((rc = ws_get(ctx, WS_STATES, sizeof(IcpState) * count, (void **)&d_states))) return rc;
    if ((rc = ws_get(ctx, WS_PARAMS, sizeof(IcpParamsDev), (void **)&d_prm))) return rc;
    if ((rc = ws_get(ctx, WS_TGT_SOA, tot_groups * kGroup * 3 * sizeof(float), (void **)&d_soa))) return rc;
    if ((rc = ws_get(ctx, WS_PM1, tot_n * splits * sizeof(float), (void **)&d_pm1))) return rc;
    if ((rc = ws_get(ctx, WS_PM2, tot_n * splits * sizeof(float), (void **)&d_pm2))) return rc;
    if ((rc = ws_get(ctx, WS_PG, tot_n * splits * sizeof(int), (void **)&d_pg))) return rc;
    if ((rc = ws_get(ctx, WS_PA, tot_n * sizeof(float), (void **)&d_pa))) return rc;
    if ((rc = ws_get(ctx, WS_PM3, tot_n * splits * sizeof(float), (void **)&d_pm3))) return rc;
    if ((rc = ws_get(ctx, WS_PG2, tot_n * splits * sizeof(int), (void **)&d_pg2))) return rc;
    if ((rc = ws_get(ctx, WS_IDX, tot_n * sizeof(int), (void **)&d_idx))) return rc;
    if ((rc = ws_get(ctx, WS_DIST, tot_n * sizeof(float), (void **)&d_dist))) return rc;
    if ((rc = ws_get(ctx, WS_CHUNKS, tot_chunks * kTerms * sizeof(double), (void **)&d_chunks))) return rc;
    if ((rc = ws_get(ctx, WS_ALT, tot_n * sizeof(float4), (void **)&d_alt))) return rc;
    // key-point variant (single registration): motion log, per-pass reject records
    const bool kp_mode = count == 1 && (regs[0].carry || regs[0].nonassoc);
    float *d_mlog = nullptr;
    int *d_rej_flag = nullptr;
    float4 *d_rej_pts = nullptr;
    if (kp_mode) {
        if (regs[0].nonassoc && regs[0].nonassoc->capacity < passes * regs[0].data->n)
            return fail(ctx, ICPB_ERR_CAPACITY, "non-association cloud needs capacity (max_iterations+1) * key-points");
        if ((rc = ws_get(ctx, WS_MLOG, sizeof(float) * 12 * (size_t)std::max(passes, 1), (void **)&d_mlog))) return rc;
        if (regs[0].nonassoc) {
            if ((rc = ws_get(ctx, WS_REJ_FLAG, sizeof(int) * (size_t)passes * max_n, (void **)&d_rej_flag))) return rc;
            if ((rc = ws_get(ctx, WS_REJ_PTS, sizeof(float4) * (size_t)passes * max_n, (void **)&d_rej_pts))) return rc;
        }
    }
    int *d_idx_trace = nullptr;
    float *d_dist_trace = nullptr;
    if (prm->idx_trace) {
        if ((rc = ws_get(ctx, WS_IDX_TRACE, (size_t)passes * max_n * sizeof(int), (void **)&d_idx_trace))) return rc;
        CU(ctx, cudaMemsetAsync(d_idx_trace, 0xff, (size_t)passes * max_n * sizeof(int), ctx->stream));
    }
    if (prm->dist_trace) {
        if ((rc = ws_get(ctx, WS_DIST_TRACE, (size_t)passes * max_n * sizeof(float), (void **)&d_dist_trace)))
            return rc;
        CU(ctx, cudaMemsetAsync(d_dist_trace, 0, (size_t)passes * max_n * sizeof(float), ctx->stream));
    }

    int *d_perm = nullptr;
    const bool grid_sorted = env_int("ICPB_GRID_SORT", 1) != 0;
    if (filter == kFilterWarp || grid_sorted) {
        if ((rc = ws_get(ctx, WS_PERM, sizeof(int) * tot_n, (void **)&d_perm))) return rc;
    }

    // ---- ICPB_NN_GRID: bucket the (fixed) targets once per registration.  Batches take the cooperative search only
    //      (the per-thread walk of round 1 stays a single-registration path).
    const bool coop_mode = env_int("ICPB_GRID_COOP_CM", 1000) > 0;
    // a carried cloud (icpb_icp_register_carry) only replays the motions afterwards and goes with either search; the
    // reject lists of the key-point loop are filled by the scan's finalize kernel alone
    const bool rejects_wanted = count == 1 && regs[0].nonassoc;
    bool grid_mode = (prm->nn_mode == ICPB_NN_GRID) && (count == 1 || coop_mode) && !rejects_wanted && (!kp_mode || coop_mode);
    if (prm->nn_mode == ICPB_NN_AUTO && !rejects_wanted && (!kp_mode || coop_mode)) {
        // measured crossover regions: ~45k x 45k for one registration (set-up of ~0.3 ms against a scan of milliseconds);
        // in a batch the set-up is shared and the search wins from a few thousand points per cloud on
        if (count == 1) grid_mode = (double)max_n * (double)max_m >= 2.0e9;
        else grid_mode = coop_mode && max_n >= 4096 && max_m >= 4096;
    }
    std::vector<GridMeta> gms;
    std::vector<size_t> cell_off, n_off, m_off; // per registration: first entry in the counts array, first query, first target
    GridMeta *d_gmeta = nullptr;
    int *d_gcounts = nullptr, *d_gcursor = nullptr, *d_gsums = nullptr;
    float4 *d_gsorted = nullptr;
    int *d_gheavy = nullptr;
    float4 *d_gnb = nullptr, *d_gseed = nullptr, *d_gbox = nullptr, *d_gboxc = nullptr;
    int *d_gord = nullptr, *d_gordcnt = nullptr;
    float4 *d_gq = nullptr;
    // Work order of the cooperative search (grid.cu): heaviest warps first from the second pass on, when the pass is more
    // than one wave of warps deep (full resolution: 3.78 -> 3.50 ms; a single 10k-point cloud is a tenth of a wave and
    // only pays the two extra loads).  ICPB_GRID_ORDER=0 / 1 forces natural order / ordering.
    const int order_knob = env_int("ICPB_GRID_ORDER", 2);
    const bool grid_ordered = order_knob == 1 || (order_knob == 2 && tot_n / 32 > (size_t)ctx->sm_count * 20);
    long long grid_launches = 0;
    size_t total_entries = 0, total_coarse = 0, tot_m = 0;
    int max_ncells = 0;
    bool any_children = false;
    if (grid_mode) {
        // bounding boxes of all targets: one batched kernel, ONE host round trip
        unsigned int *d_bbox;
        TgtRef *d_refs;
        if ((rc = ws_get(ctx, WS_GRID_BBOX, (sizeof(unsigned int) * 6 + sizeof(TgtRef)) * (size_t)count + 64, (void **)&d_bbox))) return rc;
        d_refs = (TgtRef *)(d_bbox + 6 * (size_t)count + 2);
        void *hpb;
        if ((rc = pinned_reg_get(ctx, (sizeof(unsigned int) * 6 + sizeof(TgtRef)) * (size_t)count + 64, &hpb))) return rc;
        unsigned int *hb = (unsigned int *)hpb;
        TgtRef *hrefs = (TgtRef *)(hb + 6 * (size_t)count + 2);
        for (int b = 0; b < count; ++b) {
            for (int k = 0; k < 3; ++k) { hb[6 * b + k] = 0xffffffffu; hb[6 * b + 3 + k] = 0u; }
            hrefs[b].pts = regs[b].target->d_pts;
            hrefs[b].m = regs[b].target->n;
        }
        CU(ctx, cudaMemcpyAsync(d_bbox, hb, (sizeof(unsigned int) * 6 + sizeof(TgtRef)) * (size_t)count + 64, cudaMemcpyHostToDevice, ctx->stream));
        launch_grid_bbox_batch(d_refs, count, max_m, d_bbox, ctx->stream);
        CU(ctx, cudaMemcpyAsync(hb, d_bbox, sizeof(unsigned int) * 6 * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        gms.resize((size_t)count);
        cell_off.resize((size_t)count); n_off.resize((size_t)count); m_off.resize((size_t)count);
        size_t acc_n = 0;
        for (int b = 0; b < count; ++b) {
            GridMeta &gm = gms[(size_t)b];
            memset(&gm, 0, sizeof(gm));
            const int m = regs[b].target->n;
            float lo[3], hi[3];
            for (int k = 0; k < 6; ++k) {
                unsigned int u = hb[6 * b + k];
                u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; // inverse of the ordered-uint map
                float f;
                memcpy(&f, &u, 4);
                (k < 3 ? lo[k] : hi[k - 3]) = f;
            }
            // default cell: ~90 points per occupied cell for the cooperative search (its cost per cell is a fixed few dozen
            // instructions per warp; measured optimum 0.10 - 0.125 m on the full-resolution pair), ~50 for the per-thread
            // walk, at most 0.2 m; a third of the bounding box's surface stands for the area the (surface-sampled) cloud covers
            float h = prm->grid_cell;
            bool clamped = false;
            if (!(h > 0.f)) {
                const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
                const float area = 2.f * (ex * ey + ey * ez + ez * ex) / 3.f;
                const float per_cell = coop_mode ? (float)env_int("ICPB_GRID_CELL_PTS", 90) : 50.f;
                h = sqrtf(per_cell * std::max(area, 1e-6f) / (float)m);
                clamped = h > 0.2f;
                h = std::min(std::max(h, 0.01f), 0.2f);
            }
            // the cooperative search splits every cell into 2 x 2 x 2 children (ICPB_GRID_SUB=1: no children); a sparse
            // cloud (cell edge clamped: a handful of points per cell) has nothing to split
            gm.sub = (coop_mode && !clamped && env_int("ICPB_GRID_SUB", 8) == 8) ? 8 : 1;
            const double cell_cap = (gm.sub == 8 ? 4.0e6 : 48.0e6) / std::max(1, std::min(count, 64)); // tables stay in the GB range
            for (;;) {
                double cells = 1.0;
                for (int k = 0; k < 3; ++k) {
                    gm.dim[k] = std::max(1, (int)floorf((hi[k] - lo[k]) / h) + 1);
                    cells *= gm.dim[k];
                }
                if (cells <= cell_cap) break;
                h *= 1.26f; // coarser cells until the tables fit
            }
            for (int k = 0; k < 3; ++k) gm.mn[k] = lo[k];
            gm.h = h;
            gm.ncells = gm.dim[0] * gm.dim[1] * gm.dim[2];
            gm.max_nn = prm->max_nn_distance;
            gm.max_r = (int)ceilf(prm->max_nn_distance / h) + 2;
            // per-thread shells out to ~ICPB_GRID_LIGHT_CM (default 45 cm): typical ICP residuals resolve there;
            // the few queries still open (no overlap, far from the target) are finished by one warp each
            gm.light_r = std::max(2, (int)ceilf(0.01f * env_int("ICPB_GRID_LIGHT_CM", 45) / h));
            // probe rounds of the cooperative search (grid.cu): ICPB_GRID_PROBE_PCT = ball radius in percent of the cell
            // edge from which a ball is probed first; 0 = never
            {
                const int pp = env_int("ICPB_GRID_PROBE_PCT", 150);
                gm.probe_r = pp > 0 ? 0.01f * (float)pp * h : 3.0e38f;
                const int pp2 = env_int("ICPB_GRID_PROBE2_PCT", 550);
                gm.probe_r2 = pp2 > 0 ? 0.01f * (float)pp2 * h : 3.0e38f;
            }
            cell_off[(size_t)b] = total_entries;
            n_off[(size_t)b] = acc_n;
            m_off[(size_t)b] = tot_m;
            total_entries += (size_t)gm.ncells * gm.sub + 1;
            total_coarse += (size_t)gm.ncells;
            tot_m += (size_t)m;
            acc_n += ((size_t)regs[b].data->n + 31) / 32 * 32;
            max_ncells = std::max(max_ncells, gm.ncells);
            any_children |= gm.sub == 8;
        }
        if (total_entries >= (size_t)1 << 31) return fail(ctx, ICPB_ERR_CAPACITY, "cell tables of the batch exceed 2^31 entries");
        const size_t cbytes = sizeof(int) * total_entries;
        if ((rc = ws_get(ctx, WS_GRID_META, sizeof(GridMeta) * (size_t)count, (void **)&d_gmeta))) return rc;
        if ((rc = ws_get(ctx, WS_GRID_COUNTS, cbytes, (void **)&d_gcounts))) return rc;
        if ((rc = ws_get(ctx, WS_GRID_CURSOR, cbytes, (void **)&d_gcursor))) return rc;
        if ((rc = ws_get(ctx, WS_GRID_SUMS, sizeof(int) * (total_entries / 4096 + 16), (void **)&d_gsums))) return rc;
        if ((rc = ws_get(ctx, WS_GRID_SORTED, sizeof(float4) * tot_m, (void **)&d_gsorted))) return rc;
        if ((rc = ws_get(ctx, WS_GRID_HEAVY, sizeof(int) * (tot_n + (size_t)count * (passes + 8)), (void **)&d_gheavy))) return rc;
        if ((rc = ws_get(ctx, WS_GRID_NB, sizeof(float4) * tot_n, (void **)&d_gnb))) return rc;
        if ((rc = ws_get(ctx, WS_GRID_BOX, sizeof(float4) * 2 * total_entries, (void **)&d_gbox))) return rc;
        if (any_children && (rc = ws_get(ctx, WS_GRID_BOXC, sizeof(float4) * 2 * total_coarse, (void **)&d_gboxc))) return rc;
        if ((rc = ws_get(ctx, WS_GRID_SEED, sizeof(float4) * (tot_n + 32), (void **)&d_gseed))) return rc;
        if (coop_mode && (rc = ws_get(ctx, WS_GRID_Q, sizeof(float4) * (tot_n + 32), (void **)&d_gq))) return rc;
        if (grid_ordered) {
            if ((rc = ws_get(ctx, WS_GRID_ORD, sizeof(int) * (2 * kOrderBins + 1) * (tot_n / 32), (void **)&d_gord))) return rc;
            if ((rc = ws_get(ctx, WS_GRID_ORDCNT, sizeof(int) * (size_t)count * passes * kOrderBins, (void **)&d_gordcnt))) return rc;
        }
        grid_launches = 8 + (any_children ? 1 : 0);
    }

    // host staging: descs | states | params in one pinned block
    const size_t hb = sizeof(RegDesc) * count + sizeof(IcpState) * count + sizeof(IcpParamsDev) + 64 +
                      sizeof(GridMeta) * (size_t)count + 64;
    void *hp;
    if ((rc = pinned_reg_get(ctx, hb, &hp))) return rc;
    RegDesc *h_descs = (RegDesc *)hp;
    IcpState *h_states = (IcpState *)(h_descs + count);
    IcpParamsDev *h_prm = (IcpParamsDev *)(h_states + count);
    void *hp_metas = (void *)(((uintptr_t)(h_prm + 1) + 63) & ~(uintptr_t)63);
    unsigned long long *h_pairs = (unsigned long long *)(((uintptr_t)((GridMeta *)hp_metas + count) + 15) & ~(uintptr_t)15);
    *h_pairs = 0;

    size_t off_n = 0, off_g = 0, off_c = 0;
    for (int b = 0; b < count; ++b) {
        const int n = regs[b].data->n, m = regs[b].target->n;
        const int n_stride = (n + 31) / 32 * 32;
        const int ngroups = (m + kGroup - 1) / kGroup;
        RegDesc &d = h_descs[b];
        d.D[0] = regs[b].data->d_pts;
        d.D[1] = d_alt + off_n;
        d.tgt = regs[b].target->d_pts;
        d.tgt_soa = d_soa + off_g * kGroup * 3;
        d.n = n; d.m = m; d.ngroups = ngroups; d.n_stride = n_stride;
        d.pm1 = d_pm1 + off_n * splits;
        d.pm2 = d_pm2 + off_n * splits;
        d.pg = d_pg + off_n * splits;
        d.pa = d_pa + off_n;
        // only handed to the kernels when the sort below fills it
        d.perm = (d_perm && ((filter == kFilterWarp && !grid_mode) || (grid_mode && grid_sorted))) ? d_perm + off_n : nullptr;
        d.pm3 = d_pm3 + off_n * splits;
        d.pg2 = d_pg2 + off_n * splits;
        d.idx = d_idx + off_n;
        d.dist = d_dist + off_n;
        d.chunk_sums = d_chunks + off_c * kTerms;
        d.st = d_states + b;
        d.idx_trace = (b == trace_reg) ? d_idx_trace : nullptr;
        d.dist_trace = (b == trace_reg) ? d_dist_trace : nullptr;
        if (grid_mode) {
            size_t coarse_before = 0;
            for (int bb = 0; bb < b; ++bb) coarse_before += (size_t)gms[(size_t)bb].ncells;
            d.grid = d_gmeta + b;
            d.gsorted = d_gsorted; // one sorted array for the whole batch: the start values are positions in it
            d.gstart = d_gcounts + cell_off[(size_t)b];
            d.gcursor = d_gcursor + cell_off[(size_t)b];
            d.gbox = d_gbox + 2 * cell_off[(size_t)b];
            d.gboxc = d_gboxc ? d_gboxc + 2 * coarse_before : nullptr;
            d.gnb = d_gnb + off_n;
            d.gseed = d_gseed + off_n;
            d.gq = d_gq ? d_gq + off_n : nullptr;
            // pair counter (profiling mode): the eight bytes behind the scan's block sums, 8-byte aligned
            d.gpairs = ctx->profiling ? (unsigned long long *)(d_gsums + ((total_entries / 4096 + 8 + 1) & ~(size_t)1)) : nullptr;
            d.gheavy_count = d_gheavy + (size_t)b * (passes + 8);
            d.gheavy = d_gheavy + (size_t)count * (passes + 8) + off_n;
            d.gord = d_gord ? d_gord + 2 * kOrderBins * (off_n / 32) : nullptr;
            d.gord_flat = d_gord ? d_gord + 2 * kOrderBins * (tot_n / 32) + off_n / 32 : nullptr;
            d.gord_count = d_gordcnt ? d_gordcnt + (size_t)b * passes * kOrderBins : nullptr;
        } else {
            d.grid = nullptr; d.gsorted = nullptr; d.gstart = nullptr; d.gcursor = nullptr; d.gbox = nullptr; d.gboxc = nullptr;
            d.gnb = nullptr; d.gseed = nullptr; d.gq = nullptr; d.gheavy = nullptr; d.gheavy_count = nullptr; d.gpairs = nullptr;
            d.gord = nullptr; d.gord_count = nullptr; d.gord_flat = nullptr;
        }
        d.carry = (kp_mode && regs[b].carry) ? regs[b].carry->d_pts : nullptr;
        d.n_carry = (kp_mode && regs[b].carry) ? regs[b].carry->n : 0;
        d.mlog = d_mlog;
        d.rej_flag = d_rej_flag;
        d.rej_pts = d_rej_pts;
        d.nonassoc = (kp_mode && regs[b].nonassoc) ? regs[b].nonassoc->d_pts : nullptr;
        d.nonassoc_capacity = (kp_mode && regs[b].nonassoc) ? regs[b].nonassoc->capacity : 0;
        IcpState &s = h_states[b];
        memset(&s, 0, sizeof(s));
        const float I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        for (int k = 0; k < 9; ++k) { s.rigid[k] = I3[k]; s.camR[k] = I3[k]; s.PR[k] = I3[k]; s.Rf[k] = I3[k]; }
        off_n += n_stride;
        off_g += ngroups;
        off_c += (n + kChunk - 1) / kChunk;
    }
    h_prm->max_iterations = prm->max_iterations;
    h_prm->threshold = prm->threshold;
    h_prm->max_nn_distance = prm->max_nn_distance;
    h_prm->solve_mode = prm->solve_mode;
    for (int k = 0; k < 3; ++k) h_prm->last_translation[k] = prm->last_translation[k];

    cudaStream_t st = ctx->stream;
    CU(ctx, cudaMemcpyAsync(d_descs, h_descs, sizeof(RegDesc) * count, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(d_states, h_states, sizeof(IcpState) * count, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(d_prm, h_prm, sizeof(IcpParamsDev), cudaMemcpyHostToDevice, st));

    // programmatic dependent launch for the kernels of the pass loop (icpb_internal.h).  Measured: -7 % on a 10k-point
    // registration (brute force), -9 % with the cell-grid search; +3 % on the full-resolution pair when every kernel
    // takes part, because the successor's CTAs take residency from a search kernel that runs three waves deep.  So:
    // level 2 (every launch) for small jobs, level 1 for large ones -- only the launches whose predecessor is small
    // (sums after the fall-back, solve after the sums, the next search after the one-CTA solve).
    // ICPB_PDL=0 / 1 forces none / all.
    {
        const int knob = env_int("ICPB_PDL", 2);
        pdl_set(knob == 0 ? 0 : (knob == 1 || tot_n <= 65536) ? 2 : 1);
    }
    long long launches = 0;
    CU(ctx, cudaEventRecord(ctx->ev0, st));
    const bool sorted_queries = (filter == kFilterWarp && !grid_mode) || (grid_mode && grid_sorted);
    int *d_swork = nullptr;
    if (sorted_queries &&
        (rc = ws_get(ctx, WS_SORT_COUNTS, sizeof(int) * spatial_sort_work_ints(max_n, count), (void **)&d_swork))) return rc;
    if (grid_mode) {
        // The cells of the targets and the Morton order of the queries have nothing to do with each other and are two
        // dozen short launches each: the cell build runs on the context's side stream while the sort runs on the main
        // one (fork / join by events; ICPB_GRID_FORK=0: one after the other).
        const bool fork = sorted_queries && env_int("ICPB_GRID_FORK", 1) != 0;
        cudaStream_t sb = fork ? ctx->stream2 : st;
        if (fork) {
            CU(ctx, cudaEventRecord(ctx->ev_fork, st));
            CU(ctx, cudaStreamWaitEvent(sb, ctx->ev_fork, 0));
        } else if (sorted_queries) {
            launch_spatial_sort(d_descs, count, max_n, d_swork, st);
        }
        // the metas ride in pinned memory behind the descriptors' block (the call returns only after the stream is done)
        GridMeta *h_gm = (GridMeta *)hp_metas;
        for (int b = 0; b < count; ++b) h_gm[b] = gms[(size_t)b];
        CU(ctx, cudaMemcpyAsync(d_gmeta, h_gm, sizeof(GridMeta) * (size_t)count, cudaMemcpyHostToDevice, sb));
        CU(ctx, cudaMemsetAsync(d_gcounts, 0, sizeof(int) * total_entries, sb));
        CU(ctx, cudaMemsetAsync(d_gheavy, 0, sizeof(int) * (size_t)count * ((size_t)passes + 8), sb));
        if (d_gordcnt) CU(ctx, cudaMemsetAsync(d_gordcnt, 0, sizeof(int) * (size_t)count * passes * kOrderBins, sb));
        launch_grid_build_batch(d_descs, count, max_m, max_ncells, d_gcounts, d_gcursor, (long long)total_entries, d_gsums,
                                d_gsorted, d_gbox, any_children, sb);
        if (ctx->profiling) // after the scans are done with the block sums
            CU(ctx, cudaMemsetAsync(d_gsums + ((total_entries / 4096 + 8 + 1) & ~(size_t)1), 0, sizeof(unsigned long long), sb));
        if (fork) {
            CU(ctx, cudaEventRecord(ctx->ev_join, sb));
            launch_spatial_sort(d_descs, count, max_n, d_swork, st);
            CU(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
        }
        if (sorted_queries) launches += 15;
        launches += grid_launches;
    } else if (sorted_queries) {
        launch_spatial_sort(d_descs, count, max_n, d_swork, st); // Morton order of every data cloud, once per registration
        launches += 15;
    }
    if (!grid_mode) {
        for (int b = 0; b < count; ++b) {
            launch_target_prep(h_descs[b].tgt, h_descs[b].m, const_cast<float *>(h_descs[b].tgt_soa), h_descs[b].ngroups, st);
            ++launches;
        }
    }
    // ICPB_GRID_COOP_CM: widest ball (cm) the warp-cooperative search handles itself (0 = the per-thread shell walk)
    const float coop_r = 0.01f * (float)env_int("ICPB_GRID_COOP_CM", 1000);
    const bool prof = ctx->profiling;
    if (prof) {
        while ((int)ctx->prof_events.size() < 2 * passes) {
            cudaEvent_t e;
            CU(ctx, cudaEventCreate(&e));
            ctx->prof_events.push_back(e);
        }
    }
    for (int pass = 0; pass < passes; ++pass) {
        if (prof) CU(ctx, cudaEventRecord(ctx->prof_events[2 * pass], st));
        if (grid_mode) {
            const int sp = span_begin(ctx, ICPB_PROF_NN_GRID);
            launch_nn_grid(d_descs, d_states, d_gmeta, count, max_n, pass, ctx->sm_count, st, coop_r, &h_descs[0], &gms[0]);
            span_end(ctx, sp);
        } else launch_nn_partial(d_descs, d_states, count, max_n, qpt, splits, pass, filter, st);
        if (prof) CU(ctx, cudaEventRecord(ctx->prof_events[2 * pass + 1], st));
        const int spf = span_begin(ctx, ICPB_PROF_NN_FINALIZE);
        launch_nn_finalize(d_descs, d_states, d_prm, count, max_n, grid_mode ? (coop_r > 0.f ? -1 : 0) : splits, pass, filter, st,
                           grid_mode ? h_descs[0].D[(pass + 1) & 1] : nullptr, d_gnb);
        span_end(ctx, spf);
        launches += grid_mode ? (coop_r > 0.f ? 4 : 3) : 2;
    }
    launch_pending_translate(d_descs, count, max_n, st);
    ++launches;
    if (kp_mode) {
        launch_keypoint_epilogue(d_descs, count, h_descs[0].n_carry, st);
        ++launches;
    }
    if (keep_transformed) {
        // the cloud the last pass associated goes back into the caller's buffer when it sits in the alternate one
        // (decided on the device from the loop state: no host round trip before the copy)
        launch_copy_back(d_descs, count, max_n, st);
        ++launches;
    }
    CU(ctx, cudaEventRecord(ctx->ev1, st));
    CU(ctx, cudaGetLastError());

    CU(ctx, cudaMemcpyAsync(h_states, d_states, sizeof(IcpState) * count, cudaMemcpyDeviceToHost, st));
    if (grid_mode && ctx->profiling)
        CU(ctx, cudaMemcpyAsync(h_pairs, d_gsums + ((total_entries / 4096 + 8 + 1) & ~(size_t)1), sizeof(unsigned long long),
                                cudaMemcpyDeviceToHost, st));
    if (trace) {
        const int n = regs[trace_reg].data->n;
        if (prm->idx_trace)
            CU(ctx, cudaMemcpyAsync(prm->idx_trace, d_idx_trace, (size_t)passes * n * sizeof(int),
                                    cudaMemcpyDeviceToHost, st));
        if (prm->dist_trace)
            CU(ctx, cudaMemcpyAsync(prm->dist_trace, d_dist_trace, (size_t)passes * n * sizeof(float),
                                    cudaMemcpyDeviceToHost, st));
    }
    CU(ctx, cudaEventRecord(ctx->ev_reg_done, st));
    ctx->launches += launches;

    // ---- everything that needs the results on the host: run now (blocking call) or by the wait (async call)
    std::vector<RegHost> regs_copy(regs, regs + count);
    std::vector<float> cells((size_t)count, 0.f);
    if (grid_mode)
        for (int b = 0; b < count; ++b) cells[(size_t)b] = gms[(size_t)b].h;
    auto finish = [ctx, count, regs_copy, cells, kp_mode, h_states, h_pairs, prof, qpt, splits, grid_mode, filter,
                   launches](icpb_icp_result *results) -> int {
        CU(ctx, cudaEventSynchronize(ctx->ev_reg_done));
        float ms = 0.f;
        CU(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        float nn_ms = 0.f;
        int nn_launches = 0;
        if (prof) {
            // passes after convergence exit at once; count only the ones that did the scan
            int executed = 0;
            for (int b = 0; b < count; ++b) executed = std::max(executed, h_states[b].passes);
            for (int pass = 0; pass < executed; ++pass) {
                float t = 0.f;
                CU(ctx, cudaEventElapsedTime(&t, ctx->prof_events[2 * pass], ctx->prof_events[2 * pass + 1]));
                nn_ms += t;
                ++nn_launches;
            }
        }
        for (int b = 0; b < count; ++b) {
            const IcpState &s = h_states[b];
            if (kp_mode && regs_copy[(size_t)b].nonassoc)
                regs_copy[(size_t)b].nonassoc->n = std::min(s.n_nonassoc, regs_copy[(size_t)b].nonassoc->capacity);
            if (!results) continue;
            icpb_icp_result &r = results[b];
            memset(&r, 0, sizeof(r));
            r.iterations = s.iterations;
            r.nn_passes = s.passes;
            r.n_assoc = s.n_assoc;
            r.mse = s.mse;
            for (int rr = 0; rr < 3; ++rr) {
                for (int c = 0; c < 3; ++c) r.rigid[4 * rr + c] = s.rigid[3 * rr + c];
                r.rigid[4 * rr + 3] = s.offset[rr]; // icp.cpp:266-268
            }
            r.rigid[15] = 1.f; // row 3 is never written by the reference (icp.cpp:29)
            memcpy(r.cam_rotation, s.camR, sizeof(s.camR));
            memcpy(r.cam_position, s.camP, sizeof(s.camP));
            memcpy(r.offset, s.offset, sizeof(s.offset));
            memcpy(r.pose_R, s.PR, sizeof(s.PR));
            memcpy(r.pose_t, s.Pt, sizeof(s.Pt));
            r.small_assoc_exit = s.small_exit;
            r.exact_rescans = s.rescans;
            r.gpu_ms = ms;
            r.kernel_launches = (int)launches;
            r.nn_partial_ms = nn_ms;
            r.nn_partial_launches = nn_launches;
            r.nn_qpt = qpt;
            r.nn_splits = splits;
            r.nn_mode_used = grid_mode ? ICPB_NN_GRID : ICPB_NN_BRUTE;
            r.grid_cell_used = cells[(size_t)b];
            r.nn_filter_used = filter;
            r.n_nonassoc = s.n_nonassoc;
            r.grid_pairs = grid_mode ? (long long)*h_pairs : 0; // the whole call's (a batch shares one counter)
            r.nn_grid_ms = grid_mode ? nn_ms : 0.f;
        }
        return ICPB_OK;
    };
    if (!async_out) return finish(results);
    icpb_pending *p = new icpb_pending;
    p->ctx = ctx;
    p->count = count;
    p->finish = finish;
    ctx->inflight = p;
    *async_out = p;
    return ICPB_OK;
}

} // namespace

namespace icpb {
// per calling thread: run_registrations decides per call (the launches happen on the caller's thread)
static thread_local int tl_pdl = 0;
int pdl_level() { return tl_pdl; }
void pdl_set(int level) { tl_pdl = level; }

int api_fail(icpb_ctx *ctx, int status, const char *what, cudaError_t ce) { return fail(ctx, status, what, ce); }

int api_ws_get(icpb_ctx *ctx, int id, size_t bytes, void **out, bool zero_new) { return ws_get(ctx, id, bytes, out, zero_new); }

int api_span_begin(icpb_ctx *ctx, int kernel) { return span_begin(ctx, kernel); }

void api_span_end(icpb_ctx *ctx, int id) { span_end(ctx, id); }

// Rows [row0, row1) of a device-resident depth frame -> world-space band (header row + points) on `stream`;
// tile_state: 8 * (3 * tiles + 2) bytes of scratch, zeroed once by its owner.
int api_lift_band(icpb_ctx *ctx, cudaStream_t stream, void *tile_state, const void *d_depth, int w, int h, int row0,
                  int row1, const icpb_intrinsics *K, const float *R, const float *t, void *d_band, int band_capacity)
{
    if (row0 < 0 || row1 <= row0 || row1 > h) return fail(ctx, ICPB_ERR_INVALID, "row band outside the image");
    if ((long long)(row1 - row0) * w > band_capacity) return fail(ctx, ICPB_ERR_CAPACITY, "band smaller than its pixel count");
    const uint16_t *band_depth = (const uint16_t *)d_depth + (size_t)row0 * w;
    if (((uintptr_t)band_depth & 15) != 0) return fail(ctx, ICPB_ERR_INVALID, "band depth rows must start 16-byte aligned");
    BackprojectArgs a;
    a.depth = band_depth;
    a.bgr = nullptr;
    a.w = w; a.h = row1 - row0; a.K = *K;
    a.rule = ICPB_SUB_NONE; a.rule_arg = 1; a.seed = 0;
    a.keep_stream = nullptr; a.keep_stream_len = 0;
    a.out = (float4 *)d_band + 1;           // row 0 is the header: point count in its first word
    a.capacity = band_capacity;
    a.n_tiles = backproject_tiles(w, a.h);
    a.frames = 1;
    a.v_offset = row0;
    a.depth_stride = a.bgr_stride = a.out_stride = a.state_stride = 0;
    a.out_count = (int *)d_band;
    a.tile_state = (unsigned long long *)tile_state + 2;
    CU(ctx, cudaMemsetAsync(d_band, 0, sizeof(float4), stream));
    launch_backproject(a, stream);
    if (R || t) launch_transform(a.out, band_capacity, R, t, R != nullptr, t != nullptr, stream, (const int *)d_band);
    ctx->launches += 1 + ((R || t) ? 1 : 0);
    CU(ctx, cudaGetLastError());
    return ICPB_OK;
}

} // namespace icpb

extern "C" {

int icpb_version(void) { return ICPB_VERSION; }

const char *icpb_status_string(int status)
{
    switch (status) {
    case ICPB_OK: return "ok";
    case ICPB_ERR_INVALID: return "invalid argument";
    case ICPB_ERR_EMPTY: return "empty cloud";
    case ICPB_ERR_CUDA: return "CUDA error";
    case ICPB_ERR_CAPACITY: return "capacity exceeded";
    case ICPB_ERR_NCCL: return "NCCL error";
    default: return "unknown status";
    }
}

const char *icpb_last_error(const icpb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int icpb_device_count(int *count)
{
    if (!count) return ICPB_ERR_INVALID;
    int c = 0;
    cudaError_t ce = cudaGetDeviceCount(&c);
    if (ce != cudaSuccess) {
        *count = 0;
        return fail(nullptr, ICPB_ERR_CUDA, "cudaGetDeviceCount", ce);
    }
    *count = c;
    return ICPB_OK;
}

static int ctx_create_common(int device, void *stream, bool own, icpb_ctx **out)
{
    if (!out) return fail(nullptr, ICPB_ERR_INVALID, "out == NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0)
        return fail(nullptr, ICPB_ERR_CUDA, "no CUDA device (libicpb200 has no CPU fallback)", ce);
    if (device < 0 || device >= count) return fail(nullptr, ICPB_ERR_INVALID, "device index out of range");
    CU(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(nullptr, ICPB_ERR_CUDA, "device is not sm_100-class; this library is built for sm_100a only");
    icpb_ctx *ctx = new (std::nothrow) icpb_ctx();
    if (!ctx) return fail(nullptr, ICPB_ERR_INVALID, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->ws.resize(WS_COUNT);
    ctx->own_stream = own;
    if (own) {
        ce = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (ce != cudaSuccess) { delete ctx; return fail(nullptr, ICPB_ERR_CUDA, "cudaStreamCreate", ce); }
    } else {
        ctx->stream = (cudaStream_t)stream;
    }
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    cudaEventCreate(&ctx->evt0);
    cudaEventCreate(&ctx->evt1);
    cudaEventCreateWithFlags(&ctx->ev_reg_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
    *out = ctx;
    return ICPB_OK;
}

int icpb_ctx_create(int device, icpb_ctx **out) { return ctx_create_common(device, nullptr, true, out); }

int icpb_ctx_create_on_stream(int device, void *cuda_stream, icpb_ctx **out)
{
    return ctx_create_common(device, cuda_stream, false, out);
}

int icpb_ctx_destroy(icpb_ctx *ctx)
{
    if (!ctx) return ICPB_OK;
    cudaSetDevice(ctx->device);
    drain_inflight(ctx); // an un-waited async registration keeps its results in its handle
    cudaStreamSynchronize(ctx->stream);
    for (auto &b : ctx->ws)
        if (b.p) cudaFree(b.p);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->pinned_reg) cudaFreeHost(ctx->pinned_reg);
    cudaEventDestroy(ctx->ev_reg_done);
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaEventDestroy(ctx->evt0);
    cudaEventDestroy(ctx->evt1);
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    for (const icpb_ctx::Span &sp : ctx->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ICPB_OK;
}

int icpb_ctx_sync(icpb_ctx *ctx)
{
    if (!ctx) return ICPB_ERR_INVALID;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ICPB_OK;
}

void *icpb_ctx_stream(icpb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int icpb_timer_start(icpb_ctx *ctx)
{
    if (!ctx) return ICPB_ERR_INVALID;
    CU(ctx, cudaEventRecord(ctx->evt0, ctx->stream));
    return ICPB_OK;
}

int icpb_timer_stop(icpb_ctx *ctx, float *elapsed_ms)
{
    if (!ctx || !elapsed_ms) return ICPB_ERR_INVALID;
    CU(ctx, cudaEventRecord(ctx->evt1, ctx->stream));
    CU(ctx, cudaEventSynchronize(ctx->evt1));
    CU(ctx, cudaEventElapsedTime(elapsed_ms, ctx->evt0, ctx->evt1));
    return ICPB_OK;
}

int icpb_ctx_set_profiling(icpb_ctx *ctx, int enabled)
{
    if (!ctx) return ICPB_ERR_INVALID;
    ctx->profiling = enabled != 0;
    return ICPB_OK;
}

int icpb_ctx_profile_read(icpb_ctx *ctx, int kernel, float *ms, int *launches)
{
    if (!ctx || kernel < 0 || kernel >= ICPB_PROF_COUNT) return ICPB_ERR_INVALID;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    float total = 0.f;
    int count = 0;
    std::vector<icpb_ctx::Span> keep;
    for (const icpb_ctx::Span &sp : ctx->spans) {
        if (sp.kernel != kernel) { keep.push_back(sp); continue; }
        float t = 0.f;
        if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) { total += t; ++count; }
        ctx->ev_pool.push_back(sp.a);
        ctx->ev_pool.push_back(sp.b);
    }
    ctx->spans.swap(keep);
    if (ms) *ms = total;
    if (launches) *launches = count;
    return ICPB_OK;
}

int icpb_ctx_launch_count(icpb_ctx *ctx, long long *count)
{
    if (!ctx || !count) return ICPB_ERR_INVALID;
    *count = ctx->launches;
    return ICPB_OK;
}

int icpb_measure_fp32_peak(icpb_ctx *ctx, int repeats, double *tflops, float *ms_out)
{
    if (!ctx || !tflops) return ICPB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
    float *d_out;
    int rc;
    if ((rc = ws_get(ctx, WS_MISC, (size_t)blocks * threads * sizeof(float), (void **)&d_out))) return rc;
    float best = 1e30f;
    if (repeats < 1) repeats = 1;
    launch_fp32_peak(d_out, blocks, threads, iters, ctx->stream); // warm-up
    for (int r = 0; r < repeats; ++r) {
        CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        launch_fp32_peak(d_out, blocks, threads, iters, ctx->stream);
        CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CU(ctx, cudaEventSynchronize(ctx->ev1));
        float ms;
        CU(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        best = std::min(best, ms);
    }
    ctx->launches += repeats + 1;
    const double flops = (double)blocks * threads * iters * 64.0 * 2.0; // 8 unrolled x 8 chains FMAs
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return ICPB_OK;
}

// ---- clouds -------------------------------------------------------------------

int icpb_cloud_create(icpb_ctx *ctx, int capacity, icpb_cloud **out)
{
    if (!ctx || !out || capacity < 0) return fail(ctx, ICPB_ERR_INVALID, "icpb_cloud_create: bad argument");
    CU(ctx, cudaSetDevice(ctx->device));
    icpb_cloud *c = new (std::nothrow) icpb_cloud();
    if (!c) return fail(ctx, ICPB_ERR_INVALID, "out of host memory");
    c->ctx = ctx;
    c->capacity = std::max(capacity, 1);
    cudaError_t ce = cudaMalloc((void **)&c->d_pts, sizeof(float4) * (size_t)c->capacity);
    if (ce != cudaSuccess) { delete c; return fail(ctx, ICPB_ERR_CUDA, "cudaMalloc(cloud)", ce); }
    *out = c;
    return ICPB_OK;
}

int icpb_cloud_destroy(icpb_cloud *cloud)
{
    if (!cloud) return ICPB_OK;
    cudaSetDevice(cloud->ctx->device);
    cudaStreamSynchronize(cloud->ctx->stream);
    cudaFree(cloud->d_pts);
    delete cloud;
    return ICPB_OK;
}

int icpb_cloud_size(const icpb_cloud *cloud, int *n)
{
    if (!cloud || !n) return ICPB_ERR_INVALID;
    *n = cloud->n;
    return ICPB_OK;
}

int icpb_cloud_upload(icpb_cloud *cloud, const icpb_point *points, int n)
{
    if (!cloud || n < 0 || (n > 0 && !points)) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = cloud->ctx;
    if (n > cloud->capacity) return fail(ctx, ICPB_ERR_CAPACITY, "icpb_cloud_upload: n > capacity");
    static_assert(sizeof(icpb_point) == sizeof(float4), "color_point_t is 16 bytes");
    CU(ctx, cudaSetDevice(ctx->device));
    if (n) CU(ctx, cudaMemcpyAsync(cloud->d_pts, points, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cloud->n = n;
    return ICPB_OK;
}

int icpb_cloud_upload_xyz(icpb_cloud *cloud, const float *xyz, int n)
{
    if (!cloud || n < 0 || (n > 0 && !xyz)) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = cloud->ctx;
    if (n > cloud->capacity) return fail(ctx, ICPB_ERR_CAPACITY, "icpb_cloud_upload_xyz: n > capacity");
    void *hp;
    int rc;
    if ((rc = pinned_get(ctx, sizeof(icpb_point) * (size_t)std::max(n, 1), &hp))) return rc;
    icpb_point *p = (icpb_point *)hp;
    for (int i = 0; i < n; ++i) {
        p[i].x = xyz[3 * i]; p[i].y = xyz[3 * i + 1]; p[i].z = xyz[3 * i + 2];
        p[i].c0 = p[i].c1 = p[i].c2 = p[i].pad = 0;
    }
    return icpb_cloud_upload(cloud, p, n);
}

int icpb_cloud_download(icpb_cloud *cloud, icpb_point *out, int capacity, int *n)
{
    if (!cloud || (!out && capacity > 0)) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = cloud->ctx;
    if (n) *n = cloud->n;
    if (capacity < cloud->n) return fail(ctx, ICPB_ERR_CAPACITY, "icpb_cloud_download: capacity < size");
    CU(ctx, cudaSetDevice(ctx->device));
    if (cloud->n)
        CU(ctx, cudaMemcpyAsync(out, cloud->d_pts, sizeof(float4) * (size_t)cloud->n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ICPB_OK;
}

int icpb_cloud_copy(icpb_cloud *dst, const icpb_cloud *src)
{
    if (!dst || !src) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = dst->ctx;
    if (src->n > dst->capacity) return fail(ctx, ICPB_ERR_CAPACITY, "icpb_cloud_copy: capacity");
    CU(ctx, cudaSetDevice(ctx->device));
    if (src->n)
        CU(ctx, cudaMemcpyAsync(dst->d_pts, src->d_pts, sizeof(float4) * (size_t)src->n, cudaMemcpyDeviceToDevice, ctx->stream));
    dst->n = src->n;
    return ICPB_OK;
}

int icpb_cloud_upload_device(icpb_cloud *cloud, const void *device_points, int n)
{
    if (!cloud || n < 0 || (n > 0 && !device_points)) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = cloud->ctx;
    if (n > cloud->capacity) return fail(ctx, ICPB_ERR_CAPACITY, "icpb_cloud_upload_device: n > capacity");
    CU(ctx, cudaSetDevice(ctx->device));
    if (n)
        CU(ctx, cudaMemcpyAsync(cloud->d_pts, device_points, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    cloud->n = n;
    return ICPB_OK;
}

const void *icpb_cloud_device_ptr(const icpb_cloud *cloud) { return cloud ? cloud->d_pts : nullptr; }

int icpb_cloud_download_device(icpb_cloud *cloud, void *device_dst, int capacity)
{
    if (!cloud || (!device_dst && cloud->n > 0)) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = cloud->ctx;
    if (capacity < cloud->n) return fail(ctx, ICPB_ERR_CAPACITY, "icpb_cloud_download_device: capacity < size");
    CU(ctx, cudaSetDevice(ctx->device));
    if (cloud->n)
        CU(ctx, cudaMemcpyAsync(device_dst, cloud->d_pts, sizeof(float4) * (size_t)cloud->n, cudaMemcpyDeviceToDevice, ctx->stream));
    return ICPB_OK;
}

void icpb_intrinsics_reference_v1(icpb_intrinsics *K)
{
    if (!K) return;
    K->fx_u = 468.60f; K->cx_u = 318.27f; K->fx_v = 468.60f; K->cx_v = 318.27f; K->depth_scale = 5000.0f;
}

void icpb_intrinsics_reference_v2(icpb_intrinsics *K)
{
    if (!K) return;
    K->fx_u = 363.58f; K->cx_u = 250.32f; K->fx_v = 363.58f; K->cx_v = 250.32f; K->depth_scale = 5000.0f;
}

int icpb_cloud_from_depth_device(icpb_cloud *cloud, const void *d_depth, const void *d_bgr, int w, int h,
                                 const icpb_intrinsics *K, int rule, uint32_t rule_arg, uint32_t seed,
                                 const void *d_keep_stream, int keep_stream_len)
{
    if (!cloud || !d_depth || !K || w <= 0 || h <= 0) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = cloud->ctx;
    if (rule < ICPB_SUB_NONE || rule > ICPB_SUB_STREAM) return fail(ctx, ICPB_ERR_INVALID, "unknown subsample rule");
    if (rule == ICPB_SUB_STREAM && !d_keep_stream) return fail(ctx, ICPB_ERR_INVALID, "keep_stream == NULL");
    if (((uintptr_t)d_depth & 15) != 0) return fail(ctx, ICPB_ERR_INVALID, "depth must be 16-byte aligned");
    CU(ctx, cudaSetDevice(ctx->device));
    BackprojectArgs a;
    a.depth = (const uint16_t *)d_depth;
    a.bgr = (const uint8_t *)d_bgr;
    a.w = w; a.h = h; a.K = *K;
    a.rule = rule; a.rule_arg = rule_arg; a.seed = seed;
    a.keep_stream = (const uint8_t *)d_keep_stream;
    a.keep_stream_len = keep_stream_len;
    a.out = cloud->d_pts;
    a.capacity = cloud->capacity;
    a.n_tiles = backproject_tiles(w, h);
    a.frames = 1;
    a.v_offset = 0;
    a.depth_stride = a.bgr_stride = a.out_stride = a.state_stride = 0;
    void *ts;
    int rc;
    // [reserved word][out_count (i32) | pad][2*n_tiles epoch-tagged status words][n_tiles words of per-tile counts]
    // (+ n_tiles words: the per-tile counts of the two-pass kernels, kept apart from the epoch-tagged status words)
    if ((rc = ws_get(ctx, WS_TILESTATE, sizeof(unsigned long long) * (3 * (size_t)a.n_tiles + 2), &ts, true))) return rc;
    a.out_count = (int *)((unsigned long long *)ts + 1);
    a.tile_state = (unsigned long long *)ts + 2;
    launch_backproject(a, ctx->stream);
    ctx->launches += 1;
    void *hp;
    if ((rc = pinned_get(ctx, 64, &hp))) return rc;
    CU(ctx, cudaMemcpyAsync(hp, a.out_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    int n = *(int *)hp;
    if (n < 0 && (rule == ICPB_SUB_NONE || rule == ICPB_SUB_HASH)) {
        // a predecessor tile never published (CTAs not dispatched in tile order): redo with the kernels that never wait
        launch_backproject(a, ctx->stream, true);
        ctx->launches += 2;
        CU(ctx, cudaMemcpyAsync(hp, a.out_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        CU(ctx, cudaGetLastError());
        n = *(int *)hp;
    }
    if (n < 0) {
        cloud->n = 0;
        CU(ctx, cudaMemsetAsync(ts, 0, sizeof(unsigned long long) * 2, ctx->stream));
        return fail(ctx, ICPB_ERR_CUDA, "icpb_cloud_from_depth: tile scan timed out (internal error)");
    }
    if (n > cloud->capacity) {
        cloud->n = cloud->capacity;
        return fail(ctx, ICPB_ERR_CAPACITY, "icpb_cloud_from_depth: more points than the cloud's capacity");
    }
    cloud->n = n;
    return ICPB_OK;
}

int icpb_backproject_batch_device(icpb_ctx *ctx, const void *d_depth, const void *d_bgr, int frames, int w, int h,
                                  const icpb_intrinsics *K, void *d_points, int capacity_per_frame, int *d_counts)
{
    if (!ctx || !d_depth || !K || !d_points || !d_counts || frames <= 0 || w <= 0 || h <= 0 || capacity_per_frame <= 0)
        return ICPB_ERR_INVALID;
    if ((((size_t)w * h * sizeof(uint16_t)) & 15) != 0 || ((uintptr_t)d_depth & 15) != 0)
        return fail(ctx, ICPB_ERR_INVALID, "batched depth frames must be 16-byte aligned");
    CU(ctx, cudaSetDevice(ctx->device));
    BackprojectArgs a;
    a.depth = (const uint16_t *)d_depth;
    a.bgr = (const uint8_t *)d_bgr;
    a.w = w; a.h = h; a.K = *K;
    a.rule = ICPB_SUB_NONE; a.rule_arg = 1; a.seed = 0;
    a.keep_stream = nullptr; a.keep_stream_len = 0;
    a.out = (float4 *)d_points;
    a.capacity = capacity_per_frame;
    a.n_tiles = backproject_tiles(w, h);
    a.frames = frames;
    a.v_offset = 0;
    a.depth_stride = (long long)w * h;
    a.bgr_stride = (long long)w * h * 3;
    a.out_stride = capacity_per_frame;
    a.state_stride = 3LL * a.n_tiles + 2; // reserved, count, 2 n_tiles status words, n_tiles words of per-tile counts
    void *ts;
    int rc;
    if ((rc = ws_get(ctx, WS_BATCHSTATE, sizeof(unsigned long long) * (size_t)a.state_stride * frames, &ts, true))) return rc;
    a.out_count = (int *)((unsigned long long *)ts + 1);
    a.tile_state = (unsigned long long *)ts + 2;
    launch_backproject(a, ctx->stream);
    ctx->launches += 1;
    // counts: word 1 of every frame's state block -> d_counts[frame]
    CU(ctx, cudaMemcpy2DAsync(d_counts, sizeof(int), (const char *)ts + sizeof(unsigned long long),
                              sizeof(unsigned long long) * (size_t)a.state_stride, sizeof(int), (size_t)frames,
                              cudaMemcpyDeviceToDevice, ctx->stream));
    CU(ctx, cudaGetLastError());
    return ICPB_OK;
}

int icpb_cloud_from_depth(icpb_cloud *cloud, const uint16_t *depth, const uint8_t *bgr, int w, int h,
                          const icpb_intrinsics *K, int rule, uint32_t rule_arg, uint32_t seed,
                          const uint8_t *keep_stream, int keep_stream_len)
{
    if (!cloud || !depth || !K || w <= 0 || h <= 0) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = cloud->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t npx = (size_t)w * h;
    void *d_depth, *d_bgr = nullptr, *d_keep = nullptr;
    int rc;
    if ((rc = ws_get(ctx, WS_DEPTH, npx * sizeof(uint16_t), &d_depth))) return rc;
    CU(ctx, cudaMemcpyAsync(d_depth, depth, npx * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    if (bgr) {
        if ((rc = ws_get(ctx, WS_BGR, npx * 3, &d_bgr))) return rc;
        CU(ctx, cudaMemcpyAsync(d_bgr, bgr, npx * 3, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (rule == ICPB_SUB_STREAM) {
        if (!keep_stream || keep_stream_len <= 0) return fail(ctx, ICPB_ERR_INVALID, "keep_stream missing");
        if ((rc = ws_get(ctx, WS_KEEP, (size_t)keep_stream_len, &d_keep))) return rc;
        CU(ctx, cudaMemcpyAsync(d_keep, keep_stream, (size_t)keep_stream_len, cudaMemcpyHostToDevice, ctx->stream));
    }
    return icpb_cloud_from_depth_device(cloud, d_depth, d_bgr, w, h, K, rule, rule_arg, seed, d_keep, keep_stream_len);
}

int icpb_cloud_transform(icpb_cloud *cloud, const float R[9], const float t[3])
{
    if (!cloud) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = cloud->ctx;
    if (cloud->n == 0 || (!R && !t)) return ICPB_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    launch_transform(cloud->d_pts, cloud->n, R, t, R != nullptr, t != nullptr, ctx->stream);
    ctx->launches += 1;
    CU(ctx, cudaGetLastError());
    return ICPB_OK;
}

int icpb_cloud_pack_band_device(icpb_cloud *cloud, void *device_dst, int band_capacity)
{
    if (!cloud || !device_dst) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = cloud->ctx;
    if (cloud->n > band_capacity) return fail(ctx, ICPB_ERR_CAPACITY, "icpb_cloud_pack_band_device: band too small");
    CU(ctx, cudaSetDevice(ctx->device));
    launch_pack_band(cloud->d_pts, cloud->n, (float4 *)device_dst, ctx->stream);
    ctx->launches += 1;
    CU(ctx, cudaGetLastError());
    return ICPB_OK;
}

int icpb_cloud_assemble_bands_device(icpb_cloud *cloud, const void *device_bands, int world, int band_capacity)
{
    if (!cloud || !device_bands || world <= 0 || band_capacity <= 0) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = cloud->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    void *misc;
    int rc;
    if ((rc = ws_get(ctx, WS_MISC, 64, &misc, true))) return rc;
    int *d_total = (int *)((char *)misc + 56);
    launch_assemble_bands((const float4 *)device_bands, world, band_capacity, cloud->d_pts, cloud->capacity, d_total, ctx->stream);
    ctx->launches += 1;
    int total = 0;
    CU(ctx, cudaMemcpyAsync(&total, d_total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    if (total > cloud->capacity) {
        cloud->n = cloud->capacity;
        return fail(ctx, ICPB_ERR_CAPACITY, "icpb_cloud_assemble_bands_device: cloud capacity exceeded");
    }
    cloud->n = total;
    return ICPB_OK;
}

int icpb_cloud_center(icpb_cloud *cloud, double center[3])
{
    if (!cloud || !center) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = cloud->ctx;
    if (cloud->n <= 0) return fail(ctx, ICPB_ERR_EMPTY, "icpb_cloud_center: empty cloud");
    CU(ctx, cudaSetDevice(ctx->device));
    const int nchunks = (cloud->n + kChunk - 1) / kChunk;
    double *d_chunks;
    void *misc;
    int rc;
    if ((rc = ws_get(ctx, WS_CHUNKS, (size_t)nchunks * kTerms * sizeof(double), (void **)&d_chunks))) return rc;
    if ((rc = ws_get(ctx, WS_MISC, 64, &misc, true))) return rc;
    double *d_out = (double *)misc;
    unsigned int *d_counter = (unsigned int *)((char *)misc + 32);
    CU(ctx, cudaMemsetAsync(d_counter, 0, sizeof(unsigned int), ctx->stream));
    launch_center(cloud->d_pts, cloud->n, d_chunks, d_out, d_counter, ctx->stream);
    ctx->launches += 1;
    CU(ctx, cudaMemcpyAsync(center, d_out, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    return ICPB_OK;
}

// ---- image-space stages ----------------------------------------------------------

int icpb_normals_from_depth(icpb_ctx *ctx, const uint16_t *depth, int w, int h, float *normals)
{
    if (!ctx || !depth || !normals || w <= 0 || h <= 0) return ICPB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t npx = (size_t)w * h;
    void *d_depth, *d_n;
    int rc;
    if ((rc = ws_get(ctx, WS_DEPTH, npx * sizeof(uint16_t), &d_depth))) return rc;
    if ((rc = ws_get(ctx, WS_NORMALS, npx * 3 * sizeof(float), &d_n))) return rc;
    CU(ctx, cudaMemcpyAsync(d_depth, depth, npx * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    launch_normals((const uint16_t *)d_depth, w, h, (float *)d_n, ctx->stream);
    ctx->launches += 1;
    CU(ctx, cudaMemcpyAsync(normals, d_n, npx * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    return ICPB_OK;
}

int icpb_normals_batch_device(icpb_ctx *ctx, const void *d_depth, int frames, int w, int h, void *d_normals)
{
    if (!ctx || !d_depth || !d_normals || frames <= 0 || w <= 0 || h <= 0) return ICPB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    launch_normals((const uint16_t *)d_depth, w, h, (float *)d_normals, ctx->stream, frames);
    ctx->launches += 1;
    CU(ctx, cudaGetLastError());
    return ICPB_OK;
}

int icpb_depth_filter(icpb_ctx *ctx, const uint16_t *depth, int w, int h, int min_d, int max_d, uint16_t *out)
{
    if (!ctx || !depth || !out || w <= 0 || h <= 0) return ICPB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)w * h * sizeof(uint16_t);
    void *d_in, *d_a, *d_b;
    int rc;
    if ((rc = ws_get(ctx, WS_DEPTH, bytes, &d_in))) return rc;
    if ((rc = ws_get(ctx, WS_IMG_A, bytes, &d_a))) return rc;
    if ((rc = ws_get(ctx, WS_IMG_B, bytes, &d_b))) return rc;
    CU(ctx, cudaMemcpyAsync(d_in, depth, bytes, cudaMemcpyHostToDevice, ctx->stream));
    launch_depth_filter((const uint16_t *)d_in, (uint16_t *)d_a, nullptr, (uint16_t *)d_b, w, h, min_d, max_d, ctx->stream);
    ctx->launches += 2;
    CU(ctx, cudaMemcpyAsync(out, d_b, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    return ICPB_OK;
}

// ---- registration -------------------------------------------------------------------

int icpb_nn_search_device(icpb_ctx *ctx, const icpb_cloud *data, const icpb_cloud *target, const int32_t **d_idx,
                          const float **d_dist)
{
    if (!ctx || !data || !target) return ICPB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    icpb_icp_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.max_iterations = 0;
    prm.threshold = 0.f;
    prm.max_nn_distance = 0.75f;
    RegHost r{const_cast<icpb_cloud *>(data), target};
    icpb_icp_result res;
    int rc = run_registrations(ctx, &r, 1, &prm, &res, false);
    if (rc) return rc;
    if (d_idx) *d_idx = (const int32_t *)ctx->ws[WS_IDX].p;
    if (d_dist) *d_dist = (const float *)ctx->ws[WS_DIST].p;
    return ICPB_OK;
}

int icpb_nn_search(icpb_ctx *ctx, const icpb_cloud *data, const icpb_cloud *target, int32_t *idx, float *dist,
                   int *exact_rescans)
{
    if (!ctx || !data || !target) return ICPB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    icpb_icp_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.max_nn_distance = 0.75f;
    RegHost r{const_cast<icpb_cloud *>(data), target};
    icpb_icp_result res;
    int rc = run_registrations(ctx, &r, 1, &prm, &res, false);
    if (rc) return rc;
    if (idx)
        CU(ctx, cudaMemcpyAsync(idx, ctx->ws[WS_IDX].p, sizeof(int32_t) * (size_t)data->n, cudaMemcpyDeviceToHost, ctx->stream));
    if (dist)
        CU(ctx, cudaMemcpyAsync(dist, ctx->ws[WS_DIST].p, sizeof(float) * (size_t)data->n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (exact_rescans) *exact_rescans = res.exact_rescans;
    return ICPB_OK;
}

int icpb_icp_register(icpb_ctx *ctx, icpb_cloud *data, const icpb_cloud *target, const icpb_icp_params *params,
                      icpb_icp_result *result)
{
    if (!ctx || !data || !target || !params) return ICPB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    RegHost r{data, target};
    return run_registrations(ctx, &r, 1, params, result, true);
}

int icpb_icp_register_async(icpb_ctx *ctx, icpb_cloud *data, const icpb_cloud *target, const icpb_icp_params *params,
                            icpb_pending **out)
{
    if (!ctx || !data || !target || !params || !out) return ICPB_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    RegHost r{data, target};
    return run_registrations(ctx, &r, 1, params, nullptr, true, out);
}

int icpb_icp_register_batch_async(icpb_ctx *ctx, icpb_cloud *const *data, const icpb_cloud *const *target, int count,
                                  const icpb_icp_params *params, icpb_pending **out)
{
    if (!ctx || !data || !target || !params || !out || count <= 0) return ICPB_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    std::vector<RegHost> regs((size_t)count);
    for (int b = 0; b < count; ++b) regs[b] = RegHost{data[b], target[b]};
    return run_registrations(ctx, regs.data(), count, params, nullptr, true, out);
}

int icpb_icp_pending_ready(icpb_pending *p, int *ready)
{
    if (!p || !ready) return ICPB_ERR_INVALID;
    if (p->finished) { *ready = 1; return ICPB_OK; }
    cudaError_t e = cudaEventQuery(p->ctx->ev_reg_done);
    if (e != cudaSuccess && e != cudaErrorNotReady) return fail(p->ctx, ICPB_ERR_CUDA, "cudaEventQuery", e);
    *ready = e == cudaSuccess;
    return ICPB_OK;
}

int icpb_icp_pending_wait(icpb_pending *p, icpb_icp_result *results)
{
    if (!p) return ICPB_ERR_INVALID;
    if (!p->finished) {
        cudaSetDevice(p->ctx->device);
        drain_inflight(p->ctx); // p is the context's one registration in flight
    }
    const int rc = p->status;
    if (rc == ICPB_OK && results)
        for (int b = 0; b < p->count; ++b) results[b] = p->results[(size_t)b];
    delete p;
    return rc;
}

int icpb_icp_register_carry(icpb_ctx *ctx, icpb_cloud *data, const icpb_cloud *target, icpb_cloud *carry,
                            const icpb_icp_params *params, icpb_icp_result *result)
{
    if (!ctx || !data || !target || !params) return ICPB_ERR_INVALID;
    if (carry && carry->ctx != ctx) return fail(ctx, ICPB_ERR_INVALID, "cloud belongs to another context");
    CU(ctx, cudaSetDevice(ctx->device));
    RegHost r{data, target};
    r.carry = (carry && carry->n > 0) ? carry : nullptr;
    return run_registrations(ctx, &r, 1, params, result, true);
}

int icpb_icp_register_batch(icpb_ctx *ctx, icpb_cloud *const *data, const icpb_cloud *const *target, int count,
                            const icpb_icp_params *params, icpb_icp_result *results)
{
    if (!ctx || !data || !target || !params || count <= 0) return ICPB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    std::vector<RegHost> regs((size_t)count);
    for (int b = 0; b < count; ++b) regs[b] = RegHost{data[b], target[b]};
    return run_registrations(ctx, regs.data(), count, params, results, true);
}

// 8f-2: the loop as the reference runs it (icp.cpp:98,155-258) -- key-points against the map's key-points
int icpb_icp_register_keypoints(icpb_ctx *ctx, icpb_cloud *keypoints, icpb_cloud *points, const icpb_cloud *map_keypoints,
                                const icpb_icp_params *params, icpb_icp_result *result, icpb_cloud *non_associations)
{
    if (!ctx || !keypoints || !map_keypoints || !params) return ICPB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    if ((points && points->ctx != ctx) || (non_associations && non_associations->ctx != ctx))
        return fail(ctx, ICPB_ERR_INVALID, "cloud belongs to another context");
    if (non_associations) non_associations->n = 0;
    if (keypoints->n <= 0 || map_keypoints->n <= 0) {
        // icp.cpp:490-491 (empty map: nothing is touched) / no key-points to associate: zero passes, identity motion
        if (result) {
            memset(result, 0, sizeof(*result));
            for (int k = 0; k < 3; ++k) {
                result->rigid[5 * k] = 1.f; result->cam_rotation[4 * k] = 1.f; result->pose_R[4 * k] = 1.0;
            }
            result->rigid[15] = 1.f;
        }
        return ICPB_OK;
    }
    RegHost r{keypoints, map_keypoints};
    r.carry = (points && points->n > 0) ? points : nullptr;
    r.nonassoc = non_associations;
    if (!r.carry && !r.nonassoc) return run_registrations(ctx, &r, 1, params, result, true);
    icpb_icp_params p = *params;
    p.nn_mode = ICPB_NN_BRUTE; // a few thousand key-points: the brute-force scan is the fast path
    return run_registrations(ctx, &r, 1, &p, result, true);
}

// ---- certainty map ---------------------------------------------------------------------

int icpb_map_create(icpb_ctx *ctx, const int dims[3], float cell, int z_lo, int z_hi, icpb_map **out)
{
    if (!ctx || !dims || !out) return ICPB_ERR_INVALID;
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 || !(cell > 0.f) || z_lo < 0 || z_hi > dims[2] || z_lo >= z_hi)
        return fail(ctx, ICPB_ERR_INVALID, "icpb_map_create: bad dims / cell / slab");
    CU(ctx, cudaSetDevice(ctx->device));
    icpb_map *m = new (std::nothrow) icpb_map();
    if (!m) return fail(ctx, ICPB_ERR_INVALID, "out of host memory");
    m->ctx = ctx;
    for (int k = 0; k < 3; ++k) m->dev.dims[k] = dims[k];
    m->dev.cell = cell;
    m->dev.z_lo = z_lo;
    m->dev.z_hi = z_hi;
    m->dev.zs = z_hi - z_lo;
    m->bytes = (long long)dims[0] * dims[1] * m->dev.zs;
    const size_t alloc = ((size_t)m->bytes + 3) / 4 * 4;
    cudaError_t ce = cudaMalloc((void **)&m->dev.grid, alloc);
    if (ce != cudaSuccess) { delete m; return fail(ctx, ICPB_ERR_CUDA, "cudaMalloc(map)", ce); }
    ce = cudaMemsetAsync(m->dev.grid, 0, alloc, ctx->stream); // map.cpp:23-30
    if (ce != cudaSuccess) { cudaFree(m->dev.grid); delete m; return fail(ctx, ICPB_ERR_CUDA, "cudaMemset(map)", ce); }
    // brick occupancy bits (icpb_internal.h), all clear = "every voxel is zero"
    const long long nbx = (dims[0] + kBrick - 1) / kBrick;
    m->dev.nby = (dims[1] + kBrick - 1) / kBrick;
    m->dev.nbz = (m->dev.zs + kBrick - 1) / kBrick;
    const long long words1 = (nbx * m->dev.nby * m->dev.nbz + 31) / 32;
    const long long nbx2 = (dims[0] + kBrick2 - 1) / kBrick2;
    m->dev.nby2 = (dims[1] + kBrick2 - 1) / kBrick2;
    m->dev.nbz2 = (m->dev.zs + kBrick2 - 1) / kBrick2;
    m->brick_words = words1 + (nbx2 * m->dev.nby2 * m->dev.nbz2 + 31) / 32; // both levels in one allocation
    ce = cudaMalloc((void **)&m->dev.bricks, sizeof(uint32_t) * (size_t)m->brick_words);
    m->dev.bricks2 = m->dev.bricks + words1;
    if (ce == cudaSuccess) ce = cudaMemsetAsync(m->dev.bricks, 0, sizeof(uint32_t) * (size_t)m->brick_words, ctx->stream);
    if (ce != cudaSuccess) {
        cudaFree(m->dev.grid);
        if (m->dev.bricks) cudaFree(m->dev.bricks);
        delete m;
        return fail(ctx, ICPB_ERR_CUDA, "cudaMalloc(map occupancy)", ce);
    }
    *out = m;
    return ICPB_OK;
}

int icpb_map_destroy(icpb_map *map)
{
    if (!map) return ICPB_OK;
    cudaSetDevice(map->ctx->device);
    cudaStreamSynchronize(map->ctx->stream);
    cudaFree(map->dev.grid);
    cudaFree(map->dev.bricks);
    if (map->table) cudaFree(map->table);
    delete map;
    return ICPB_OK;
}

int icpb_map_clear(icpb_map *map)
{
    if (!map) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = map->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemsetAsync(map->dev.grid, 0, ((size_t)map->bytes + 3) / 4 * 4, ctx->stream));
    CU(ctx, cudaMemsetAsync(map->dev.bricks, 0, sizeof(uint32_t) * (size_t)map->brick_words, ctx->stream));
    if (map->table) CU(ctx, cudaMemsetAsync(map->table, 0xff, sizeof(int) * (size_t)map->bytes, ctx->stream));
    return ICPB_OK;
}

int icpb_map_update_endpoints(icpb_map *map, const icpb_cloud *points, int rule, int delta, int max_conf)
{
    if (!map || !points) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = map->ctx;
    if (rule != ICPB_RULE_A && rule != ICPB_RULE_C) return fail(ctx, ICPB_ERR_INVALID, "unknown update rule");
    if (delta < 0 || delta > 255) return fail(ctx, ICPB_ERR_INVALID, "delta out of [0,255]");
    CU(ctx, cudaSetDevice(ctx->device));
    launch_map_endpoints(map->dev, flat_src(points->d_pts, points->n), rule, delta, max_conf, ctx->stream);
    ctx->launches += points->n > 0;
    CU(ctx, cudaGetLastError());
    return ICPB_OK;
}

int icpb_map_update_tracked(icpb_map *map, const icpb_cloud *points, int variant, int delta, int max_conf,
                            icpb_cloud *map_cloud, int *n_appended)
{
    if (!map_cloud) return ICPB_ERR_INVALID;
    return icpb_map_update_tracked_base(map, points, variant, delta, max_conf, map_cloud, map_cloud->n, n_appended);
}

int icpb_map_table_entry(icpb_map *map, const int v[3], int *entry)
{
    if (!map || !v || !entry) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = map->ctx;
    *entry = -1;
    for (int k = 0; k < 3; ++k)
        if (v[k] < 0 || v[k] >= map->dev.dims[k]) return fail(ctx, ICPB_ERR_INVALID, "icpb_map_table_entry: voxel outside the grid");
    if (v[2] < map->dev.z_lo || v[2] >= map->dev.z_hi) return fail(ctx, ICPB_ERR_INVALID, "icpb_map_table_entry: voxel outside the slab");
    if (!map->table) return ICPB_OK;
    const size_t lin = ((size_t)v[0] * map->dev.dims[1] + v[1]) * map->dev.zs + (v[2] - map->dev.z_lo);
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(entry, map->table + lin, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ICPB_OK;
}

int icpb_map_update_tracked_base(icpb_map *map, const icpb_cloud *points, int variant, int delta, int max_conf,
                                 icpb_cloud *map_cloud, int table_base, int *n_appended)
{
    if (!map || !points || !map_cloud) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = map->ctx;
    if (variant < ICPB_TRACK_INIT || variant > ICPB_TRACK_NONASSOC) return fail(ctx, ICPB_ERR_INVALID, "unknown variant");
    if (delta < 0 || delta > 255) return fail(ctx, ICPB_ERR_INVALID, "delta out of [0,255]");
    if (map->dev.z_lo != 0 || map->dev.z_hi != map->dev.dims[2])
        return fail(ctx, ICPB_ERR_INVALID, "icpb_map_update_tracked needs a whole-map handle");
    if (points->n > 65536) return fail(ctx, ICPB_ERR_CAPACITY, "icpb_map_update_tracked: more than 65536 points");
    if (n_appended) *n_appended = 0;
    if (points->n == 0) return ICPB_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!map->table) {
        const size_t tb = sizeof(int) * (size_t)map->bytes;
        CU(ctx, cudaMalloc((void **)&map->table, tb));
        CU(ctx, cudaMemsetAsync(map->table, 0xff, tb, ctx->stream)); // -1 = empty (map.cpp:27)
    }
    void *wsp;
    int rc;
    if ((rc = ws_get(ctx, WS_TRACK, 16 + (sizeof(long long) + sizeof(int)) * 65536, &wsp))) return rc;
    int *d_app = (int *)wsp;
    launch_map_tracked(map->dev, map->table, points->d_pts, points->n, variant, delta, max_conf, map_cloud->d_pts,
                       map_cloud->n, map_cloud->capacity, d_app, table_base, ctx->stream);
    ctx->launches += 1;
    int app = 0;
    CU(ctx, cudaMemcpyAsync(&app, d_app, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    if (map_cloud->n + app > map_cloud->capacity) {
        map_cloud->n = map_cloud->capacity;
        return fail(ctx, ICPB_ERR_CAPACITY, "icpb_map_update_tracked: map cloud capacity exceeded");
    }
    map_cloud->n += app;
    if (n_appended) *n_appended = app;
    return ICPB_OK;
}

int icpb_map_has_entry(icpb_map *map, const float p[3], int *has_entry)
{
    if (!map || !p || !has_entry) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = map->ctx;
    *has_entry = 0;
    if (!map->table) return ICPB_OK;
    int v[3];
    icpb_map_voxel_coords(map, p, v);
    const size_t lin = ((size_t)v[0] * map->dev.dims[1] + v[1]) * map->dev.zs + (v[2] - map->dev.z_lo);
    int e = -1;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(&e, map->table + lin, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    *has_entry = e >= 0;
    return ICPB_OK;
}

int icpb_map_integrate_rays(icpb_map *map, const icpb_cloud *points, const float origin[3], int delta_dec,
                            int delta_inc, long long *voxels_visited)
{
    if (!map || !points || !origin) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = map->ctx;
    if (delta_dec < 0 || delta_dec > 255 || delta_inc < 0 || delta_inc > 255)
        return fail(ctx, ICPB_ERR_INVALID, "delta out of [0,255]");
    CU(ctx, cudaSetDevice(ctx->device));
    unsigned long long *d_vis = nullptr;
    void *misc;
    int rc;
    if ((rc = ws_get(ctx, WS_MISC, 64, &misc, true))) return rc;
    unsigned int *d_next = (unsigned int *)((char *)misc + 40); // ray counter of the persistent walk kernel
    if (voxels_visited) {
        d_vis = (unsigned long long *)((char *)misc + 48);
        CU(ctx, cudaMemsetAsync(d_vis, 0, sizeof(unsigned long long), ctx->stream));
    }
    int sp = span_begin(ctx, ICPB_PROF_MAP_RAYS);
    launch_map_rays(map->dev, flat_src(points->d_pts, points->n), origin, delta_dec, d_vis, d_next, ctx->sm_count, ctx->stream); // phase 1
    span_end(ctx, sp);
    sp = span_begin(ctx, ICPB_PROF_MAP_ENDPOINTS);
    launch_map_endpoints(map->dev, flat_src(points->d_pts, points->n), ICPB_RULE_A, delta_inc, 0, ctx->stream);  // phase 2
    span_end(ctx, sp);
    ctx->launches += 2 * (points->n > 0);
    CU(ctx, cudaGetLastError());
    if (voxels_visited) {
        unsigned long long v = 0;
        CU(ctx, cudaMemcpyAsync(&v, d_vis, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        *voxels_visited = (long long)v;
    }
    return ICPB_OK;
}

int icpb_map_integrate_rays_profiled(icpb_map *map, const icpb_cloud *points, const float origin[3], int delta_dec,
                                     int delta_inc, unsigned long long *layer_work)
{
    if (!map || !points || !origin || !layer_work) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = map->ctx;
    if (delta_dec < 0 || delta_dec > 255 || delta_inc < 0 || delta_inc > 255)
        return fail(ctx, ICPB_ERR_INVALID, "delta out of [0,255]");
    if (map->dev.z_lo != 0 || map->dev.z_hi != map->dev.dims[2])
        return fail(ctx, ICPB_ERR_INVALID, "icpb_map_integrate_rays_profiled needs a whole-map handle");
    CU(ctx, cudaSetDevice(ctx->device));
    const int layers = map->dev.dims[2];
    void *misc, *hist;
    int rc;
    if ((rc = ws_get(ctx, WS_MISC, 64, &misc, true))) return rc;
    if ((rc = ws_get(ctx, WS_FRAME, sizeof(unsigned int) * (size_t)layers, &hist))) return rc;
    unsigned int *d_next = (unsigned int *)((char *)misc + 40);
    CU(ctx, cudaMemsetAsync(hist, 0, sizeof(unsigned int) * (size_t)layers, ctx->stream));
    launch_map_rays(map->dev, flat_src(points->d_pts, points->n), origin, delta_dec, nullptr, d_next, ctx->sm_count, ctx->stream,
                    (unsigned int *)hist);
    launch_map_endpoints(map->dev, flat_src(points->d_pts, points->n), ICPB_RULE_A, delta_inc, 0, ctx->stream);
    ctx->launches += 2 * (points->n > 0);
    std::vector<unsigned int> h((size_t)layers);
    CU(ctx, cudaMemcpyAsync(h.data(), hist, sizeof(unsigned int) * (size_t)layers, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    for (int z = 0; z < layers; ++z) layer_work[z] += h[(size_t)z];
    return ICPB_OK;
}

int icpb_slab_bounds_from_work(const unsigned long long *work, int layers, int world, int *bounds)
{
    if (!work || !bounds || layers <= 0 || world <= 0 || world > layers) return ICPB_ERR_INVALID;
    long double total = 0;
    for (int z = 0; z < layers; ++z) total += (long double)work[z] + 1e-3L; // tiny floor: empty layers still get an owner
    bounds[0] = 0;
    long double acc = 0;
    int z = 0;
    for (int g = 1; g < world; ++g) {
        const long double want = total * g / world;
        while (z < layers && acc + (long double)work[z] + 1e-3L <= want) { acc += (long double)work[z] + 1e-3L; ++z; }
        // the boundary sits where the running sum is closest to the quantile, leaving room for the slabs still to come
        int b = z;
        if (z < layers && (want - acc) > ((long double)work[z] + 1e-3L) / 2) b = z + 1;
        b = std::max(b, bounds[g - 1] + 1);
        b = std::min(b, layers - (world - g));
        bounds[g] = b;
        while (z < b) { acc += (long double)work[z] + 1e-3L; ++z; }
    }
    bounds[world] = layers;
    return ICPB_OK;
}

// ---- sync-free frame path (multi-GPU z-slab map and the single-GPU sequence alike) --------------------------------
// The point count of a lifted frame never visits the host: it travels in the band's header row and every consumer
// kernel reads it from device memory, so a sequence of frames is enqueued without a single host synchronisation.

int icpb_frame_lift_band_device(icpb_ctx *ctx, const void *d_depth, int w, int h, int row0, int row1,
                                const icpb_intrinsics *K, const float R[9], const float t[3], void *d_band,
                                int band_capacity)
{
    if (!ctx || !d_depth || !K || !d_band || w <= 0 || h <= 0) return ICPB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    void *ts;
    int rc;
    const int tiles = backproject_tiles(w, row1 > row0 ? row1 - row0 : 1);
    if ((rc = ws_get(ctx, WS_TILESTATE, sizeof(unsigned long long) * (3 * (size_t)tiles + 2), &ts, true))) return rc;
    const int sp = span_begin(ctx, ICPB_PROF_LIFT);
    rc = icpb::api_lift_band(ctx, ctx->stream, ts, d_depth, w, h, row0, row1, K, R, t, d_band, band_capacity);
    span_end(ctx, sp);
    return rc;
}

int icpb_map_integrate_bands_device(icpb_map *map, const void *d_bands, int world, int band_capacity, const float origin[3],
                                    int delta_dec, int delta_inc)
{
    if (!map || !d_bands || !origin || world <= 0 || band_capacity <= 0) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = map->ctx;
    if (delta_dec < 0 || delta_dec > 255 || delta_inc < 0 || delta_inc > 255)
        return fail(ctx, ICPB_ERR_INVALID, "delta out of [0,255]");
    CU(ctx, cudaSetDevice(ctx->device));
    void *misc;
    int rc;
    if ((rc = ws_get(ctx, WS_MISC, 64, &misc, true))) return rc;
    unsigned int *d_next = (unsigned int *)((char *)misc + 40);
    if (world > kMaxBands) return fail(ctx, ICPB_ERR_INVALID, "more bands than kMaxBands");
    // the kernels read the bands where they lie (header row + points, band_capacity + 1 rows apart): no assembly pass
    const PointSrc src = band_src((const float4 *)d_bands, world, band_capacity, (long long)band_capacity + 1);
    int sp = span_begin(ctx, ICPB_PROF_MAP_RAYS);
    launch_map_rays(map->dev, src, origin, delta_dec, nullptr, d_next, ctx->sm_count, ctx->stream); // phase 1
    span_end(ctx, sp);
    sp = span_begin(ctx, ICPB_PROF_MAP_ENDPOINTS);
    launch_map_endpoints(map->dev, src, ICPB_RULE_A, delta_inc, 0, ctx->stream);                   // phase 2
    span_end(ctx, sp);
    ctx->launches += 2;
    CU(ctx, cudaGetLastError());
    return ICPB_OK;
}

int icpb_map_voxel_coords(const icpb_map *map, const float p[3], int v[3])
{
    if (!map || !p || !v) return ICPB_ERR_INVALID;
    for (int k = 0; k < 3; ++k) { // map.cpp:55-85
        int q = (int)(p[k] / map->dev.cell);
        if (q < 0) q = 0;
        if (q >= map->dev.dims[k]) q = map->dev.dims[k] - 1;
        v[k] = q;
    }
    return ICPB_OK;
}

int icpb_map_size_bytes(const icpb_map *map, long long *size)
{
    if (!map || !size) return ICPB_ERR_INVALID;
    *size = map->bytes;
    return ICPB_OK;
}

int icpb_map_download(icpb_map *map, uint8_t *out, long long capacity)
{
    if (!map || !out) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = map->ctx;
    if (capacity < map->bytes) return fail(ctx, ICPB_ERR_CAPACITY, "icpb_map_download: capacity < size");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(out, map->dev.grid, (size_t)map->bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ICPB_OK;
}

int icpb_map_upload(icpb_map *map, const uint8_t *in, long long size)
{
    if (!map || !in) return ICPB_ERR_INVALID;
    icpb_ctx *ctx = map->ctx;
    if (size != map->bytes) return fail(ctx, ICPB_ERR_INVALID, "icpb_map_upload: size mismatch");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(map->dev.grid, in, (size_t)size, cudaMemcpyHostToDevice, ctx->stream));
    launch_map_rebuild_bricks(map->dev, map->brick_words, ctx->stream);
    ctx->launches += 1;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaGetLastError());
    return ICPB_OK;
}

// ---- pose reporting (8f-4): scalar host arithmetic, as in the reference ------------------------------------
// Separately rounded float operations throughout (host code of this file is built with -ffp-contract=off).
namespace {
inline float sign_of(float v) { return v >= 0.0f ? 1.0f : -1.0f; } // SIGN, quaternion.hpp:22
constexpr float kPiF = 3.14159265358979f;                            // PI, icp.hpp:4
} // namespace

int icpb_pose_quat_from_rotation(const float R[9], float q[4]) // quaternion.cpp:23-79
{
    if (!R || !q) return ICPB_ERR_INVALID;
    const float tr[4] = {(R[0] + R[4] + R[8] + 1.0f) / 4.0f, (R[0] - R[4] - R[8] + 1.0f) / 4.0f,
                         (-R[0] + R[4] - R[8] + 1.0f) / 4.0f, (-R[0] - R[4] + R[8] + 1.0f) / 4.0f};
    float c[4]; // magnitudes of w, x, y, z
    for (int k = 0; k < 4; ++k) c[k] = sqrtf(tr[k] < 0.0f ? 0.0f : tr[k]);
    // antisymmetric / symmetric off-diagonal combinations that carry the signs
    const float a_x = R[7] - R[5], a_y = R[2] - R[6], a_z = R[3] - R[1];
    const float s_xy = R[3] + R[1], s_xz = R[2] + R[6], s_yz = R[7] + R[5];
    if (c[0] >= c[1] && c[0] >= c[2] && c[0] >= c[3]) {        // w largest: signs from the antisymmetric part
        c[1] *= sign_of(a_x); c[2] *= sign_of(a_y); c[3] *= sign_of(a_z);
    } else if (c[1] >= c[0] && c[1] >= c[2] && c[1] >= c[3]) { // x largest
        c[0] *= sign_of(a_x); c[2] *= sign_of(s_xy); c[3] *= sign_of(s_xz);
    } else if (c[2] >= c[0] && c[2] >= c[1] && c[2] >= c[3]) { // y largest
        c[0] *= sign_of(a_y); c[1] *= sign_of(s_xy); c[3] *= sign_of(s_yz);
    } else if (c[3] >= c[0] && c[3] >= c[1] && c[3] >= c[2]) { // z largest
        c[0] *= sign_of(a_z); c[1] *= sign_of(s_xz); c[2] *= sign_of(s_yz);
    }
    const float len = sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3]); // NORM, quaternion.hpp:23
    for (int k = 0; k < 4; ++k) q[k] = c[k] / len;
    return ICPB_OK;
}

int icpb_pose_quat_mul(const float a[4], const float b[4], float out[4]) // quaternion.cpp:184-192
{
    if (!a || !b || !out) return ICPB_ERR_INVALID;
    const float t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                        a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3],
                        a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]};
    for (int k = 0; k < 4; ++k) out[k] = t[k];
    return ICPB_OK;
}

int icpb_pose_quat_inverse(const float q[4], float out[4]) // quaternion.cpp:325-328: conjugate scaled by 1 / squared norm
{
    if (!q || !out) return ICPB_ERR_INVALID;
    const float inv = 1 / (q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const float t[4] = {q[0] * inv, -q[1] * inv, -q[2] * inv, -q[3] * inv};
    for (int k = 0; k < 4; ++k) out[k] = t[k];
    return ICPB_OK;
}

int icpb_pose_quat_to_euler_deg(const float q[4], float e[3]) // SLAM.cpp:613-636
{
    if (!q || !e) return ICPB_ERR_INVALID;
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    const float yy = y * y;
    const float roll = atan2f(2.0f * (w * x + y * z), 1.0f - 2.0f * (x * x + yy));
    float sp = 2.0f * (w * y - z * x);
    sp = sp > 1.0f ? 1.0f : sp;
    sp = sp < -1.0f ? -1.0f : sp;
    const float pitch = asinf(sp);
    const float yaw = atan2f(2.0f * (w * z + x * y), 1.0f - 2.0f * (yy + z * z));
    e[0] = roll * 180.0f / kPiF; e[1] = pitch * 180.0f / kPiF; e[2] = yaw * 180.0f / kPiF;
    return ICPB_OK;
}

int icpb_pose_matrix_to_euler_deg(const float R[9], float e[3]) // SLAM.cpp:638-648
{
    if (!R || !e) return ICPB_ERR_INVALID;
    const float c2 = (float)sqrt((double)R[0] * (double)R[0] + (double)R[1] * (double)R[1]); // pow(float, 2): double
    e[0] = atan2f(R[5], R[8]) * 180 / kPiF;
    e[1] = atan2f(-R[2], c2) * 180 / kPiF;
    e[2] = atan2f(R[1], R[0]) * 180 / kPiF;
    return ICPB_OK;
}

} // extern "C"

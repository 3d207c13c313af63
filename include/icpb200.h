/*
 * icpb200.h -- C-ABI of libicpb200.so: the B200-native (sm_100a) registration
 * + mapping hot path of BenniG123/icp-slam-prototype.
 *
 * The reference has no FFI: its boundary is the C++ declarations in icp.hpp,
 * pointcloud.hpp and map.hpp (cv::Mat / std::vector by value).  Each entry
 * point below names the reference interface it replaces (file:line relative
 * to the reference repository).  Header-compatible C++ wrappers with the
 * reference's own names live in include/icpb200/{icp,pointcloud,map}.hpp and
 * marshal to these calls.
 *
 * Conventions: plain pointers + sizes, POD structs, opaque handles that own
 * device memory, caller-owned host buffers, int status returns (0 = ok), no
 * exceptions, no C++ types.  One context = one device + one CUDA stream;
 * calls on one context are serialised by the caller.  There is NO CPU
 * fallback: without a usable CUDA device every compute call fails with
 * ICPB_ERR_CUDA.
 */
#ifndef ICPB200_H
#define ICPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICPB_VERSION 100

enum {
    ICPB_OK = 0,
    ICPB_ERR_INVALID = 1,   /* bad argument */
    ICPB_ERR_EMPTY = 2,     /* empty cloud where the reference would dereference begin() (icp.cpp:572) */
    ICPB_ERR_CUDA = 3,      /* CUDA runtime / no device */
    ICPB_ERR_CAPACITY = 4,  /* output does not fit the handle's capacity */
    ICPB_ERR_NCCL = 5       /* NCCL missing (libnccl.so.2 could not be loaded) or a collective failed */
};

typedef struct icpb_ctx icpb_ctx;
typedef struct icpb_cloud icpb_cloud;
typedef struct icpb_map icpb_map;
typedef struct icpb_pending icpb_pending; /* a registration that has been enqueued and not yet waited for */
typedef struct icpb_comm icpb_comm;       /* one rank of a multi-GPU job: an NCCL communicator bound to a context */
typedef struct icpb_slabmap icpb_slabmap; /* the rank's z-slab of a certainty map shared by the job */

/* color_point_t, pointcloud.hpp:13-19 (cv::Point3f + cv::Vec3b + 1 pad = 16 B). */
typedef struct {
    float x, y, z;
    uint8_t c0, c1, c2, pad;
} icpb_point;

/* x = (u - cx_u) * z / fx_u ; y = (v - cx_v) * z / fx_v ; z = d / depth_scale.
 * The reference uses CX and FX on both axes (pointcloud.cpp:38-39):
 * icpb_intrinsics_reference_v1() reproduces that. */
typedef struct {
    float fx_u, cx_u, fx_v, cx_v, depth_scale;
} icpb_intrinsics;

/* Stand-ins for `rand() % SUBSAMPLE_FACTOR` (pointcloud.cpp:28,125). */
enum {
    ICPB_SUB_NONE = 0,   /* keep every non-zero pixel (factor 1) */
    ICPB_SUB_STRIDE = 1, /* keep non-zero pixel k iff k % arg == 0 */
    ICPB_SUB_HASH = 2,   /* keep iff hash32(seed, pixel) % arg == 0 */
    ICPB_SUB_STREAM = 3  /* keep non-zero pixel k iff keep_stream[k] != 0 (replays libc rand()) */
};

enum { ICPB_SOLVE_REFERENCE = 0, /* icp.cpp:199-246: uncentred SVD, offset = mean(a-b) */
       ICPB_SOLVE_KABSCH = 1 };  /* rigid_transform_3D.py:9-40 */

/* Association scan of the registration loop.  Both return the same associations (icp.cpp:541-563):
 * BRUTE scans all N x M pairs (icp.cpp:566-593); GRID is an exact cell-grid search (the reference's intended
 * voxel-indexed scan, icp.cpp:347-486) that reports idx -1 / dist +inf for queries with no neighbour inside
 * max_nn_distance - those are rejected by icp.cpp:553 in either mode. */
enum { ICPB_NN_BRUTE = 0, ICPB_NN_GRID = 1,
       ICPB_NN_AUTO = 2 }; /* GRID when n*m is large enough for the bucketing to pay off, else BRUTE */
/* AUTO: one registration takes GRID from n*m >= 2e9 on (about 45k x 45k points); a batch builds the cells of all its
 * registrations together and takes GRID when its clouds hold >= 4,096 points.  icpb_icp_result.nn_mode_used says which
 * one ran (same associations either way).  The key-point variants always scan (a few thousand points). */

/* Approximate FP32 filter in front of the exact re-evaluation of the BRUTE scan (never visible in the results):
 * CENTRED evaluates |t'|^2 - 2a'.t' about one centre per THREAD (3 FMA per pair + the centring of the targets per
 * thread); WARP does the same about one centre per warp - the queries are first put into a spatial (Morton) order so
 * that a warp's queries are neighbours, and the targets are centred once per warp in shared memory; DIRECT
 * evaluates (a-t)^2 (6 FP32 operations per pair; kept for A/B measurements).
 * AUTO = WARP for a single registration of >= 50,000 queries and for batches of >= 4,096-query registrations,
 * CENTRED otherwise (small single registrations: the ordering would cost what it saves). */
enum { ICPB_FILTER_AUTO = 0, ICPB_FILTER_DIRECT = 1, ICPB_FILTER_WARP = 2, ICPB_FILTER_CENTRED = 3 };

enum { ICPB_RULE_A = 0,  /* map.cpp:249-253 / 104-113 */
       ICPB_RULE_C = 1 };/* map.cpp:139-149 */

typedef struct {
    int max_iterations;        /* SLAM.cpp:277: 16 */
    float threshold;           /* SLAM.cpp:277: 1e-4 */
    float max_nn_distance;     /* MAX_NN_COLOR_DISTANCE, icp.hpp:8 */
    int solve_mode;            /* ICPB_SOLVE_* */
    float last_translation[3]; /* icp.cpp:25, consumed by the <3 associations rule (icp.cpp:163-182) */
    int32_t *idx_trace;        /* optional host buffer (max_iterations+1)*n: nearest index per pass */
    float *dist_trace;         /* optional host buffer, same shape */
    int nn_mode;               /* ICPB_NN_BRUTE (default 0), ICPB_NN_GRID or ICPB_NN_AUTO */
    float grid_cell;           /* ICPB_NN_GRID: cell edge in metres (0 = chosen from the target's density) */
    int nn_filter;             /* ICPB_FILTER_* (default 0 = AUTO); env ICPB_NN_FILTER overrides */
} icpb_icp_params;

typedef struct {
    int iterations;
    int nn_passes;
    int n_assoc;
    float mse;
    float rigid[16];        /* getTransformation's return value, icp.cpp:227-233,266-268 */
    float cam_rotation[9];  /* delta applied to cameraRotation (icp.cpp:237), starting from identity */
    float cam_position[3];  /* delta applied to cameraPosition (icp.cpp:246), starting from 0 */
    float offset[3];        /* last offset (icp.cpp:240) */
    double pose_R[9];       /* composed transform applied to the data cloud */
    double pose_t[3];
    int small_assoc_exit;
    int exact_rescans;      /* queries that needed the full exact FP64 rescan (diagnostic) */
    float gpu_ms;           /* device time of the registration loop (CUDA events on the context stream) */
    int kernel_launches;    /* kernels launched by this call */
    float nn_partial_ms;    /* profiling mode only: summed device time of the nn_partial launches */
    int nn_partial_launches;
    int nn_qpt, nn_splits;  /* work decomposition chosen for nn_partial */
    int nn_mode_used;       /* ICPB_NN_BRUTE or ICPB_NN_GRID */
    float grid_cell_used;
    int nn_filter_used;     /* ICPB_FILTER_* of the BRUTE scan */
    int n_nonassoc;         /* icpb_icp_register_keypoints: length of the accumulated reject list */
    long long grid_pairs;   /* profiling mode, ICPB_NN_GRID: (query, candidate) pairs the cooperative search evaluated over
                             * all passes of the call (8 flop each: the work behind its roofline figure) */
    float nn_grid_ms;       /* profiling mode: summed device time of the search kernels of the call */
} icpb_icp_result;

/* ---- library / context ------------------------------------------------- */
int icpb_version(void);
const char *icpb_status_string(int status);
/* Last error text of a context (or of the failed icpb_ctx_create when ctx == NULL). */
const char *icpb_last_error(const icpb_ctx *ctx);
int icpb_device_count(int *count);
/* Replaces the file-scope globals of icp.cpp:22-26 with an explicit handle. */
int icpb_ctx_create(int device, icpb_ctx **out);
/* Same, but work is enqueued on an existing cudaStream_t (e.g. torch's). */
int icpb_ctx_create_on_stream(int device, void *cuda_stream, icpb_ctx **out);
int icpb_ctx_destroy(icpb_ctx *ctx);
int icpb_ctx_sync(icpb_ctx *ctx);
void *icpb_ctx_stream(icpb_ctx *ctx);
/* CUDA-event stopwatch on the context stream (used by bench.py). */
int icpb_timer_start(icpb_ctx *ctx);
int icpb_timer_stop(icpb_ctx *ctx, float *elapsed_ms);
/* Profiling mode: bracket every nn_partial launch with CUDA events on the context
 * stream and report their summed duration in icpb_icp_result (roofline evidence). */
int icpb_ctx_set_profiling(icpb_ctx *ctx, int enabled);
/* Profiling mode also brackets every launch of the kernels below with CUDA events on the context stream;
 * icpb_ctx_profile_read synchronises the stream, returns the summed duration and the number of launches of `kernel`
 * since the last read, and forgets them (the per-kernel times behind bench.py's roofline figures; the stage names
 * follow the reference's logDeltaTime keys, SLAM.hpp:4-13, where one exists). */
enum { ICPB_PROF_MAP_RAYS = 0,     /* Map::rayTrace, map.cpp:272-439 */
       ICPB_PROF_MAP_ENDPOINTS = 1, /* Map::update, map.cpp:220-269 */
       ICPB_PROF_NN_GRID = 2,      /* LOG_NEAREST_NEIGHBOR: the cell-grid search kernels of one pass */
       ICPB_PROF_NN_FINALIZE = 3,  /* LOG_SVD / LOG_ROTATE / LOG_TRANSLATE: sums + solve */
       ICPB_PROF_LIFT = 4,         /* LOG_GEN_POINT_CLOUD: back-projection (+ transform) of one frame */
       ICPB_PROF_COUNT = 5 };
int icpb_ctx_profile_read(icpb_ctx *ctx, int kernel, float *ms, int *launches);
/* Kernels launched on this context since creation. */
int icpb_ctx_launch_count(icpb_ctx *ctx, long long *count);
/* Dense FP32 FMA micro-benchmark: the measured roofline denominator for the
 * compute-bound NN kernel. */
int icpb_measure_fp32_peak(icpb_ctx *ctx, int repeats, double *tflops, float *ms);

/* ---- point clouds: class icp::PointCloud, pointcloud.hpp:27-49 ---------- */
int icpb_cloud_create(icpb_ctx *ctx, int capacity, icpb_cloud **out);
int icpb_cloud_destroy(icpb_cloud *cloud);
int icpb_cloud_size(const icpb_cloud *cloud, int *n);
/* PointCloud(std::vector<cv::Point3f>) pointcloud.cpp:256-287 / direct point_list_t upload. */
int icpb_cloud_upload(icpb_cloud *cloud, const icpb_point *points, int n);
int icpb_cloud_upload_xyz(icpb_cloud *cloud, const float *xyz, int n);
int icpb_cloud_download(icpb_cloud *cloud, icpb_point *out, int capacity, int *n);
int icpb_cloud_copy(icpb_cloud *dst, const icpb_cloud *src);
/* Adopt n points already in device memory (16 B each), e.g. after an all-gather. */
int icpb_cloud_upload_device(icpb_cloud *cloud, const void *device_points, int n);
const void *icpb_cloud_device_ptr(const icpb_cloud *cloud);
/* z-slab exchange without host round trips: a "band" is one 16-byte header row (the point count) followed by
 * band_capacity point rows.  pack writes this cloud as a band into device memory (e.g. the send buffer of an
 * all-gather); assemble concatenates `world` bands (rank order) into this cloud. */
int icpb_cloud_pack_band_device(icpb_cloud *cloud, void *device_dst, int band_capacity);
int icpb_cloud_assemble_bands_device(icpb_cloud *cloud, const void *device_bands, int world, int band_capacity);
/* Copy the cloud's points into caller-owned DEVICE memory (e.g. a torch tensor feeding an all-gather). */
int icpb_cloud_download_device(icpb_cloud *cloud, void *device_dst, int capacity);
/* PointCloud(cv::Mat& data, cv::Mat colorMat), pointcloud.cpp:109-165 (and :11-58):
 * host depth (u16, h*w) and optional BGR (u8, h*w*3). */
int icpb_cloud_from_depth(icpb_cloud *cloud, const uint16_t *depth, const uint8_t *bgr, int w, int h,
                          const icpb_intrinsics *K, int rule, uint32_t rule_arg, uint32_t seed,
                          const uint8_t *keep_stream, int keep_stream_len);
/* Same with depth / bgr already resident on the device. */
int icpb_cloud_from_depth_device(icpb_cloud *cloud, const void *d_depth, const void *d_bgr, int w, int h,
                                 const icpb_intrinsics *K, int rule, uint32_t rule_arg, uint32_t seed,
                                 const void *d_keep_stream, int keep_stream_len);
/* Batched form of the constructor for throughput work (no host synchronisation): `frames` depth images (and
 * optional BGR images) back to back in device memory -> frame f's points at d_points + f*capacity_per_frame
 * (16 B each), its point count in d_counts[f].  Keeps every non-zero pixel (ICPB_SUB_NONE). */
int icpb_backproject_batch_device(icpb_ctx *ctx, const void *d_depth, const void *d_bgr, int frames, int w, int h,
                                  const icpb_intrinsics *K, void *d_points, int capacity_per_frame, int *d_counts);
/* PointCloud::rotate pointcloud.cpp:321-331 then PointCloud::translate :349-359 (either may be NULL). */
int icpb_cloud_transform(icpb_cloud *cloud, const float R[9], const float t[3]);
/* PointCloud::center (pointcloud.cpp:43-45,100-102), canonical FP64 block-ordered mean. */
int icpb_cloud_center(icpb_cloud *cloud, double center[3]);
void icpb_intrinsics_reference_v1(icpb_intrinsics *K); /* pointcloud.hpp:7-10 */
void icpb_intrinsics_reference_v2(icpb_intrinsics *K); /* SLAM.cpp:26-29 */

/* ---- image-space stages ------------------------------------------------- */
/* getNormalMap, SLAM.cpp:412-430: host u16 depth -> host float h*w*3. */
int icpb_normals_from_depth(icpb_ctx *ctx, const uint16_t *depth, int w, int h, float *normals);
/* The same for `frames` device-resident frames laid out back to back (u16 in, 3 floats per pixel out); no host
 * synchronisation. */
int icpb_normals_batch_device(icpb_ctx *ctx, const void *d_depth, int frames, int w, int h, void *d_normals);
/* filterDepthImage, SLAM.cpp:553-573 (range threshold + 5x5 close, anchor (3,3)). */
int icpb_depth_filter(icpb_ctx *ctx, const uint16_t *depth, int w, int h, int min_d, int max_d,
                      uint16_t *out);

/* ---- registration: namespace icp, icp.hpp:22-45 -------------------------- */
/* findGlobalNearestNeighborAssociations icp.cpp:541-563 + getNearestPoint :566-593 +
 * distance :606-620, un-compacted: idx[i] = lowest index of the nearest target,
 * dist[i] = its distance; the caller applies `dist < MAX_NN_COLOR_DISTANCE`.
 * idx / dist are HOST buffers of n entries (either may be NULL). */
int icpb_nn_search(icpb_ctx *ctx, const icpb_cloud *data, const icpb_cloud *target, int32_t *idx,
                   float *dist, int *exact_rescans);
/* Device-resident variant: results stay on the device (pointers owned by the ctx). */
int icpb_nn_search_device(icpb_ctx *ctx, const icpb_cloud *data, const icpb_cloud *target,
                          const int32_t **d_idx, const float **d_dist);
/* getTransformation icp.cpp:28-285 with the all-point association (icp.cpp:149/253):
 * the data cloud is transformed in place; the whole loop runs on the device. */
int icpb_icp_register(icpb_ctx *ctx, icpb_cloud *data, const icpb_cloud *target,
                      const icpb_icp_params *params, icpb_icp_result *result);
/* The same loop with a second cloud in tow: `carry` (e.g. the data cloud's key-points) receives every motion the data
 * cloud receives, in the same order and arithmetic -- what dataCloud.rotate / translate do to points AND key-points
 * (pointcloud.cpp:321-359) while only the points are associated (icp.cpp:149/253). */
int icpb_icp_register_carry(icpb_ctx *ctx, icpb_cloud *data, const icpb_cloud *target, icpb_cloud *carry,
                            const icpb_icp_params *params, icpb_icp_result *result);
/* `count` independent registrations (BASELINE config 4); results[i] for pair i. */
int icpb_icp_register_batch(icpb_ctx *ctx, icpb_cloud *const *data, const icpb_cloud *const *target,
                            int count, const icpb_icp_params *params, icpb_icp_result *results);
/* Non-blocking forms of icpb_icp_register / icpb_icp_register_batch.  The reference's getTransformation
 * (icp.cpp:28-285) blocks its caller for the whole loop; here the loop runs on the device without host involvement,
 * so the call returns as soon as the kernels and the read-back of the result block are enqueued on the context's
 * stream and the host may prepare the next frame (upload, back-projection: they queue behind it on the same stream)
 * or run another context.  One registration can be in flight per context; whatever registration call comes next on
 * the context first completes it.  The clouds must stay alive, and params' trace buffers valid, until the wait.
 * (With ICPB_NN_GRID / ICPB_NN_AUTO on large clouds the call still waits once, ~30 us, for the bounding box of
 * the targets that sizes the cell tables.)
 * icpb_icp_pending_wait blocks until the work is done, fills `results` (count entries, nullable) and frees the
 * handle; icpb_icp_pending_ready returns 1 in *ready when the wait would not block. */
int icpb_icp_register_async(icpb_ctx *ctx, icpb_cloud *data, const icpb_cloud *target,
                            const icpb_icp_params *params, icpb_pending **out);
int icpb_icp_register_batch_async(icpb_ctx *ctx, icpb_cloud *const *data, const icpb_cloud *const *target,
                                  int count, const icpb_icp_params *params, icpb_pending **out);
int icpb_icp_pending_ready(icpb_pending *pending, int *ready);
int icpb_icp_pending_wait(icpb_pending *pending, icpb_icp_result *results);
/* The loop as the reference runs it (icp.cpp:98,155-258): the data cloud's KEY-POINTS are
 * associated with the map cloud's key-points (findGlobalKeyPointAssociations, icp.cpp:488-539; pass
 * max_nn_distance = MAX_NN_KEYPOINT_DISTANCE 0.1, icp.hpp:10), `points` (nullable) follow every motion like
 * dataCloud.points do (pointcloud.cpp:321-359), and the key-points rejected by ANY pass accumulate, in order, in
 * `non_associations` (nullable; capacity >= (max_iterations+1) * key-points; icp.cpp:96,508) - the list
 * Map::update(assoc, errors, nonAssoc, delta) consumes (icpb_map_update_tracked, ICPB_TRACK_NONASSOC).
 * An empty map cloud (icp.cpp:490-491) or an empty key-point list: zero passes, nothing is touched. */
int icpb_icp_register_keypoints(icpb_ctx *ctx, icpb_cloud *keypoints, icpb_cloud *points,
                                const icpb_cloud *map_keypoints, const icpb_icp_params *params,
                                icpb_icp_result *result, icpb_cloud *non_associations);

/* ---- certainty map: class map::Map, map.hpp:20-37 ------------------------ */
/* Map::Map map.cpp:17-31; world[x][y][z] z-fastest (map.hpp:25).  The reference
 * macros are dims 300^3, cell 10/300 m (map.hpp:9-10,17); README.md:8-12 is
 * 300x300x250 at 0.02 m.  [z_lo, z_hi) is the z-slab this handle owns
 * (0, dims[2] for the whole map). */
int icpb_map_create(icpb_ctx *ctx, const int dims[3], float cell, int z_lo, int z_hi, icpb_map **out);
int icpb_map_destroy(icpb_map *map);
int icpb_map_clear(icpb_map *map);
/* Map::update overloads, map.cpp:88-119 / 122-151 / 220-269: saturating endpoint increments. */
int icpb_map_update_endpoints(icpb_map *map, const icpb_cloud *points, int rule, int delta, int max_conf);
/* The same updates WITH the lookup-table / mapCloud bookkeeping of the reference: the point whose hit first
 * satisfies the variant's condition on a voxel without a table entry is recorded in the table and appended to
 * `map_cloud` (in point order, like the push_back of the sequential loop).
 *   ICPB_TRACK_INIT     Map::update(PointCloud, delta, win) map.cpp:246-259: rule A; insert when the value AFTER
 *                       the hit is >= max_conf
 *   ICPB_TRACK_ASSOC    Map::update(assoc, delta) map.cpp:101-113: rule A; insert in the saturating branch
 *                       (value before the hit > 255 - delta)
 *   ICPB_TRACK_NONASSOC Map::update(assoc, errors, nonAssoc, delta) map.cpp:136-149: rule C; insert in the
 *                       promoting branch (value before the hit >= max_conf - delta)
 * At most 65536 points per call (the reference feeds key-points here).  Whole-map handles only. */
enum { ICPB_TRACK_INIT = 0, ICPB_TRACK_ASSOC = 1, ICPB_TRACK_NONASSOC = 2 };
int icpb_map_update_tracked(icpb_map *map, const icpb_cloud *points, int variant, int delta, int max_conf,
                            icpb_cloud *map_cloud, int *n_appended);
/* The same with the caller's numbering: the lookup-table entries written by this call are table_base, table_base + 1,
 * ... in append order (e.g. table_base = the length of the caller's own copy of the map cloud before the call), so
 * that icpb_map_table_entry leads back to the stored point (pointLookupTable, map.hpp:24). */
int icpb_map_update_tracked_base(icpb_map *map, const icpb_cloud *points, int variant, int delta, int max_conf,
                                 icpb_cloud *map_cloud, int table_base, int *n_appended);
/* Lookup-table entry of voxel v (-1 = empty, map.cpp:27). */
int icpb_map_table_entry(icpb_map *map, const int v[3], int *entry);
/* 1 when the voxel of p holds a lookup-table entry (pointLookupTable[..] != empty, map.hpp:24). */
int icpb_map_has_entry(icpb_map *map, const float p[3], int *has_entry);
/* Map::rayTrace map.cpp:272-439 (semantics in DESIGN.md "M4"): integer ray walk from the
 * origin voxel, decrements with clamp at 0, then rule-A endpoint increments. */
int icpb_map_integrate_rays(icpb_map *map, const icpb_cloud *points, const float origin[3],
                            int delta_dec, int delta_inc, long long *voxels_visited);
/* The same integration, and in addition the WORK the ray walk spends per z-layer is accumulated into layer_work (host,
 * dims[2] entries, added to): 1 per voxel step, 4 per brick jump, 6 per ray set-up -- their relative cost.  Equal shares of
 * the histogram are equal shares of the walk: icpb_slab_bounds_from_work turns it into z-slab boundaries.  Calibration
 * call (it synchronises and uses atomics per step); whole-map handles only. */
int icpb_map_integrate_rays_profiled(icpb_map *map, const icpb_cloud *points, const float origin[3], int delta_dec,
                                     int delta_inc, unsigned long long *layer_work);
/* bounds[0..world]: bounds[0] = 0, bounds[world] = layers, every slab at least one layer, equal shares of `work`. */
int icpb_slab_bounds_from_work(const unsigned long long *work, int layers, int world, int *bounds);
/* Map::getVoxelCoordinates map.cpp:55-85 (host-side scalar helper). */
/* Sync-free frame path (the z-slab map of SURVEY.md 8e, and frame sequences on one GPU).  A "band" is one 16-byte
 * header row (point count in its first word) followed by band_capacity point rows, in device memory.
 * icpb_frame_lift_band_device: image rows [row0, row1) of a device-resident depth frame -> world-space points of the
 * band (pointcloud.cpp:109-165 followed by rotate / translate :321-359), count into the header; nothing returns to
 * the host.  icpb_map_integrate_bands_device: `world` bands laid out back to back (one all-gather of the ranks' bands;
 * rank order = raster order) -> M4 ray decrements then rule-A endpoint increments on this handle's slab, every kernel
 * reading the point count from device memory. */
int icpb_frame_lift_band_device(icpb_ctx *ctx, const void *d_depth, int w, int h, int row0, int row1,
                                const icpb_intrinsics *K, const float R[9], const float t[3], void *d_band,
                                int band_capacity);
int icpb_map_integrate_bands_device(icpb_map *map, const void *d_bands, int world, int band_capacity,
                                    const float origin[3], int delta_dec, int delta_inc);
int icpb_map_voxel_coords(const icpb_map *map, const float p[3], int v[3]);
/* Slab download in the reference's linear order restricted to the slab:
 * out[(x*dimY + y)*(z_hi-z_lo) + (z - z_lo)]. */
int icpb_map_download(icpb_map *map, uint8_t *out, long long capacity);
int icpb_map_upload(icpb_map *map, const uint8_t *in, long long size);
int icpb_map_size_bytes(const icpb_map *map, long long *size);

/* ---- multi-GPU: one process (or host thread) per GPU, NCCL over NVLink ----------------------------
 * The reference is single-process (its map is a file-scope global, icp.cpp:26); these entry points are what a host
 * written against map.hpp needs to shard that map by z-slab, and to gather the results of registrations it has split
 * over the ranks.  NCCL is loaded at run time (dlopen of libnccl.so.2): single-GPU users need no NCCL.
 *
 * Bootstrap: rank 0 calls icpb_comm_unique_id and hands the 128 bytes to every rank by any means (a file, MPI, a
 * socket, torch.distributed); every rank then calls icpb_comm_create with the same bytes. */
#define ICPB_COMM_ID_BYTES 128
int icpb_comm_unique_id(uint8_t id[ICPB_COMM_ID_BYTES]);
int icpb_comm_create(icpb_ctx *ctx, int world, int rank, const uint8_t id[ICPB_COMM_ID_BYTES], icpb_comm **out);
int icpb_comm_destroy(icpb_comm *comm);
int icpb_comm_rank(const icpb_comm *comm, int *world, int *rank);
/* [lo, hi) of n_items owned by this rank: contiguous blocks, sizes differ by at most one (batches of registrations). */
int icpb_comm_shard_range(const icpb_comm *comm, long long n_items, long long *lo, long long *hi);
/* All-gather of `bytes` host bytes per rank (e.g. the poses of the rank's share of a batch), rank order.  Blocking. */
int icpb_comm_allgather_host(icpb_comm *comm, const void *send, void *recv, long long bytes);

/* The certainty map of map.hpp:20-37 sharded by z: rank g owns layers [bounds[g], bounds[g+1]) (bounds == NULL:
 * equal layer counts) of a dims / cell grid.  Every rank lifts its band of image rows, the bands are all-gathered
 * (the only exchange: <= 16 bytes per valid pixel per frame), and every rank walks every ray clipped to its slab
 * (Map::rayTrace, map.cpp:272-439, semantics of DESIGN.md M4) and applies the endpoints it owns.  comm == NULL: a
 * single rank owning the whole map, same code path without the collective. */
int icpb_slabmap_create(icpb_ctx *ctx, icpb_comm *comm, const int dims[3], float cell, const int *bounds, int w, int h,
                        icpb_slabmap **out);
int icpb_slabmap_destroy(icpb_slabmap *sm);
/* The rank's slab as an ordinary map handle (download / upload / clear); owned by the slab map. */
int icpb_slabmap_local(icpb_slabmap *sm, icpb_map **map, int *z_lo, int *z_hi);
/* `frames` device-resident w x h depth frames (u16, back to back) with their camera poses (R: frames x 9, t: frames x 3,
 * camera-to-world, host arrays) -> ray decrements + endpoint increments of every frame, in order.  Nothing returns to
 * the host: the call enqueues `frames_per_exchange` frames per all-gather, double buffered on three streams, so that
 * the lift and the exchange of group g+1 overlap the walks of group g; it returns when everything is enqueued.
 * Synchronise with icpb_ctx_sync (the slab map's streams are joined to the context stream at the end of the call). */
int icpb_slabmap_integrate_sequence_device(icpb_slabmap *sm, const void *d_depths, int frames, const icpb_intrinsics *K,
                                           const float *R, const float *t, int delta_dec, int delta_inc,
                                           int frames_per_exchange);

/* ---- pose reporting ------------------------------------
 * Scalar per-frame host arithmetic in the reference (SLAM.cpp:284-293); it stays host code here: no context,
 * no device.  Quaternions are {w, x, y, z}; rotation matrices are row-major 3x3 float; angles are degrees. */
int icpb_pose_quat_from_rotation(const float R[9], float q[4]);           /* Quaternion(cv::Mat), quaternion.cpp:23-79 */
int icpb_pose_quat_mul(const float a[4], const float b[4], float out[4]); /* operator*, quaternion.cpp:184-192 */
int icpb_pose_quat_inverse(const float q[4], float out[4]);               /* inverse(), quaternion.cpp:325-328 */
int icpb_pose_quat_to_euler_deg(const float q[4], float e[3]);            /* toEulerianAngle, SLAM.cpp:613-636 */
int icpb_pose_matrix_to_euler_deg(const float R[9], float e[3]);          /* transformationMatToEulerianAngle, SLAM.cpp:638-648 */

#ifdef __cplusplus
}
#endif
#endif /* ICPB200_H */

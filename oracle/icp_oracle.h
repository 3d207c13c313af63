/*
 * icp_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the per-frame registration + mapping hot path of
 * BenniG123/icp-slam-prototype.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (libicpb200.so) never links, includes or calls anything in oracle/.
 *
 * Parity pinning: the reference ships no golden vectors.  The
 * oracle is pinned (a) against the reference's own sources compiled by path
 * with an OpenCV type shim (oracle/refshim -> oracle/_ref, see
 * tests/test_oracle_vs_ref.py and tests/golden/), and (b) for the OpenCV
 * arithmetic that is not under /root/reference (3x3 gemm, invert,
 * determinant, SVD) against cv2 4.13 (tests/test_oracle_cv2.py).
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference).
 */
#ifndef ICP_ORACLE_H
#define ICP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* color_point_t, pointcloud.hpp:13-19: cv::Point3f (12 B) + cv::Vec3b (3 B) + 1 pad. */
typedef struct {
    float x, y, z;
    uint8_t c0, c1, c2, pad;
} orc_point;

/* Back-projection intrinsics.  The reference uses CX and FX for BOTH image
 * axes (pointcloud.cpp:38-39), so the "reference" preset passes
 * fx_u = fx_v = FX, cx_u = cx_v = CX. */
typedef struct {
    float fx_u, cx_u; /* x = (u - cx_u) * z / fx_u */
    float fx_v, cx_v; /* y = (v - cx_v) * z / fx_v */
    float depth_scale; /* 5000.0f, pointcloud.cpp:37 */
} orc_intrinsics;

enum { ORC_SUB_NONE = 0, ORC_SUB_STRIDE = 1, ORC_SUB_HASH = 2, ORC_SUB_STREAM = 3 };
enum { ORC_SOLVE_REFERENCE = 0, ORC_SOLVE_KABSCH = 1 };
enum { ORC_RULE_A = 0, ORC_RULE_C = 1 };

typedef struct {
    int max_iterations;    /* SLAM.cpp:277 passes 16; BASELINE configs use 20 */
    float threshold;       /* SLAM.cpp:277 passes 1e-4 */
    float max_nn_distance; /* MAX_NN_COLOR_DISTANCE 0.75f, icp.hpp:8 */
    int solve_mode;
    float last_translation[3]; /* icp.cpp:25 (used only by the <3 associations rule) */
    int n_threads;         /* OpenMP threads for the NN scan (1 = the reference) */
} orc_icp_params;

typedef struct {
    int iterations;     /* value of i when the loop ended (icp.cpp:152-258) */
    int nn_passes;      /* association passes executed */
    int n_assoc;        /* associations.size() after the last pass */
    float mse;          /* meanSquareError(errors) of the last pass */
    float rigid[16];    /* reference mode: icp.cpp:227-233,266-268 (row 3 = 0,0,0,1 here) */
    float cam_rotation[9]; /* reference mode: cameraRotation, icp.cpp:237 */
    float cam_position[3]; /* reference mode: cameraPosition, icp.cpp:246 */
    float offset[3];       /* last offset (icp.cpp:240) */
    double pose_R[9];   /* composed map data0 -> dataFinal: p_final ~= pose_R p + pose_t */
    double pose_t[3];
    int small_assoc_exit; /* 1 when the <3 associations branch ran (icp.cpp:163-182) */
} orc_icp_result;

/* ---- P1: pointcloud.cpp:11-58,109-165 ---- */
int orc_backproject(const uint16_t *depth, const uint8_t *bgr, int w, int h,
                    const orc_intrinsics *K, int rule, uint32_t rule_arg, uint32_t seed,
                    const uint8_t *keep_stream, orc_point *out, int *n_out,
                    double center_canon[3], float center_ref[3]);
uint32_t orc_hash32(uint32_t seed, uint32_t pixel);

/* ---- P2: pointcloud.cpp:321-331 (rotate), 349-359 (translate) ---- */
void orc_rotate(orc_point *pts, int n, const float R[9]);
void orc_translate(orc_point *pts, int n, const float t[3]);

/* ---- P3: SLAM.cpp:412-430 ---- */
void orc_normals(const uint16_t *depth, int w, int h, float *normals /* h*w*3 */);

/* ---- 8f-1: SLAM.cpp:553-573 ---- */
void orc_depth_filter(const uint16_t *in, int w, int h, int min_d, int max_d, uint16_t *out);

/* ---- N1-N3: icp.cpp:606-620, 566-593, 541-563 ---- */
float orc_distance(const orc_point *a, const orc_point *b);
void orc_nn(const orc_point *data, int n, const orc_point *target, int m,
            int32_t *idx, float *dist, int n_threads);

/* ---- canonical FP64 block-ordered reduction ---- */
void orc_canon_reduce(const double *terms, int n, int k, double *out /* k */);

/* ---- S1: 3x3 SVD (stands in for cv::SVD, icp.cpp:215) ---- */
void orc_svd3(const double A[9], double U[9], double w[3], double Vt[9]);
void orc_gemm33f(const float A[9], const float B[9], float C[9]);
int orc_inv33f(const float S[9], float D[9]);
double orc_det33f(const float m[9]);

/* ---- S1-S3 + loop: icp.cpp:28-285 with the all-point association (icp.cpp:149/253) ---- */
int orc_icp(orc_point *data /* in/out, n */, int n, const orc_point *target, int m,
            const orc_icp_params *prm, orc_icp_result *res,
            int32_t *idx_trace /* nullable, (max_iterations+1)*n */,
            float *dist_trace /* nullable, same shape */);

/* 8f-2: the same loop with the KEY-POINT association the reference runs (icp.cpp:98,255; :488-539): `keypoints`
 * (in/out) are associated with `map_keypoints`, `points` (in/out, nullable) only follow the motion, the rejected
 * key-points of every pass accumulate in `nonassoc` (capacity (max_iterations+1)*k, icp.cpp:508). */
int orc_icp_carry(orc_point *data, int n, orc_point *carry, int n_carry, const orc_point *target, int m,
                  const orc_icp_params *prm, orc_icp_result *res);
int orc_icp_keypoints(orc_point *keypoints, int k, orc_point *points, int n, const orc_point *map_keypoints, int mk,
                      const orc_icp_params *prm, orc_icp_result *res, orc_point *nonassoc, int *n_nonassoc);

/* ---- M1-M4: map.hpp:20-37, map.cpp:55-85, 88-151, 220-269, 272-439 ---- */
void orc_voxel_coords(const float p[3], float cell, const int dims[3], int v[3]);
void orc_map_update_endpoints(uint8_t *grid, const int dims[3], float cell,
                              const orc_point *pts, int n, int rule, int delta, int max_conf);
/* Map::update WITH the pointLookupTable / mapCloud bookkeeping, restated sequentially:
 * variant 0 = update(PointCloud, delta, win) map.cpp:220-269; 1 = update(assoc, delta) map.cpp:88-119;
 * 2 = update(assoc, errors, nonAssoc, delta) map.cpp:122-151.  `table` holds -1 (== empty, map.cpp:27) or the
 * index the point received in the map cloud; appended[] lists the inserted points' indices in insertion order. */
int orc_map_update_tracked(uint8_t *grid, int32_t *table, const int dims[3], float cell, const orc_point *pts, int n,
                           int variant, int delta, int max_conf, int map_cloud_size, int32_t *appended);
/* returns number of voxels visited (decrement candidates) */
long long orc_map_integrate_rays(uint8_t *grid, const int dims[3], float cell,
                                 const orc_point *pts, int n, const float origin[3],
                                 int delta_dec, int delta_inc, int z_lo, int z_hi);

/* ---- 8f-4 pose reporting: quaternion.cpp:23-79,184-192,294-341; SLAM.cpp:613-648 (q = {w, x, y, z}) ---- */
void orc_quat_from_rot(const float R[9], float q[4]);
void orc_quat_mul(const float a[4], const float b[4], float out[4]);
void orc_quat_inverse(const float q[4], float out[4]);
void orc_quat_to_euler_deg(const float q[4], float e[3]);
void orc_mat_to_euler_deg(const float R[9], float e[3]);

#ifdef __cplusplus
}
#endif
#endif

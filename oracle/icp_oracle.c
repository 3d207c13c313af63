/*
 * icp_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See icp_oracle.h for the rules on who may load this.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (the reference was built
 * /O2 /fp:precise, build/SLAM.exe.dir/RelWithDebInfo/SLAM.exe.tlog/CL.command.1.tlog).
 * FMA contraction must stay off: the canonical FP64 arithmetic below is
 * matched bit for bit by the CUDA kernels.
 */
#include "icp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* P1: depth image -> list of color_point_t                                   */
/* ------------------------------------------------------------------------- */

/* Deterministic stand-in for `rand() % SUBSAMPLE_FACTOR` (pointcloud.cpp:28):
 * murmur3 finaliser over seed and raster pixel index. */
uint32_t orc_hash32(uint32_t seed, uint32_t pixel)
{
    uint32_t h = seed ^ (pixel * 0x9E3779B9u);
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

/* pointcloud.cpp:19-58 / 115-155.  Raster order (y outer, x inner); skip
 * d == 0 (:22); subsample decision consumed once per NON-ZERO pixel (:28);
 * z = float(d)/5000.0f; x = (u - CX) * z / FX; y = (v - CX) * z / FX, all
 * float, left to right (:37-39); colour = colorMat(v,u) (:47).
 * center_ref is the reference's float running sum / count (:43-45,100-102);
 * center_canon is the canonical FP64 block-ordered mean used for parity. */
int orc_backproject(const uint16_t *depth, const uint8_t *bgr, int w, int h,
                    const orc_intrinsics *K, int rule, uint32_t rule_arg, uint32_t seed,
                    const uint8_t *keep_stream, orc_point *out, int *n_out,
                    double center_canon[3], float center_ref[3])
{
    int count = 0;
    uint32_t ordinal = 0; /* index among non-zero pixels */
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (rule_arg == 0) rule_arg = 1;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            uint16_t d = depth[(size_t)y * w + x];
            if (d == 0) continue;
            uint32_t ord = ordinal++;
            int keep;
            switch (rule) {
            case ORC_SUB_NONE: keep = 1; break;
            case ORC_SUB_STRIDE: keep = (ord % rule_arg) == 0; break;
            case ORC_SUB_HASH: keep = (orc_hash32(seed, (uint32_t)(y * w + x)) % rule_arg) == 0; break;
            case ORC_SUB_STREAM: keep = keep_stream[ord] != 0; break;
            default: return -1;
            }
            if (!keep) continue;
            float p_z = ((float)d) / K->depth_scale;
            float p_x = ((float)x - K->cx_u) * p_z / K->fx_u;
            float p_y = ((float)y - K->cx_v) * p_z / K->fx_v;
            cx += p_x; cy += p_y; cz += p_z;
            orc_point p;
            p.x = p_x; p.y = p_y; p.z = p_z;
            if (bgr) {
                const uint8_t *c = bgr + ((size_t)y * w + x) * 3;
                p.c0 = c[0]; p.c1 = c[1]; p.c2 = c[2];
            } else {
                p.c0 = p.c1 = p.c2 = 0;
            }
            p.pad = 0;
            out[count++] = p;
        }
    }
    *n_out = count;
    if (center_ref) {
        /* 0/0 = NaN when no point survives, as in the reference (:100-102) */
        center_ref[0] = cx / count; center_ref[1] = cy / count; center_ref[2] = cz / count;
    }
    if (center_canon) {
        double *terms = (double *)malloc(sizeof(double) * 3 * (size_t)(count > 0 ? count : 1));
        for (int i = 0; i < count; i++) {
            terms[3 * i + 0] = out[i].x; terms[3 * i + 1] = out[i].y; terms[3 * i + 2] = out[i].z;
        }
        double s[3];
        orc_canon_reduce(terms, count, 3, s);
        for (int k = 0; k < 3; k++) center_canon[k] = s[k] / (double)count;
        free(terms);
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* P2: rotate / translate                                                      */
/* ------------------------------------------------------------------------- */

/* pointcloud.cpp:321-331: points <- (R * M^T)^T through cv::Mat operator*.
 * For a 3x3 by 3xN CV_32F product OpenCV's gemm computes each element as
 * fl(fl(fl(r0*x) + fl(r1*y)) + fl(r2*z)) in float without FMA (pinned
 * against cv2.gemm in tests/test_oracle_cv2.py).  Rotation is about the world
 * origin; `center` is not rotated. */
void orc_rotate(orc_point *pts, int n, const float R[9])
{
    for (int i = 0; i < n; i++) {
        float x = pts[i].x, y = pts[i].y, z = pts[i].z;
        float nx = (R[0] * x + R[1] * y) + R[2] * z;
        float ny = (R[3] * x + R[4] * y) + R[5] * z;
        float nz = (R[6] * x + R[7] * y) + R[8] * z;
        pts[i].x = nx; pts[i].y = ny; pts[i].z = nz;
    }
}

/* pointcloud.cpp:349-359: p += offset in float. */
void orc_translate(orc_point *pts, int n, const float t[3])
{
    for (int i = 0; i < n; i++) {
        pts[i].x += t[0]; pts[i].y += t[1]; pts[i].z += t[2];
    }
}

/* ------------------------------------------------------------------------- */
/* P3: normals from raw depth                                                  */
/* ------------------------------------------------------------------------- */

/* SLAM.cpp:412-430.  dzdx = (I(r+1,c) - I(r-1,c)) / 2, dzdy = (I(r,c+1) -
 * I(r,c-1)) / 2 on raw depth units converted to float; n = normalize(-dzdx,
 * -dzdy, 1) where cv::normalize(Vec3f) is v * (1.0 / sqrt(double sum of
 * squares)) evaluated in double and rounded to float.  The reference never
 * writes row 0 / col 0 and reads one past the end at the last row / col
 * (:416-422); the defined border here is zeros on r in {0,h-1}, c in {0,w-1}. */
void orc_normals(const uint16_t *depth, int w, int h, float *normals)
{
    memset(normals, 0, sizeof(float) * 3 * (size_t)w * h);
    for (int r = 1; r < h - 1; r++) {
        for (int c = 1; c < w - 1; c++) {
            float up = (float)depth[(size_t)(r - 1) * w + c];
            float dn = (float)depth[(size_t)(r + 1) * w + c];
            float lf = (float)depth[(size_t)r * w + c - 1];
            float rt = (float)depth[(size_t)r * w + c + 1];
            float dzdx = (dn - up) / 2.0f;
            float dzdy = (rt - lf) / 2.0f;
            float v0 = -dzdx, v1 = -dzdy, v2 = 1.0f;
            double nv = sqrt((double)v0 * v0 + (double)v1 * v1 + (double)v2 * v2);
            double inv = 1.0 / nv;
            float *o = normals + ((size_t)r * w + c) * 3;
            o[0] = (float)((double)v0 * inv);
            o[1] = (float)((double)v1 * inv);
            o[2] = (float)((double)v2 * inv);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* 8f-1: depth range filter + 5x5 close with anchor (3,3)                     */
/* ------------------------------------------------------------------------- */

/* SLAM.cpp:553-573.  Values > max or < min become 0 (:559-565); then
 * cv::dilate and cv::erode with a 5x5 MORPH_RECT element anchored at (3,3)
 * (:567-573): the window covers offsets [-3,+1] on both axes; out-of-image
 * pixels do not take part (OpenCV's default morphology border; pinned
 * against cv2 in tests/test_oracle_cv2.py). */
void orc_depth_filter(const uint16_t *in, int w, int h, int min_d, int max_d, uint16_t *out)
{
    size_t npx = (size_t)w * h;
    uint16_t *a = (uint16_t *)malloc(npx * sizeof(uint16_t));
    uint16_t *b = (uint16_t *)malloc(npx * sizeof(uint16_t));
    for (size_t i = 0; i < npx; i++) {
        int v = in[i];
        a[i] = (v > max_d || v < min_d) ? 0 : (uint16_t)v;
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int best = 0;
            for (int dy = -3; dy <= 1; dy++) {
                int yy = y + dy;
                if (yy < 0 || yy >= h) continue;
                for (int dx = -3; dx <= 1; dx++) {
                    int xx = x + dx;
                    if (xx < 0 || xx >= w) continue;
                    int v = a[(size_t)yy * w + xx];
                    if (v > best) best = v;
                }
            }
            b[(size_t)y * w + x] = (uint16_t)best;
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int best = 65535;
            for (int dy = -3; dy <= 1; dy++) {
                int yy = y + dy;
                if (yy < 0 || yy >= h) continue;
                for (int dx = -3; dx <= 1; dx++) {
                    int xx = x + dx;
                    if (xx < 0 || xx >= w) continue;
                    int v = b[(size_t)yy * w + xx];
                    if (v < best) best = v;
                }
            }
            out[(size_t)y * w + x] = (uint16_t)best;
        }
    free(a);
    free(b);
}

/* ------------------------------------------------------------------------- */
/* N1-N3: brute-force nearest neighbour                                        */
/* ------------------------------------------------------------------------- */

/* icp.cpp:606-620 with icp.hpp:6-7 expanded (COLOR_WEIGHT 0.0f): float
 * differences (:607-609); pow(float,2) promotes to double so each square is
 * exact in double, summed left to right in double and rounded ONCE to float
 * on assignment (:611); the return expression reduces to sqrt(xyz)
 * (:619), the correctly rounded float square root. */
float orc_distance(const orc_point *a, const orc_point *b)
{
    float x = a->x - b->x;
    float y = a->y - b->y;
    float z = a->z - b->z;
    float xyz = (float)((double)x * (double)x + (double)y * (double)y + (double)z * (double)z);
    return sqrtf(xyz);
}

/* icp.cpp:566-593 per query (best starts at target[0]; strict `<` at :578,
 * so the lowest index wins ties judged on the sqrt-ed float distance), and
 * icp.cpp:541-563 over all queries in order.  Output is the un-compacted
 * (idx, dist) per query; the `dist < MAX_NN_COLOR_DISTANCE` filter (:553)
 * is applied by the caller.  Queries are independent, so OpenMP over
 * queries does not change any result. */
void orc_nn(const orc_point *data, int n, const orc_point *target, int m,
            int32_t *idx, float *dist, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads)
#endif
    for (int i = 0; i < n; i++) {
        const orc_point q = data[i];
        int best_j = 0;
        float best = orc_distance(&q, &target[0]);
        for (int j = 1; j < m; j++) {
            float d = orc_distance(&q, &target[j]);
            if (d < best) {
                best = d;
                best_j = j;
            }
        }
        idx[i] = best_j;
        dist[i] = best;
    }
}

/* ------------------------------------------------------------------------- */
/* Canonical FP64 block-ordered reduction                                      */
/* ------------------------------------------------------------------------- */

/* The one summation order shared with the CUDA path (DESIGN.md "CANON-3").
 * reduce256: 256 values = 8 groups of 32; each group is folded with offsets
 * 16, 8, 4, 2, 1 (v[l] += v[l+off] for l < off); the 8 group sums are then
 * added left to right.  Level 1 folds every chunk of 256 consecutive terms
 * (zero padded).  Level 2 gives slot t in [0,256) the chunk sums t, t+256,
 * t+512, ... added in that order starting from 0.0, then folds the 256
 * slots with reduce256. */
static double reduce256(double *v)
{
    double g[8];
    for (int wv = 0; wv < 8; wv++) {
        double *x = v + 32 * wv;
        for (int off = 16; off >= 1; off >>= 1)
            for (int l = 0; l < off; l++) x[l] = x[l] + x[l + off];
        g[wv] = x[0];
    }
    double s = g[0];
    for (int wv = 1; wv < 8; wv++) s = s + g[wv];
    return s;
}

void orc_canon_reduce(const double *terms, int n, int k, double *out)
{
    int chunks = (n + 255) / 256;
    double *partial = (double *)malloc(sizeof(double) * (size_t)(chunks > 0 ? chunks : 1));
    double v[256];
    for (int q = 0; q < k; q++) {
        for (int c = 0; c < chunks; c++) {
            for (int l = 0; l < 256; l++) {
                int i = c * 256 + l;
                v[l] = (i < n) ? terms[(size_t)i * k + q] : 0.0;
            }
            partial[c] = reduce256(v);
        }
        for (int t = 0; t < 256; t++) {
            double acc = 0.0;
            for (int c = t; c < chunks; c += 256) acc = acc + partial[c];
            v[t] = acc;
        }
        out[q] = reduce256(v);
    }
    free(partial);
}

/* ------------------------------------------------------------------------- */
/* S1: small dense helpers standing in for OpenCV 3.2 calls                    */
/* ------------------------------------------------------------------------- */

/* cv::Mat operator* on two 3x3 CV_32F (icp.cpp:218,231,237): float, no FMA,
 * left to right (pinned against cv2.gemm). */
void orc_gemm33f(const float A[9], const float B[9], float C[9])
{
    float T[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            T[3 * i + j] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j]) + A[3 * i + 2] * B[6 + j];
    memcpy(C, T, sizeof(T));
}

/* cv::determinant on 3x3 CV_32F (icp.cpp:220): evaluated in double
 * (pinned against cv2.determinant). */
double orc_det33f(const float m[9])
{
    double m00 = m[0], m01 = m[1], m02 = m[2];
    double m10 = m[3], m11 = m[4], m12 = m[5];
    double m20 = m[6], m21 = m[7], m22 = m[8];
    return m00 * (m11 * m22 - m12 * m21) - m01 * (m10 * m22 - m12 * m20) +
           m02 * (m10 * m21 - m11 * m20);
}

/* Mat::inv() on 3x3 CV_32F (icp.cpp:235): closed-form adjugate in double
 * times 1/det, rounded to float (pinned against cv2.invert). */
int orc_inv33f(const float Sf[9], float D[9])
{
    double d = orc_det33f(Sf);
    if (d == 0.0) { memset(D, 0, 9 * sizeof(float)); return 0; }
    d = 1.0 / d;
    double S[9];
    for (int i = 0; i < 9; i++) S[i] = Sf[i];
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
    memcpy(D, T, sizeof(T));
    return 1;
}

/* One-sided (Hestenes) Jacobi SVD of a 3x3 in double: A = U diag(w) Vt, w
 * descending.  Stands in for cv::SVD (icp.cpp:215), whose JacobiSVD
 * internals are third-party (OpenCV 3.2, not under /root/reference).  Only
 * + - * / sqrt and comparisons are used so that the CUDA twin is bit
 * identical.  Column pair order (0,1),(0,2),(1,2); at most 30 sweeps. */
void orc_svd3(const double Ain[9], double U[9], double w[3], double Vt[9])
{
    double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(A, Ain, sizeof(A));
    static const int P[3] = {0, 0, 1}, Q[3] = {1, 2, 2};
    const double eps = 2.220446049250313e-16;
    for (int sweep = 0; sweep < 30; sweep++) {
        int changed = 0;
        for (int k = 0; k < 3; k++) {
            int p = P[k], q = Q[k];
            double alpha = (A[p] * A[p] + A[3 + p] * A[3 + p]) + A[6 + p] * A[6 + p];
            double beta = (A[q] * A[q] + A[3 + q] * A[3 + q]) + A[6 + q] * A[6 + q];
            double gamma = (A[p] * A[q] + A[3 + p] * A[3 + q]) + A[6 + p] * A[6 + q];
            if (fabs(gamma) <= eps * sqrt(alpha * beta)) continue;
            changed = 1;
            double zeta = (beta - alpha) / (2.0 * gamma);
            double az = fabs(zeta);
            double t = 1.0 / (az + sqrt(1.0 + zeta * zeta));
            if (zeta < 0.0) t = -t;
            double c = 1.0 / sqrt(1.0 + t * t);
            double s = c * t;
            for (int r = 0; r < 3; r++) {
                double a = A[3 * r + p], b = A[3 * r + q];
                A[3 * r + p] = c * a - s * b;
                A[3 * r + q] = s * a + c * b;
                double va = V[3 * r + p], vb = V[3 * r + q];
                V[3 * r + p] = c * va - s * vb;
                V[3 * r + q] = s * va + c * vb;
            }
        }
        if (!changed) break;
    }
    for (int k = 0; k < 3; k++)
        w[k] = sqrt((A[k] * A[k] + A[3 + k] * A[3 + k]) + A[6 + k] * A[6 + k]);
    /* selection sort, descending, stable for equal values */
    for (int i = 0; i < 2; i++) {
        int best = i;
        for (int j = i + 1; j < 3; j++)
            if (w[j] > w[best]) best = j;
        if (best != i) {
            double tw = w[i]; w[i] = w[best]; w[best] = tw;
            for (int r = 0; r < 3; r++) {
                double ta = A[3 * r + i]; A[3 * r + i] = A[3 * r + best]; A[3 * r + best] = ta;
                double tv = V[3 * r + i]; V[3 * r + i] = V[3 * r + best]; V[3 * r + best] = tv;
            }
        }
    }
    for (int k = 0; k < 3; k++) {
        if (w[k] > 0.0) {
            for (int r = 0; r < 3; r++) U[3 * r + k] = A[3 * r + k] / w[k];
        } else {
            for (int r = 0; r < 3; r++) U[3 * r + k] = 0.0;
        }
    }
    /* rank deficiency: complete U with cross products (degenerate input only) */
    if (!(w[2] > 0.0) && w[1] > 0.0) {
        U[2] = U[3] * U[7] - U[6] * U[4];
        U[5] = U[6] * U[1] - U[0] * U[7];
        U[8] = U[0] * U[4] - U[3] * U[1];
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Vt[3 * i + j] = V[3 * j + i];
}

/* ------------------------------------------------------------------------- */
/* S1-S3 and the registration loop                                             */
/* ------------------------------------------------------------------------- */

#define ORC_NTERMS 20 /* a(3) b(3) b*a^T(9) d(1) a-b(3) count(1) */

static void assoc_sums(const orc_point *data, int n, const orc_point *target,
                       const int32_t *idx, const float *dist, float max_d, double *terms,
                       double sums[ORC_NTERMS])
{
    for (int i = 0; i < n; i++) {
        double *t = terms + (size_t)i * ORC_NTERMS;
        if (dist[i] < max_d) { /* icp.cpp:553 */
            const orc_point *a = &data[i];
            const orc_point *b = &target[idx[i]];
            t[0] = a->x; t[1] = a->y; t[2] = a->z;
            t[3] = b->x; t[4] = b->y; t[5] = b->z;
            const float av[3] = {a->x, a->y, a->z};
            const float bv[3] = {b->x, b->y, b->z};
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) t[6 + 3 * r + c] = (double)bv[r] * (double)av[c];
            t[15] = dist[i];
            /* calculateOffset, icp.cpp:325-331: float a - b per component */
            t[16] = (double)(float)(a->x - b->x);
            t[17] = (double)(float)(a->y - b->y);
            t[18] = (double)(float)(a->z - b->z);
            t[19] = 1.0;
        } else {
            for (int k = 0; k < ORC_NTERMS; k++) t[k] = 0.0;
        }
    }
    orc_canon_reduce(terms, n, ORC_NTERMS, sums);
}

/* meanSquareError, icp.cpp:622-638: (sum(errors)/n)^2, i.e. the square of the
 * MEAN distance.  The sum is the canonical FP64 one; mean and square are
 * rounded to float as the reference's float `error_sum` is. */
static float mse_from_sums(const double sums[ORC_NTERMS])
{
    if (!(sums[19] > 0.0)) return 0.f;
    float e = (float)(sums[15] / sums[19]);
    return (float)((double)e * (double)e);
}

/* Rejected queries of one association pass, in query order (icp.cpp:507-509: nonAssociations.push_back, never
 * cleared between the passes of one getTransformation call). */
static void append_rejects(const orc_point *data, int n, const float *dist, float max_nn, orc_point *out, int *n_out)
{
    if (!out) return;
    for (int q = 0; q < n; q++)
        if (!(dist[q] < max_nn)) out[(*n_out)++] = data[q];
}

/* The loop of icp.cpp:98-258.  `data` are the associated points (all points: :149/:253; key-points: :98/:255);
 * `carry` (nullable) are points that only follow the motion (dataCloud.rotate / translate move points AND
 * key-points, pointcloud.cpp:321-359); `nonassoc` (nullable, capacity (max_iterations+1)*n) collects the rejects. */
static int icp_core(orc_point *data, int n, orc_point *carry, int n_carry, const orc_point *target, int m,
                    const orc_icp_params *prm, orc_icp_result *res, int32_t *idx_trace, float *dist_trace,
                    orc_point *nonassoc, int *n_nonassoc)
{
    if (n <= 0 || m <= 0) return -1;
    if (n_nonassoc) *n_nonassoc = 0;
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    float *dist = (float *)malloc(sizeof(float) * (size_t)n);
    double *terms = (double *)malloc(sizeof(double) * ORC_NTERMS * (size_t)n);
    double sums[ORC_NTERMS];
    memset(res, 0, sizeof(*res));

    float rigid[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float camR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float camP[3] = {0, 0, 0};
    float offset[3] = {0, 0, 0};
    double PR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Pt[3] = {0, 0, 0};
    int passes = 0;

    /* icp.cpp:149 (all-point variant of :98) */
    orc_nn(data, n, target, m, idx, dist, prm->n_threads);
    if (idx_trace) memcpy(idx_trace, idx, sizeof(int32_t) * (size_t)n);
    if (dist_trace) memcpy(dist_trace, dist, sizeof(float) * (size_t)n);
    passes++;
    assoc_sums(data, n, target, idx, dist, prm->max_nn_distance, terms, sums);
    append_rejects(data, n, dist, prm->max_nn_distance, nonassoc, n_nonassoc);

    int i = 0;
    /* icp.cpp:155 */
    while (mse_from_sums(sums) > prm->threshold && i < prm->max_iterations) {
        int n_assoc = (int)sums[19];
        if (n_assoc < 3) {
            /* icp.cpp:163-182: replay the last motion (lastRotation is the
             * identity: the `R` assigned at :261 is the outer, shadowed one) */
            i = prm->max_iterations;
            offset[0] = -prm->last_translation[0];
            offset[1] = -prm->last_translation[1];
            offset[2] = -prm->last_translation[2];
            orc_translate(data, n, prm->last_translation);
            if (carry) orc_translate(carry, n_carry, prm->last_translation);
            for (int k = 0; k < 3; k++) Pt[k] += (double)prm->last_translation[k];
            res->small_assoc_exit = 1;
            break;
        }
        float Rf[9], tf[3];
        if (prm->solve_mode == ORC_SOLVE_REFERENCE) {
            /* icp.cpp:212: M = previousMat.t() * dataMat = sum b a^T, UNcentred */
            double U[9], w[3], Vt[9], Rd[9];
            orc_svd3(&sums[6], U, w, Vt);
            /* icp.cpp:218: R = vt.t() * u.t()  ->  R[r][c] = sum_k Vt[k][r] U[c][k] */
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++)
                    Rd[3 * r + c] = (Vt[r] * U[3 * c] + Vt[3 + r] * U[3 * c + 1]) + Vt[6 + r] * U[3 * c + 2];
            float R[9];
            for (int k = 0; k < 9; k++) R[k] = (float)Rd[k];
            /* icp.cpp:220-223 */
            if (orc_det33f(R) < 0) { R[2] *= -1; R[5] *= -1; R[8] *= -1; }
            /* icp.cpp:227-233 */
            if (i == 0) memcpy(rigid, R, sizeof(rigid));
            else orc_gemm33f(R, rigid, rigid);
            /* icp.cpp:235-237 */
            orc_inv33f(R, Rf);
            orc_rotate(data, n, Rf);
            if (carry) orc_rotate(carry, n_carry, Rf);
            orc_gemm33f(camR, Rf, camR);
            /* icp.cpp:240 / 314-344: mean of (a - b) over the associations,
             * which hold pre-rotation copies of a */
            for (int k = 0; k < 3; k++) offset[k] = (float)(sums[16 + k] / sums[19]);
            /* icp.cpp:245-246 */
            tf[0] = -offset[0]; tf[1] = -offset[1]; tf[2] = -offset[2];
            orc_translate(data, n, tf);
            if (carry) orc_translate(carry, n_carry, tf);
            camP[0] -= offset[0]; camP[1] -= offset[1]; camP[2] -= offset[2];
        } else {
            /* rigid_transform_3D.py:14-36 with A = data, B = matches */
            double cnt = sums[19];
            double cA[3] = {sums[0] / cnt, sums[1] / cnt, sums[2] / cnt};
            double cB[3] = {sums[3] / cnt, sums[4] / cnt, sums[5] / cnt};
            double H[9];
            /* H = sum (a-cA)(b-cB)^T = sum a b^T - n cA cB^T; sums[6+3r+c] = sum b_r a_c */
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) H[3 * r + c] = sums[6 + 3 * c + r] - (cnt * cA[r]) * cB[c];
            double U[9], w[3], Vt[9], Rd[9];
            orc_svd3(H, U, w, Vt);
            for (int pass = 0; pass < 2; pass++) {
                /* R = Vt.T * U.T */
                for (int r = 0; r < 3; r++)
                    for (int c = 0; c < 3; c++)
                        Rd[3 * r + c] = (Vt[r] * U[3 * c] + Vt[3 + r] * U[3 * c + 1]) + Vt[6 + r] * U[3 * c + 2];
                double det = Rd[0] * (Rd[4] * Rd[8] - Rd[5] * Rd[7]) - Rd[1] * (Rd[3] * Rd[8] - Rd[5] * Rd[6]) +
                             Rd[2] * (Rd[3] * Rd[7] - Rd[4] * Rd[6]);
                if (pass == 0 && det < 0) { Vt[6] *= -1; Vt[7] *= -1; Vt[8] *= -1; } /* :32-35 */
                else break;
            }
            for (int k = 0; k < 9; k++) Rf[k] = (float)Rd[k];
            /* t = -R cA + cB (:37) */
            for (int r = 0; r < 3; r++)
                tf[r] = (float)(cB[r] - ((Rd[3 * r] * cA[0] + Rd[3 * r + 1] * cA[1]) + Rd[3 * r + 2] * cA[2]));
            orc_rotate(data, n, Rf);
            orc_translate(data, n, tf);
            if (carry) { orc_rotate(carry, n_carry, Rf); orc_translate(carry, n_carry, tf); }
            offset[0] = -tf[0]; offset[1] = -tf[1]; offset[2] = -tf[2];
        }
        /* composed pose in double from the float R, t actually applied */
        {
            double NR[9], Nt[3];
            for (int r = 0; r < 3; r++) {
                for (int c = 0; c < 3; c++)
                    NR[3 * r + c] = ((double)Rf[3 * r] * PR[c] + (double)Rf[3 * r + 1] * PR[3 + c]) +
                                    (double)Rf[3 * r + 2] * PR[6 + c];
                Nt[r] = (((double)Rf[3 * r] * Pt[0] + (double)Rf[3 * r + 1] * Pt[1]) +
                         (double)Rf[3 * r + 2] * Pt[2]) + (double)tf[r];
            }
            memcpy(PR, NR, sizeof(PR));
            memcpy(Pt, Nt, sizeof(Pt));
        }
        /* icp.cpp:253 */
        orc_nn(data, n, target, m, idx, dist, prm->n_threads);
        passes++;
        if (idx_trace) memcpy(idx_trace + (size_t)passes * n - n, idx, sizeof(int32_t) * (size_t)n);
        if (dist_trace) memcpy(dist_trace + (size_t)passes * n - n, dist, sizeof(float) * (size_t)n);
        assoc_sums(data, n, target, idx, dist, prm->max_nn_distance, terms, sums);
        append_rejects(data, n, dist, prm->max_nn_distance, nonassoc, n_nonassoc);
        i++;
    }

    res->iterations = i;
    res->nn_passes = passes;
    res->n_assoc = (int)sums[19];
    res->mse = mse_from_sums(sums);
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) res->rigid[4 * r + c] = rigid[3 * r + c];
        res->rigid[4 * r + 3] = offset[r]; /* icp.cpp:266-268 */
    }
    res->rigid[12] = res->rigid[13] = res->rigid[14] = 0.f;
    res->rigid[15] = 1.f; /* row 3 is never written by the reference (icp.cpp:29) */
    memcpy(res->cam_rotation, camR, sizeof(camR));
    memcpy(res->cam_position, camP, sizeof(camP));
    memcpy(res->offset, offset, sizeof(offset));
    memcpy(res->pose_R, PR, sizeof(PR));
    memcpy(res->pose_t, Pt, sizeof(Pt));
    free(idx); free(dist); free(terms);
    return 0;
}

int orc_icp(orc_point *data, int n, const orc_point *target, int m,
            const orc_icp_params *prm, orc_icp_result *res, int32_t *idx_trace,
            float *dist_trace)
{
    return icp_core(data, n, NULL, 0, target, m, prm, res, idx_trace, dist_trace, NULL, NULL);
}

/* The all-point loop with a second cloud in tow: `carry` follows every motion of `data` (dataCloud.rotate /
 * translate move points and key-points alike, pointcloud.cpp:321-359). */
int orc_icp_carry(orc_point *data, int n, orc_point *carry, int n_carry, const orc_point *target, int m,
                  const orc_icp_params *prm, orc_icp_result *res)
{
    return icp_core(data, n, n_carry > 0 ? carry : NULL, n_carry, target, m, prm, res, NULL, NULL, NULL, NULL);
}

/* 8f-2, the loop as the reference runs it (icp.cpp:98,155-258): the data cloud's KEY-POINTS are associated with
 * the map cloud's key-points (findGlobalKeyPointAssociations :488-515, prm->max_nn_distance = 0.1 m, icp.hpp:10);
 * the cloud's points follow every motion; the rejected key-points of ALL passes accumulate in `nonassoc`
 * (:508, :96).  An empty map (:490-491) or an empty key-point list leaves everything untouched. */
int orc_icp_keypoints(orc_point *keypoints, int k, orc_point *points, int n, const orc_point *map_keypoints, int mk,
                      const orc_icp_params *prm, orc_icp_result *res, orc_point *nonassoc, int *n_nonassoc)
{
    *n_nonassoc = 0;
    if (k <= 0 || mk <= 0) {
        memset(res, 0, sizeof(*res));
        for (int d = 0; d < 3; d++) { res->rigid[5 * d] = 1.f; res->cam_rotation[4 * d] = 1.f; res->pose_R[4 * d] = 1.0; }
        res->rigid[15] = 1.f;
        return 0;
    }
    return icp_core(keypoints, k, points, n, map_keypoints, mk, prm, res, NULL, NULL, nonassoc, n_nonassoc);
}

/* ------------------------------------------------------------------------- */
/* M1-M4: certainty grid                                                       */
/* ------------------------------------------------------------------------- */

/* map.cpp:55-85: per axis int(p / c) with float true division and truncation
 * toward zero, then clamp into [0, dim-1] (out-of-range points are clamped
 * in, not rejected). */
void orc_voxel_coords(const float p[3], float cell, const int dims[3], int v[3])
{
    for (int k = 0; k < 3; k++) {
        int q = (int)(p[k] / cell);
        if (q < 0) q = 0;
        if (q >= dims[k]) q = dims[k] - 1;
        v[k] = q;
    }
}

/* Rule A, map.cpp:249-253 (also :104-113): c > 255-delta ? 255 : c + delta.
 * Rule C, map.cpp:139-149: c >= max_conf-delta ? 255 : c + delta.
 * Grid layout is world[x][y][z], z fastest (map.hpp:25).  Points are applied
 * in order; both rules are pure functions of c so any order gives the same
 * grid. */
void orc_map_update_endpoints(uint8_t *grid, const int dims[3], float cell,
                              const orc_point *pts, int n, int rule, int delta, int max_conf)
{
    for (int i = 0; i < n; i++) {
        float p[3] = {pts[i].x, pts[i].y, pts[i].z};
        int v[3];
        orc_voxel_coords(p, cell, dims, v);
        uint8_t *c = &grid[((size_t)v[0] * dims[1] + v[1]) * dims[2] + v[2]];
        if (rule == ORC_RULE_A) {
            if (*c > 255 - delta) *c = 255;
            else *c = (uint8_t)(*c + delta);
        } else {
            if (*c >= max_conf - delta) *c = 255;
            else *c = (uint8_t)(*c + delta);
        }
    }
}

int orc_map_update_tracked(uint8_t *grid, int32_t *table, const int dims[3], float cell, const orc_point *pts, int n,
                           int variant, int delta, int max_conf, int map_cloud_size, int32_t *appended)
{
    int n_app = 0;
    for (int i = 0; i < n; i++) {
        float p[3] = {pts[i].x, pts[i].y, pts[i].z};
        int v[3];
        orc_voxel_coords(p, cell, dims, v);
        size_t lin = ((size_t)v[0] * dims[1] + v[1]) * dims[2] + v[2];
        uint8_t *c = &grid[lin];
        int insert = 0;
        if (variant == 0) {                       /* map.cpp:249-259 */
            if (*c > 255 - delta) *c = 255;
            else *c = (uint8_t)(*c + delta);
            insert = (table[lin] < 0) && (*c >= max_conf);
        } else if (variant == 1) {                /* map.cpp:104-113 */
            if (*c > 255 - delta) { *c = 255; insert = table[lin] < 0; }
            else *c = (uint8_t)(*c + delta);
        } else {                                  /* map.cpp:139-149 */
            if (*c >= max_conf - delta) { *c = 255; insert = table[lin] < 0; }
            else *c = (uint8_t)(*c + delta);
        }
        if (insert) {
            table[lin] = map_cloud_size + n_app;
            appended[n_app++] = i;
        }
    }
    return n_app;
}

/* Ray integration.  Origin of the semantics: Map::rayTrace, map.cpp:272-439
 * (Amanatides-Woo, dead and buggy in the reference: call sites commented at
 * :99,:231; "TODO - This is wrong" :363; unsigned wrap :424-427).  The
 * builder-defined semantics (DESIGN.md "M4"), identical in the CUDA kernel:
 *   - the ray runs from the CENTRE of the origin voxel to the CENTRE of the
 *     endpoint voxel (the Point3i interface of map.hpp:32);
 *   - exact integer Amanatides-Woo: axis k crosses its i-th cell wall at
 *     t = (2i+1)/(2 n_k), n_k = |delta_k|; walls are ordered by
 *     cross-multiplied integers, ties broken x before y before z; the walk
 *     takes n_x+n_y+n_z steps and ends on the endpoint voxel;
 *   - every voxel entered EXCEPT the endpoint voxel (and never the origin
 *     voxel) is decremented: c = max(0, c - delta_dec) (clamp, not the
 *     reference's mod-256 wrap);
 *   - per frame all decrements happen first (phase 1), then every endpoint
 *     voxel gets rule A with delta_inc (phase 2);
 *   - only voxels with z in [z_lo, z_hi) are touched (z-slab ownership).
 */
long long orc_map_integrate_rays(uint8_t *grid, const int dims[3], float cell,
                                 const orc_point *pts, int n, const float origin[3],
                                 int delta_dec, int delta_inc, int z_lo, int z_hi)
{
    long long visited = 0;
    int o[3];
    orc_voxel_coords(origin, cell, dims, o);
    for (int i = 0; i < n; i++) {
        float p[3] = {pts[i].x, pts[i].y, pts[i].z};
        int e[3];
        orc_voxel_coords(p, cell, dims, e);
        long long nx = llabs((long long)e[0] - o[0]), ny = llabs((long long)e[1] - o[1]),
                  nz = llabs((long long)e[2] - o[2]);
        int sx = (e[0] > o[0]) - (e[0] < o[0]);
        int sy = (e[1] > o[1]) - (e[1] < o[1]);
        int sz = (e[2] > o[2]) - (e[2] < o[2]);
        /* wall time of axis k scaled by 2 nx' ny' nz' (n' = max(n,1)) */
        long long mx = nx ? nx : 1, my = ny ? ny : 1, mz = nz ? nz : 1;
        const long long INF = (long long)1 << 62;
        long long ex = nx ? my * mz : INF, ey = ny ? mx * mz : INF, ez = nz ? mx * my : INF;
        long long dxs = 2 * my * mz, dys = 2 * mx * mz, dzs = 2 * mx * my;
        long long steps = nx + ny + nz;
        int x = o[0], y = o[1], z = o[2];
        long long cx = 0, cy = 0, cz = 0; /* steps taken per axis */
        for (long long s = 0; s + 1 < steps; s++) {
            if (ex <= ey && ex <= ez) { x += sx; cx++; ex = (cx < nx) ? ex + dxs : INF; }
            else if (ey <= ez) { y += sy; cy++; ey = (cy < ny) ? ey + dys : INF; }
            else { z += sz; cz++; ez = (cz < nz) ? ez + dzs : INF; }
            visited++;
            if (z < z_lo || z >= z_hi) continue;
            uint8_t *c = &grid[((size_t)x * dims[1] + y) * dims[2] + z];
            if (*c > 0) { /* map.cpp:423 */
                int v = (int)*c - delta_dec;
                *c = (uint8_t)(v < 0 ? 0 : v);
            }
        }
    }
    for (int i = 0; i < n; i++) {
        float p[3] = {pts[i].x, pts[i].y, pts[i].z};
        int e[3];
        orc_voxel_coords(p, cell, dims, e);
        if (e[2] < z_lo || e[2] >= z_hi) continue;
        uint8_t *c = &grid[((size_t)e[0] * dims[1] + e[1]) * dims[2] + e[2]];
        if (*c > 255 - delta_inc) *c = 255;
        else *c = (uint8_t)(*c + delta_inc);
    }
    return visited;
}


/* ======================================================================== */
/* 8f-4 pose reporting (scalar host code in the reference)                   */
/* ======================================================================== */
#define ORC_PI 3.14159265358979f /* icp.hpp:4 */

static float orc_sign(float x) { return (x >= 0.0f) ? +1.0f : -1.0f; } /* quaternion.hpp:22 */

/* Quaternion::Quaternion(cv::Mat), quaternion.cpp:23-79.  q = {w, x, y, z}. */
void orc_quat_from_rot(const float R[9], float q[4])
{
    float r11 = R[0], r12 = R[1], r13 = R[2], r21 = R[3], r22 = R[4], r23 = R[5], r31 = R[6], r32 = R[7], r33 = R[8];
    float w = (r11 + r22 + r33 + 1.0f) / 4.0f; /* :35-38 */
    float x = (r11 - r22 - r33 + 1.0f) / 4.0f;
    float y = (-r11 + r22 - r33 + 1.0f) / 4.0f;
    float z = (-r11 - r22 + r33 + 1.0f) / 4.0f;
    if (w < 0.0f) w = 0.0f; /* :40-43 */
    if (x < 0.0f) x = 0.0f;
    if (y < 0.0f) y = 0.0f;
    if (z < 0.0f) z = 0.0f;
    w = sqrtf(w); x = sqrtf(x); y = sqrtf(y); z = sqrtf(z); /* :44-47 */
    if (w >= x && w >= y && w >= z) { /* :49-69 */
        x *= orc_sign(r32 - r23); y *= orc_sign(r13 - r31); z *= orc_sign(r21 - r12);
    } else if (x >= w && x >= y && x >= z) {
        w *= orc_sign(r32 - r23); y *= orc_sign(r21 + r12); z *= orc_sign(r13 + r31);
    } else if (y >= w && y >= x && y >= z) {
        w *= orc_sign(r13 - r31); x *= orc_sign(r21 + r12); z *= orc_sign(r32 + r23);
    } else if (z >= w && z >= x && z >= y) {
        w *= orc_sign(r21 - r12); x *= orc_sign(r31 + r13); y *= orc_sign(r32 + r23);
    }
    float r = sqrtf(w * w + x * x + y * y + z * z); /* NORM, quaternion.hpp:23; :73-77 */
    q[0] = w / r; q[1] = x / r; q[2] = y / r; q[3] = z / r;
}

/* Quaternion::operator*, quaternion.cpp:184-192 */
void orc_quat_mul(const float a[4], const float b[4], float out[4])
{
    float w = a[0], x = a[1], y = a[2], z = a[3], qw = b[0], qx = b[1], qy = b[2], qz = b[3];
    float o0 = w * qw - x * qx - y * qy - z * qz;
    float o1 = w * qx + x * qw + y * qz - z * qy;
    float o2 = w * qy + y * qw + z * qx - x * qz;
    float o3 = w * qz + z * qw + x * qy - y * qx;
    out[0] = o0; out[1] = o1; out[2] = o2; out[3] = o3;
}

/* Quaternion::inverse = conjugate().scale(1/norm()), quaternion.cpp:294-341 (norm() is the SQUARED norm, :294-297) */
void orc_quat_inverse(const float q[4], float out[4])
{
    float n = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    float s = 1 / n;
    out[0] = q[0] * s; out[1] = -q[1] * s; out[2] = -q[2] * s; out[3] = -q[3] * s;
}

/* toEulerianAngle, SLAM.cpp:613-636: degrees */
void orc_quat_to_euler_deg(const float q[4], float e[3])
{
    float qw = q[0], qx = q[1], qy = q[2], qz = q[3];
    float ysqr = qy * qy;
    float t0 = 2.0f * (qw * qx + qy * qz);
    float t1 = 1.0f - 2.0f * (qx * qx + ysqr);
    float x = atan2f(t0, t1);
    float t2 = +2.0f * (qw * qy - qz * qx);
    t2 = t2 > 1.0f ? 1.0f : t2;
    t2 = t2 < -1.0f ? -1.0f : t2;
    float y = asinf(t2);
    float t3 = +2.0f * (qw * qz + qx * qy);
    float t4 = +1.0f - 2.0f * (ysqr + qz * qz);
    float z = atan2f(t3, t4);
    e[0] = x * 180.0f / ORC_PI; e[1] = y * 180.0f / ORC_PI; e[2] = z * 180.0f / ORC_PI;
}

/* transformationMatToEulerianAngle, SLAM.cpp:638-648: pow(float,2) promotes to double, sqrt(double), stored to float;
 * `x * 180 / PI` is float * int -> float */
void orc_mat_to_euler_deg(const float R[9], float e[3])
{
    float c2 = (float)sqrt((double)R[0] * (double)R[0] + (double)R[1] * (double)R[1]);
    float x = atan2f(R[5], R[8]);
    float y = atan2f(-R[2], c2);
    float z = atan2f(R[1], R[0]);
    e[0] = x * 180 / ORC_PI; e[1] = y * 180 / ORC_PI; e[2] = z * 180 / ORC_PI;
}

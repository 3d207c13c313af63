// ref_harness.cpp -- TEST INFRASTRUCTURE (SURVEY.md 8c "Route B").
//
// extern "C" entry points around the reference's OWN functions, compiled unmodified and by path
// from /root/reference against oracle/refshim/opencv2/cvshim.hpp.  Used only by tests/ (and the
// golden-vector generator) to pin the hand-written oracle.  Nothing here is shipped or timed.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <new>

#include "opencv2/imgproc/imgproc.hpp"
#include "opencv2/viz/vizcore.hpp"
#include "map.hpp"
#include "SLAM.hpp"
#include "icp.hpp"

// SLAM.cpp:493-510 (out of scope: timers that are never printed); icp.cpp calls it.
void logDeltaTime(int, int) {}

namespace icp {
extern map::Map map;               // icp.cpp:26
extern cv::Mat cameraRotation;     // icp.cpp:22
extern cv::Point3f cameraPosition; // icp.cpp:24
}

struct ref_point { float x, y, z; unsigned char c0, c1, c2, pad; };

static color_point_t to_cp(const ref_point &p)
{
    color_point_t c;
    c.point = cv::Point3f(p.x, p.y, p.z);
    c.color = cv::Vec3b(p.c0, p.c1, p.c2);
    return c;
}
static ref_point from_cp(const color_point_t &c)
{
    ref_point p;
    p.x = c.point.x; p.y = c.point.y; p.z = c.point.z;
    p.c0 = c.color[0]; p.c1 = c.color[1]; p.c2 = c.color[2]; p.pad = 0;
    return p;
}
static void fill(icp::PointCloud &pc, const ref_point *pts, int n, bool as_keypoints = false)
{
    for (int i = 0; i < n; ++i) (as_keypoints ? pc.keypoints : pc.points).push_back(to_cp(pts[i]));
}

extern "C" {

float ref_distance(const ref_point *a, const ref_point *b) { return icp::distance(to_cp(*a), to_cp(*b)); } // icp.cpp:606

// PointCloud(cv::Mat&, cv::Mat) pointcloud.cpp:109-165.  `decisions[k]` = (rand() % SUBSAMPLE_FACTOR == 0) for
// the k-th non-zero pixel, replayed from the same srand(seed) the constructor then consumes.
int ref_backproject(const uint16_t *depth, const uint8_t *bgr, int w, int h, unsigned seed, ref_point *out,
                    unsigned char *decisions, float *center3)
{
    cv::Mat d(h, w, CV_16UC1, (void *)depth);
    cv::Mat col(h, w, CV_8UC3);
    for (int i = 0; i < w * h * 3; ++i) col.data[i] = bgr ? bgr[i] : 0;
    int nz = 0;
    for (int i = 0; i < w * h; ++i) nz += depth[i] != 0;
    srand(seed);
    for (int k = 0; k < nz; ++k) decisions[k] = (rand() % SUBSAMPLE_FACTOR) == 0;
    srand(seed);
    icp::PointCloud pc(d, col);
    for (size_t i = 0; i < pc.points.size(); ++i) out[i] = from_cp(pc.points[i]);
    center3[0] = pc.center.x; center3[1] = pc.center.y; center3[2] = pc.center.z;
    return (int)pc.points.size();
}

void ref_rotate(ref_point *pts, int n, const float *R9) // pointcloud.cpp:321
{
    icp::PointCloud pc; fill(pc, pts, n);
    float r[9]; for (int i = 0; i < 9; ++i) r[i] = R9[i];
    cv::Mat R(3, 3, CV_32FC1, r);
    pc.rotate(R);
    for (int i = 0; i < n; ++i) pts[i] = from_cp(pc.points[i]);
}

void ref_translate(ref_point *pts, int n, const float *t3) // pointcloud.cpp:349
{
    icp::PointCloud pc; fill(pc, pts, n);
    pc.translate(cv::Point3f(t3[0], t3[1], t3[2]));
    for (int i = 0; i < n; ++i) pts[i] = from_cp(pc.points[i]);
}

// findGlobalNearestNeighborAssociations icp.cpp:541-563: compacted (a, b, error) triples.
int ref_nn_assoc(const ref_point *data, int n, const ref_point *target, int m, ref_point *a_out, ref_point *b_out,
                 float *err_out)
{
    icp::PointCloud dc, tc; fill(dc, data, n); fill(tc, target, m);
    std::vector<float> errors; associations_t assoc;
    icp::findGlobalNearestNeighborAssociations(dc, tc, errors, assoc);
    for (size_t i = 0; i < assoc.size(); ++i) { a_out[i] = from_cp(assoc[i].first); b_out[i] = from_cp(assoc[i].second); err_out[i] = errors[i]; }
    return (int)assoc.size();
}

// getNearestPoint icp.cpp:566-593 for every query, no distance filter.
void ref_nearest(const ref_point *data, int n, const ref_point *target, int m, ref_point *b_out, float *d_out)
{
    icp::PointCloud tc; fill(tc, target, m);
    for (int i = 0; i < n; ++i) { color_point_t nn; d_out[i] = icp::getNearestPoint(to_cp(data[i]), nn, tc); b_out[i] = from_cp(nn); }
}

float ref_mse(const float *errors, int n) { return icp::meanSquareError(std::vector<float>(errors, errors + n)); } // icp.cpp:622

void ref_calculate_offset(const ref_point *a, const ref_point *b, int n, float *out3) // icp.cpp:314-344
{
    associations_t as;
    for (int i = 0; i < n; ++i) as.push_back(std::make_pair(to_cp(a[i]), to_cp(b[i])));
    cv::Point3f o = icp::calculateOffset(as);
    out3[0] = o.x; out3[1] = o.y; out3[2] = o.z;
}

void ref_make_rotation(float x, float y, float z, float *out9) // icp.cpp:640
{
    cv::Mat R = icp::makeRotationMatrix(x, y, z);
    for (int i = 0; i < 9; ++i) out9[i] = R.at<float>(i / 3, i % 3);
}

void ref_voxel(const float *p3, int *v3) // map.cpp:55
{
    cv::Point3i v = icp::map.getVoxelCoordinates(cv::Point3f(p3[0], p3[1], p3[2]));
    v3[0] = v.x; v3[1] = v.y; v3[2] = v.z;
}

void ref_map_reset() { icp::map.~Map(); new (&icp::map) map::Map(); } // map.cpp:17-31

// Map::update(PointCloud, delta, win) map.cpp:220-269 (iterates key-points, rule A)
void ref_map_update_cloud(const ref_point *pts, int n, int delta)
{
    icp::PointCloud pc; fill(pc, pts, n, true);
    cv::viz::Viz3d win;
    icp::map.update(pc, delta, win);
}
// Map::update(assoc, errors, nonAssoc, delta) map.cpp:122-151 (rule C on the non-associated points)
void ref_map_update_nonassoc(const ref_point *pts, int n, int delta)
{
    associations_t assoc; assoc.push_back(std::make_pair(to_cp(pts[0]), to_cp(pts[0]))); // non-empty, or it returns at :124
    point_list_t non; for (int i = 0; i < n; ++i) non.push_back(to_cp(pts[i]));
    icp::map.update(assoc, std::vector<float>(1, 0.f), non, delta);
}
// Map::update(assoc, delta) map.cpp:88-119 (rule A on association firsts)
void ref_map_update_assoc(const ref_point *pts, int n, int delta)
{
    associations_t assoc; for (int i = 0; i < n; ++i) assoc.push_back(std::make_pair(to_cp(pts[i]), to_cp(pts[i])));
    icp::map.update(assoc, delta);
}
void ref_map_world(unsigned char *out) { std::memcpy(out, icp::map.world, sizeof(icp::map.world)); }
int ref_map_keypoints() { return (int)icp::map.mapCloud.keypoints.size(); }
// mapCloud.keypoints (kind 0) or mapCloud.points (kind 1) as appended by the Map::update overloads
int ref_map_cloud(int kind, ref_point *out, int capacity)
{
    const point_list_t &l = kind == 0 ? icp::map.mapCloud.keypoints : icp::map.mapCloud.points;
    int n = (int)l.size();
    for (int i = 0; i < n && i < capacity; ++i) out[i] = from_cp(l[i]);
    return n;
}

// The while loop of icp.cpp:155-258 with the all-point association of :149/:253 swapped in for the
// key-point one, every step being the reference's own function (the loop skeleton is restated
// here because the reference's call sites are commented out).
int ref_icp_allpoints(ref_point *data, int n, const ref_point *target, int m, int maxIterations, float threshold,
                      float *rigid16, float *camR9, float *camP3, float *mse_out, int *n_assoc_out)
{
    icp::PointCloud dataCloud, previousCloud; fill(dataCloud, data, n); fill(previousCloud, target, m);
    std::vector<float> errors; associations_t associations;
    cv::Mat rigid(4, 4, CV_32FC1);
    for (int i = 0; i < 16; ++i) rigid.at<float>(i / 4, i % 4) = (i % 5 == 0) ? 1.f : 0.f;
    cv::Mat cameraRotation = icp::makeRotationMatrix(0, 0, 0);
    cv::Point3f cameraPosition(0, 0, 0), offset(0, 0, 0);
    icp::findGlobalNearestNeighborAssociations(dataCloud, previousCloud, errors, associations);
    int i = 0;
    while (icp::meanSquareError(errors) > threshold && i < maxIterations) {
        if (associations.size() < 3) break;
        icp::PointCloud tempDataCloud, tempMapCloud;
        for (size_t k = 0; k < associations.size(); ++k) { tempDataCloud.points.push_back(associations[k].first); tempMapCloud.points.push_back(associations[k].second); }
        cv::Mat dataMat = tempDataCloud.centered_matrix();
        cv::Mat previousMat = tempMapCloud.centered_matrix();
        cv::Mat M = previousMat.t() * dataMat;
        cv::SVD svd(M);
        cv::Mat R = svd.vt.t() * svd.u.t();
        if (cv::determinant(R) < 0) R.col(2) *= -1;
        if (i == 0) R.copyTo(rigid(cv::Rect(0, 0, 3, 3)));
        else { cv::Mat Rp = R * rigid(cv::Rect(0, 0, 3, 3)); Rp.copyTo(rigid(cv::Rect(0, 0, 3, 3))); }
        R = R.inv();
        dataCloud.rotate(R);
        cameraRotation *= R;
        offset = icp::calculateOffset(associations);
        dataCloud.translate(-offset);
        cameraPosition -= offset;
        icp::findGlobalNearestNeighborAssociations(dataCloud, previousCloud, errors, associations);
        i++;
    }
    rigid.at<float>(0, 3) = offset.x; rigid.at<float>(1, 3) = offset.y; rigid.at<float>(2, 3) = offset.z;
    for (int k = 0; k < 16; ++k) rigid16[k] = rigid.at<float>(k / 4, k % 4);
    for (int k = 0; k < 9; ++k) camR9[k] = cameraRotation.at<float>(k / 3, k % 3);
    camP3[0] = cameraPosition.x; camP3[1] = cameraPosition.y; camP3[2] = cameraPosition.z;
    *mse_out = icp::meanSquareError(errors);
    *n_assoc_out = (int)associations.size();
    for (int k = 0; k < n; ++k) data[k] = from_cp(dataCloud.points[k]);
    return i;
}

// The reference's OWN icp::getTransformation (icp.cpp:28-285), live key-point variant, called as SLAM.cpp:277 calls
// it.  Key-points arrive as pixel coordinates (cv::FAST is out of scope).  srand(seed) first: the two PointCloud
// constructors (:38-39) draw rand() per non-zero pixel (pointcloud.cpp:28), data cloud first.
void ref_get_transformation(const uint16_t *depth_cur, const uint16_t *depth_prev, const uint8_t *bgr, int w, int h,
                            const float *kp_xy, int n_kp, int maxIterations, float threshold, unsigned seed,
                            float *rigid16, float *camR9, float *camP3)
{
    cv::Mat data(h, w, CV_16UC1, (void *)depth_cur), previous(h, w, CV_16UC1, (void *)depth_prev);
    cv::Mat color(h, w, CV_8UC3, (void *)bgr);
    std::vector<cv::KeyPoint> kps((size_t)n_kp);
    for (int i = 0; i < n_kp; ++i) kps[(size_t)i].pt = cv::Point2f(kp_xy[2 * i], kp_xy[2 * i + 1]);
    cv::Mat rotation;
    cv::viz::Viz3d win("ref");
    srand(seed);
    cv::Mat T = icp::getTransformation(data, previous, color, kps, rotation, maxIterations, threshold, win);
    for (int k = 0; k < 16; ++k) rigid16[k] = T.at<float>(k / 4, k % 4);
    for (int k = 0; k < 9; ++k) camR9[k] = icp::cameraRotation.at<float>(k / 3, k % 3);
    camP3[0] = icp::cameraPosition.x; camP3[1] = icp::cameraPosition.y; camP3[2] = icp::cameraPosition.z;
}

// CPU baseline of bench.py: getNearestPoint (icp.cpp:566-593) for every query, the query loop spread over `threads`
// OpenMP threads (the reference itself is single threaded; every call is its own unmodified function).
void ref_nearest_mt(const ref_point *data, int n, const ref_point *target, int m, float *d_out, int threads)
{
    icp::PointCloud tc; fill(tc, target, m);
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads > 0 ? threads : 1)
    for (int i = 0; i < n; ++i) { color_point_t nn; d_out[i] = icp::getNearestPoint(to_cp(data[i]), nn, tc); }
}

void ref_flush_stdout() { std::cout.flush(); fflush(stdout); }

// ---- 8f-4: the reference's own Quaternion class (quaternion.cpp, compiled by path); q = {w, x, y, z}
void ref_quat_from_rot(const float *R9, float *q4) // quaternion.cpp:23-79
{
    cv::Mat R(3, 3, CV_32FC1);
    for (int i = 0; i < 9; ++i) R.at<float>(i / 3, i % 3) = R9[i];
    Quaternion q(R);
    q4[0] = q.w; q4[1] = q.x; q4[2] = q.y; q4[3] = q.z;
}
void ref_quat_mul(const float *a, const float *b, float *o) // quaternion.cpp:184-192
{
    Quaternion qa(a[0], a[1], a[2], a[3]), qb(b[0], b[1], b[2], b[3]);
    Quaternion r = qa * qb;
    o[0] = r.w; o[1] = r.x; o[2] = r.y; o[3] = r.z;
}
void ref_quat_inverse(const float *a, float *o) // quaternion.cpp:325-328
{
    Quaternion qa(a[0], a[1], a[2], a[3]);
    Quaternion r = qa.inverse();
    o[0] = r.w; o[1] = r.x; o[2] = r.y; o[3] = r.z;
}

} // extern "C"

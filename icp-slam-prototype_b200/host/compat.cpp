// Host side of the drop-in headers (include/icpb200/{icp,pointcloud,map}.hpp): the reference's C++ entry
// points, marshalled onto the C-ABI of libicpb200.so.  Bulk paths (depth -> XYZ, rotate / translate,
// association scans, the registration loop, grid updates) run on the B200; only the scalar helpers the
// reference also evaluates per call (distance, meanSquareError, calculateOffset, makeRotationMatrix,
// getVoxelCoordinates, key-point lifting, mapCloud bookkeeping) are host code.  No CPU fallback: without a
// usable device the first bulk call prints the library's error and aborts.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

#include "icpb200.h"
#include "icpb200/map.hpp"

namespace {

struct Host {
    icpb_ctx *ctx = nullptr;
    // device scratch clouds: 0-4 are transients of the PointCloud / Map methods and the association scans,
    // 5-6 hold the frame's clouds for the length of one icp::getTransformation call
    icpb_cloud *slot[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int cap[7] = {0, 0, 0, 0, 0, 0, 0};
};

Host &H()
{
    static Host h;
    if (!h.ctx) {
        const char *dev = getenv("ICPB_DEVICE");
        int rc = icpb_ctx_create(dev ? atoi(dev) : 0, &h.ctx);
        if (rc != ICPB_OK) {
            fprintf(stderr, "icpb200: %s (%s)\n", icpb_last_error(nullptr), icpb_status_string(rc));
            abort();
        }
    }
    return h;
}

void check(int rc, const char *what)
{
    if (rc != ICPB_OK) {
        fprintf(stderr, "icpb200: %s failed: %s\n", what, icpb_last_error(H().ctx));
        abort();
    }
}

// Device scratch cloud `i` with room for n points.
icpb_cloud *scratch(int i, int n)
{
    Host &h = H();
    if (h.cap[i] < n || !h.slot[i]) {
        if (h.slot[i]) icpb_cloud_destroy(h.slot[i]);
        int cap = n + n / 4 + 16;
        check(icpb_cloud_create(h.ctx, cap, &h.slot[i]), "icpb_cloud_create");
        h.cap[i] = cap;
    }
    return h.slot[i];
}

icpb_cloud *upload(int i, const point_list_t &pts)
{
    icpb_cloud *c = scratch(i, (int)pts.size());
    check(icpb_cloud_upload(c, reinterpret_cast<const icpb_point *>(pts.data()), (int)pts.size()), "icpb_cloud_upload");
    return c;
}

void download(icpb_cloud *c, point_list_t &pts)
{
    int n = 0;
    check(icpb_cloud_size(c, &n), "icpb_cloud_size");
    pts.resize((size_t)n);
    if (n) check(icpb_cloud_download(c, reinterpret_cast<icpb_point *>(pts.data()), n, &n), "icpb_cloud_download");
}

// cv::Mat operator* on 3x3 CV_32F: float, no FMA, left to right (what OpenCV's small-matrix gemm does).
void mul33(const float *A, const float *B, float *C)
{
    float T[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[3 * i + j] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j]) + A[3 * i + 2] * B[6 + j];
    for (int k = 0; k < 9; ++k) C[k] = T[k];
}

cv::Mat mat33(const float *v)
{
    cv::Mat m(3, 3, CV_32FC1);
    for (int k = 0; k < 9; ++k) m.at<float>(k / 3, k % 3) = v[k];
    return m;
}

// file-scope state of icp.cpp:22-26
struct IcpGlobals {
    float cameraRotation[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float lastTranslation[3] = {0, 0, 0};
    cv::Point3f cameraPosition;
    map::Map *map = nullptr;
    bool started = false;
    int mode = 0; // icp::ASSOCIATE_ALL_POINTS
};
IcpGlobals &G()
{
    static IcpGlobals g;
    if (!g.map) g.map = new map::Map();
    return g;
}

void lift(const cv::Mat &data, const cv::Mat &colorMat, int x, int y, color_point_t &out)
{
    // pointcloud.cpp:37-39 / 85-87 (CX and FX on both axes)
    std::memset(static_cast<void *>(&out), 0, sizeof out); // the 16th byte is padding: keep it defined (and equal to the device's 0)
    float p_z = ((float)data.at<uint16_t>(y, x)) / 5000.0f;
    float p_x = (x - CX) * p_z / FX;
    float p_y = (y - CX) * p_z / FX;
    out.point = cv::Point3f(p_x, p_y, p_z);
    out.color = colorMat.empty() ? cv::Vec3b() : colorMat.at<cv::Vec3b>(y, x);
}

// ---- the subsample draws of pointcloud.cpp:22-28 -------------------------------------------------------------------
// The reference draws one rand() per non-zero pixel, in raster order, from the process-wide generator; the caller's
// srand() and any rand() it draws itself belong to the same stream, so the drop-in must leave that generator exactly
// where the reference's loop leaves it.  Calling rand() 290,000 times costs 6-7 ms a frame (glibc takes a lock per
// call) -- more than everything else icp::getTransformation does.  glibc's rand() is random(): an additive-feedback
// generator r[i] = r[i-deg] + r[i-deg+sep] over a state array that setstate() hands to the caller.  So the draws are
// made in bulk on that array: park the generator on a spare state, advance the real one m steps here, hand it back.
// Same values, same final state (tests/cpp/test_compat_host.cpp checks both against rand() itself); a start-up
// self-check on a private state falls back to rand() per pixel if this libc's generator is not the one described.
// Like rand() itself in a multi-threaded caller, the order of draws across threads is not defined.
#if defined(__GLIBC__)
// m steps of glibc's random_r() on the state block `word` (word[0]: 5 * rear + type, then the state array);
// out[i] = (draw % factor) == 0.  Returns false for a layout it does not know.
bool advance_glibc_state(int32_t *word, uint8_t *out, size_t m, int factor)
{
    static const int kDeg[5] = {0, 7, 15, 31, 63}, kSep[5] = {0, 3, 1, 3, 1};
    const int type = word[0] % 5, rear = word[0] / 5;
    if (word[0] < 0 || type < 0 || type > 4) return false;
    uint32_t *s = reinterpret_cast<uint32_t *>(word + 1);
    if (type == 0) { // TYPE_0: the linear congruential generator, one state word
        uint32_t x = s[0];
        for (size_t i = 0; i < m; ++i) {
            x = (x * 1103515245u + 12345u) & 0x7fffffffu;
            out[i] = ((int)x % factor) == 0;
        }
        s[0] = x;
        return true;
    }
    const int deg = kDeg[type], sep = kSep[type];
    if (rear >= deg) return false;
    // x[n] = x[n-deg] + x[n-sep], draw = x[n] >> 1.  The circular state is unrolled, oldest first, in front of a
    // linear block buffer, so that the recurrence is a plain streaming loop (no wrap tests, three independent add
    // chains for sep = 3) and the decisions are a second pass over the block.
    enum { kBlock = 8192 };
    uint32_t G[kBlock + 64];
    const int f = (rear + sep) % deg, lag = deg - sep;
    for (int j = 0; j < deg; ++j) G[j] = s[(f + j) % deg];
    for (size_t done = 0; done < m;) {
        const size_t n = std::min<size_t>(kBlock, m - done);
        size_t i = 0;
        if (deg == 31) { // TYPE_3, glibc's default: the three newest values ride in registers (no store-to-load stall)
            uint32_t a = G[28], b = G[29], c = G[30];
            for (; i + 3 <= n; i += 3) {
                a += G[i]; b += G[i + 1]; c += G[i + 2];
                G[i + 31] = a; G[i + 32] = b; G[i + 33] = c;
            }
        }
        for (; i < n; ++i) G[i + deg] = G[i] + G[i + lag];
        const uint32_t *x = G + deg;
        uint8_t *o = out + done;
        if (factor == SUBSAMPLE_FACTOR) // the one factor the reference uses: a compile-time divisor (no idiv per draw)
            for (size_t i = 0; i < n; ++i) o[i] = ((x[i] >> 1) % (uint32_t)SUBSAMPLE_FACTOR) == 0;
        else
            for (size_t i = 0; i < n; ++i) o[i] = ((int)(x[i] >> 1) % factor) == 0;
        for (int j = 0; j < deg; ++j) G[j] = G[n + j]; // the newest deg values lead the next block (n >= 1: no overlap hazard going up)
        done += n;
    }
    const int f_new = (int)((f + m) % (size_t)deg);
    for (int j = 0; j < deg; ++j) s[(f_new + j) % deg] = G[j];
    word[0] = 5 * (int)((rear + m) % (size_t)deg) + type;
    return true;
}

int32_t g_parked[34] = {3}; // a valid TYPE_3 state (rear 0) for glibc to hold while the real one is advanced here

// One-time check on private states that never touch the caller's stream: 64 draws of rand() against 64 bulk draws.
bool bulk_draws_usable()
{
    static int usable = -1;
    if (usable >= 0) return usable == 1;
    alignas(8) static char a[128], b[128];
    char *caller = initstate(20161u, a, sizeof a); // glibc now on `a`; the caller's state is set aside untouched
    if (!caller) { usable = 0; return false; }
    uint8_t want[64], got[64];
    for (int i = 0; i < 64; ++i) want[i] = (rand() % 7) == 0;
    const int tail = rand();
    initstate(20161u, b, sizeof b);                                   // the same seed on a second block
    char *blk = setstate(reinterpret_cast<char *>(g_parked));        // ... which is handed back here
    bool ok = blk == b && advance_glibc_state(reinterpret_cast<int32_t *>(blk), got, 64, 7);
    if (ok) {
        setstate(blk);
        ok = rand() == tail; // the state handed back continues where rand() itself would
        for (int i = 0; i < 64; ++i) ok = ok && want[i] == got[i];
    }
    setstate(caller);
    usable = ok ? 1 : 0;
    return ok;
}
#endif

// keep[i] = (rand() % factor) == 0 for m successive draws of the process-wide generator.
void draw_keep(uint8_t *keep, size_t m, int factor)
{
#if defined(__GLIBC__)
    static const bool forced_off = getenv("ICPB_COMPAT_RAND_CALLS") != nullptr; // A/B switch: one rand() call per draw
    if (m >= 64 && !forced_off && bulk_draws_usable()) {
        char *blk = setstate(reinterpret_cast<char *>(g_parked));
        if (blk) {
            const bool ok = advance_glibc_state(reinterpret_cast<int32_t *>(blk), keep, m, factor);
            setstate(blk);
            if (ok) return; // (a refused layout leaves the state untouched: fall through to rand())
        }
    }
#endif
    for (size_t i = 0; i < m; ++i) keep[i] = (rand() % factor) == 0;
}

} // namespace

// test hook (tests/cpp/test_compat_host.cpp): the subsample decisions of one depth image, as build_from_depth draws them
extern "C" int icpb_compat_draw_keep(const uint16_t *depth, int n_px, int factor, uint8_t *keep)
{
    size_t m = 0;
    for (int i = 0; i < n_px; ++i) m += depth[i] != 0;
    draw_keep(keep, m, factor);
    return (int)m;
}

// ================================================================ pointcloud.hpp
namespace icp {

// pointcloud.cpp:19-52 on the device: the subsampled cloud of a depth image, left in scratch slot `slot`.
static icpb_cloud *device_cloud_from_depth(int slot, cv::Mat &data, cv::Mat &colorMat, int *count)
{
    const int w = data.cols, h = data.rows;
    const uint16_t *d = reinterpret_cast<const uint16_t *>(data.data);
    // One rand() per non-zero pixel in raster order, exactly where the reference draws it (pointcloud.cpp:22-28)
    size_t m = 0;
    for (int i = 0; i < w * h; ++i) m += d[i] != 0;
    std::vector<uint8_t> keep(std::max<size_t>(m, 1), 0);
    draw_keep(keep.data(), m, SUBSAMPLE_FACTOR);
    icpb_intrinsics K;
    icpb_intrinsics_reference_v1(&K);
    icpb_cloud *c = scratch(slot, w * h);
    check(icpb_cloud_from_depth(c, d, colorMat.empty() ? nullptr : colorMat.data, w, h, &K, ICPB_SUB_STREAM, SUBSAMPLE_FACTOR, 0,
                                keep.data(), (int)keep.size()),
          "icpb_cloud_from_depth");
    if (count) check(icpb_cloud_size(c, count), "icpb_cloud_size");
    return c;
}

// The draws of a PointCloud whose points nothing reads (the reference constructs it all the same, icp.cpp:39).
static void draw_only(cv::Mat &data)
{
    const uint16_t *d = reinterpret_cast<const uint16_t *>(data.data);
    size_t m = 0;
    for (int i = 0; i < data.cols * data.rows; ++i) m += d[i] != 0;
    std::vector<uint8_t> keep(std::max<size_t>(m, 1), 0);
    draw_keep(keep.data(), m, SUBSAMPLE_FACTOR);
}

static void build_from_depth(PointCloud &pc, cv::Mat &data, cv::Mat &colorMat)
{
    icpb_cloud *c = device_cloud_from_depth(0, data, colorMat, nullptr);
    download(c, pc.points);
    // center: float running sum, then / count (pointcloud.cpp:43-45,100-102; 0/0 = NaN when empty, as there)
    pc.center = cv::Point3f(0, 0, 0);
    for (const color_point_t &p : pc.points) { pc.center.x += p.point.x; pc.center.y += p.point.y; pc.center.z += p.point.z; }
    int index = (int)pc.points.size();
    pc.center.x /= index; pc.center.y /= index; pc.center.z /= index;
}

static void lift_keypoints(cv::Mat &data, cv::Mat &colorMat, const std::vector<cv::KeyPoint> &keypointsList, point_list_t &out)
{
    for (const cv::KeyPoint &kp : keypointsList) { // pointcloud.cpp:64-97
        int x = (int)std::lrint(kp.pt.x), y = (int)std::lrint(kp.pt.y); // Point2f -> Point2i rounds (saturate_cast)
        if (data.at<uint16_t>(y, x) == 0) continue;
        color_point_t p;
        lift(data, colorMat, x, y, p);
        out.push_back(p);
    }
}

PointCloud::PointCloud(cv::Mat &data, cv::Mat colorMat, std::vector<cv::KeyPoint> keypointsList)
{
    build_from_depth(*this, data, colorMat);
    lift_keypoints(data, colorMat, keypointsList, keypoints);
    center_points();
}

PointCloud::PointCloud(cv::Mat &data, cv::Mat colorMat)
{
    build_from_depth(*this, data, colorMat);
    center_points();
}

PointCloud::PointCloud(std::vector<cv::Point3f> data) // pointcloud.cpp:256-287
{
    center = cv::Point3f(0, 0, 0);
    int index = 0;
    for (const cv::Point3f &p : data) {
        color_point_t c;
        c.point = p;
        points.push_back(c);
        center.x += p.x; center.y += p.y; center.z += p.z;
        index++;
    }
    center.x /= index; center.y /= index; center.z /= index;
    center_points();
}

PointCloud::PointCloud() { center = cv::Point3f(0, 0, 0); }

void PointCloud::center_points() {} // pointcloud.cpp:296-319 mutates local copies only: a no-op

void PointCloud::rotate(cv::Mat &R) // pointcloud.cpp:321-346 (about the world origin; center untouched)
{
    float r[9];
    for (int k = 0; k < 9; ++k) r[k] = R.at<float>(k / 3, k % 3);
    if (!points.empty()) {
        icpb_cloud *c = upload(0, points);
        check(icpb_cloud_transform(c, r, nullptr), "icpb_cloud_transform");
        download(c, points);
    }
    if (!keypoints.empty()) {
        icpb_cloud *c = upload(0, keypoints);
        check(icpb_cloud_transform(c, r, nullptr), "icpb_cloud_transform");
        download(c, keypoints);
    }
}

void PointCloud::translate(cv::Point3f offset) // pointcloud.cpp:349-359
{
    const float t[3] = {offset.x, offset.y, offset.z};
    if (!points.empty()) {
        icpb_cloud *c = upload(0, points);
        check(icpb_cloud_transform(c, nullptr, t), "icpb_cloud_transform");
        download(c, points);
    }
    if (!keypoints.empty()) {
        icpb_cloud *c = upload(0, keypoints);
        check(icpb_cloud_transform(c, nullptr, t), "icpb_cloud_transform");
        download(c, keypoints);
    }
    center += offset;
}

static cv::Mat list_matrix(const point_list_t &pts, cv::Point3f add)
{
    cv::Mat M((int)pts.size(), 3, CV_32FC1);
    for (size_t i = 0; i < pts.size(); ++i) {
        M.at<float>((int)i, 0) = pts[i].point.x + add.x;
        M.at<float>((int)i, 1) = pts[i].point.y + add.y;
        M.at<float>((int)i, 2) = pts[i].point.z + add.z;
    }
    return M;
}
cv::Mat PointCloud::centered_matrix() { return list_matrix(points, cv::Point3f(0, 0, 0)); }            // :361-371 (no centring)
cv::Mat PointCloud::centered_keypoint_matrix() { return list_matrix(keypoints, cv::Point3f(0, 0, 0)); } // :373-383
cv::Mat PointCloud::matrix() { return list_matrix(points, center); }                                    // :385-395
void PointCloud::displayColorPoints(cv::viz::Viz3d &, std::string, int) {}
void PointCloud::displayKeyPoints(cv::viz::Viz3d &, std::string, int, cv::viz::Color) {}
void PointCloud::displayAll(cv::viz::Viz3d &, std::string, int, cv::viz::Color) {}

// ================================================================ icp.hpp

float distance(cv::Point3f a, cv::Point3f b) // icp.cpp:595-602
{
    float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z;
    return (float)std::sqrt(std::pow((double)x, 2) + std::pow((double)y, 2) + std::pow((double)z, 2));
}

float distance(color_point_t a, color_point_t b) // icp.cpp:606-620 (COLOR_WEIGHT 0)
{
    float x = a.point.x - b.point.x, y = a.point.y - b.point.y, z = a.point.z - b.point.z;
    float xyz = (float)((double)x * (double)x + (double)y * (double)y + (double)z * (double)z);
    return std::sqrt(xyz);
}

float meanSquareError(std::vector<float> errors) // icp.cpp:622-638
{
    float error_sum = 0;
    for (size_t i = 0; i < errors.size(); i++) error_sum += errors[i];
    if (errors.size() > 0) {
        error_sum /= errors.size();
        error_sum = (float)std::pow((double)error_sum, 2);
    }
    return error_sum;
}

cv::Point3f calculateOffset(associations_t associations) // icp.cpp:314-344
{
    cv::Point3f offset(0, 0, 0);
    int offset_count = 0;
    for (const auto &pr : associations) {
        offset += pr.first.point - pr.second.point;
        offset_count++;
    }
    if (offset_count > 0) { offset.x /= offset_count; offset.y /= offset_count; offset.z /= offset_count; }
    return offset;
}

cv::Mat makeRotationMatrix(float x, float y, float z) // icp.cpp:640-653
{
    double rotX = x * PI / 180, rotY = y * PI / 180, rotZ = z * PI / 180;
    float d[9] = {1, 0, 0, 0, (float)cos(rotX), (float)sin(rotX), 0, (float)-sin(rotX), (float)cos(rotX)};
    float f[9] = {(float)cos(rotY), 0, (float)-sin(rotY), 0, 1, 0, (float)sin(rotY), 0, (float)cos(rotY)};
    float g[9] = {(float)cos(rotZ), (float)sin(rotZ), 0, (float)-sin(rotZ), (float)cos(rotZ), 0, 0, 0, 1};
    float ab[9], abc[9];
    mul33(d, f, ab);
    mul33(ab, g, abc);
    return mat33(abc);
}

void showAssocations(associations_t, std::vector<float>, cv::viz::Viz3d &) {}

// Shared body of the association scans: exact brute-force NN on the device, compaction on the host.
static void associate(const point_list_t &queries, const point_list_t &targets, float max_d, std::vector<float> &errors,
                      associations_t &associations, point_list_t *rejects)
{
    const int n = (int)queries.size();
    if (n == 0) return;
    icpb_cloud *dc = upload(0, queries);
    icpb_cloud *tc = upload(1, targets);
    std::vector<int32_t> idx((size_t)n);
    std::vector<float> dist((size_t)n);
    check(icpb_nn_search(H().ctx, dc, tc, idx.data(), dist.data(), nullptr), "icpb_nn_search");
    for (int i = 0; i < n; ++i) {
        if (dist[i] < max_d) {
            associations.push_back(std::make_pair(queries[i], targets[(size_t)idx[i]]));
            errors.push_back(dist[i]);
        } else if (rejects) {
            rejects->push_back(queries[i]);
        }
    }
}

void findGlobalNearestNeighborAssociations(PointCloud &data, PointCloud &previous, std::vector<float> &errors,
                                           associations_t &associations) // icp.cpp:541-563
{
    errors.clear();
    associations.clear();
    associate(data.points, previous.points, MAX_NN_COLOR_DISTANCE, errors, associations, nullptr);
}

void findGlobalKeyPointAssociations(PointCloud &data, std::vector<float> &errors, associations_t &associations,
                                    point_list_t &nonAssociations) // icp.cpp:488-515
{
    map::Map &m = *G().map;
    if (m.mapCloud.keypoints.size() == 0) return; // :490-491 (nothing is cleared)
    errors.clear();
    associations.clear();
    associate(data.keypoints, m.mapCloud.keypoints, MAX_NN_KEYPOINT_DISTANCE, errors, associations, &nonAssociations);
}

// icp.cpp:347-369.  The reference's voxel-table search (:371-486) is dead code with out-of-bounds reads; an exact
// indexed search must return what the brute-force scan over the mapped points returns, so that is what runs.
void findMappedNearestNeighborAssociations(PointCloud &data, std::vector<float> &errors, associations_t &associations)
{
    errors.clear();
    associations.clear();
    map::Map &m = *G().map;
    if (m.mapCloud.points.empty()) return;
    associate(data.points, m.mapCloud.points, MAX_NN_COLOR_DISTANCE, errors, associations, nullptr);
}

static float nearest_one(const color_point_t &point, color_point_t &nearest, const point_list_t &targets)
{
    point_list_t q(1, point);
    icpb_cloud *dc = upload(0, q);
    icpb_cloud *tc = upload(1, targets);
    int32_t idx = 0;
    float d = 0.f;
    check(icpb_nn_search(H().ctx, dc, tc, &idx, &d, nullptr), "icpb_nn_search");
    nearest = targets[(size_t)idx];
    return d;
}
// icp.cpp:476-486.  The reference reads pointLookupTable[x][y][z]; here the table lives on the device as an index into
// the map cloud, and a voxel's entry is the single map-cloud point recorded in it -- found on the host by its voxel.
void processVoxel(color_point_t point, color_point_t &nearest, float &shortestDistance, int x, int y, int z)
{
    map::Map &m = *G().map;
    for (const color_point_t &p : m.mapCloud.points) {
        const cv::Point3i v = m.getVoxelCoordinates(p.point);
        if (v.x != x || v.y != y || v.z != z) continue;
        const float d = distance(point, p);
        if (d < shortestDistance) {
            shortestDistance = d;
            nearest = p;
        }
        return; // one entry per voxel
    }
}

// icp.cpp:371-474 made exact: the expanding-cube walk is replaced by the exact scan of the map cloud on the device.
float getNearestMappedPoint(color_point_t point, color_point_t &nearest)
{
    map::Map &m = *G().map;
    if (m.mapCloud.points.empty()) return MAX_NN_COLOR_DISTANCE;
    color_point_t best;
    const float d = nearest_one(point, best, m.mapCloud.points);
    if (!(d < MAX_NN_COLOR_DISTANCE)) return MAX_NN_COLOR_DISTANCE;
    nearest = best;
    return d;
}

float getNearestPoint(color_point_t point, color_point_t &nearest, PointCloud &cloud) { return nearest_one(point, nearest, cloud.points); }       // :566-593
float getNearestKeyPoint(color_point_t point, color_point_t &nearest, PointCloud &cloud) { return nearest_one(point, nearest, cloud.keypoints); } // :517-539

void resetState()
{
    IcpGlobals &g = G();
    const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int k = 0; k < 9; ++k) g.cameraRotation[k] = I[k];
    for (int k = 0; k < 3; ++k) g.lastTranslation[k] = 0.f;
    g.cameraPosition = cv::Point3f(0, 0, 0);
    g.started = false;
    g.map->clear();
    g.map->mapCloud = PointCloud();
}
void setAssociationMode(int mode) { G().mode = mode; }
map::Map &mapState() { return *G().map; }
cv::Mat cameraRotationState() { return mat33(G().cameraRotation); }
cv::Point3f cameraPositionState() { return G().cameraPosition; }

// icp.cpp:28-285 exactly as the reference runs it (ASSOCIATE_KEYPOINTS): the data cloud's key-points
// against the growing map cloud's key-points (:98,:255), the whole cloud moved along, rule-C map update from the
// accumulated rejects (:271).  The loop, the motion of all points and the reject list stay on the device
// (icpb_icp_register_keypoints); the map update is icpb_map_update_tracked.
static cv::Mat getTransformationKeyPoints(cv::Mat &data, cv::Mat &previous, cv::Mat color, std::vector<cv::KeyPoint> keypoints,
                                          int maxIterations, float threshold, cv::viz::Viz3d &depthWindow)
{
    IcpGlobals &g = G();
    map::Map &m = *g.map;
    // :38 -- the data cloud stays on the device (the loop moves it along with the key-points and nothing reads it back);
    // its key-points are lifted on the host as in the reference
    int n_points = 0;
    icpb_cloud *pc = device_cloud_from_depth(5, data, color, &n_points);
    point_list_t dataKeypoints;
    lift_keypoints(data, color, keypoints, dataKeypoints);
    if (m.mapCloud.points.size() == 0) {                  // :47-68
        PointCloud previousCloud(previous, color, keypoints); // :39
        const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        for (int k = 0; k < 9; ++k) g.cameraRotation[k] = I[k];
        g.cameraPosition = cv::Point3f(5, 5, 5);
        for (int k = 0; k < 3; ++k) g.lastTranslation[k] = 0.f;
        cv::Mat R0 = mat33(g.cameraRotation);
        previousCloud.rotate(R0);
        previousCloud.translate(g.cameraPosition);
        m.update(previousCloud, MAX_CONFIDENCE, depthWindow);
        m.mapCloud.points = previousCloud.points;
        g.started = true;
    } else {
        draw_only(previous); // :39 -- with the map seeded nothing reads previousCloud; its rand() draws are all that shows
    }
    cv::Mat rigid(4, 4, CV_32FC1);
    for (int k = 0; k < 16; ++k) rigid.at<float>(k / 4, k % 4) = (k % 5 == 0) ? 1.f : 0.f;
    // :70-71 and the loop, on the device
    const float cp[3] = {g.cameraPosition.x, g.cameraPosition.y, g.cameraPosition.z};
    icpb_cloud *kc = upload(0, dataKeypoints);
    icpb_cloud *mc = upload(3, m.mapCloud.keypoints);
    icpb_cloud *nc = scratch(4, (maxIterations + 1) * (int)std::max<size_t>(dataKeypoints.size(), 1));
    if (!dataKeypoints.empty()) check(icpb_cloud_transform(kc, g.cameraRotation, cp), "icpb_cloud_transform");
    if (n_points > 0) check(icpb_cloud_transform(pc, g.cameraRotation, cp), "icpb_cloud_transform");
    icpb_icp_params prm;
    prm.max_iterations = maxIterations;
    prm.threshold = threshold;
    prm.max_nn_distance = MAX_NN_KEYPOINT_DISTANCE; // :503
    prm.solve_mode = ICPB_SOLVE_REFERENCE;
    for (int k = 0; k < 3; ++k) prm.last_translation[k] = g.lastTranslation[k];
    prm.idx_trace = nullptr;
    prm.dist_trace = nullptr;
    prm.nn_mode = ICPB_NN_BRUTE;
    prm.grid_cell = 0.f;
    prm.nn_filter = ICPB_FILTER_AUTO;
    icpb_icp_result res;
    check(icpb_icp_register_keypoints(H().ctx, kc, n_points > 0 ? pc : nullptr, mc, &prm, &res, nc),
          "icpb_icp_register_keypoints");
    mul33(g.cameraRotation, res.cam_rotation, g.cameraRotation);                                  // :237
    g.cameraPosition += cv::Point3f(res.cam_position[0], res.cam_position[1], res.cam_position[2]); // :246
    for (int k = 0; k < 3; ++k) g.lastTranslation[k] = -res.offset[k];                              // :260
    std::cout << res.mse;                                                                           // :264
    if (res.n_assoc > 0) { // map.cpp:124-126; rule C on the rejects of every pass (:271, map.cpp:130-151)
        point_list_t non;
        download(nc, non);
        associations_t some(1);
        m.update(some, std::vector<float>(), non, DELTA_CONFIDENCE);
    }
    std::cout << std::endl << m.mapCloud.points.size() << std::endl;                                // :279
    for (int k = 0; k < 16; ++k) rigid.at<float>(k / 4, k % 4) = res.rigid[k];
    return rigid;
}

// icp.cpp:28-285 with the all-point association (:149/:253) against the previous frame's cloud placed at the
// current camera pose.  Depth -> XYZ, the 20-iteration loop and the pose update all stay on the device.
cv::Mat getTransformation(cv::Mat &data, cv::Mat &previous, cv::Mat color, std::vector<cv::KeyPoint> keypoints,
                          cv::Mat &rotation, int maxIterations, float threshold, cv::viz::Viz3d &depthWindow)
{
    (void)rotation;
    IcpGlobals &g = G();
    if (g.mode == ASSOCIATE_KEYPOINTS)
        return getTransformationKeyPoints(data, previous, color, keypoints, maxIterations, threshold, depthWindow);
    // :38-39 -- both clouds are built on the device and stay there: the loop reads and moves them in HBM, and of the
    // registered data cloud only the key-points come back (for the map update of :270)
    int n_data = 0, n_prev = 0;
    icpb_cloud *dc = device_cloud_from_depth(5, data, color, &n_data);
    point_list_t dataKeypoints;
    lift_keypoints(data, color, keypoints, dataKeypoints);
    const float cp0[3] = {g.cameraPosition.x, g.cameraPosition.y, g.cameraPosition.z};
    icpb_cloud *tc = nullptr;
    if (!g.started) {                                   // :47-68
        PointCloud previousCloud(previous, color, keypoints); // :39 (previous depth with the CURRENT colour / key-points)
        const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        for (int k = 0; k < 9; ++k) g.cameraRotation[k] = I[k];
        g.cameraPosition = cv::Point3f(5, 5, 5);
        for (int k = 0; k < 3; ++k) g.lastTranslation[k] = 0.f;
        cv::Mat R0 = mat33(g.cameraRotation);
        previousCloud.rotate(R0);
        previousCloud.translate(g.cameraPosition);
        g.map->update(previousCloud, MAX_CONFIDENCE, depthWindow);
        g.map->mapCloud.points = previousCloud.points;
        g.started = true;
        tc = upload(6, previousCloud.points);
        n_prev = (int)previousCloud.points.size();
    } else {
        // the previous frame's cloud placed at the current camera pose: rotate, then translate (pointcloud.cpp:321-359)
        tc = device_cloud_from_depth(6, previous, color, &n_prev);
        if (n_prev > 0) check(icpb_cloud_transform(tc, g.cameraRotation, cp0), "icpb_cloud_transform");
    }
    // :70-71 on the device, then the loop
    const float cp[3] = {g.cameraPosition.x, g.cameraPosition.y, g.cameraPosition.z};
    if (n_data > 0) check(icpb_cloud_transform(dc, g.cameraRotation, cp), "icpb_cloud_transform");
    icpb_icp_params prm;
    prm.max_iterations = maxIterations;
    prm.threshold = threshold;
    prm.max_nn_distance = MAX_NN_COLOR_DISTANCE;
    prm.solve_mode = ICPB_SOLVE_REFERENCE;
    for (int k = 0; k < 3; ++k) prm.last_translation[k] = g.lastTranslation[k];
    prm.idx_trace = nullptr;
    prm.dist_trace = nullptr;
    prm.nn_mode = ICPB_NN_AUTO; // the exact cell-grid search from ~45k x 45k points on, the scan below: same associations
    prm.grid_cell = 0.f;
    prm.nn_filter = ICPB_FILTER_AUTO;
    icpb_icp_result res;
    cv::Mat rigid(4, 4, CV_32FC1);
    if (n_data == 0 || n_prev == 0) {
        for (int k = 0; k < 16; ++k) rigid.at<float>(k / 4, k % 4) = (k % 5 == 0) ? 1.f : 0.f;
        return rigid;
    }
    // the key-points ride along: dataCloud.rotate / translate move points and key-points together (pointcloud.cpp:321-359),
    // so they receive (cameraRotation, cameraPosition) of :70-71 and then every motion of the loop, in order
    icpb_cloud *kc = nullptr;
    if (!dataKeypoints.empty()) {
        kc = upload(2, dataKeypoints);
        check(icpb_cloud_transform(kc, g.cameraRotation, cp), "icpb_cloud_transform");
    }
    check(icpb_icp_register_carry(H().ctx, dc, tc, kc, &prm, &res), "icpb_icp_register_carry");
    // cameraRotation *= R (:237) and cameraPosition -= offset (:246), accumulated over the iterations
    mul33(g.cameraRotation, res.cam_rotation, g.cameraRotation);
    g.cameraPosition += cv::Point3f(res.cam_position[0], res.cam_position[1], res.cam_position[2]);
    for (int k = 0; k < 3; ++k) g.lastTranslation[k] = -res.offset[k]; // :260
    std::cout << res.mse;                                               // :264
    // :270 (the all-point twin of :271): certainty update from the registered key-points
    if (kc) {
        PointCloud registered;
        download(kc, registered.keypoints);
        g.map->update(registered, DELTA_CONFIDENCE, depthWindow);
    }
    std::cout << std::endl << g.map->mapCloud.points.size() << std::endl; // :279
    for (int k = 0; k < 16; ++k) rigid.at<float>(k / 4, k % 4) = res.rigid[k];
    return rigid;
}

} // namespace icp

// ================================================================ map.hpp
namespace map {

Map::Map() : pointLookupTable{this}, world(nullptr), dev_(nullptr) {}

Map::~Map()
{
    if (dev_) icpb_map_destroy(dev_);
    delete[] reinterpret_cast<unsigned char *>(world);
}

void Map::ensure()
{
    if (dev_) return;
    const int dims[3] = {MAP_HEIGHT, MAP_HEIGHT, MAP_HEIGHT};
    check(icpb_map_create(H().ctx, dims, float(CELL_PHYSICAL_HEIGHT), 0, MAP_HEIGHT, &dev_), "icpb_map_create");
    world = reinterpret_cast<unsigned char (*)[MAP_HEIGHT][MAP_HEIGHT]>(new unsigned char[(size_t)MAP_HEIGHT * MAP_HEIGHT * MAP_HEIGHT]());
}

void Map::clear()
{
    ensure();
    check(icpb_map_clear(dev_), "icpb_map_clear");
    std::memset(world, 0, (size_t)MAP_HEIGHT * MAP_HEIGHT * MAP_HEIGHT);
}

void Map::syncWorld()
{
    ensure();
    check(icpb_map_download(dev_, reinterpret_cast<uint8_t *>(world), (long long)MAP_HEIGHT * MAP_HEIGHT * MAP_HEIGHT), "icpb_map_download");
}

cv::Point3i Map::getVoxelCoordinates(cv::Point3f point) // map.cpp:55-85
{
    cv::Point3i p;
    float c = float(CELL_PHYSICAL_HEIGHT);
    p.x = int(point.x / c); p.y = int(point.y / c); p.z = int(point.z / c);
    if (p.x < 0) p.x = 0;
    if (p.x >= MAP_HEIGHT) p.x = MAP_HEIGHT - 1;
    if (p.y < 0) p.y = 0;
    if (p.y >= MAP_HEIGHT) p.y = MAP_HEIGHT - 1;
    if (p.z < 0) p.z = 0;
    if (p.z >= MAP_HEIGHT) p.z = MAP_HEIGHT - 1;
    return p;
}

// The three Map::update overloads, certainty grid AND pointLookupTable / mapCloud bookkeeping, on the device
// (icpb_map_update_tracked); the points it appends come back in point order and join the host-side list.
// Table entries number the stored points of mapCloud.keypoints from 0 and those of mapCloud.points from kPointsList.
static const int kPointsList = 1 << 30;

static void tracked_update(icpb_map *dev, const point_list_t &pts, int variant, int delta, point_list_t &append_to)
{
    const size_t kMax = 65536;
    for (size_t off = 0; off < pts.size(); off += kMax) {
        point_list_t part(pts.begin() + (long)off, pts.begin() + (long)std::min(pts.size(), off + kMax));
        icpb_cloud *src = upload(2, part);
        icpb_cloud *dst = scratch(1, (int)part.size());
        check(icpb_cloud_upload(dst, nullptr, 0), "icpb_cloud_upload");
        int appended = 0;
        const int base = (int)append_to.size() + (variant == ICPB_TRACK_ASSOC ? kPointsList : 0);
        check(icpb_map_update_tracked_base(dev, src, variant, delta, MAX_CONFIDENCE, dst, base, &appended), "icpb_map_update_tracked");
        if (appended) {
            point_list_t got;
            download(dst, got);
            append_to.insert(append_to.end(), got.begin(), got.end());
        }
    }
}

void Map::update(icp::PointCloud data, int delta_confidence, cv::viz::Viz3d &) // map.cpp:220-269
{
    ensure();
    tracked_update(dev_, data.keypoints, ICPB_TRACK_INIT, delta_confidence, mapCloud.keypoints);
}

void Map::update(associations_t associations, int delta_confidence) // map.cpp:88-119 (never called by the reference)
{
    ensure();
    point_list_t firsts;
    for (const auto &pr : associations) firsts.push_back(pr.first);
    tracked_update(dev_, firsts, ICPB_TRACK_ASSOC, delta_confidence, mapCloud.points);
}

void Map::update(associations_t keyPointAssociations, std::vector<float>, point_list_t nonAssociations, int delta_confidence)
{
    if (keyPointAssociations.size() == 0) return; // map.cpp:124-126
    ensure();
    tracked_update(dev_, nonAssociations, ICPB_TRACK_NONASSOC, delta_confidence, mapCloud.keypoints); // map.cpp:130-151
}

// map.cpp:272-439 with the semantics of DESIGN.md "M4": one ray between two voxels.
void Map::rayTrace(cv::Point3i point, cv::Point3i origin, cv::viz::Viz3d &)
{
    ensure();
    const float c = float(CELL_PHYSICAL_HEIGHT);
    point_list_t one(1);
    one[0].point = cv::Point3f((point.x + 0.5f) * c, (point.y + 0.5f) * c, (point.z + 0.5f) * c);
    icpb_cloud *cl = upload(2, one);
    const float o[3] = {(origin.x + 0.5f) * c, (origin.y + 0.5f) * c, (origin.z + 0.5f) * c};
    check(icpb_map_integrate_rays(dev_, cl, o, DELTA_CONFIDENCE, 0, nullptr), "icpb_map_integrate_rays");
}

void Map::integrateRays(icp::PointCloud &cloud, cv::Point3f origin, int delta_dec, int delta_inc)
{
    ensure();
    if (cloud.points.empty()) return;
    icpb_cloud *cl = upload(2, cloud.points);
    const float o[3] = {origin.x, origin.y, origin.z};
    check(icpb_map_integrate_rays(dev_, cl, o, delta_dec, delta_inc, nullptr), "icpb_map_integrate_rays");
}

void Map::drawCertaintyMap(cv::viz::Viz3d &) {}

color_point_t Map::lookup(int x, int y, int z) const // pointLookupTable[x][y][z], map.hpp:24
{
    if (!dev_) return empty;
    const int v[3] = {x, y, z};
    int e = -1;
    check(icpb_map_table_entry(dev_, v, &e), "icpb_map_table_entry");
    if (e < 0) return empty;
    const point_list_t &list = (e >= kPointsList) ? mapCloud.points : mapCloud.keypoints;
    const size_t k = (size_t)(e >= kPointsList ? e - kPointsList : e);
    return k < list.size() ? list[k] : empty;
}

color_point_t LookupZ::operator[](int z) const { return m->lookup(x, y, z); }

bool Map::isOccupied(cv::Point3f p) // map.cpp:441-444
{
    syncWorld();
    cv::Point3i v = getVoxelCoordinates(p);
    return world[v.x][v.y][v.z] >= MAX_CONFIDENCE;
}

int Map::bound(int t, int ds) // map.cpp:209-217 (always 1/ds; unused by the reference)
{
    if (ds < 0) return bound(-t, -ds);
    t = ((t % 1) + 1) % 1;
    return (1 - t) / ds;
}

} // namespace map

// ================================================================ quaternion.hpp / SLAM.hpp pose reporting (8f-4)
#include "icpb200/quaternion.hpp"

Quaternion::Quaternion(void) : x(0), y(0), z(0), w(0) {}                                          // quaternion.cpp:15-21
Quaternion::Quaternion(float wi, float xi, float yi, float zi) : x(xi), y(yi), z(zi), w(wi) {}    // :87-93
Quaternion::Quaternion(float v[4]) : x(v[1]), y(v[2]), z(v[3]), w(v[0]) {}                         // :100-106
Quaternion::Quaternion(cv::Mat rotationMatrix)                                                   // :23-79
{
    float R[9], q[4];
    for (int k = 0; k < 9; ++k) R[k] = rotationMatrix.at<float>(k / 3, k % 3);
    icpb_pose_quat_from_rotation(R, q);
    w = q[0]; x = q[1]; y = q[2]; z = q[3];
}
Quaternion Quaternion::operator*(const Quaternion &q) // :184-192
{
    const float a[4] = {w, x, y, z}, b[4] = {q.w, q.x, q.y, q.z};
    float o[4];
    icpb_pose_quat_mul(a, b, o);
    return Quaternion(o[0], o[1], o[2], o[3]);
}
bool Quaternion::operator==(const Quaternion &q) { return w == q.w && x == q.x && y == q.y && z == q.z; } // :285-288
float Quaternion::norm() { return (w * w + x * x + y * y + z * z); }                                       // :294-297 (squared)
float Quaternion::magnitude() { return sqrtf(norm()); }                                                    // :304-307
Quaternion Quaternion::scale(float s) { return Quaternion(w * s, x * s, y * s, z * s); }                   // :314-317
Quaternion Quaternion::conjugate() { return Quaternion(w, -x, -y, -z); }                                   // :335-338
Quaternion Quaternion::inverse()                                                                           // :325-328
{
    const float a[4] = {w, x, y, z};
    float o[4];
    icpb_pose_quat_inverse(a, o);
    return Quaternion(o[0], o[1], o[2], o[3]);
}

void toEulerianAngle(Quaternion q, float &x, float &y, float &z) // SLAM.cpp:613-636
{
    const float a[4] = {q.w, q.x, q.y, q.z};
    float e[3];
    icpb_pose_quat_to_euler_deg(a, e);
    x = e[0]; y = e[1]; z = e[2];
}

void transformationMatToEulerianAngle(cv::Mat t, float &x, float &y, float &z) // SLAM.cpp:638-648
{
    float R[9], e[3];
    for (int k = 0; k < 9; ++k) R[k] = t.at<float>(k / 3, k % 3);
    icpb_pose_matrix_to_euler_deg(R, e);
    x = e[0]; y = e[1]; z = e[2];
}

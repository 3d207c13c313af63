// Host-only part of the C++ drop-in layer: the scalar helpers that never touch the device (they stay
// host-side restatements).  Runs without a GPU; prints "ok" or the first failed check.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "icpb200/icp.hpp"
#include "icpb200/map.hpp"
#include "icpb200/pointcloud.hpp"
#include "icpb200/quaternion.hpp"

// test hook of the drop-in layer (host/compat.cpp): the subsample decisions of pointcloud.cpp:22-28 for one depth image
extern "C" int icpb_compat_draw_keep(const uint16_t *depth, int n_px, int factor, uint8_t *keep);

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static color_point_t cp(float x, float y, float z)
{
    color_point_t p;
    p.point = cv::Point3f(x, y, z);
    p.color = cv::Vec3b(1, 2, 3);
    return p;
}

// "dump" mode: the helpers on a seeded sweep, one value per line, for the comparison against the reference's own
// compiled functions (tests/test_compat_host.py, Route B)
static int dump()
{
    unsigned int st = 12345u;
    auto rnd = [&st]() { st = st * 1664525u + 1013904223u; return (float)(st >> 8) / 16777216.0f; };
    for (int k = 0; k < 64; ++k) {
        const float ax = rnd() * 360.f - 180.f, ay = rnd() * 360.f - 180.f, az = rnd() * 360.f - 180.f;
        cv::Mat r = icp::makeRotationMatrix(ax, ay, az);
        std::printf("R %.9g %.9g %.9g", ax, ay, az);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) std::printf(" %.9g", r.at<float>(i, j));
        std::printf("\n");
    }
    map::Map &m = icp::mapState();
    for (int k = 0; k < 256; ++k) {
        const cv::Point3f a(rnd() * 12.f - 1.f, rnd() * 12.f - 1.f, rnd() * 12.f - 1.f), b(rnd() * 10.f, rnd() * 10.f, rnd() * 10.f);
        const cv::Point3i v = m.getVoxelCoordinates(a);
        color_point_t ca = cp(a.x, a.y, a.z), cb = cp(b.x, b.y, b.z);
        std::printf("P %.9g %.9g %.9g %.9g %.9g %.9g %d %d %d %.9g\n", a.x, a.y, a.z, b.x, b.y, b.z, v.x, v.y, v.z,
                    icp::distance(ca, cb));
    }
    for (int n = 1; n < 5; ++n) { // calculateOffset (icp.cpp:314-344): the sequential float sums of the reference
        associations_t as;
        std::printf("O %d", n * 53);
        for (int i = 0; i < n * 53; ++i) {
            const color_point_t a = cp(rnd() * 10.f, rnd() * 10.f, rnd() * 10.f), b = cp(rnd() * 10.f, rnd() * 10.f, rnd() * 10.f);
            as.push_back(std::make_pair(a, b));
            std::printf(" %.9g %.9g %.9g %.9g %.9g %.9g", a.point.x, a.point.y, a.point.z, b.point.x, b.point.y, b.point.z);
        }
        const cv::Point3f o = icp::calculateOffset(as);
        std::printf(" %.9g %.9g %.9g\n", o.x, o.y, o.z);
    }
    for (int n = 0; n < 6; ++n) {
        std::vector<float> e;
        for (int i = 0; i < n * 37; ++i) e.push_back(rnd() * 0.75f);
        std::printf("E %d %.9g", (int)e.size(), icp::meanSquareError(e));
        for (float x : e) std::printf(" %.9g", x);
        std::printf("\n");
    }
    return 0;
}

// The subsample draws (pointcloud.cpp:22-28): the drop-in makes them in bulk on the generator's own state array; the
// decisions AND the stream the caller sees afterwards must be those of one rand() per non-zero pixel.
static int check_draws(int n_px, int zero_every, int factor)
{
    std::vector<uint16_t> depth((size_t)n_px);
    for (int i = 0; i < n_px; ++i) depth[(size_t)i] = (zero_every && i % zero_every == 0) ? 0 : (uint16_t)(1 + i % 4000);
    // where the generator stands now is unknown to this function: remember it by cloning the caller-visible stream
    const unsigned seed = 100u + (unsigned)n_px;
    srand(seed);
    std::vector<uint8_t> want;
    for (int i = 0; i < n_px; ++i)
        if (depth[(size_t)i] != 0) want.push_back((rand() % factor) == 0);
    const int w0 = rand(), w1 = rand(), w2 = rand();
    srand(seed);
    std::vector<uint8_t> got((size_t)n_px + 1, 0xff);
    const int m = icpb_compat_draw_keep(depth.data(), n_px, factor, got.data());
    CHECK(m == (int)want.size());
    for (int i = 0; i < m; ++i) CHECK(got[(size_t)i] == want[(size_t)i]);
    CHECK(got[(size_t)m] == 0xff);
    CHECK(rand() == w0 && rand() == w1 && rand() == w2);
    return 0;
}

static int check_draws_all()
{
    if (check_draws(640 * 480, 20, SUBSAMPLE_FACTOR)) return 1; // a Kinect v1 frame with holes
    if (check_draws(512 * 424, 0, SUBSAMPLE_FACTOR)) return 1;
    if (check_draws(1000, 3, 40) || check_draws(63, 0, 40) || check_draws(0, 0, 40)) return 1; // below and above the bulk threshold
    if (check_draws(31 * 7 + 5, 0, 3) || check_draws(100, 0, 40) || check_draws(8192 * 2 + 1, 0, 40)) return 1; // block and unroll tails
    // the other generator types a caller may have selected with initstate(): 8 / 32 / 64 / 256-byte state blocks
    alignas(8) static char blocks[4][256];
    const size_t sizes[4] = {8, 32, 64, 256};
    for (int k = 0; k < 4; ++k) {
        char *before = initstate(7u + (unsigned)k, blocks[k], sizes[k]);
        const int rc = check_draws(5000 + k, 5, SUBSAMPLE_FACTOR);
        setstate(before);
        if (rc) { std::printf("generator type %d\n", k); return 1; }
    }
    return 0;
}

int main(int argc, char **argv)
{
    if (argc > 1 && std::string(argv[1]) == "dump") return dump();
    if (check_draws_all()) return 1;
    // distance (icp.cpp:595-620), meanSquareError (:622-638), calculateOffset (:314-344)
    CHECK(icp::distance(cv::Point3f(0, 0, 0), cv::Point3f(3, 4, 0)) == 5.0f);
    CHECK(icp::distance(cp(1, 1, 1), cp(1, 1, 3)) == 2.0f);
    std::vector<float> errs = {1.f, 3.f};
    CHECK(icp::meanSquareError(errs) == 4.0f);
    CHECK(icp::meanSquareError(std::vector<float>()) == 0.0f);
    associations_t as;
    as.push_back(std::make_pair(cp(1, 2, 3), cp(0, 0, 0)));
    as.push_back(std::make_pair(cp(3, 2, 1), cp(0, 0, 0)));
    const cv::Point3f off = icp::calculateOffset(as);
    CHECK(off.x == 2.f && off.y == 2.f && off.z == 2.f);

    // makeRotationMatrix (:640-653): degrees, Rx * Ry * Rz with the reference's sign layout
    cv::Mat r = icp::makeRotationMatrix(90.f, 0.f, 0.f);
    CHECK(std::fabs(r.at<float>(1, 2) - 1.f) < 1e-6f && std::fabs(r.at<float>(2, 1) + 1.f) < 1e-6f);
    CHECK(r.at<float>(0, 0) == 1.f);

    // getVoxelCoordinates (map.cpp:55-85): truncation and clamping
    map::Map &m = icp::mapState();
    const float c = float(CELL_PHYSICAL_HEIGHT);
    cv::Point3i v = m.getVoxelCoordinates(cv::Point3f(5.f, 5.f, 5.f));
    CHECK(v.x == int(5.f / c) && v.y == v.x && v.z == v.x);
    v = m.getVoxelCoordinates(cv::Point3f(-1.f, 100.f, 0.f));
    CHECK(v.x == 0 && v.y == MAP_HEIGHT - 1 && v.z == 0);

    // processVoxel (icp.cpp:476-486): the entry of a voxel against the running best
    m.mapCloud.points.clear();
    m.mapCloud.points.push_back(cp(5.01f, 5.01f, 5.01f));
    m.mapCloud.points.push_back(cp(6.0f, 6.0f, 6.0f));
    const cv::Point3i v0 = m.getVoxelCoordinates(m.mapCloud.points[0].point);
    const cv::Point3i v1 = m.getVoxelCoordinates(m.mapCloud.points[1].point);
    color_point_t nearest = cp(0, 0, 0);
    float best = MAX_NN_COLOR_DISTANCE;
    icp::processVoxel(cp(5.0f, 5.0f, 5.0f), nearest, best, v0.x, v0.y, v0.z);
    CHECK(best == icp::distance(cp(5.0f, 5.0f, 5.0f), m.mapCloud.points[0]));
    CHECK(nearest == m.mapCloud.points[0]);
    const float kept = best;
    icp::processVoxel(cp(5.0f, 5.0f, 5.0f), nearest, best, v1.x, v1.y, v1.z); // farther: the best stays
    CHECK(best == kept && nearest == m.mapCloud.points[0]);
    icp::processVoxel(cp(5.0f, 5.0f, 5.0f), nearest, best, 1, 2, 3);          // empty voxel: nothing happens
    CHECK(best == kept);
    m.mapCloud.points.clear();
    // with an empty map cloud nothing is closer than the acceptance radius (icp.cpp:379,474) -- and no device is touched
    CHECK(icp::getNearestMappedPoint(cp(5, 5, 5), nearest) == MAX_NN_COLOR_DISTANCE);

    // pose reporting (quaternion.cpp:23-79, SLAM.cpp:613-648): identity
    float eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    Quaternion q(cv::Mat(3, 3, CV_32FC1, eye));
    CHECK(std::fabs(q.w - 1.f) < 1e-6f && std::fabs(q.x) < 1e-6f);

    std::printf("ok\n");
    return 0;
}

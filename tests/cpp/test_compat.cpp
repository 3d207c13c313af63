// Exercises the drop-in headers (reference names, reference call shapes) end to end on a GPU.
// usage: test_compat <in.bin> <out_dir>   -- driven by tests/test_gpu_compat.py, which checks every output
// against the CPU oracle.  in.bin: int32 w, h; u16 depth_prev[h*w]; u16 depth_cur[h*w]; u8 bgr[h*w*3].
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "icpb200/icp.hpp"
#include "icpb200/map.hpp"
#include "icpb200/pointcloud.hpp"
#include "icpb200/quaternion.hpp"

static void dump(const std::string &path, const void *p, size_t bytes)
{
    std::ofstream f(path, std::ios::binary);
    f.write((const char *)p, (std::streamsize)bytes);
}

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    std::ifstream in(argv[1], std::ios::binary);
    int w = 0, h = 0;
    in.read((char *)&w, 4); in.read((char *)&h, 4);
    cv::Mat prev(h, w, CV_16UC1), cur(h, w, CV_16UC1), bgr(h, w, CV_8UC3);
    in.read((char *)prev.data, (std::streamsize)w * h * 2);
    in.read((char *)cur.data, (std::streamsize)w * h * 2);
    in.read((char *)bgr.data, (std::streamsize)w * h * 3);
    if (!in) return 3;
    const std::string out = argv[2];
    cv::viz::Viz3d win;

    // P1: constructors consume libc rand() like the reference (pointcloud.cpp:125)
    srand(1);
    icp::PointCloud data(cur, bgr);
    icp::PointCloud target(prev, bgr);
    dump(out + "/data_pts.bin", data.points.data(), data.points.size() * sizeof(color_point_t));
    dump(out + "/target_pts.bin", target.points.data(), target.points.size() * sizeof(color_point_t));
    float centers[6] = {data.center.x, data.center.y, data.center.z, target.center.x, target.center.y, target.center.z};
    dump(out + "/centers.bin", centers, sizeof(centers));

    // P2
    cv::Mat R = icp::makeRotationMatrix(3.f, -2.f, 1.f);
    dump(out + "/rotation.bin", R.data, 9 * sizeof(float));
    data.rotate(R);
    data.translate(cv::Point3f(5, 5, 5));
    target.translate(cv::Point3f(5, 5, 5));
    dump(out + "/data_moved.bin", data.points.data(), data.points.size() * sizeof(color_point_t));

    // N1-N3
    std::vector<float> errors;
    associations_t assoc;
    icp::findGlobalNearestNeighborAssociations(data, target, errors, assoc);
    dump(out + "/errors.bin", errors.data(), errors.size() * sizeof(float));
    std::vector<color_point_t> firsts, seconds;
    for (auto &pr : assoc) { firsts.push_back(pr.first); seconds.push_back(pr.second); }
    dump(out + "/assoc_first.bin", firsts.data(), firsts.size() * sizeof(color_point_t));
    dump(out + "/assoc_second.bin", seconds.data(), seconds.size() * sizeof(color_point_t));
    color_point_t nn;
    float d0 = icp::getNearestPoint(data.points[0], nn, target);
    cv::Point3f off = icp::calculateOffset(assoc);
    float scal[6] = {icp::meanSquareError(errors), off.x, off.y, off.z, d0, icp::distance(data.points[0], nn)};
    dump(out + "/scalars.bin", scal, sizeof(scal));

    // M2-M4
    map::Map m;
    m.update(assoc, DELTA_CONFIDENCE);
    m.update(assoc, DELTA_CONFIDENCE);
    m.integrateRays(data, cv::Point3f(5, 5, 5), DELTA_CONFIDENCE, DELTA_CONFIDENCE);
    {   // lookup-table / mapCloud bookkeeping through the reference's three update overloads
        map::Map m2;
        icp::PointCloud kp;
        kp.keypoints.assign(data.points.begin(), data.points.begin() + 1500);
        m2.update(kp, MAX_CONFIDENCE, win);                                   // map.cpp:220-269
        point_list_t non(target.points.begin(), target.points.begin() + 1200);
        for (int rep = 0; rep < 7; ++rep) m2.update(assoc, errors, non, DELTA_CONFIDENCE); // map.cpp:122-151
        dump(out + "/mapcloud_kp.bin", m2.mapCloud.keypoints.data(), m2.mapCloud.keypoints.size() * sizeof(color_point_t));
        {   // pointLookupTable (map.hpp:24): the voxel of every stored key-point leads back to a stored point of that voxel
            int bad = 0;
            for (const color_point_t &kp : m2.mapCloud.keypoints) {
                cv::Point3i v = m2.getVoxelCoordinates(kp.point);
                color_point_t got = m2.pointLookupTable[v.x][v.y][v.z];
                cv::Point3i gv = m2.getVoxelCoordinates(got.point);
                if (got == m2.empty || gv.x != v.x || gv.y != v.y || gv.z != v.z) ++bad;
            }
            color_point_t none = m2.pointLookupTable[0][0][0];
            int look[2] = {bad, (none == m2.empty) ? 1 : 0};
            dump(out + "/lookup.bin", look, sizeof(look));
        }
        m2.syncWorld();
        dump(out + "/world2.bin", m2.world, (size_t)MAP_HEIGHT * MAP_HEIGHT * MAP_HEIGHT);
    }
    m.syncWorld();
    dump(out + "/world.bin", m.world, (size_t)MAP_HEIGHT * MAP_HEIGHT * MAP_HEIGHT);
    cv::Point3i v = m.getVoxelCoordinates(data.points[0].point);
    int vox[4] = {v.x, v.y, v.z, m.isOccupied(data.points[0].point) ? 1 : 0};
    dump(out + "/voxel.bin", vox, sizeof(vox));

    // getTransformation twice (first call initialises the globals, icp.cpp:47-68)
    srand(7);
    cv::Mat rot;
    std::vector<cv::KeyPoint> kps;
    for (int i = 0; i < 50; ++i) { cv::KeyPoint k; k.pt = cv::Point2f((float)(20 + (i * 37) % (w - 40)), (float)(20 + (i * 53) % (h - 40))); kps.push_back(k); }
    cv::Mat T1 = icp::getTransformation(cur, prev, bgr, kps, rot, 3, 0.f, win);
    cv::Mat T2 = icp::getTransformation(prev, cur, bgr, kps, rot, 3, 0.f, win);
    dump(out + "/T1.bin", T1.data, 16 * sizeof(float));
    dump(out + "/T2.bin", T2.data, 16 * sizeof(float));
    cv::Mat cr = icp::cameraRotationState();
    cv::Point3f cp = icp::cameraPositionState();
    float pose[12] = {0};
    for (int k = 0; k < 9; ++k) pose[k] = cr.at<float>(k / 3, k % 3);
    pose[9] = cp.x; pose[10] = cp.y; pose[11] = cp.z;
    dump(out + "/pose.bin", pose, sizeof(pose));
    {   // the map after two all-point frames: the registered key-points must land where the points went
        map::Map &am = icp::mapState();
        dump(out + "/ap_mapkp.bin", am.mapCloud.keypoints.data(), am.mapCloud.keypoints.size() * sizeof(color_point_t));
        am.syncWorld();
        dump(out + "/ap_world.bin", am.world, (size_t)MAP_HEIGHT * MAP_HEIGHT * MAP_HEIGHT);
    }
    // pose reporting exactly as the frame loop writes it (SLAM.cpp:284-293)
    {
        Quaternion rotationQ = Quaternion(cr);
        float e[10];
        toEulerianAngle(rotationQ, e[0], e[1], e[2]);
        transformationMatToEulerianAngle(cr, e[3], e[4], e[5]);
        Quaternion d = rotationQ * rotationQ.inverse();
        e[6] = d.w; e[7] = d.x; e[8] = d.y; e[9] = d.z;
        dump(out + "/euler.bin", e, sizeof(e));
    }
    // the loop as the reference runs it (key-points against the growing map cloud), two frames from a fresh state
    {
        icp::resetState();
        icp::setAssociationMode(icp::ASSOCIATE_KEYPOINTS);
        std::vector<cv::KeyPoint> kp2;
        for (int i = 0; i < 700; ++i) {
            int x = 10 + (i * 97) % (w - 20), y = 10 + (i * 61) % (h - 20);
            if (prev.at<uint16_t>(y, x) == 0 || cur.at<uint16_t>(y, x) == 0) continue;
            cv::KeyPoint k; k.pt = cv::Point2f((float)x, (float)y); kp2.push_back(k);
        }
        srand(11);
        cv::Mat L1 = icp::getTransformation(cur, prev, bgr, kp2, rot, 16, 1e-4f, win);
        cv::Mat L2 = icp::getTransformation(prev, cur, bgr, kp2, rot, 16, 1e-4f, win);
        dump(out + "/L1.bin", L1.data, 16 * sizeof(float));
        dump(out + "/L2.bin", L2.data, 16 * sizeof(float));
        cv::Mat lr = icp::cameraRotationState();
        cv::Point3f lp = icp::cameraPositionState();
        float lpose[12];
        for (int k = 0; k < 9; ++k) lpose[k] = lr.at<float>(k / 3, k % 3);
        lpose[9] = lp.x; lpose[10] = lp.y; lpose[11] = lp.z;
        dump(out + "/live_pose.bin", lpose, sizeof(lpose));
        std::vector<float> kxy;
        for (const cv::KeyPoint &k : kp2) { kxy.push_back(k.pt.x); kxy.push_back(k.pt.y); }
        dump(out + "/live_kxy.bin", kxy.data(), kxy.size() * sizeof(float));
        map::Map &lm = icp::mapState();
        dump(out + "/live_mapkp.bin", lm.mapCloud.keypoints.data(), lm.mapCloud.keypoints.size() * sizeof(color_point_t));
        lm.syncWorld();
        dump(out + "/live_world.bin", lm.world, (size_t)MAP_HEIGHT * MAP_HEIGHT * MAP_HEIGHT);
        icp::setAssociationMode(icp::ASSOCIATE_ALL_POINTS);
    }
    printf("\ncompat ok: n_data=%zu n_target=%zu assoc=%zu\n", data.points.size(), target.points.size(), assoc.size());
    return 0;
}

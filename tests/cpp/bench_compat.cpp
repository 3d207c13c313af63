// Times the drop-in C++ layer the way SLAM.cpp:277 calls it: icp::getTransformation(depth, previous, colour,
// key-points, rotation, 16, 1e-4, window) once per frame, in both association modes.
// usage: bench_compat <in.bin> <frames>   -- in.bin: int32 w, h, n_frames, n_kp; u16 depth[n_frames][h*w];
// u8 bgr[h*w*3]; float kxy[n_kp][2].  Prints one line: "compat_bench all_points_ms=<per frame> keypoints_ms=<per frame>".
// The reference's own getTransformation on the same frames is timed by bench_extra.run_live (oracle/_ref).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <vector>

#include "icpb200/icp.hpp"
#include "icpb200/map.hpp"
#include "icpb200/pointcloud.hpp"

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    std::ifstream in(argv[1], std::ios::binary);
    int w = 0, h = 0, nf = 0, nk = 0;
    in.read((char *)&w, 4); in.read((char *)&h, 4); in.read((char *)&nf, 4); in.read((char *)&nk, 4);
    std::vector<cv::Mat> depth;
    for (int f = 0; f < nf; ++f) {
        cv::Mat d(h, w, CV_16UC1);
        in.read((char *)d.data, (std::streamsize)w * h * 2);
        depth.push_back(d);
    }
    cv::Mat bgr(h, w, CV_8UC3);
    in.read((char *)bgr.data, (std::streamsize)w * h * 3);
    std::vector<float> kxy((size_t)nk * 2);
    in.read((char *)kxy.data(), (std::streamsize)kxy.size() * 4);
    if (!in) return 3;
    std::vector<cv::KeyPoint> kps;
    for (int i = 0; i < nk; ++i) { cv::KeyPoint k; k.pt = cv::Point2f(kxy[2 * i], kxy[2 * i + 1]); kps.push_back(k); }
    cv::viz::Viz3d win;
    cv::Mat rot;
    // the reference prints the MSE and the map size per frame (icp.cpp:264,279): keep that out of the timing's way
    std::ostringstream sink;
    std::streambuf *old = std::cout.rdbuf(sink.rdbuf());
    double ms[2] = {0, 0};
    for (int mode = 0; mode < 2; ++mode) {
        icp::resetState();
        icp::setAssociationMode(mode == 0 ? icp::ASSOCIATE_ALL_POINTS : icp::ASSOCIATE_KEYPOINTS);
        srand(100);
        icp::getTransformation(depth[1], depth[0], bgr, kps, rot, 16, 1e-4f, win); // first call seeds the map (icp.cpp:47-68)
        const auto t0 = std::chrono::steady_clock::now();
        for (int f = 2; f < nf; ++f) icp::getTransformation(depth[f], depth[f - 1], bgr, kps, rot, 16, 1e-4f, win);
        const auto t1 = std::chrono::steady_clock::now();
        ms[mode] = std::chrono::duration<double, std::milli>(t1 - t0).count() / (nf - 2);
    }
    std::cout.rdbuf(old);
    printf("compat_bench all_points_ms=%.4f keypoints_ms=%.4f frames=%d keypoints=%d\n", ms[0], ms[1], nf - 2, nk);
    return 0;
}

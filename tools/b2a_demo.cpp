// b2a_demo.cpp -- the reference's two calls (src/aruco_slam.cpp:313-314) through the C++ shim, no ROS:
//   ./b2a_demo frame.pgm [dict_id=10] [marker_length=0.27] [slam]
// reads a binary PGM (P5, 8 bit), runs detectMarkers + estimatePoseSingleMarkers on GPU 0 and prints
// ids, corners and poses.  With a fourth argument "slam" the frame goes three times through b2a::ArucoSlam
// (the reference's class interface: addEncoder / addImage / toRosPose / toRosMappedMarkers / toRosDetectedMarkers /
// getMarkedImg) with fixed encoder readings and time steps, and the records are printed.  Build:  g++ -std=c++17 -Iinclude tools/b2a_demo.cpp -Laruco_slam_b200/csrc -lb2aruco -o b2a_demo
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "b2aruco.hpp"

static bool read_pgm(const char *path, std::vector<uint8_t> &px, int &w, int &h)
{
    std::ifstream f(path, std::ios::binary);
    std::string magic;
    int maxv = 0;
    if (!(f >> magic) || magic != "P5") return false;
    auto skip = [&]() { while (f >> std::ws && f.peek() == '#') { std::string l; std::getline(f, l); } };
    skip(); f >> w; skip(); f >> h; skip(); f >> maxv;
    f.get();
    if (!f || w <= 0 || h <= 0 || maxv != 255) return false;
    px.resize((size_t)w * h);
    f.read((char *)px.data(), (std::streamsize)px.size());
    return (bool)f;
}

int main(int argc, char **argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s frame.pgm [dict_id] [marker_length]\n", argv[0]); return 2; }
    std::vector<uint8_t> px;
    int w = 0, h = 0;
    if (!read_pgm(argv[1], px, w, h)) { std::fprintf(stderr, "cannot read %s as 8-bit P5\n", argv[1]); return 2; }
    const int dict_id = argc > 2 ? std::atoi(argv[2]) : b2a::aruco::DICT_6X6_250;
    const float len = argc > 3 ? (float)std::atof(argv[3]) : 0.27f;
    if (argc > 4 && std::string(argv[4]) == "slam") {
        try {
            b2a::ArucoSlamIniteData init;
            init.markers_dictionary = dict_id; init.marker_length = len; init.image_width = w; init.image_height = h;
            init.r2c_translation[0] = 0.12; init.r2c_translation[2] = 0.25;
            init.R_x = 0.1; init.R_y = 0.1; init.R_theta = 0.01;          // observation noise small enough for integer-pixel corners to pass the covariance gate (:367-368)
            b2a::ArucoSlam slam(init);
            const double K[9] = {1400.0, 0, w / 2.0, 0, 1400.0, h / 2.0, 0, 0, 1};
            slam.setCameraParameters(K, {});
            slam.addEncoder(0.0, 0.0, 0.0);                              // the first message only starts the clock
            for (int k = 0; k < 3; ++k) {
                slam.addEncoder(2.0, 2.5, 0.05);
                slam.addImage(b2a::Image{px.data(), w, h, 1, 0});
            }
            const auto pose = slam.toRosPose();
            std::printf("pose %.9f %.9f %.9f q %.9f %.9f %.9f %.9f cov %.9e %.9e %.9e\n", pose.position[0], pose.position[1], pose.position[2], pose.orientation[0],
                        pose.orientation[1], pose.orientation[2], pose.orientation[3], pose.covariance[0], pose.covariance[7], pose.covariance[35]);
            for (const auto &m : slam.toRosMappedMarkers()) std::printf("mapped %d %.9f %.9f %.9f\n", m.id, m.x, m.y, m.yaw);
            for (const auto &m : slam.toRosDetectedMarkers()) std::printf("detected %d %.9f %.9f %.9f q %.9f %.9f %.9f %.9f\n", m.id, m.x, m.y, m.z, m.q[0], m.q[1], m.q[2], m.q[3]);
            const auto img = slam.getMarkedImg();
            unsigned long long sum = 0;
            for (uint8_t v : img.data) sum += v;
            std::printf("marked %d x %d x %d sum %llu dim %zu\n", img.cols, img.rows, img.channels, sum, slam.mu().size());
        } catch (const b2a::Exception &e) {
            std::fprintf(stderr, "b2aruco error %d: %s\n", e.code, e.what());
            return 1;
        }
        return 0;
    }
    try {
        auto dict = b2a::aruco::getPredefinedDictionary(dict_id);
        std::vector<std::vector<b2a::Point2f>> corners, rejected;
        std::vector<int> ids;
        b2a::aruco::detectMarkers(b2a::Image{px.data(), w, h, 1, 0}, dict, corners, ids, b2a::aruco::DetectorParameters::create(), &rejected);
        const double K[9] = {1400.0, 0, w / 2.0, 0, 1400.0, h / 2.0, 0, 0, 1};
        std::vector<b2a::Vec3d> rvecs, tvecs;
        b2a::aruco::estimatePoseSingleMarkers(corners, len, K, {}, rvecs, tvecs);
        std::printf("%zu markers, %zu rejected candidates\n", ids.size(), rejected.size());
        for (size_t i = 0; i < ids.size(); ++i) {
            std::printf("id %d corners", ids[i]);
            for (const auto &p : corners[i]) std::printf(" (%.1f,%.1f)", p.x, p.y);
            std::printf(" rvec (%.5f %.5f %.5f) tvec (%.5f %.5f %.5f)\n", rvecs[i][0], rvecs[i][1], rvecs[i][2], tvecs[i][0], tvecs[i][1], tvecs[i][2]);
        }
    } catch (const b2a::Exception &e) {
        std::fprintf(stderr, "b2aruco error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}

// b2aruco.hpp -- header-only C++ shim over the C ABI (b2aruco.h) with the call surface the
// reference uses at src/aruco_slam.cpp:11-12,313-314:
//
//     cv::aruco::getPredefinedDictionary(int)                                   -> b2a::aruco::getPredefinedDictionary
//     cv::aruco::detectMarkers(img, dictionary, corners, ids[, params, rejected]) -> b2a::aruco::detectMarkers
//     cv::aruco::estimatePoseSingleMarkers(corners, len, K, D, rvecs, tvecs)     -> b2a::aruco::estimatePoseSingleMarkers
//
// Same argument order and meaning, same output container shapes
// (std::vector<std::vector<Point2f>>, std::vector<int>, std::vector<Vec3d>, aruco_slam.cpp:309-311),
// same error behaviour (an exception where OpenCV's CV_Assert would throw cv::Exception; zero
// detections is not an error).  OpenCV's headers are not available in this image, so cv::Mat /
// cv::Point2f / cv::Vec3d are replaced by the POD stand-ins below; with OpenCV present the
// adaptor is `Image{mat.data, mat.cols, mat.rows, mat.channels(), mat.step}` and a
// reinterpret_cast of the Point2f / Vec3d vectors (identical layouts).  INTEGRATION.md shows
// the three-line patch of aruco_slam.cpp.
//
// The detector handle (device memory, streams) is cached per (device, dictionary, frame size):
// the reference calls detectMarkers once per camera frame from a single thread
// (aruco_slam_node.cpp:79,96), so the first call creates the handle and later calls reuse it.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "b2aruco.h"

namespace b2a {

struct Point2f { float x, y; };
struct Vec3d { double v[3]; double &operator[](int i) { return v[i]; } const double &operator[](int i) const { return v[i]; } };

// 8-bit image, 1 (gray) or 3 (bgr8) channels; `step` = bytes per row (0 = cols*channels)
struct Image {
    const uint8_t *data = nullptr;
    int cols = 0, rows = 0, channels = 1;
    size_t step = 0;
    bool empty() const { return !data || cols <= 0 || rows <= 0; }
};

struct Exception : std::runtime_error {
    int code;
    Exception(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc)
{
    if (rc != B2A_OK) throw Exception(rc, b2a_last_error());
}

namespace aruco {

// cv::aruco::PREDEFINED_DICTIONARY_NAME values used by the reference / the benchmark configs
enum PREDEFINED_DICTIONARY_NAME { DICT_4X4_50 = 0, DICT_5X5_50 = 4, DICT_6X6_50 = 8, DICT_6X6_250 = 10, DICT_7X7_50 = 12, DICT_ARUCO_ORIGINAL = 16 };

struct Dictionary {
    b2a_dictionary c{};
    int id = -1;
    int markerSize() const { return c.markerSize; }
    int maxCorrectionBits() const { return c.maxCorrectionBits; }
};
using DictionaryPtr = std::shared_ptr<Dictionary>;

inline DictionaryPtr getPredefinedDictionary(int name)
{
    auto d = std::make_shared<Dictionary>();
    check(b2a_get_predefined_dictionary(name, &d->c));
    d->id = name;
    return d;
}

struct DetectorParameters : b2a_detector_params {
    DetectorParameters() { b2a_default_detector_params(this); }
    static std::shared_ptr<DetectorParameters> create() { return std::make_shared<DetectorParameters>(); }
};
using DetectorParametersPtr = std::shared_ptr<DetectorParameters>;

namespace detail {
struct Handle {
    b2a_detector *h = nullptr;
    DetectorParameters prm;
    ~Handle() { if (h) b2a_detector_destroy(h); }
};
inline int &device() { static int dev = 0; return dev; }
inline std::map<std::tuple<int, const void *, int, int>, std::unique_ptr<Handle>> &cache()
{
    static thread_local std::map<std::tuple<int, const void *, int, int>, std::unique_ptr<Handle>> c;
    return c;
}
inline bool same_params(const b2a_detector_params &a, const b2a_detector_params &b)
{
    return a.adaptiveThreshWinSizeMin == b.adaptiveThreshWinSizeMin && a.adaptiveThreshWinSizeMax == b.adaptiveThreshWinSizeMax &&
           a.adaptiveThreshWinSizeStep == b.adaptiveThreshWinSizeStep && a.adaptiveThreshConstant == b.adaptiveThreshConstant &&
           a.minMarkerPerimeterRate == b.minMarkerPerimeterRate && a.maxMarkerPerimeterRate == b.maxMarkerPerimeterRate &&
           a.polygonalApproxAccuracyRate == b.polygonalApproxAccuracyRate && a.minCornerDistanceRate == b.minCornerDistanceRate &&
           a.minDistanceToBorder == b.minDistanceToBorder && a.minMarkerDistanceRate == b.minMarkerDistanceRate &&
           a.minGroupDistance == b.minGroupDistance && a.markerBorderBits == b.markerBorderBits &&
           a.perspectiveRemovePixelPerCell == b.perspectiveRemovePixelPerCell &&
           a.perspectiveRemoveIgnoredMarginPerCell == b.perspectiveRemoveIgnoredMarginPerCell &&
           a.maxErroneousBitsInBorderRate == b.maxErroneousBitsInBorderRate && a.minOtsuStdDev == b.minOtsuStdDev &&
           a.errorCorrectionRate == b.errorCorrectionRate && a.cornerRefinementMethod == b.cornerRefinementMethod &&
           a.cornerRefinementWinSize == b.cornerRefinementWinSize && a.relativeCornerRefinmentWinSize == b.relativeCornerRefinmentWinSize &&
           a.cornerRefinementMaxIterations == b.cornerRefinementMaxIterations && a.cornerRefinementMinAccuracy == b.cornerRefinementMinAccuracy &&
           a.detectInvertedMarker == b.detectInvertedMarker && a.useAruco3Detection == b.useAruco3Detection &&
           a.minSideLengthCanonicalImg == b.minSideLengthCanonicalImg && a.minMarkerLengthRatioOriginalImg == b.minMarkerLengthRatioOriginalImg;
}
inline b2a_detector *handle_for(const Dictionary &dict, const DetectorParameters &prm, int cols, int rows)
{
    auto key = std::make_tuple(device(), (const void *)dict.c.table, cols, rows);
    auto &slot = cache()[key];
    if (slot && !same_params(slot->prm, prm)) slot.reset();
    if (!slot) {
        slot.reset(new Handle());
        slot->prm = prm;
        b2a_detector_config cfg{device(), cols, rows, 1, 0, 0};
        check(b2a_detector_create(&cfg, &dict.c, &prm, &slot->h));
    }
    return slot->h;
}
inline b2a_detector *&last_handle() { static thread_local b2a_detector *h = nullptr; return h; }
}  // namespace detail

// which CUDA device the shim's handles live on (default 0)
inline void setDevice(int dev) { detail::device() = dev; }

// cv::aruco::detectMarkers (aruco_slam.cpp:313).  `rejectedImgPoints` may be null (= cv::noArray()).
inline void detectMarkers(const Image &image, const DictionaryPtr &dictionary, std::vector<std::vector<Point2f>> &corners, std::vector<int> &ids,
                          const DetectorParametersPtr &parameters = DetectorParameters::create(),
                          std::vector<std::vector<Point2f>> *rejectedImgPoints = nullptr)
{
    if (image.empty()) throw Exception(B2A_ERR_INVALID, "detectMarkers: empty image");               // CV_Assert(!_image.empty())
    if (!dictionary) throw Exception(B2A_ERR_INVALID, "detectMarkers: null dictionary");
    b2a_detector *h = detail::handle_for(*dictionary, parameters ? *parameters : DetectorParameters(), image.cols, image.rows);
    detail::last_handle() = h;
    b2a_frames fr{image.data, 0, 1, image.cols, image.rows, image.channels, image.step, 0};
    b2a_detections det{};
    check(b2a_detect(h, &fr, &det));
    const int na = det.n_accepted[0], nr = det.n_rejected[0];
    corners.assign(na, std::vector<Point2f>(4));
    ids.assign(det.ids, det.ids + na);
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < 4; ++j) corners[i][j] = Point2f{det.corners[(i * 4 + j) * 2], det.corners[(i * 4 + j) * 2 + 1]};
    if (rejectedImgPoints) {
        rejectedImgPoints->assign(nr, std::vector<Point2f>(4));
        for (int i = 0; i < nr; ++i)
            for (int j = 0; j < 4; ++j) (*rejectedImgPoints)[i][j] = Point2f{det.rejected[(i * 4 + j) * 2], det.rejected[(i * 4 + j) * 2 + 1]};
    }
}

// cv::aruco::estimatePoseSingleMarkers (aruco_slam.cpp:314).  cameraMatrix: 9 doubles row-major
// (CV_64F 3x3, aruco_slam_node.cpp:121-130); distCoeffs: 0, 4 or 5 doubles.
inline void estimatePoseSingleMarkers(const std::vector<std::vector<Point2f>> &corners, float markerLength, const double *cameraMatrix,
                                      const std::vector<double> &distCoeffs, std::vector<Vec3d> &rvecs, std::vector<Vec3d> &tvecs)
{
    if (!(markerLength > 0)) throw Exception(B2A_ERR_INVALID, "estimatePoseSingleMarkers: markerLength must be > 0");   // CV_Assert(markerLength > 0)
    if (distCoeffs.size() != 0 && distCoeffs.size() != 4 && distCoeffs.size() != 5)
        throw Exception(B2A_ERR_UNSUPPORTED, "estimatePoseSingleMarkers: distCoeffs must hold 0, 4 or 5 values");
    const int n = (int)corners.size();
    rvecs.assign(n, Vec3d{});
    tvecs.assign(n, Vec3d{});
    if (n == 0) return;
    std::vector<float> flat((size_t)n * 8);
    for (int i = 0; i < n; ++i) {
        if (corners[i].size() != 4) throw Exception(B2A_ERR_INVALID, "estimatePoseSingleMarkers: a marker needs 4 corners");
        for (int j = 0; j < 4; ++j) { flat[(i * 4 + j) * 2] = corners[i][j].x; flat[(i * 4 + j) * 2 + 1] = corners[i][j].y; }
    }
    b2a_camera cam{};
    for (int i = 0; i < 9; ++i) cam.K[i] = cameraMatrix[i];
    cam.nD = (int)distCoeffs.size();
    for (int i = 0; i < cam.nD; ++i) cam.D[i] = distCoeffs[i];
    cam.marker_length = markerLength;
    b2a_detector *h = detail::last_handle();
    std::unique_ptr<detail::Handle> tmp;
    if (!h) {                                   // pose without a preceding detect on this thread: a minimal handle
        tmp.reset(new detail::Handle());
        b2a_dictionary d{};
        check(b2a_get_predefined_dictionary(DICT_4X4_50, &d));
        b2a_detector_config cfg{detail::device(), 64, 64, 1, 0, 0};
        check(b2a_detector_create(&cfg, &d, nullptr, &tmp->h));
        h = tmp->h;
    }
    check(b2a_estimate_pose_single_markers(h, flat.data(), n, &cam, &rvecs[0].v[0], &tvecs[0].v[0]));
}

}  // namespace aruco

// ---- wire / on-disk formats (reference src/map_loader.cpp, ArucoSlam::toRosPose) as plain records ----
// MapLoader::loadMap(file_path): the cubes a map file defines, under the reference loader's acceptance rules (see b2aruco.h).
// Like the reference, an unreadable file gives an empty map instead of an error.
inline std::vector<b2a_map_marker> loadMap(const std::string &file_path)
{
    std::vector<b2a_map_marker> out(256);
    int n = 0;
    int rc = b2a_map_load(file_path.c_str(), out.data(), (int)out.size(), &n);
    if (rc == B2A_ERR_CAPACITY) { out.resize((size_t)n); rc = b2a_map_load(file_path.c_str(), out.data(), (int)out.size(), &n); }
    if (rc != B2A_OK) return {};
    out.resize((size_t)n);
    return out;
}
// ArucoSlam::toRosPose(): position, yaw quaternion and the 6x6 covariance packing of the filter's robot state
inline b2a_pose_with_covariance robotPose(b2a_slam *slam)
{
    b2a_pose_with_covariance p;
    check(b2a_slam_robot_pose(slam, &p));
    return p;
}
}  // namespace b2a

// b2aruco.hpp -- header-only C++ shim over the C ABI (b2aruco.h) with the call surface the
// reference uses at src/aruco_slam.cpp:11-12,313-314:
//
//     cv::aruco::getPredefinedDictionary(int)                                   -> b2a::aruco::getPredefinedDictionary
//     cv::aruco::detectMarkers(img, dictionary, corners, ids[, params, rejected]) -> b2a::aruco::detectMarkers
//     cv::aruco::estimatePoseSingleMarkers(corners, len, K, D, rvecs, tvecs)     -> b2a::aruco::estimatePoseSingleMarkers
//     cv::aruco::drawDetectedMarkers(image, corners, ids)            (:318-319)  -> b2a::aruco::drawDetectedMarkers
//
// Same argument order and meaning, same output container shapes
// (std::vector<std::vector<Point2f>>, std::vector<int>, std::vector<Vec3d>, aruco_slam.cpp:309-311),
// same error behaviour (an exception where OpenCV's CV_Assert would throw cv::Exception; zero
// detections is not an error).  OpenCV's headers are not available in this image, so cv::Mat /
// cv::Point2f / cv::Vec3d are replaced by the POD stand-ins below; with OpenCV present the
// adaptor is `Image{mat.data, mat.cols, mat.rows, mat.channels(), mat.step}` and a
// reinterpret_cast of the Point2f / Vec3d vectors (identical layouts).  INTEGRATION.md shows
// the three-line patch of aruco_slam.cpp.  b2a::ArucoSlam (end of this file) is the reference's ArucoSlam class itself over the C
// ABI: ArucoSlam(init data), addEncoder, addImage, setCameraParameters, toRosMappedMarkers, toRosDetectedMarkers, toRosPose,
// getMarkedImg (include/aruco_slam/aruco_slam.h:101-152) with plain records where the reference returns ROS messages.
//
// The detector handle (device memory, streams) is cached per (device, dictionary, frame size):
// the reference calls detectMarkers once per camera frame from a single thread
// (aruco_slam_node.cpp:79,96), so the first call creates the handle and later calls reuse it.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstring>
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

// cv::aruco::drawDetectedMarkers(image, corners, ids, borderColor) (aruco_slam.cpp:318-319): the overlay is drawn into `image` in
// place.  `image` is a writable 8-bit image (1 or 3 channels) no larger than the frame of the last detectMarkers call on this
// thread, whose handle does the drawing; ids may be empty (no labels); borderColor = (b, g, r), default green.
struct MutableImage {
    uint8_t *data = nullptr;
    int cols = 0, rows = 0, channels = 1;
    size_t step = 0;
    bool empty() const { return !data || cols <= 0 || rows <= 0; }
};
inline void drawDetectedMarkers(const MutableImage &image, const std::vector<std::vector<Point2f>> &corners, const std::vector<int> &ids = {},
                                const uint8_t *borderColor = nullptr)
{
    if (image.empty()) throw Exception(B2A_ERR_INVALID, "drawDetectedMarkers: empty image");
    if (!ids.empty() && ids.size() != corners.size()) throw Exception(B2A_ERR_INVALID, "drawDetectedMarkers: ids and corners differ in length");   // CV_Assert
    b2a_detector *h = detail::last_handle();
    if (!h) throw Exception(B2A_ERR_INVALID, "drawDetectedMarkers: no detectMarkers call on this thread yet");
    const int n = (int)corners.size();
    std::vector<float> flat((size_t)n * 8);
    for (int i = 0; i < n; ++i) {
        if (corners[i].size() != 4) throw Exception(B2A_ERR_INVALID, "drawDetectedMarkers: a marker needs 4 corners");
        for (int j = 0; j < 4; ++j) { flat[(i * 4 + j) * 2] = corners[i][j].x; flat[(i * 4 + j) * 2 + 1] = corners[i][j].y; }
    }
    check(b2a_draw_detected_markers(h, image.data, image.cols, image.rows, image.channels, image.step, flat.data(), ids.empty() ? nullptr : ids.data(), n, borderColor));
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
// ---- the reference's ArucoSlam class (include/aruco_slam/aruco_slam.h:101-193) over the C ABI: same public methods, no ROS types ----
// ArucoSlamIniteData (aruco_slam.h:40-60) without the frame / topic / file names the node consumes itself; the robot-to-camera
// transform is a plain (rotation quaternion x y z w, translation) pair; image_width / image_height / max_landmarks size the buffers.
struct ArucoSlamIniteData {
    double Q_k = 0.01, R_x = 100.0, R_y = 100.0, R_theta = 10.0;      // parameters.yaml:5-8
    double kl = 0.05, kr = 0.05, b = 0.09;                            // parameters.yaml:11-13
    int markers_dictionary = aruco::DICT_ARUCO_ORIGINAL;
    double marker_length = 0.27;
    double r2c_rotation[4] = {0, 0, 0, 1};
    double r2c_translation[3] = {0, 0, 0};
    float USEFUL_DISTANCE_THRESHOLD = 3;
    int image_width = 640, image_height = 480;
    int max_landmarks = 0;              // 0 = 512
    int device = 0;
};
// an owned 8-bit image (what getMarkedImg returns: cv::Mat markered_img_, aruco_slam.h:152)
struct OwnedImage {
    std::vector<uint8_t> data;
    int cols = 0, rows = 0, channels = 1;
    bool empty() const { return data.empty(); }
};

class ArucoSlam {
public:
    explicit ArucoSlam(const ArucoSlamIniteData &d) : init_(d)
    {
        b2a_slam_params p;
        b2a_default_slam_params(&p);
        p.Q_k = d.Q_k; p.R_x = d.R_x; p.R_y = d.R_y; p.R_theta = d.R_theta; p.kl = d.kl; p.kr = d.kr; p.b = d.b;
        p.r2c_tx = d.r2c_translation[0]; p.r2c_ty = d.r2c_translation[1];
        p.useful_distance_threshold = d.USEFUL_DISTANCE_THRESHOLD; p.max_landmarks = d.max_landmarks;
        check(b2a_slam_create(d.device, &p, &slam_));
        dictionary_ = aruco::getPredefinedDictionary(d.markers_dictionary);                 // aruco_slam.cpp:11-12
        b2a_detector_config cfg{d.device, d.image_width, d.image_height, 1, 0, 0};
        const int rc = b2a_detector_create(&cfg, &dictionary_->c, nullptr, &det_);
        if (rc != B2A_OK) { const std::string m = b2a_last_error(); b2a_slam_destroy(slam_); slam_ = nullptr; throw Exception(rc, m); }
    }
    ~ArucoSlam() { if (det_) b2a_detector_destroy(det_); if (slam_) b2a_slam_destroy(slam_); }
    ArucoSlam(const ArucoSlam &) = delete;
    ArucoSlam &operator=(const ArucoSlam &) = delete;

    // addEncoder(wl, wr) (aruco_slam.cpp:21-74): the time step is what a steady clock measured since the previous call, as the
    // reference takes it from ros::Time::now() (:26-32); the first call only starts the clock (:24-29).  The three-argument
    // form takes dt from the caller (replayed logs, tests).
    void addEncoder(const double &wl, const double &wr)
    {
        const auto now = std::chrono::steady_clock::now();
        const double dt = started_ ? std::chrono::duration<double>(now - last_time_).count() : 0.0;
        last_time_ = now; started_ = true;
        check(b2a_slam_add_encoder(slam_, wl, wr, dt));
    }
    void addEncoder(const double &wl, const double &wr, double dt) { started_ = true; last_time_ = std::chrono::steady_clock::now(); check(b2a_slam_add_encoder(slam_, wl, wr, dt)); }

    // addImage(img) (aruco_slam.cpp:76-263): detection, poses, observations and the EKF loop of one frame; the frame is kept for getMarkedImg
    void addImage(const Image &img)
    {
        if (img.empty()) throw Exception(B2A_ERR_INVALID, "addImage: empty image");
        if (!have_cam_) throw Exception(B2A_ERR_INVALID, "addImage: setCameraParameters first");
        b2a_frames fr{img.data, 0, 1, img.cols, img.rows, img.channels, img.step, 0};
        check(b2a_slam_add_image(slam_, det_, &fr, &cam_));
        const size_t row = (size_t)img.cols * img.channels, step = img.step ? img.step : row;
        last_.cols = img.cols; last_.rows = img.rows; last_.channels = img.channels;
        last_.data.resize(row * img.rows);
        for (int y = 0; y < img.rows; ++y) std::memcpy(last_.data.data() + y * row, img.data + y * step, row);
    }
    // setCameraParameters (aruco_slam.h:129-133): K 9 doubles row-major, dist 0, 4 or 5 doubles
    void setCameraParameters(const double *cameraMatrix, const std::vector<double> &distCoeffs)
    {
        if (distCoeffs.size() != 0 && distCoeffs.size() != 4 && distCoeffs.size() != 5) throw Exception(B2A_ERR_UNSUPPORTED, "setCameraParameters: distCoeffs must hold 0, 4 or 5 values");
        cam_ = b2a_camera{};
        for (int i = 0; i < 9; ++i) cam_.K[i] = cameraMatrix[i];
        cam_.nD = (int)distCoeffs.size();
        for (int i = 0; i < cam_.nD; ++i) cam_.D[i] = distCoeffs[i];
        cam_.marker_length = (float)init_.marker_length;
        have_cam_ = true;
    }
    // toRosMappedMarkers (detected_map_, :265-281): one cube per landmark of the map
    std::vector<b2a_map_marker> toRosMappedMarkers()
    {
        const int n = (b2a_slam_dim(slam_) - 3) / 3;
        std::vector<b2a_map_marker> out((size_t)(n > 0 ? n : 1));
        int cnt = 0;
        check(b2a_slam_detected_map(slam_, init_.marker_length, out.data(), (int)out.size(), &cnt));
        out.resize((size_t)cnt);
        return out;
    }
    // toRosDetectedMarkers (detected_markers_, :324-347): the cubes of the last frame's markers inside the useful range, robot frame
    std::vector<b2a_map_marker> toRosDetectedMarkers()
    {
        b2a_detections det{};
        if (b2a_detector_last_detections(det_, &det) != B2A_OK || !det.rvecs) return {};
        const int n = det.n_accepted[0];
        std::vector<b2a_map_marker> out((size_t)(n > 0 ? n : 1));
        int cnt = 0;
        check(b2a_pack_detected_markers(det.ids, det.rvecs, det.tvecs, n, init_.marker_length, init_.USEFUL_DISTANCE_THRESHOLD, init_.r2c_rotation,
                                        init_.r2c_translation, out.data(), (int)out.size(), &cnt));
        out.resize((size_t)cnt);
        return out;
    }
    // toRosPose (:376-407)
    b2a_pose_with_covariance toRosPose() { return robotPose(slam_); }
    // getMarkedImg (aruco_slam.h:152): the last frame with drawDetectedMarkers applied (:318-319)
    OwnedImage getMarkedImg()
    {
        OwnedImage out = last_;
        if (out.empty()) return out;
        b2a_detections det{};
        if (b2a_detector_last_detections(det_, &det) != B2A_OK) return out;     // addImage before the first encoder message does not look at the frame (:84-85)
        const int n = det.n_accepted[0];
        if (n > 0) check(b2a_draw_detected_markers(det_, out.data.data(), out.cols, out.rows, out.channels, 0, det.corners, det.ids, n, nullptr));
        return out;
    }
    // mu_ and sigma_ (aruco_slam.h:182-183), row-major
    std::vector<double> mu() { std::vector<double> m((size_t)b2a_slam_dim(slam_)); check(b2a_slam_get_state(slam_, m.data(), nullptr, nullptr)); return m; }
    std::vector<double> sigma() { const size_t n = (size_t)b2a_slam_dim(slam_); std::vector<double> s(n * n); check(b2a_slam_get_state(slam_, nullptr, s.data(), nullptr)); return s; }
    b2a_slam *handle() { return slam_; }
    b2a_detector *detector() { return det_; }

private:
    ArucoSlamIniteData init_;
    b2a_slam *slam_ = nullptr;
    b2a_detector *det_ = nullptr;
    aruco::DictionaryPtr dictionary_;
    b2a_camera cam_{};
    bool have_cam_ = false, started_ = false;
    std::chrono::steady_clock::time_point last_time_{};
    OwnedImage last_;
};

}  // namespace b2a

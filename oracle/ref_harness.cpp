// ref_harness.cpp -- harness around the reference's own src/aruco_slam.cpp + src/map_loader.cpp (compiled unmodified
// from /root/reference, see oracle/Makefile target _ref): the cv:: entry points declared in ref_stubs/b2a_cv_stub.h
// forward to hooks, the C interface of ref_backend.h drives ArucoSlam / MapLoader.
// TEST INFRASTRUCTURE ONLY (see oracle.h).
#include <cmath>
#include <cstring>
#include <map>
#include <queue>
#include <string>
#include <vector>

#include "ref_stubs/b2a_cv_stub.h"
#include "ref_stubs/b2a_eigen_stub.h"
#include "ref_stubs/b2a_ros_stub.h"
// the harness reads ArucoSlam's private state (mu_, sigma_, obs_, aruco_id_map, last_observed_marker_); access
// specifiers do not change the class layout, and the reference's own translation units are compiled without this
#define private public
#include "aruco_slam/aruco_slam.h"
#include "aruco_slam/map_loader.h"
#undef private

#include "oracle.h"
#include "ref_backend.h"

extern "C" double b2a_ref_clock = 0.0;

namespace {
ref_detect_fn g_detect = nullptr;
ref_pose_fn g_pose = nullptr;
ref_rodrigues_fn g_rodrigues = nullptr;
ref_project_fn g_project = nullptr;
orc_dict g_dict = {0, 0, 0, 0, nullptr};
std::vector<uint8_t> g_dict_table;
bool g_replay = false;
std::vector<float> g_rp_corners;
std::vector<int32_t> g_rp_ids;
std::vector<double> g_rp_rvecs, g_rp_tvecs;

void mat_to_cam(const cv::Mat &K, const cv::Mat &D, double *K9, double *D5, int *nD)
{
    for (int i = 0; i < 9; ++i) K9[i] = K.at<double>(i);
    const int n = (int)D.total();
    for (int i = 0; i < 5; ++i) D5[i] = i < n ? D.at<double>(i) : 0.0;
    *nD = n < 5 ? n : 5;
}
}  // namespace

// ---- the OpenCV calls of aruco_slam.cpp ----
namespace cv {
void Rodrigues(const Vec3d &rvec, Mat &R)
{
    double r[3] = {rvec[0], rvec[1], rvec[2]}, R9[9];
    if (g_rodrigues) g_rodrigues(r, R9); else orc_rodrigues(r, R9);
    R = Mat(3, 3, CV_64FC1);
    for (int i = 0; i < 9; ++i) R.at<double>(i) = R9[i];
}

void projectPoints(const std::vector<Point3f> &obj, const Vec3d &rvec, const Vec3d &tvec, const Mat &K, const Mat &D, std::vector<Point2f> &out)
{
    double K9[9], D5[5]; int nD;
    mat_to_cam(K, D, K9, D5, &nD);
    const int n = (int)obj.size();
    const double r[3] = {rvec[0], rvec[1], rvec[2]}, t[3] = {tvec[0], tvec[1], tvec[2]};
    std::vector<float> o((size_t)n * 3), img((size_t)n * 2);
    for (int i = 0; i < n; ++i) { o[3 * i] = obj[i].x; o[3 * i + 1] = obj[i].y; o[3 * i + 2] = obj[i].z; }
    if (g_project) g_project(o.data(), n, r, t, K9, D5, nD, img.data());
    else {
        std::vector<double> od(o.begin(), o.end()), id((size_t)n * 2);
        orc_project_points(od.data(), n, r, t, K9, D5, id.data());
        for (int i = 0; i < 2 * n; ++i) img[i] = (float)id[i];          // float object points give Point2f output
    }
    out.resize(n);
    for (int i = 0; i < n; ++i) out[i] = Point2f(img[2 * i], img[2 * i + 1]);
}

namespace aruco {
Ptr<Dictionary> getPredefinedDictionary(PREDEFINED_DICTIONARY_NAME name)
{
    auto d = std::make_shared<Dictionary>();
    d->predefined_id = (int)name;
    return d;
}

void detectMarkers(const Mat &image, const Ptr<Dictionary> &dictionary, std::vector<std::vector<Point2f>> &corners, std::vector<int> &ids)
{
    corners.clear(); ids.clear();
    std::vector<float> c; std::vector<int32_t> id;
    int n = 0;
    if (g_replay) { c = g_rp_corners; id = g_rp_ids; n = (int)id.size(); }
    else if (g_detect) {
        const int cap = 1024;
        c.resize((size_t)cap * 8); id.resize(cap);
        n = g_detect(image.ptr(), image.cols, image.rows, image.channels(), dictionary->predefined_id, c.data(), id.data(), cap);
    } else {
        orc_params p; orc_default_params(&p);
        orc_detections det; std::memset(&det, 0, sizeof(det));
        if (!g_dict.table) return;
        orc_detect(image.ptr(), image.cols, image.rows, image.channels(), &g_dict, &p, &det);
        n = det.n_acc;
        c.assign(det.corners, det.corners + (size_t)n * 8); id.assign(det.ids, det.ids + n);
        orc_free_detections(&det);
    }
    for (int i = 0; i < n; ++i) {
        std::vector<Point2f> q(4);
        for (int k = 0; k < 4; ++k) q[k] = Point2f(c[(size_t)i * 8 + 2 * k], c[(size_t)i * 8 + 2 * k + 1]);
        corners.push_back(q); ids.push_back(id[i]);
    }
}

void estimatePoseSingleMarkers(const std::vector<std::vector<Point2f>> &corners, float markerLength, const Mat &K, const Mat &D,
                               std::vector<Vec3d> &rvecs, std::vector<Vec3d> &tvecs)
{
    const int n = (int)corners.size();
    rvecs.assign(n, Vec3d()); tvecs.assign(n, Vec3d());
    if (!n) return;
    std::vector<double> r((size_t)n * 3), t((size_t)n * 3);
    if (g_replay) { r = g_rp_rvecs; t = g_rp_tvecs; }
    else {
        double K9[9], D5[5]; int nD;
        mat_to_cam(K, D, K9, D5, &nD);
        std::vector<float> c((size_t)n * 8);
        for (int i = 0; i < n; ++i) for (int k = 0; k < 4; ++k) { c[(size_t)i * 8 + 2 * k] = corners[i][k].x; c[(size_t)i * 8 + 2 * k + 1] = corners[i][k].y; }
        if (g_pose) g_pose(c.data(), n, markerLength, K9, D5, nD, r.data(), t.data());
        else orc_estimate_pose_single_markers(c.data(), n, (double)markerLength, K9, D5, nD, r.data(), t.data());
    }
    for (int i = 0; i < n; ++i) { rvecs[i] = Vec3d(r[3 * i], r[3 * i + 1], r[3 * i + 2]); tvecs[i] = Vec3d(t[3 * i], t[3 * i + 1], t[3 * i + 2]); }
}

void drawDetectedMarkers(Mat &, const std::vector<std::vector<Point2f>> &, const std::vector<int> &) {}   // overlay: not part of the state
}  // namespace aruco
}  // namespace cv

// ---- C interface ----
struct ref_slam { ArucoSlam *a; };

static cv::Mat make_image(const uint8_t *img, int w, int h, int channels)
{
    return cv::Mat(h, w, channels == 3 ? CV_8UC3 : CV_8UC1, img);
}

static int dump_markers(const visualization_msgs::MarkerArray &arr, int cap, int32_t *id, double *scale, double *position, double *orientation,
                        double *color, double *lifetime)
{
    const int n = (int)arr.markers.size();
    for (int i = 0; i < n && i < cap; ++i) {
        const visualization_msgs::Marker &m = arr.markers[i];
        if (id) id[i] = m.id;
        if (scale) { scale[3 * i] = m.scale.x; scale[3 * i + 1] = m.scale.y; scale[3 * i + 2] = m.scale.z; }
        if (position) { position[3 * i] = m.pose.position.x; position[3 * i + 1] = m.pose.position.y; position[3 * i + 2] = m.pose.position.z; }
        if (orientation) { orientation[4 * i] = m.pose.orientation.x; orientation[4 * i + 1] = m.pose.orientation.y; orientation[4 * i + 2] = m.pose.orientation.z; orientation[4 * i + 3] = m.pose.orientation.w; }
        if (color) { color[4 * i] = m.color.r; color[4 * i + 1] = m.color.g; color[4 * i + 2] = m.color.b; color[4 * i + 3] = m.color.a; }
        if (lifetime) lifetime[i] = m.lifetime.toSec();
    }
    return n;
}

extern "C" {

void ref_set_hooks(ref_detect_fn d, ref_pose_fn p, ref_rodrigues_fn r, ref_project_fn j) { g_detect = d; g_pose = p; g_rodrigues = r; g_project = j; }

void ref_set_dictionary(int markerSize, int maxCorrectionBits, int nMarkers, int nBytes, const uint8_t *table)
{
    g_dict_table.assign(table, table + (size_t)nMarkers * 4 * nBytes);
    g_dict.markerSize = markerSize; g_dict.maxCorrectionBits = maxCorrectionBits; g_dict.nMarkers = nMarkers; g_dict.nBytes = nBytes;
    g_dict.table = g_dict_table.data();
}

void ref_set_replay(const float *corners, const int32_t *ids, int n, const double *rvecs, const double *tvecs)
{
    g_replay = n >= 0;
    if (n < 0) n = 0;
    g_rp_corners.assign(corners, corners + (size_t)n * 8); g_rp_ids.assign(ids, ids + n);
    g_rp_rvecs.assign(rvecs, rvecs + (size_t)n * 3); g_rp_tvecs.assign(tvecs, tvecs + (size_t)n * 3);
}

void ref_set_clock(double t) { b2a_ref_clock = t; }

ref_slam *ref_create(const ref_init *in)
{
    ArucoSlamIniteData d;
    d.Q_k = in->Q_k; d.R_x = in->R_x; d.R_y = in->R_y; d.R_theta = in->R_theta;
    d.kl = in->kl; d.kr = in->kr; d.b = in->b;
    d.markers_dictionary = in->markers_dictionary; d.marker_length = in->marker_length;
    d.transformStamped_r2c.transform.translation.x = in->r2c_t[0];
    d.transformStamped_r2c.transform.translation.y = in->r2c_t[1];
    d.transformStamped_r2c.transform.translation.z = in->r2c_t[2];
    d.transformStamped_r2c.transform.rotation.x = in->r2c_q[0]; d.transformStamped_r2c.transform.rotation.y = in->r2c_q[1];
    d.transformStamped_r2c.transform.rotation.z = in->r2c_q[2]; d.transformStamped_r2c.transform.rotation.w = in->r2c_q[3];
    d.USEFUL_DISTANCE_THRESHOLD = in->useful_distance_threshold;
    ref_slam *s = new ref_slam;
    s->a = new ArucoSlam(d);
    return s;
}

void ref_destroy(ref_slam *s) { if (s) { delete s->a; delete s; } }

void ref_set_camera(ref_slam *s, const double *K9, const double *D, int nD)
{
    cv::Mat K(3, 3, CV_64FC1), Dm(nD, 1, CV_64FC1);
    for (int i = 0; i < 9; ++i) K.at<double>(i) = K9[i];
    for (int i = 0; i < nD; ++i) Dm.at<double>(i) = D[i];
    s->a->setCameraParameters(std::pair<cv::Mat, cv::Mat>(K, Dm));
}

void ref_add_encoder(ref_slam *s, double wl, double wr) { s->a->addEncoder(wl, wr); }

void ref_add_image(ref_slam *s, const uint8_t *img, int w, int h, int channels) { s->a->addImage(make_image(img, w, h, channels)); }

int ref_get_observations(ref_slam *s, const uint8_t *img, int w, int h, int channels, int cap, int32_t *ids, int32_t *index, double *xyt, double *cov)
{
    s->a->getObservations(make_image(img, w, h, channels));
    int n = 0;
    while (!s->a->obs_.empty()) {
        const ArucoMarker ob = s->a->obs_.top();
        s->a->obs_.pop();
        if (n < cap) {
            ids[n] = ob.aruco_id_; index[n] = ob.aruco_index_;
            xyt[3 * n] = ob.x_; xyt[3 * n + 1] = ob.y_; xyt[3 * n + 2] = ob.theta_;
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cov[9 * n + 3 * i + j] = ob.observe_covariance_(i, j);
        }
        ++n;
    }
    return n;
}

int ref_dim(const ref_slam *s) { return (int)s->a->mu_.rows(); }
int ref_is_init(const ref_slam *s) { return s->a->is_init_ ? 1 : 0; }

void ref_get_state(const ref_slam *s, double *mu, double *sigma)
{
    const int N = (int)s->a->mu_.rows();
    if (mu) for (int i = 0; i < N; ++i) mu[i] = s->a->mu_(i);
    if (sigma) for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) sigma[(size_t)i * N + j] = s->a->sigma_(i, j);
}

int ref_get_ids(const ref_slam *s, int32_t *ids_by_index, int cap)
{
    for (const auto &kv : s->a->aruco_id_map) if (kv.second >= 0 && kv.second < cap) ids_by_index[kv.second] = kv.first;
    return (int)s->a->aruco_id_map.size();
}

void ref_set_state(ref_slam *s, int N, const double *mu, const double *sigma, const int32_t *ids_by_index, int is_init)
{
    s->a->mu_.resize(N);
    s->a->sigma_.resize(N, N);
    for (int i = 0; i < N; ++i) s->a->mu_(i) = mu[i];
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) s->a->sigma_(i, j) = sigma[(size_t)i * N + j];
    s->a->aruco_id_map.clear();
    for (int k = 0; k < (N - 3) / 3; ++k) s->a->aruco_id_map.insert(std::pair<int, int>{ids_by_index[k], k});
    s->a->last_observed_marker_.clear();
    s->a->is_init_ = is_init != 0;
}

int ref_get_last_observed(const ref_slam *s, int cap, int32_t *ids, double *last_obs)
{
    const int n = (int)s->a->last_observed_marker_.size();
    for (int i = 0; i < n && i < cap; ++i) {
        ids[i] = s->a->last_observed_marker_[i].aruco_id_;
        for (int k = 0; k < 3; ++k) last_obs[3 * i + k] = s->a->last_observed_marker_[i].last_observation_(k);
    }
    return n;
}

void ref_robot_pose(ref_slam *s, double *position3, double *orientation4, double *covariance36)
{
    const geometry_msgs::PoseWithCovarianceStamped p = s->a->toRosPose();
    position3[0] = p.pose.pose.position.x; position3[1] = p.pose.pose.position.y; position3[2] = p.pose.pose.position.z;
    orientation4[0] = p.pose.pose.orientation.x; orientation4[1] = p.pose.pose.orientation.y; orientation4[2] = p.pose.pose.orientation.z; orientation4[3] = p.pose.pose.orientation.w;
    for (int i = 0; i < 36; ++i) covariance36[i] = p.pose.covariance[i];
}

int ref_markers(ref_slam *s, int which, int cap, int32_t *id, double *scale, double *position, double *orientation, double *color, double *lifetime)
{
    return dump_markers(which == 0 ? s->a->toRosMappedMarkers() : s->a->toRosDetectedMarkers(), cap, id, scale, position, orientation, color, lifetime);
}

int ref_map_load(const char *path, int cap, int32_t *id, double *scale, double *position, double *orientation, double *color)
{
    MapLoader ml{std::string(path)};
    return dump_markers(ml.toRosRealMapMarkers(), cap, id, scale, position, orientation, color, nullptr);
}

}  // extern "C"

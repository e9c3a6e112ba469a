// b2a_cv_stub.h -- the OpenCV types and calls /root/reference/src/aruco_slam.cpp names, declared so that
// the UNMODIFIED reference source compiles without OpenCV headers (oracle/_ref build).
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  The arithmetic behind the five library calls
//   cv::aruco::detectMarkers / estimatePoseSingleMarkers   (aruco_slam.cpp:313-314)
//   cv::aruco::drawDetectedMarkers                         (:319)
//   cv::Rodrigues / cv::projectPoints                      (:354, :441, :478)
// is NOT here: each forwards to a function pointer (oracle/ref_backend.h) that the harness points either at
// the cv2 4.13.0 wheel (tools/make_golden_slam.py, Python callbacks) or at oracle/orc_*.c (pinned to cv2).
#ifndef B2A_CV_STUB_H
#define B2A_CV_STUB_H
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8UC1 0
#define CV_8UC3 16
#define CV_64FC1 6
#define CV_64F 6

namespace cv {

template <typename T, int m, int n> struct Matx {
    T val[m * n];
    Matx() { for (int i = 0; i < m * n; ++i) val[i] = T(0); }
};
template <typename T, int cn> struct Vec : Matx<T, cn, 1> {
    Vec() {}
    Vec(T a, T b, T c) { this->val[0] = a; this->val[1] = b; this->val[2] = c; }
    T &operator[](int i) { return this->val[i]; }
    const T &operator[](int i) const { return this->val[i]; }
};
typedef Vec<double, 3> Vec3d;
typedef Vec<float, 3> Vec3f;

// cv::norm(Matx): sqrt of the sum of squares accumulated in double, elements in order
template <typename T, int m, int n> double norm(const Matx<T, m, n> &M)
{
    double s = 0;
    for (int i = 0; i < m * n; ++i) s += (double)M.val[i] * (double)M.val[i];
    return std::sqrt(s);
}

template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
    Point3_(const Vec<T, 3> &v) : x(v[0]), y(v[1]), z(v[2]) {}
};
typedef Point3_<float> Point3f;

// dense 2-D array: only what the reference touches (construction, at<double>, clone, image bytes)
class Mat {
public:
    Mat() : rows(0), cols(0), type_(0) {}
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type), buf_(std::make_shared<std::vector<uint8_t>>((size_t)r * c * elem_size(type))) {}
    Mat(int r, int c, int type, const void *data) : Mat(r, c, type) { std::memcpy(buf_->data(), data, buf_->size()); }
    int rows, cols;
    int type() const { return type_; }
    int channels() const { return (type_ >> 3) + 1; }
    bool empty() const { return !buf_ || buf_->empty(); }
    template <typename T> T &at(int i, int j) { return reinterpret_cast<T *>(buf_->data())[(size_t)i * cols + j]; }
    template <typename T> const T &at(int i, int j) const { return reinterpret_cast<const T *>(buf_->data())[(size_t)i * cols + j]; }
    template <typename T> T &at(int i) { return reinterpret_cast<T *>(buf_->data())[i]; }
    template <typename T> const T &at(int i) const { return reinterpret_cast<const T *>(buf_->data())[i]; }
    const uint8_t *ptr() const { return buf_ ? buf_->data() : nullptr; }
    uint8_t *ptr() { return buf_ ? buf_->data() : nullptr; }
    size_t total() const { return (size_t)rows * cols; }
    Mat clone() const
    {
        Mat m;
        m.rows = rows; m.cols = cols; m.type_ = type_;
        if (buf_) m.buf_ = std::make_shared<std::vector<uint8_t>>(*buf_);
        return m;
    }
private:
    static size_t elem_size(int type) { const int depth = type & 7, cn = (type >> 3) + 1; return (size_t)cn * (depth == 6 ? 8 : depth == 5 ? 4 : 1); }
    int type_;
    std::shared_ptr<std::vector<uint8_t>> buf_;     // shared like cv::Mat's reference-counted header copies
};

template <typename T> using Ptr = std::shared_ptr<T>;

void Rodrigues(const Vec3d &rvec, Mat &R);
void projectPoints(const std::vector<Point3f> &objectPoints, const Vec3d &rvec, const Vec3d &tvec, const Mat &cameraMatrix,
                   const Mat &distCoeffs, std::vector<Point2f> &imagePoints);

namespace aruco {
enum PREDEFINED_DICTIONARY_NAME { DICT_4X4_50 = 0, DICT_6X6_250 = 10, DICT_ARUCO_ORIGINAL = 16 };
struct Dictionary { int predefined_id; };
Ptr<Dictionary> getPredefinedDictionary(PREDEFINED_DICTIONARY_NAME name);
void detectMarkers(const Mat &image, const Ptr<Dictionary> &dictionary, std::vector<std::vector<Point2f>> &corners, std::vector<int> &ids);
void estimatePoseSingleMarkers(const std::vector<std::vector<Point2f>> &corners, float markerLength, const Mat &cameraMatrix,
                               const Mat &distCoeffs, std::vector<Vec3d> &rvecs, std::vector<Vec3d> &tvecs);
void drawDetectedMarkers(Mat &image, const std::vector<std::vector<Point2f>> &corners, const std::vector<int> &ids);
}  // namespace aruco
}  // namespace cv
#endif

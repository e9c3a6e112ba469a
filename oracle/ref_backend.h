/* ref_backend.h -- C interface of oracle/_ref/libarucoslam_ref.so: the reference's own
 * src/aruco_slam.cpp and src/map_loader.cpp, compiled UNMODIFIED from /root/reference against the stand-in
 * headers in oracle/ref_stubs/ (the image has no Eigen / OpenCV / ROS headers), plus this thin harness.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h): it exists to PIN the EKF / observation half of the path
 * (aruco_slam.cpp:21-74, 76-287, 307-376, 399-407, 423-471) -- tools/make_golden_slam.py runs it here and
 * writes tests/golden/slam_*.npz -- and to serve as the CPU baseline of the SLAM workloads in bench.py.
 *
 * The five OpenCV calls of that file are forwarded to the hooks below.  NULL hooks = oracle/orc_*.c (pinned
 * to cv2 4.13.0 on tests/golden); tools/make_golden_slam.py installs Python callbacks into the cv2 wheel
 * itself, so the committed vectors are "reference source + real OpenCV".
 */
#ifndef B2A_REF_BACKEND_H
#define B2A_REF_BACKEND_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    double Q_k, R_x, R_y, R_theta, kl, kr, b, marker_length;   /* ArucoSlamIniteData, aruco_slam.h:40-60 */
    double r2c_t[3], r2c_q[4];                                 /* transformStamped_r2c: translation, rotation (x,y,z,w) */
    int    markers_dictionary;
    float  useful_distance_threshold;
} ref_init;

/* detectMarkers: returns the number of markers written (corners [cap][4][2], ids [cap]) */
typedef int  (*ref_detect_fn)(const uint8_t *img, int w, int h, int channels, int dict_id, float *corners, int32_t *ids, int cap);
typedef void (*ref_pose_fn)(const float *corners, int n, float marker_length, const double *K9, const double *D, int nD, double *rvecs, double *tvecs);
typedef void (*ref_rodrigues_fn)(const double *rvec, double *R9);
typedef void (*ref_project_fn)(const float *obj, int n, const double *rvec, const double *tvec, const double *K9, const double *D, int nD, float *out);
void ref_set_hooks(ref_detect_fn d, ref_pose_fn p, ref_rodrigues_fn r, ref_project_fn j);
/* dictionary table for the default (orc_detect) hook: [nMarkers][4][nBytes], copied */
void ref_set_dictionary(int markerSize, int maxCorrectionBits, int nMarkers, int nBytes, const uint8_t *table);
/* replay: the next detectMarkers / estimatePoseSingleMarkers calls return these arrays (copied) instead of calling a hook;
 * n < 0 switches replay off */
void ref_set_replay(const float *corners, const int32_t *ids, int n, const double *rvecs, const double *tvecs);
/* the value ros::Time::now() returns (seconds) */
void ref_set_clock(double t);

typedef struct ref_slam ref_slam;
ref_slam *ref_create(const ref_init *init);
void ref_destroy(ref_slam *s);
void ref_set_camera(ref_slam *s, const double *K9, const double *D, int nD);          /* setCameraParameters */
void ref_add_encoder(ref_slam *s, double wl, double wr);                              /* ArucoSlam::addEncoder at the injected clock */
void ref_add_image(ref_slam *s, const uint8_t *img, int w, int h, int channels);      /* ArucoSlam::addImage */
/* private getObservations (:307-376) alone: the queue is drained in pop order into the arrays (capacity cap) and left empty;
 * returns the number of observations.  index = aruco_index_ (-1 = new). */
int  ref_get_observations(ref_slam *s, const uint8_t *img, int w, int h, int channels, int cap,
                          int32_t *ids, int32_t *index, double *xyt /*[cap][3]*/, double *cov /*[cap][9]*/);
int  ref_dim(const ref_slam *s);
int  ref_is_init(const ref_slam *s);
void ref_get_state(const ref_slam *s, double *mu, double *sigma /* row-major N x N */);
int  ref_get_ids(const ref_slam *s, int32_t *ids_by_index, int cap);                  /* aruco_id_map inverted; returns n landmarks */
void ref_set_state(ref_slam *s, int N, const double *mu, const double *sigma, const int32_t *ids_by_index, int is_init);
/* last_observed_marker_ (:263): ids and last_observation_ (NaN where the reference never wrote it); returns the count */
int  ref_get_last_observed(const ref_slam *s, int cap, int32_t *ids, double *last_obs /*[cap][3]*/);
void ref_robot_pose(ref_slam *s, double *position3, double *orientation4, double *covariance36);   /* toRosPose :378-410 */
/* MarkerArray taps: per marker id, scale[3], position[3], orientation[4] (x,y,z,w), color[4] (r,g,b,a), lifetime seconds.
 * which = 0: toRosMappedMarkers (detected_map_, :265-281), 1: toRosDetectedMarkers (:336-347).  Returns the count. */
int  ref_markers(ref_slam *s, int which, int cap, int32_t *id, double *scale, double *position, double *orientation, double *color, double *lifetime);
/* MapLoader (map_loader.cpp:7-117) on a file: same record layout as ref_markers; returns the count */
int  ref_map_load(const char *path, int cap, int32_t *id, double *scale, double *position, double *orientation, double *color);

#ifdef __cplusplus
}
#endif
#endif

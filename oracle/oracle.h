/* oracle.h -- CPU restatement of the reference's detect + pose + EKF path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.
 *
 * The reference (gitAugust/Aruco_Slam) performs this path by calling the
 * un-vendored third-party library OpenCV (cv::aruco::detectMarkers and
 * cv::aruco::estimatePoseSingleMarkers, reference src/aruco_slam.cpp:313-314;
 * cv::Rodrigues / cv::projectPoints at :354,:441) and Eigen; its own arithmetic
 * is src/aruco_slam.cpp:21-287,307-376,412-471.  OpenCV's sources are not in
 * /root/reference, so the detector / pose parts restate OpenCV 4.13.0's
 * published algorithm in plain C and are PINNED against
 * outputs of the `cv2 4.13.0` wheel: tests/golden/ (written by
 * tools/make_golden.py) -- every stage (gray, masks, contours, polygons,
 * candidates, ids/corners/rejected, poses).  The observation / EKF part follows the
 * reference file line by line and is PINNED against runs of the reference ITSELF:
 * oracle/_ref is /root/reference/src/aruco_slam.cpp + src/map_loader.cpp compiled
 * unmodified against stand-in headers (oracle/ref_stubs, oracle/Makefile target
 * `ref`); tools/make_golden_slam.py drives it with cv2 behind its OpenCV calls and
 * writes tests/golden/slam_*.npz and map_txt.npz (tests/test_slam_golden.py).
 */
#ifndef B2A_ORACLE_H
#define B2A_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* cv::aruco::DetectorParameters subset (defaults of cv2 4.13.0) */
typedef struct {
    int    adaptiveThreshWinSizeMin;          /* 3 */
    int    adaptiveThreshWinSizeMax;          /* 23 */
    int    adaptiveThreshWinSizeStep;         /* 10 */
    double adaptiveThreshConstant;            /* 7 */
    double minMarkerPerimeterRate;            /* 0.03 */
    double maxMarkerPerimeterRate;            /* 4.0 */
    double polygonalApproxAccuracyRate;       /* 0.03 */
    double minCornerDistanceRate;             /* 0.05 */
    int    minDistanceToBorder;               /* 3 */
    double minMarkerDistanceRate;             /* 0.125 */
    float  minGroupDistance;                  /* 0.21f */
    int    markerBorderBits;                  /* 1 */
    int    perspectiveRemovePixelPerCell;     /* 4 */
    double perspectiveRemoveIgnoredMarginPerCell; /* 0.13 */
    double maxErroneousBitsInBorderRate;      /* 0.35 */
    double minOtsuStdDev;                     /* 5.0 */
    double errorCorrectionRate;               /* 0.6 */
    int    cornerRefinementMethod;            /* 0 none, 1 subpix, 2 contour */
    int    cornerRefinementWinSize;           /* 5 */
    double relativeCornerRefinmentWinSize;    /* 0.3 */
    int    cornerRefinementMaxIterations;     /* 30 */
    double cornerRefinementMinAccuracy;       /* 0.1 */
    int    detectInvertedMarker;              /* 0 */
    int    useAruco3Detection;                /* 0 */
    int    minSideLengthCanonicalImg;         /* 32 */
    float  minMarkerLengthRatioOriginalImg;   /* 0 */
} orc_params;

typedef struct {
    int markerSize, maxCorrectionBits, nMarkers, nBytes;
    const uint8_t *table;                     /* [nMarkers][4][nBytes] */
} orc_dict;

void orc_default_params(orc_params *p);
/* cv::pyrDown and cv::resize(INTER_LINEAR) on 8-bit gray images (orc_pyramid.c): the two image operations of ArUco3 */
void orc_pyr_down(const uint8_t *src, int W, int H, uint8_t *dst);                       /* dst is ((W + 1) / 2) x ((H + 1) / 2) */
void orc_resize_linear(const uint8_t *src, int W, int H, uint8_t *dst, int dW, int dH);

/* ---- stage functions (each pinned separately against cv2) ---- */
void orc_bgr2gray(const uint8_t *bgr, int W, int H, uint8_t *gray);
void orc_adaptive_threshold(const uint8_t *gray, int W, int H, int k, double C, uint8_t *mask);
/* findContours(RETR_LIST, CHAIN_APPROX_NONE): returns number of contours; *pts = malloc'd
 * int32 x,y pairs of all contours back to back, *offs = malloc'd nContours+1 offsets
 * (in points).  Free both with orc_free. */
int  orc_find_contours(const uint8_t *mask, int W, int H, int32_t **pts, int32_t **offs);
/* approxPolyDP(closed=true) on int points; out has room for n points; returns count */
int  orc_approx_poly_dp(const int32_t *pts, int n, double eps, int32_t *out);
int  orc_is_contour_convex(const int32_t *pts, int n);
/* pointPolygonTest(measureDist=false) for a float polygon; returns -1,0,1 */
int  orc_point_polygon_test(const float *poly, int n, float px, float py);
void orc_get_perspective_transform(const float *src4, const float *dst4, double *H9);
/* warpPerspective(INTER_NEAREST, BORDER_CONSTANT 0) into an S x S patch */
void orc_warp_nearest(const uint8_t *gray, int W, int H, const double *H9, int S, uint8_t *patch);
int  orc_otsu(const uint8_t *img, int n);
/* identifyOneCandidate: returns 1 if valid; fills id, rot; bits (n x n) optional */
int  orc_identify_one(const uint8_t *gray, int W, int H, const float *corners4,
                      const orc_dict *d, const orc_params *p, int *id, int *rot, uint8_t *bits_out);
void orc_corner_subpix(const uint8_t *gray, int W, int H, float *corners, int n,
                       int win, int maxIter, double eps);
/* CORNER_REFINE_CONTOUR of one marker (_refineCandidateLines): contour n x (x, y); corners 4 x (x, y) in / out */
int  orc_refine_candidate_lines(const int32_t *contour, int n, float *corners);

/* ---- full detector ---- */
typedef struct {
    int n_acc, n_rej;
    float  *corners;      /* n_acc*8 */
    int32_t *ids;         /* n_acc */
    float  *rejected;     /* n_rej*8 */
    /* debug / stage outputs */
    int n_cand;           /* candidates after A3/A4 (all scales, in order) */
    float *cand;          /* n_cand*8 */
    int32_t *cand_len;    /* contour length of each candidate */
    int n_sel;            /* selected candidates after A5 */
    float *sel;           /* n_sel*8 (corners before identification swap/rotation) */
    int32_t *sel_info;    /* n_sel*5: parent, depth, valid, id, rot */
    int n_scales;
    int32_t *n_contours;  /* per scale */
} orc_detections;

int  orc_detect(const uint8_t *img, int W, int H, int channels, const orc_dict *d,
                const orc_params *p, orc_detections *out);
void orc_free_detections(orc_detections *out);
void orc_free(void *p);

/* ---- pose (estimatePoseSingleMarkers == per-marker solvePnP ITERATIVE) ---- */
void orc_rodrigues(const double *rvec, double *R9);
void orc_project_points(const double *obj, int n, const double *rvec, const double *tvec,
                        const double *K9, const double *D5, double *img);
void orc_undistort_points(const double *pts, int n, const double *K9, const double *D5, double *out);
int  orc_estimate_pose_single_markers(const float *corners, int n, double markerLength,
                                      const double *K9, const double *D5, int nD,
                                      double *rvecs, double *tvecs);

/* ---- observation mapping + EKF (reference src/aruco_slam.cpp) ---- */
typedef struct {
    double Q_k, R_x, R_y, R_theta, kl, kr, b, marker_length;
    double r2c_tx, r2c_ty;                /* transformStamped_r2c_.transform.translation.{x,y} */
    float  useful_distance_threshold;     /* aruco_slam.h:58 default 3 */
} orc_slam_params;

typedef struct {
    int aruco_id, aruco_index;
    double x, y, theta;
    double cov[9];
} orc_observation;

/* getObservations post-processing, aruco_slam.cpp:325-374 (without the queue):
 * returns number of observations kept, in detection order. */
int  orc_make_observations(const float *corners, const int32_t *ids, const double *rvecs,
                           const double *tvecs, int n, const double *K9, const double *D5,
                           const orc_slam_params *sp, orc_observation *out);

typedef struct orc_ekf orc_ekf;
orc_ekf *orc_ekf_create(const orc_slam_params *sp);
void orc_ekf_destroy(orc_ekf *e);
int  orc_ekf_dim(const orc_ekf *e);
void orc_ekf_get_state(const orc_ekf *e, double *mu, double *sigma, int32_t *ids);
void orc_ekf_set_state(orc_ekf *e, int N, const double *mu, const double *sigma, const int32_t *ids);
void orc_ekf_predict(orc_ekf *e, double wl, double wr, double dt);               /* :21-74 */
/* addImage EKF loop :88-263; observations are processed in priority-queue order;
 * dense != 0 evaluates (I-K Gx) Sigma as the reference does, else the rank-3 form */
void orc_ekf_update(orc_ekf *e, const orc_observation *obs, int n, int dense);

#ifdef __cplusplus
}
#endif
#endif

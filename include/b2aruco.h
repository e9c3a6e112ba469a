/* b2aruco.h -- C ABI of the B200-native ArUco detect + pose (+ EKF landmark update) path.
 *
 * Drop-in boundary for the calls the reference makes at
 *   /root/reference/src/aruco_slam.cpp:11-12  cv::aruco::getPredefinedDictionary
 *   /root/reference/src/aruco_slam.cpp:313    cv::aruco::detectMarkers(img, dictionary_, corners, IDs)
 *   /root/reference/src/aruco_slam.cpp:314    cv::aruco::estimatePoseSingleMarkers(corners, marker_length_, K, D, rvs, tvs)
 *   /root/reference/src/aruco_slam.cpp:325-374,437-471  observation mapping + CalculateCovariance
 *   /root/reference/src/aruco_slam.cpp:21-74   ArucoSlam::addEncoder (EKF prediction)
 *   /root/reference/src/aruco_slam.cpp:88-263  ArucoSlam::addImage   (EKF correction / augmentation)
 * The reference has no plugin/FFI layer of its own; these entry points are what a
 * C++ shim (include/b2aruco.hpp), a ctypes binding (aruco_slam_b200/_lib.py) or a ROS
 * node would bind instead of the cv::aruco / Eigen calls.  Plain pointers and sizes
 * only; every function returns a b2a_status (0 = OK) unless stated otherwise; no
 * exceptions cross the boundary.  All compute runs in hand-written CUDA kernels
 * (sm_100a); there is no CPU fallback: without a usable CUDA device the create calls
 * fail with B2A_ERR_CUDA.
 */
#ifndef B2ARUCO_H
#define B2ARUCO_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    B2A_OK = 0,
    B2A_ERR_INVALID = 1,     /* bad argument (cv::Exception from CV_Assert in the reference path) */
    B2A_ERR_CUDA = 2,        /* CUDA runtime error, see b2a_last_error() */
    B2A_ERR_CAPACITY = 3,    /* an internal list overflowed its configured capacity; results incomplete */
    B2A_ERR_UNSUPPORTED = 4  /* parameter combination outside what the kernels implement */
} b2a_status;

const char *b2a_last_error(void);
const char *b2a_version(void);

/* ---- cv::aruco::DetectorParameters (fields the detect path reads; cv2 4.13.0 defaults) ---- */
typedef struct {
    int    adaptiveThreshWinSizeMin;              /* 3 */
    int    adaptiveThreshWinSizeMax;              /* 23 */
    int    adaptiveThreshWinSizeStep;             /* 10 */
    double adaptiveThreshConstant;                /* 7 */
    double minMarkerPerimeterRate;                /* 0.03 */
    double maxMarkerPerimeterRate;                /* 4.0 */
    double polygonalApproxAccuracyRate;           /* 0.03 */
    double minCornerDistanceRate;                 /* 0.05 */
    int    minDistanceToBorder;                   /* 3 */
    double minMarkerDistanceRate;                 /* 0.125 */
    float  minGroupDistance;                      /* 0.21 */
    int    markerBorderBits;                      /* 1 */
    int    perspectiveRemovePixelPerCell;         /* 4 */
    double perspectiveRemoveIgnoredMarginPerCell; /* 0.13 */
    double maxErroneousBitsInBorderRate;          /* 0.35 */
    double minOtsuStdDev;                         /* 5.0 */
    double errorCorrectionRate;                   /* 0.6 */
    int    cornerRefinementMethod;                /* 0 = CORNER_REFINE_NONE, 1 = CORNER_REFINE_SUBPIX, 2 = CORNER_REFINE_CONTOUR;
                                                     3 (CORNER_REFINE_APRILTAG) -> B2A_ERR_UNSUPPORTED */
    int    cornerRefinementWinSize;               /* 5 */
    double relativeCornerRefinmentWinSize;        /* 0.3 */
    int    cornerRefinementMaxIterations;         /* 30 */
    double cornerRefinementMinAccuracy;           /* 0.1 */
    int    detectInvertedMarker;                  /* 0 */
    /* ArUco3 ("Speeded up detection of squared fiducial markers", the mode cv2 enables with useAruco3Detection): candidates are
     * searched in a reduced copy of the frame (factor minSide / (minSide + max(W, H) * ratio), cv::resize INTER_LINEAR), every
     * candidate is identified in the level of the frame's image pyramid (cv::buildPyramid) where its contour is just longer than
     * 4 * minSide, and the accepted corners are refined level by level up to the full-size frame (cornerSubPix, window 3, or 5
     * above 1080 px); cornerRefinementMethod is then CORNER_REFINE_SUBPIX whatever was asked, and -- as in cv2 4.13 -- the
     * REJECTED quads stay in the coordinates of the reduced image.  Needs minSideLengthCanonicalImg >= 1 and a ratio >= 0. */
    int    useAruco3Detection;                    /* 0 */
    int    minSideLengthCanonicalImg;             /* 32 */
    float  minMarkerLengthRatioOriginalImg;       /* 0 */
} b2a_detector_params;

void b2a_default_detector_params(b2a_detector_params *p);

/* ---- pinned host staging memory for the frames handed to the detect calls (what the reference's node gets from cv_bridge,
 * aruco_slam_node.cpp:93, once a caller owns the buffer).  cudaHostAlloc'd, usable from every device.  write_combined != 0 asks
 * for write-combined pages: the GPU reads them over PCIe without snooping the CPU caches, which some hosts turn into a faster
 * (and, with several GPUs, less contended) host-to-device copy; the CPU should only WRITE such memory, front to back. ---- */
int  b2a_host_alloc(size_t bytes, int write_combined, void **out);
void b2a_host_free(void *p);

/* ---- dictionaries: cv::aruco::getPredefinedDictionary (aruco_slam.cpp:11-12) ---- */
typedef struct {
    int markerSize, maxCorrectionBits, nMarkers, nBytes;
    const uint8_t *table;             /* [nMarkers][4 rotations][nBytes], cv2 bytesList memory order */
} b2a_dictionary;

/* dict_id = cv::aruco::PREDEFINED_DICTIONARY_NAME value (16 = DICT_ARUCO_ORIGINAL, parameters.yaml:16).
 * The returned table is owned by the library and lives for the process. */
int b2a_get_predefined_dictionary(int dict_id, b2a_dictionary *out);

/* ---- detector handle: one per GPU; single caller at a time (like one ArucoSlam instance) ---- */
typedef struct b2a_detector b2a_detector;

typedef struct {
    int device;                 /* CUDA device ordinal */
    int max_width, max_height;  /* largest frame */
    int max_batch;              /* largest number of frames per call */
    int max_markers;            /* capacity of accepted (and of rejected) quads per frame; 0 = 256 */
    int max_candidates;         /* capacity of quad candidates per frame before grouping; 0 = 2048 */
} b2a_detector_config;

int  b2a_detector_create(const b2a_detector_config *cfg, const b2a_dictionary *dict,
                         const b2a_detector_params *params, b2a_detector **out);
void b2a_detector_destroy(b2a_detector *d);

/* Where the frames of a batch live. */
typedef struct {
    const uint8_t *data;        /* first frame */
    int    on_device;           /* 0: host memory (copied H2D inside the call), 1: device memory */
    int    batch, width, height, channels;   /* channels 1 (gray) or 3 (bgr8, aruco_slam_node.cpp:93) */
    size_t row_stride;          /* bytes between rows   (0 = width*channels) */
    size_t frame_stride;        /* bytes between frames (0 = row_stride*height) */
} b2a_frames;

/* Results of one call; pointers are library-owned pinned host memory, valid until the next
 * call on the same handle.  Arrays are [batch][max_markers]..., frame f's entries start at
 * f*max_markers.  Order inside a frame is the reference's (cv2 4.13.0) output order. */
typedef struct {
    int batch, max_markers;
    const int32_t *n_accepted;  /* [batch] */
    const int32_t *n_rejected;  /* [batch] */
    const float   *corners;     /* [batch][max_markers][4][2]  clockwise from the marker's top-left */
    const int32_t *ids;         /* [batch][max_markers] */
    const float   *rejected;    /* [batch][max_markers][4][2] */
    const double  *rvecs;       /* [batch][max_markers][3]  (only after b2a_detect_pose) */
    const double  *tvecs;       /* [batch][max_markers][3] */
    const int32_t *status;      /* [batch] per-frame b2a_status (capacity overflow) */
} b2a_detections;

/* detectMarkers over a batch of frames (aruco_slam.cpp:313). */
int b2a_detect(b2a_detector *d, const b2a_frames *frames, b2a_detections *out);

/* Camera model of the pose step: K row-major 3x3, D = (k1,k2,p1,p2[,k3]), nD in {0,4,5}. */
typedef struct {
    double K[9];
    double D[5];
    int    nD;
    float  marker_length;       /* markerLength of estimatePoseSingleMarkers (parameters.yaml:17) */
} b2a_camera;

/* detectMarkers + estimatePoseSingleMarkers without leaving the device (aruco_slam.cpp:313-314). */
int b2a_detect_pose(b2a_detector *d, const b2a_frames *frames, const b2a_camera *cam, b2a_detections *out);

/* Compact records of a call's detections (what travels when several GPUs / processes gather their results on one host):
 * header {magic, batch, has_pose, max_markers} (4 x i32), then per frame {n_accepted, n_rejected} (2 x i32), ids [na] i32,
 * corners [na][8] f32, rvecs [na][3] f64 and tvecs [na][3] f64 (if has_pose), rejected [nr][8] f32.  *n_bytes = size written;
 * B2A_ERR_CAPACITY (with *n_bytes = size needed) when cap is too small. */
int b2a_pack_detections(const b2a_detections *det, void *out, size_t cap, size_t *n_bytes);

/* The result arrays of the handle's last completed b2a_detect / b2a_detect_pose / b2a_slam_add_image call again (still valid: no call
 * since); used to draw the overlay of the frame addImage just processed (getMarkedImg, aruco_slam.h:152). */
int b2a_detector_last_detections(b2a_detector *d, b2a_detections *out);

/* The same call split in two, so that one host thread keeps two batches in flight on ONE handle: submit enqueues the H2D
 * copies and every kernel of a batch and returns at once (cam = NULL: detect only); wait blocks until that batch's results
 * are in host memory and fills `out`.  Tickets alternate between two internal pipeline contexts (the second is created on
 * the first odd ticket), so the PCIe copy of batch k+1 overlaps the kernels of batch k:
 *     submit(k+1) ... wait(k) ... submit(k+2) ... wait(k+1) ...
 * At most n batches in flight (b2a_detector_set_inflight, default two); waits in submission order; frames->data must stay valid until the matching wait; the
 * arrays `out` points to stay valid until the second submit after it.  The synchronous calls above refuse to run while a
 * batch is in flight on the first context. */
int b2a_detect_pose_submit(b2a_detector *d, const b2a_frames *frames, const b2a_camera *cam, int *ticket);
/* how many batches submit / wait keep in flight on this handle: 1 .. 8 contexts, default 2 (more pay off for single frames, whose
 * kernels leave most of the GPU idle; a 32-frame batch is bound by its PCIe copy with two).  Resets the ticket counter; not while a
 * batch is in flight. */
int b2a_detector_set_inflight(b2a_detector *d, int n);
/* Calls on up to 4 frames are launch-bound on the host, so by default their kernels are captured once per (batch, shape, camera)
 * into a CUDA graph and replayed (the frames are first copied into the handle's own buffer so that every node reads fixed
 * addresses); b2a_last_stage_times reports zeros for such calls.  on = 0 goes back to plain launches. */
int  b2a_detector_set_graph(b2a_detector *d, int on);
int b2a_detect_pose_wait(b2a_detector *d, int ticket, b2a_detections *out);

/* cv::aruco::drawDetectedMarkers(image, corners, ids, borderColor) (aruco_slam.cpp:319; the image getMarkedImg returns,
 * aruco_slam.h:152): the overlay is drawn into `image` (host memory, 8-bit, 1 or 3 channels, in place) on the device -- marker
 * sides, the anti-aliased square on the first corner, the "id=N" label -- pixel for pixel as OpenCV 4.13 draws it when the corners
 * are integer valued (CORNER_REFINE_NONE) and inside the image.  ids may be NULL (no labels); border_bgr NULL = (0, 255, 0).
 * n <= max_markers of the handle; ids 0 .. 9999. */
int b2a_draw_detected_markers(b2a_detector *d, uint8_t *image, int width, int height, int channels, size_t row_stride,
                              const float *corners, const int32_t *ids, int n, const uint8_t border_bgr[3]);

/* ---- several GPUs of one box from one process (frames are independent: no collective) ----
 * One detector handle and one host thread per listed device (a device may be listed more than once); a batch of HOST frames is
 * cut into contiguous blocks, frames [g B / G, (g+1) B / G) go to device g, and the detections are gathered on the host in frame
 * order into library-owned arrays (valid until the next call).  cfg->max_batch is the largest batch of a CALL; cfg->device is
 * ignored.  cam = NULL: detect only. */
typedef struct b2a_multi b2a_multi;
int  b2a_multi_create(const int *devices, int n_devices, const b2a_detector_config *cfg, const b2a_dictionary *dict,
                      const b2a_detector_params *params, b2a_multi **out);
void b2a_multi_destroy(b2a_multi *m);
int  b2a_multi_num_devices(const b2a_multi *m);
int  b2a_multi_detect_pose(b2a_multi *m, const b2a_frames *frames, const b2a_camera *cam, b2a_detections *out);

/* ---- cv::aruco::ArucoDetector::refineDetectedMarkers (cv2 4.13; part of the cv::aruco surface of aruco_slam.cpp:313, the reference
 * itself does not call it).  Rejected candidates that lie where the board says an undetected marker must be are moved to the
 * detected list.  board: cv::aruco::Board (ids + object points of every marker's four corners); cam null = the global-homography
 * form (all board points must share one z), else the board pose is fitted through the camera as cv::solvePnP(ITERATIVE) fits it
 * (start from the plane homography for coplanar boards, from the DLT for boards in general position, which need >= 6 matched
 * corners; then least squares on the reprojection error).  corners / ids / rejected are host arrays, edited in place like cv2's
 * InputOutputArrays: recovered markers are appended in board order with the candidate's corners rotated to the matching order
 * (and refined when the detector was created with CORNER_REFINE_SUBPIX), the recovered candidates leave `rejected`;
 * recovered_idx (optional, room for *n_rejected entries) gets their indices in the incoming rejected list.
 * The bit extraction of the candidates runs on the GPU (the identification kernels); it overwrites the detector's results of its
 * last detect call.  Errors: B2A_ERR_CAPACITY when 4 * n_rejected exceeds max_candidates or the outputs exceed `capacity`. */
typedef struct {
    int n_markers;
    const int32_t *ids;          /* [n_markers] */
    const float *obj_points;     /* [n_markers][4][3], Board::getObjPoints order */
} b2a_board;
typedef struct {                 /* cv::aruco::RefineParameters */
    float minRepDistance;        /* 10 */
    float errorCorrectionRate;   /* 3; negative = no code test */
    int   checkAllOrders;        /* 1 */
} b2a_refine_params;
void b2a_default_refine_params(b2a_refine_params *p);
int b2a_refine_detected_markers(b2a_detector *d, const b2a_frames *image, const b2a_board *board, float *corners, int32_t *ids,
                                int *n_detected, int capacity, float *rejected, int *n_rejected, const b2a_camera *cam,
                                const b2a_refine_params *params, int32_t *recovered_idx, int *n_recovered);

/* estimatePoseSingleMarkers on caller-provided corners (host arrays): corners [n][4][2] f32,
 * rvecs/tvecs [n][3] f64.  (aruco_slam.cpp:314) */
int b2a_estimate_pose_single_markers(b2a_detector *d, const float *corners, int n, const b2a_camera *cam,
                                     double *rvecs, double *tvecs);

/* Stage taps for parity tests (host output buffers):
 *   gray  [batch][H][W] u8 (after A1),  masks [batch][nScales][H][W] u8 (0/255, A2). */
int b2a_debug_threshold(b2a_detector *d, const b2a_frames *frames, uint8_t *gray, uint8_t *masks);
/* Contours that pass the perimeter gate, in cv2.findContours list order per (frame,scale):
 *   counts [batch][nScales] (all borders found, incl. short ones), n_kept [batch][nScales],
 *   kept_len / kept_start_xy filled up to cap entries per (frame,scale), points of kept contours
 *   concatenated per (frame,scale) into pts (x,y int16 pairs), pts_cap points per (frame,scale). */
int b2a_debug_contours(b2a_detector *d, const b2a_frames *frames, int32_t *counts, int32_t *n_kept,
                       int32_t *kept_len, int cap, int16_t *pts, int pts_cap);
/* Quad candidates before grouping (A3/A4), in reference order: n_cand [batch], quads [batch][cap][4][2]. */
int b2a_debug_candidates(b2a_detector *d, const b2a_frames *frames, int32_t *n_cand, float *quads, int cap);
int b2a_detector_num_scales(const b2a_detector *d);

/* timing taps: milliseconds of the last b2a_detect / b2a_detect_pose call, per stage, measured with
 * CUDA events on the handle's stream.  names[i] are static strings; returns number of stages. */
int b2a_last_stage_times(const b2a_detector *d, const char **names, float *ms, int cap);
/* number of kernel launches issued by the last detect call (for bench.py's gpu_launches) */
int b2a_last_launch_count(const b2a_detector *d);
/* A call's batch is cut into up to n sub-batches that run on separate CUDA streams (default 0 = automatic: 2 for frames already
 * in device memory, 4 for host frames; 1 = strictly serial stages, which is what per-stage timings / profiles want). */
int b2a_detector_set_streams(b2a_detector *d, int n);
/* the CUDA stream (cudaStream_t) the handle launches on (sub-batch 0; the others fork from / join into it) */
void *b2a_detector_stream(const b2a_detector *d);

/* ---- observation mapping + EKF (ArucoSlam, reference include/aruco_slam/aruco_slam.h:101-193) ---- */
typedef struct {
    double Q_k, R_x, R_y, R_theta;   /* parameters.yaml:5-8 */
    double kl, kr, b;                /* parameters.yaml:11-13 */
    double r2c_tx, r2c_ty;           /* transformStamped_r2c_.transform.translation.{x,y} */
    float  useful_distance_threshold;/* aruco_slam.h:58 (default 3) */
    int    max_landmarks;            /* capacity of the map (state dim 3+3n); 0 = 512 */
} b2a_slam_params;

void b2a_default_slam_params(b2a_slam_params *p);

typedef struct {
    int32_t aruco_id, aruco_index;
    double  x, y, theta;
    double  cov[9];
} b2a_observation;

typedef struct b2a_slam b2a_slam;
int  b2a_slam_create(int device, const b2a_slam_params *p, b2a_slam **out);
void b2a_slam_destroy(b2a_slam *s);
int  b2a_slam_dim(const b2a_slam *s);                     /* 3 + 3*landmarks */
/* mu [N], sigma [N][N] row-major, ids [n landmarks] (host). NULL pointers are skipped. */
int  b2a_slam_get_state(b2a_slam *s, double *mu, double *sigma, int32_t *ids);
int  b2a_slam_set_state(b2a_slam *s, int N, const double *mu, const double *sigma, const int32_t *ids);
/* addEncoder(wl, wr) with an explicit dt (aruco_slam.cpp:21-74 reads ros::Time::now()).  As in the reference (:24-29) the
 * FIRST call on a fresh handle only marks the filter initialised (there it latches the clock) and predicts nothing;
 * b2a_slam_set_state also marks it initialised. */
int  b2a_slam_add_encoder(b2a_slam *s, double wl, double wr, double dt);
/* getObservations' post-processing (aruco_slam.cpp:325-374) on host arrays; returns the kept
 * observations in detection order through out (capacity n), *n_out = count. */
int  b2a_slam_make_observations(b2a_slam *s, const float *corners, const int32_t *ids, const double *rvecs,
                                const double *tvecs, int n, const b2a_camera *cam,
                                b2a_observation *out, int *n_out);
/* The EKF loop of addImage (aruco_slam.cpp:88-263) for one frame's observations, given in detection order (the order
 * getObservations pushes them, :373); they are processed in the order the reference's priority queue pops them.  The
 * kernels are enqueued on the handle's stream; the call does not wait for them. */
int  b2a_slam_update(b2a_slam *s, const b2a_observation *obs, int n);
/* The EKF kernels are enqueued on the handle's stream; get_state waits for them, and so does this (used to time updates). */
int  b2a_slam_synchronize(b2a_slam *s);
/* the CUDA stream (cudaStream_t) the filter's kernels run on (to time them with CUDA events) */
void *b2a_slam_stream(const b2a_slam *s);
/* addImage(img): detect + pose + observations + EKF update for one frame (aruco_slam.cpp:76-263). */
int  b2a_slam_add_image(b2a_slam *s, b2a_detector *d, const b2a_frames *frame, const b2a_camera *cam);

/* addImage split in two (like b2a_detect_pose_submit / _wait): submit enqueues detection, pose and observation mapping of a frame on
 * one of the detector's two contexts and returns; wait blocks for that frame and runs the EKF loop on its observations.  Submitting
 * frame k + 1 before waiting for frame k keeps two frames of a camera stream in flight.  The filter's state evolves exactly as with
 * b2a_slam_add_image as long as every frame's b2a_slam_add_encoder call(s) come before that frame's wait (prediction before
 * correction, aruco_slam.cpp:21-74 then :76-263).  *ticket = -1 when the frame is ignored (no encoder message yet, :84-85). */
int  b2a_slam_add_image_submit(b2a_slam *s, b2a_detector *d, const b2a_frames *frame, const b2a_camera *cam, int *ticket);
int  b2a_slam_add_image_wait(b2a_slam *s, b2a_detector *d, int ticket);

/* ---- wire / on-disk formats either side of the path (host only, no ROS types) ---- */
/* One line of the landmark map (reference map/map.txt:1 "id length x y z roll_x pitch_y yaw_z") and, equally, one cube of the
 * reference's MarkerArray messages (visualization_msgs::Marker fields the reference fills, map_loader.cpp:96-117 and
 * aruco_slam.cpp:289-305): frame "world", scale (length, length, 0.01), pose position (x, y, z), orientation q = (qx, qy, qz, qw). */
typedef struct {
    int32_t id;
    double  length, x, y, z, roll, pitch, yaw;
    double  q[4];                    /* tf2::Quaternion::setRPY(roll, pitch, yaw) as (x, y, z, w) */
} b2a_map_marker;
/* tf2::Quaternion::setRPY (fixed axes: roll about X, then pitch about Y, then yaw about Z); q = (x, y, z, w). */
void b2a_quaternion_from_rpy(double roll, double pitch, double yaw, double q[4]);
/* MapLoader::loadMap (map_loader.cpp:7-84) on a text buffer: blank lines and lines whose first non-blank character is '#'
 * are skipped; a line whose first non-blank character is not a digit aborts the load with an EMPTY map (:44-50); a line with
 * fewer than the four fields id length x y is skipped (:52-58); missing z / roll / pitch / yaw read as 0 (the reference leaves
 * roll and yaw uninitialised in that case, :65-79 -- 0 is its evident intent).  Returns B2A_OK and *n_out = markers written
 * (B2A_ERR_CAPACITY if cap is too small; *n_out then holds the number the text defines). */
int  b2a_map_parse(const char *text, size_t len, b2a_map_marker *out, int cap, int *n_out);
/* the same from a file; B2A_ERR_INVALID when the file cannot be opened (the reference logs and leaves the map empty, :13-17) */
int  b2a_map_load(const char *path, b2a_map_marker *out, int cap, int *n_out);
/* ArucoSlam::toRosDetectedMarkers (the cubes getObservations builds for the markers of the CURRENT frame, aruco_slam.cpp:324-347):
 * one record per detection whose distance passes the range gate (float norm(tvec) <= useful_distance_threshold, :327-333), in
 * detection order; id = the marker's ArUco id, length = marker_length, orientation = r2c rotation * getRotation(Rodrigues(rvec))
 * (tf2::Matrix3x3::getRotation, fillTransform :473-489), position = r2c * tvec (tf2::doTransform with transformStamped_r2c_, frame
 * "base_link", lifetime 0.1 s); roll / pitch / yaw are left 0.  r2c_q = (x, y, z, w) (NULL = identity), r2c_t = translation
 * (NULL = 0).  Host only.  B2A_ERR_CAPACITY (with *n_out = the count) when cap is too small. */
int  b2a_pack_detected_markers(const int32_t *ids, const double *rvecs, const double *tvecs, int n, double marker_length,
                               float useful_distance_threshold, const double r2c_q[4], const double r2c_t[3],
                               b2a_map_marker *out, int cap, int *n_out);
/* ArucoSlam::toRosPose (aruco_slam.cpp:376-407): position (mu0, mu1, 0.1), orientation setRPY(0, 0, mu2), and the 6x6 row-major
 * covariance with Sigma[0:3,0:3] scattered to entries 0,1,5 / 6,7,11 / 30,31,35 (all others 0). */
typedef struct {
    double position[3];
    double orientation[4];           /* (x, y, z, w) */
    double covariance[36];
} b2a_pose_with_covariance;
int  b2a_slam_robot_pose(b2a_slam *s, b2a_pose_with_covariance *out);
/* The same record read back without stalling the host thread (a camera loop with several frames in flight): _submit enqueues the
 * 96-byte copy into pinned slot `slot` (0 .. 7) behind everything the filter has been given so far, _wait blocks only until that
 * copy has landed and packs the record. */
int  b2a_slam_robot_pose_submit(b2a_slam *s, int slot);
int  b2a_slam_robot_pose_wait(b2a_slam *s, int slot, b2a_pose_with_covariance *out);
/* the packing step alone, on host values: mu[0:3] and Sigma[0:3,0:3] row-major (no device access) */
void b2a_pack_robot_pose(const double mu3[3], const double sigma33[9], b2a_pose_with_covariance *out);
/* detected_map_ of addImage (aruco_slam.cpp:266-281): one cube per landmark i: id = i (the landmark index, not the aruco id),
 * length = marker_length, position (mu[3+3i], mu[4+3i], 0.3), orientation setRPY(0, 1.5708, mu[5+3i]). */
int  b2a_slam_detected_map(b2a_slam *s, double marker_length, b2a_map_marker *out, int cap, int *n_out);
/* one cube of that list from host values: landmark3 = (mu[3+3i], mu[4+3i], mu[5+3i]) (no device access) */
void b2a_pack_map_marker(int index, double marker_length, const double landmark3[3], b2a_map_marker *out);

#ifdef __cplusplus
}
#endif
#endif

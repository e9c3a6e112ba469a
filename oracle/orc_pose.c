/* orc_pose.c -- CPU restatement of cv::aruco::estimatePoseSingleMarkers (per-marker
 * cv::solvePnP, SOLVEPNP_ITERATIVE; the call at reference src/aruco_slam.cpp:314),
 * cv::Rodrigues / cv::projectPoints (:354,:441) and the reference's own observation
 * mapping and covariance (src/aruco_slam.cpp:325-374, 412-421, 437-471).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Pinned against cv2.solvePnP /
 * cv2.projectPoints / cv2.Rodrigues outputs in tests/golden (SURVEY App. A "Pose").
 */
#include "oracle.h"
#include <float.h>
#include <math.h>
#include <string.h>

void orc_rodrigues(const double *r, double *R)
{
    double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (th < DBL_EPSILON) { memset(R, 0, 9 * sizeof(double)); R[0] = R[4] = R[8] = 1; return; }
    double c = cos(th), s = sin(th), c1 = 1 - c, it = 1 / th;
    double x = r[0] * it, y = r[1] * it, z = r[2] * it;
    R[0] = c + c1 * x * x;     R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y;     R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}

static void rot_to_rvec(const double *R, double *r)
{
    /* log map of a (numerically) proper rotation */
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1 ? 1 : (c < -1 ? -1 : c);
    double th = acos(c);
    if (s < 1e-5) {
        if (c > 0) { r[0] = r[1] = r[2] = 0; return; }
        double t;
        t = (R[0] + 1) * 0.5; rx = sqrt(t > 0 ? t : 0);
        t = (R[4] + 1) * 0.5; ry = sqrt(t > 0 ? t : 0) * (R[1] < 0 ? -1. : 1.);
        t = (R[8] + 1) * 0.5; rz = sqrt(t > 0 ? t : 0) * (R[2] < 0 ? -1. : 1.);
        if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
        th /= sqrt(rx * rx + ry * ry + rz * rz);
        r[0] = th * rx; r[1] = th * ry; r[2] = th * rz;
        return;
    }
    double vth = 1 / (2 * s);
    vth *= th;
    r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
}

void orc_project_points(const double *obj, int n, const double *rvec, const double *tvec,
                        const double *K, const double *D, double *img)
{
    double R[9];
    orc_rodrigues(rvec, R);
    double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3], k3 = D[4];
    for (int i = 0; i < n; i++) {
        double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
        double x = R[0] * X + R[1] * Y + R[2] * Z + tvec[0];
        double y = R[3] * X + R[4] * Y + R[5] * Z + tvec[1];
        double z = R[6] * X + R[7] * Y + R[8] * Z + tvec[2];
        z = z ? 1. / z : 1;
        x *= z; y *= z;
        double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
        double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
        double cdist = 1 + k1 * r2 + k2 * r4 + k3 * r6;
        double xd = x * cdist + p1 * a1 + p2 * a2;
        double yd = y * cdist + p1 * a3 + p2 * a1;
        img[2 * i] = xd * fx + cx;
        img[2 * i + 1] = yd * fy + cy;
    }
}

void orc_undistort_points(const double *pts, int n, const double *K, const double *D, double *out)
{
    /* undistortPoints default criteria: 5 fixed-point iterations */
    double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3], k3 = D[4];
    for (int i = 0; i < n; i++) {
        double x = (pts[2 * i] - cx) / fx, y = (pts[2 * i + 1] - cy) / fy;
        double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            double r2 = x * x + y * y;
            double icdist = 1. / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
            if (icdist < 0) { x = x0; y = y0; break; }
            double dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
            double dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
            x = (x0 - dx) * icdist;
            y = (y0 - dy) * icdist;
        }
        out[2 * i] = x; out[2 * i + 1] = y;
    }
}

/* Gaussian elimination with partial pivoting, n <= 8 */
static int solve_n(double *A, double *b, int n)
{
    for (int i = 0; i < n; i++) {
        int k = i;
        for (int j = i + 1; j < n; j++) if (fabs(A[j * n + i]) > fabs(A[k * n + i])) k = j;
        if (fabs(A[k * n + i]) < 1e-300) return 0;
        if (k != i) { for (int j = 0; j < n; j++) { double t = A[i * n + j]; A[i * n + j] = A[k * n + j]; A[k * n + j] = t; } double t = b[i]; b[i] = b[k]; b[k] = t; }
        for (int j = i + 1; j < n; j++) {
            double a = A[j * n + i] / A[i * n + i];
            for (int c = i; c < n; c++) A[j * n + c] -= a * A[i * n + c];
            b[j] -= a * b[i];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i];
        for (int k = i + 1; k < n; k++) s -= A[i * n + k] * b[k];
        b[i] = s / A[i * n + i];
    }
    return 1;
}

static void nearest_rotation(double *R)
{
    /* polar decomposition by Newton iteration: R <- (R + R^-T)/2 converges to U V^T */
    for (int it = 0; it < 60; it++) {
        double a[9];
        memcpy(a, R, sizeof(a));
        double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
        double id = 1 / det, iT[9];
        /* inverse transpose = cofactor / det */
        iT[0] = (a[4] * a[8] - a[5] * a[7]) * id; iT[1] = (a[5] * a[6] - a[3] * a[8]) * id; iT[2] = (a[3] * a[7] - a[4] * a[6]) * id;
        iT[3] = (a[2] * a[7] - a[1] * a[8]) * id; iT[4] = (a[0] * a[8] - a[2] * a[6]) * id; iT[5] = (a[1] * a[6] - a[0] * a[7]) * id;
        iT[6] = (a[1] * a[5] - a[2] * a[4]) * id; iT[7] = (a[2] * a[3] - a[0] * a[5]) * id; iT[8] = (a[0] * a[4] - a[1] * a[3]) * id;
        double diff = 0;
        for (int i = 0; i < 9; i++) { double v = 0.5 * (a[i] + iT[i]); diff += fabs(v - a[i]); R[i] = v; }
        if (diff < 1e-15) break;
    }
}

/* cv::Rodrigues(rvec -> R) with the 3x9 Jacobian (row i = d R_flat / d r_i), OpenCV's formula */
static void rodrigues_jac(const double *r, double *R, double *J)
{
    double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (th < DBL_EPSILON) {
        memset(R, 0, 9 * sizeof(double)); R[0] = R[4] = R[8] = 1;
        memset(J, 0, 27 * sizeof(double));
        J[5] = J[15] = J[19] = -1; J[7] = J[11] = J[21] = 1;
        return;
    }
    double c = cos(th), s = sin(th), c1 = 1. - c, itheta = 1. / th;
    double x = r[0] * itheta, y = r[1] * itheta, z = r[2] * itheta;
    double rrt[9] = {x * x, x * y, x * z, x * y, y * y, y * z, x * z, y * z, z * z};
    double rx[9] = {0, -z, y, z, 0, -x, -y, x, 0};
    double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int k = 0; k < 9; k++) R[k] = c * I[k] + c1 * rrt[k] + s * rx[k];
    double drrt[27] = {x + x, y, z, y, 0, 0, z, 0, 0, 0, x, 0, x, y + y, z, 0, z, 0, 0, 0, x, 0, 0, y, x, y, z + z};
    double drx[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0};
    double rr[3] = {x, y, z};
    for (int i = 0; i < 3; i++) {
        double ri = rr[i], a0 = -s * ri, a1 = (s - 2 * c1 * itheta) * ri, a2 = c1 * itheta, a3 = (c - s * itheta) * ri, a4 = s * itheta;
        for (int k = 0; k < 9; k++) J[i * 9 + k] = a0 * I[k] + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * rx[k] + a4 * drx[i * 9 + k];
    }
}

/* projectPoints residuals (proj - measured) and, if J != NULL, dpdr | dpdt (8 x 6), as cv::projectPoints */
static double reproj_err(const double *obj, const double *ip, const double *p6, const double *K, const double *D, double *err, double *J)
{
    double R[9], dRdr[27];
    rodrigues_jac(p6, R, dRdr);
    double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3], k3 = D[4];
    double e2 = 0;
    for (int i = 0; i < 4; i++) {
        double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
        double x = R[0] * X + R[1] * Y + R[2] * Z + p6[3];
        double y = R[3] * X + R[4] * Y + R[5] * Z + p6[4];
        double z = R[6] * X + R[7] * Y + R[8] * Z + p6[5];
        z = z ? 1. / z : 1;
        x *= z; y *= z;
        double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
        double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
        double cdist = 1 + k1 * r2 + k2 * r4 + k3 * r6;
        double xd = x * cdist + p1 * a1 + p2 * a2, yd = y * cdist + p1 * a3 + p2 * a1;
        err[2 * i] = xd * fx + cx - ip[2 * i];
        err[2 * i + 1] = yd * fy + cy - ip[2 * i + 1];
        e2 += err[2 * i] * err[2 * i] + err[2 * i + 1] * err[2 * i + 1];
        if (!J) continue;
        for (int k = 0; k < 6; k++) {
            double dx0, dy0, dz0;
            if (k < 3) {
                const double *d = dRdr + 9 * k;
                dx0 = X * d[0] + Y * d[1] + Z * d[2];
                dy0 = X * d[3] + Y * d[4] + Z * d[5];
                dz0 = X * d[6] + Y * d[7] + Z * d[8];
            } else { dx0 = (k == 3); dy0 = (k == 4); dz0 = (k == 5); }
            double dxd = z * (dx0 - x * dz0), dyd = z * (dy0 - y * dz0);
            double dr2 = 2 * x * dxd + 2 * y * dyd;
            double dcd = k1 * dr2 + 2 * k2 * r2 * dr2 + 3 * k3 * r4 * dr2;
            double da1 = 2 * (x * dyd + y * dxd);
            J[(2 * i) * 6 + k] = fx * (dxd * cdist + x * dcd + p1 * da1 + p2 * (dr2 + 4 * x * dxd));
            J[(2 * i + 1) * 6 + k] = fy * (dyd * cdist + y * dcd + p1 * (dr2 + 4 * y * dyd) + p2 * da1);
        }
    }
    return sqrt(e2);
}

static void levmarq_step(const double *JtJ, const double *JtErr, int lambdaLg10, const double *prev, double *param)
{   /* CvLevMarq::step(): diag(JtJ) *= 1 + lambda; param = prevParam - JtJN^-1 JtErr */
    double M[36], d[6], lambda = exp(lambdaLg10 * log(10.));
    memcpy(M, JtJ, sizeof(M)); memcpy(d, JtErr, sizeof(d));
    for (int a = 0; a < 6; a++) M[a * 6 + a] *= 1. + lambda;
    if (!solve_n(M, d, 6)) memset(d, 0, sizeof(d));
    for (int a = 0; a < 6; a++) param[a] = prev[a] - d[a];
}

static void solve_pnp_planar4(const double *obj, const double *ip, const double *K, const double *D, double *rvec, double *tvec)
{
    /* init: homography obj.xy -> undistorted normalised points (SURVEY App. A pose recipe; OpenCV's
     * planar branch of findExtrinsicCameraParams2) */
    double un[8];
    orc_undistort_points(ip, 4, K, D, un);
    double A[64], b[8];
    memset(A, 0, sizeof(A));
    for (int i = 0; i < 4; i++) {
        double X = obj[3 * i], Y = obj[3 * i + 1], u = un[2 * i], v = un[2 * i + 1];
        double *r0 = A + (2 * i) * 8, *r1 = A + (2 * i + 1) * 8;
        r0[0] = X; r0[1] = Y; r0[2] = 1; r0[6] = -u * X; r0[7] = -u * Y; b[2 * i] = u;
        r1[3] = X; r1[4] = Y; r1[5] = 1; r1[6] = -v * X; r1[7] = -v * Y; b[2 * i + 1] = v;
    }
    solve_n(A, b, 8);
    double h[9] = {b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], 1.0};
    double n1 = sqrt(h[0] * h[0] + h[3] * h[3] + h[6] * h[6]);
    double n2 = sqrt(h[1] * h[1] + h[4] * h[4] + h[7] * h[7]);
    double R[9];
    double a1[3] = {h[0] / n1, h[3] / n1, h[6] / n1}, a2[3] = {h[1] / n2, h[4] / n2, h[7] / n2};
    double a3[3] = {a1[1] * a2[2] - a1[2] * a2[1], a1[2] * a2[0] - a1[0] * a2[2], a1[0] * a2[1] - a1[1] * a2[0]};
    for (int i = 0; i < 3; i++) { R[3 * i] = a1[i]; R[3 * i + 1] = a2[i]; R[3 * i + 2] = a3[i]; }
    nearest_rotation(R);
    double param[6], prev[6];
    rot_to_rvec(R, param);
    double sc = 2. / (n1 + n2);
    param[3] = h[2] * sc; param[4] = h[5] * sc; param[5] = h[8] * sc;

    /* OpenCV's CvLevMarq driven as in findExtrinsicCameraParams2: criteria (20 iterations, FLT_EPSILON) */
    int lambdaLg10 = -3, iters = 0;
    double prevErrNorm = 0, err[8], J[48], JtJ[36], JtErr[6];
    for (;;) {
        double en = reproj_err(obj, ip, param, K, D, err, J);          /* state CALC_J */
        for (int a = 0; a < 6; a++) {
            double s = 0;
            for (int i = 0; i < 8; i++) s += J[i * 6 + a] * err[i];
            JtErr[a] = s;
            for (int c = 0; c < 6; c++) { double t = 0; for (int i = 0; i < 8; i++) t += J[i * 6 + a] * J[i * 6 + c]; JtJ[a * 6 + c] = t; }
        }
        memcpy(prev, param, sizeof(prev));
        levmarq_step(JtJ, JtErr, lambdaLg10, prev, param);
        if (iters == 0) prevErrNorm = en;
        double errNorm = reproj_err(obj, ip, param, K, D, err, NULL);   /* state CHECK_ERR */
        while (errNorm > prevErrNorm && ++lambdaLg10 <= 16) {
            levmarq_step(JtJ, JtErr, lambdaLg10, prev, param);
            errNorm = reproj_err(obj, ip, param, K, D, err, NULL);
        }
        lambdaLg10 = lambdaLg10 - 1 > -16 ? lambdaLg10 - 1 : -16;
        double dn = 0, pn = 0;
        for (int a = 0; a < 6; a++) { dn += (param[a] - prev[a]) * (param[a] - prev[a]); pn += prev[a] * prev[a]; }
        if (++iters >= 20 || sqrt(dn) / sqrt(pn) < FLT_EPSILON) break;
        prevErrNorm = errNorm;
    }
    memcpy(rvec, param, 3 * sizeof(double));
    memcpy(tvec, param + 3, 3 * sizeof(double));
}

int orc_estimate_pose_single_markers(const float *corners, int n, double L, const double *K,
                                     const double *D, int nD, double *rvecs, double *tvecs)
{
    double Dd[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < nD && i < 5; i++) Dd[i] = D[i];
    double h = L / 2.0;
    /* object points: aruco_slam.h:189 (float literals there; L/2 exactly representable cases aside,
     * estimatePoseSingleMarkers itself builds them as Vec3f(-L/2.f, L/2.f, 0) ...) */
    float hf = (float)L / 2.f;
    (void)h;
    double obj[12] = {-hf, hf, 0, hf, hf, 0, hf, -hf, 0, -hf, -hf, 0};
    for (int m = 0; m < n; m++) {
        double ip[8];
        for (int i = 0; i < 8; i++) ip[i] = corners[8 * m + i];
        solve_pnp_planar4(obj, ip, K, Dd, rvecs + 3 * m, tvecs + 3 * m);
    }
    return n;
}

/* ---- observation mapping (reference src/aruco_slam.cpp:325-374) ---- */
static void norm_angle(double *a)
{   /* aruco_slam.cpp:412-421 single wrap */
    const double PI = 3.14159265358979323846, TWO_PI = 2.0 * PI;
    if (*a >= PI) *a -= TWO_PI;
    if (*a < -PI) *a += TWO_PI;
}

int orc_make_observations(const float *corners, const int32_t *ids, const double *rvecs,
                          const double *tvecs, int n, const double *K, const double *D,
                          const orc_slam_params *sp, orc_observation *out)
{
    int cnt = 0;
    float hf = (float)sp->marker_length / 2.f;                 /* aruco_slam.h:189 objectPoints_ (Point3f) */
    double obj[12] = {-hf, hf, 0, hf, hf, 0, hf, -hf, 0, -hf, -hf, 0};
    for (int i = 0; i < n; i++) {
        const double *t = tvecs + 3 * i, *r = rvecs + 3 * i;
        float dist = (float)sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);      /* :327 */
        if (dist > sp->useful_distance_threshold) continue;                       /* :329 */
        double R[9];
        orc_rodrigues(r, R);                                                      /* :354 */
        double x = t[2] + sp->r2c_tx;                                             /* :359 */
        double y = -t[0] + sp->r2c_ty;                                            /* :360 */
        double theta = atan2(-R[2], R[8]);                                        /* :361 */
        norm_angle(&theta);
        /* CalculateCovariance :437-471; projectPoints output is Point2f */
        double proj[8];
        orc_project_points(obj, 4, r, t, K, D, proj);
        const float *c = corners + 8 * i;
        double total = 0;
        for (int j = 0; j < 4; j++) {
            double ddx = (double)c[2 * j] - (double)(float)proj[2 * j], ddy = (double)c[2 * j + 1] - (double)(float)proj[2 * j + 1];
            double err = sqrt(ddx * ddx + ddy * ddy);
            total += err * err;
        }
        double rms = total / 4.0;
        double dgx = (double)c[0] - (double)c[4], dgy = (double)c[1] - (double)c[5];
        double oe = (rms / sqrt(dgx * dgx + dgy * dgy)) * (sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]) / sp->marker_length);
        double cov[9] = {oe * sp->R_x + 1e-2, 0, 0, 0, oe * sp->R_y + 1e-2, 0, 0, 0, oe * sp->R_theta + 1e-3};
        double fro = sqrt(cov[0] * cov[0] + cov[4] * cov[4] + cov[8] * cov[8]);
        if (fro > 1) continue;                                                    /* :367 */
        orc_observation *o = &out[cnt++];
        o->aruco_id = ids[i]; o->aruco_index = -1; o->x = x; o->y = y; o->theta = theta;
        memcpy(o->cov, cov, sizeof(cov));
    }
    return cnt;
}

// pose_core.h -- per-marker pose and observation arithmetic (FP64), host/device.
// Replaces cv::aruco::estimatePoseSingleMarkers (per-marker cv::solvePnP, SOLVEPNP_ITERATIVE;
// reference src/aruco_slam.cpp:314), cv::Rodrigues / cv::projectPoints (:354, :441) and the
// reference's own observation mapping and CalculateCovariance (:325-374, :412-421, :437-471).
//
// solvePnP ITERATIVE minimises the distorted reprojection error of the 4 corners starting from
// a homography decomposition (SURVEY.md App. A "Pose") with OpenCV's Levenberg-Marquardt
// schedule on (rvec, tvec).  Small markers seen nearly head-on have two close local minima
// (planar pose ambiguity), so the iteration schedule -- not just the cost -- is followed:
// same initialisation, same parametrisation, same damping rule, analytic Jacobians.
#pragma once
#include "core.h"

namespace b2a {

struct Camera {
    double fx, fy, cx, cy;
    double k1, k2, p1, p2, k3;
};

B2A_HD void rodrigues_to_R(const double *r, double *R)
{
    const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (th < DBL_EPSILON) { for (int i = 0; i < 9; ++i) R[i] = 0; R[0] = R[4] = R[8] = 1; return; }
    double s, c;
    sincos(th, &s, &c);
    const double c1 = 1 - c, it = 1 / th;
    const double x = r[0] * it, y = r[1] * it, z = r[2] * it;
    R[0] = c + c1 * x * x;     R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y;     R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}

B2A_HD void R_to_rodrigues(const double *R, double *r)
{
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
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
    const double vth = th / (2 * s);
    r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
}

// normalised camera point -> pixel, optionally with d(pixel)/d(x,y)
B2A_HD void distort_project(const Camera &cam, double x, double y, double &u, double &v, double *J /* 2x2 or null */)
{
    const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    const double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
    const double cd = 1 + cam.k1 * r2 + cam.k2 * r4 + cam.k3 * r6;
    const double xd = x * cd + cam.p1 * a1 + cam.p2 * a2;
    const double yd = y * cd + cam.p1 * a3 + cam.p2 * a1;
    u = xd * cam.fx + cam.cx;
    v = yd * cam.fy + cam.cy;
    if (J) {
        const double dcd = cam.k1 + 2 * cam.k2 * r2 + 3 * cam.k3 * r4;          // d cd / d r2
        const double dxd_dx = cd + x * dcd * 2 * x + cam.p1 * 2 * y + cam.p2 * (2 * x + 4 * x);
        const double dxd_dy = x * dcd * 2 * y + cam.p1 * 2 * x + cam.p2 * 2 * y;
        const double dyd_dx = y * dcd * 2 * x + cam.p1 * 2 * x + cam.p2 * 2 * y;
        const double dyd_dy = cd + y * dcd * 2 * y + cam.p1 * (2 * y + 4 * y) + cam.p2 * 2 * x;
        J[0] = cam.fx * dxd_dx; J[1] = cam.fx * dxd_dy; J[2] = cam.fy * dyd_dx; J[3] = cam.fy * dyd_dy;
    }
}

// cv::projectPoints for one object point
B2A_HD void project_point(const Camera &cam, const double *R, const double *t, const double *X, double &u, double &v)
{
    double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    z = z != 0 ? 1. / z : 1;
    distort_project(cam, x * z, y * z, u, v, nullptr);
}

// cv::undistortPoints default: 5 fixed-point iterations
B2A_HD void undistort_point(const Camera &cam, double u, double v, double &xo, double &yo)
{
    double x = (u - cam.cx) / cam.fx, y = (v - cam.cy) / cam.fy;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {
        const double r2 = x * x + y * y;
        const double icd = 1. / (1 + ((cam.k3 * r2 + cam.k2) * r2 + cam.k1) * r2);
        if (icd < 0) { x = x0; y = y0; break; }
        const double dx = 2 * cam.p1 * x * y + cam.p2 * (r2 + 2 * x * x);
        const double dy = cam.p1 * (r2 + 2 * y * y) + 2 * cam.p2 * x * y;
        x = (x0 - dx) * icd;
        y = (y0 - dy) * icd;
    }
    xo = x; yo = y;
}

// Gaussian elimination with partial pivoting (first largest |pivot|), written with compile-time
// indices only -- the row exchange is a predicated swap against every candidate row -- so that the
// system lives in registers instead of a dynamically indexed local-memory array.
template <int N>
B2A_HD_NOINLINE bool solve_linear(double *A, double *b)
{
    B2A_UNROLL
    for (int i = 0; i < N; ++i) {
        int k = i;
        double best = fabs(A[i * N + i]);
    B2A_UNROLL
        for (int j = i + 1; j < N; ++j) { const double v = fabs(A[j * N + i]); if (v > best) { best = v; k = j; } }
        if (best < 1e-300) return false;
    B2A_UNROLL
        for (int j = i + 1; j < N; ++j) {
            const bool sw = (k == j);
    B2A_UNROLL
            for (int c = 0; c < N; ++c) { const double x = A[i * N + c], y = A[j * N + c]; A[i * N + c] = sw ? y : x; A[j * N + c] = sw ? x : y; }
            const double x = b[i], y = b[j]; b[i] = sw ? y : x; b[j] = sw ? x : y;
        }
    B2A_UNROLL
        for (int j = i + 1; j < N; ++j) {
            const double a = A[j * N + i] / A[i * N + i];
    B2A_UNROLL
            for (int c = i; c < N; ++c) A[j * N + c] -= a * A[i * N + c];
            b[j] -= a * b[i];
        }
    }
    B2A_UNROLL
    for (int i = N - 1; i >= 0; --i) {
        double s = b[i];
    B2A_UNROLL
        for (int k = i + 1; k < N; ++k) s -= A[i * N + k] * b[k];
        b[i] = s / A[i * N + i];
    }
    return true;
}

B2A_HD void nearest_rotation(double *R)
{
    B2A_UNROLL
    for (int it = 0; it < 60; ++it) {
        double a[9];
        B2A_UNROLL
        for (int i = 0; i < 9; ++i) a[i] = R[i];
        const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
        const double id = 1 / det;
        double iT[9];
        iT[0] = (a[4] * a[8] - a[5] * a[7]) * id; iT[1] = (a[5] * a[6] - a[3] * a[8]) * id; iT[2] = (a[3] * a[7] - a[4] * a[6]) * id;
        iT[3] = (a[2] * a[7] - a[1] * a[8]) * id; iT[4] = (a[0] * a[8] - a[2] * a[6]) * id; iT[5] = (a[1] * a[6] - a[0] * a[7]) * id;
        iT[6] = (a[1] * a[5] - a[2] * a[4]) * id; iT[7] = (a[2] * a[3] - a[0] * a[5]) * id; iT[8] = (a[0] * a[4] - a[1] * a[3]) * id;
        double diff = 0;
        B2A_UNROLL
        for (int i = 0; i < 9; ++i) { const double v = 0.5 * (a[i] + iT[i]); diff += fabs(v - a[i]); R[i] = v; }
        if (diff < 1e-15) break;
    }
}

// cv::Rodrigues with its 3x9 Jacobian: dR[i*9 + k] = d R_flat[k] / d r[i]
B2A_HD void rodrigues_with_jacobian(const double *rv, double *R, double *dR)
{
    const double th = sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
    if (th < DBL_EPSILON) {
        B2A_UNROLL
        for (int i = 0; i < 9; ++i) R[i] = 0;
        R[0] = R[4] = R[8] = 1;
        B2A_UNROLL
        for (int i = 0; i < 27; ++i) dR[i] = 0;
        dR[5] = dR[15] = dR[19] = -1;
        dR[7] = dR[11] = dR[21] = 1;
        return;
    }
    double s, c;
    sincos(th, &s, &c);
    const double c1 = 1. - c, it = 1. / th;
    const double rx = rv[0] * it, ry = rv[1] * it, rz = rv[2] * it;
    const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    const double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    B2A_UNROLL
    for (int k = 0; k < 9; ++k) R[k] = c * I[k] + c1 * rrt[k] + s * r_x[k];
    const double drrt[27] = {rx + rx, ry, rz, ry, 0, 0, rz, 0, 0,
                             0, rx, 0, rx, ry + ry, rz, 0, rz, 0,
                             0, 0, rx, 0, 0, ry, rx, ry, rz + rz};
    const double d_r_x[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0,
                              0, 0, 1, 0, 0, 0, -1, 0, 0,
                              0, -1, 0, 1, 0, 0, 0, 0, 0};
    const double rr[3] = {rx, ry, rz};
    B2A_UNROLL
    for (int i = 0; i < 3; ++i) {
        const double ri = rr[i];
        const double a0 = -s * ri, a1 = (s - 2 * c1 * it) * ri, a2 = c1 * it, a3 = (c - s * it) * ri, a4 = s * it;
        B2A_UNROLL
        for (int k = 0; k < 9; ++k)
            dR[i * 9 + k] = a0 * I[k] + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * r_x[k] + a4 * d_r_x[i * 9 + k];
    }
}

// One row of the reprojection system at (rvec, tvec) = p[0..5] with R (and dR) = Rodrigues(p[0..2]):
// row r = 2 i + c is coordinate c (0 = u, 1 = v) of corner i.  out[6] = residual; with dR also
// out[0..5] = d residual / d p (cv::projectPoints' dpdr | dpdt).
B2A_HD void pose_row(const Camera &cam, double h, const float *corners, const double *p, const double *R, const double *dR, int r, double *out)
{
    const int i = r >> 1, c = r & 1;
    // object corner i of the marker square (Vec3f(-L/2, L/2, 0), (L/2, L/2, 0), (L/2, -L/2, 0), (-L/2, -L/2, 0))
    const double X[3] = {(i == 1 || i == 2) ? h : -h, (i < 2) ? h : -h, 0.0};
    const double Px = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + p[3];
    const double Py = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + p[4];
    const double Pz = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + p[5];
    const double iz = Pz != 0 ? 1. / Pz : 1;
    const double x = Px * iz, y = Py * iz;
    double u, v, Jd[4];
    distort_project(cam, x, y, u, v, dR ? Jd : nullptr);
    out[6] = (c ? v : u) - (double)corners[r];
    if (!dR) return;
    B2A_UNROLL
    for (int k = 0; k < 6; ++k) {
        double dP[3];
        if (k < 3) {
            const double *d = dR + k * 9;
            dP[0] = d[0] * X[0] + d[1] * X[1] + d[2] * X[2];
            dP[1] = d[3] * X[0] + d[4] * X[1] + d[5] * X[2];
            dP[2] = d[6] * X[0] + d[7] * X[1] + d[8] * X[2];
        } else { dP[0] = (k == 3); dP[1] = (k == 4); dP[2] = (k == 5); }
        const double dx = iz * (dP[0] - x * dP[2]);
        const double dy = iz * (dP[1] - y * dP[2]);
        out[k] = c ? (Jd[2] * dx + Jd[3] * dy) : (Jd[0] * dx + Jd[1] * dy);
    }
}

// 10^k for k in [-16, 16] (a table in constant memory on the device: building it on the stack in every LM step cost 6 % of k_pose)
#define B2A_P10_VALUES {1e-16, 1e-15, 1e-14, 1e-13, 1e-12, 1e-11, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1e0, \
                        1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16}
#if defined(__CUDACC__)
__device__ __constant__ double k_p10_device[33] = B2A_P10_VALUES;
#endif
B2A_HD double pow10_table(int k)
{
#if defined(__CUDA_ARCH__)
    return k_p10_device[k + 16];
#else
    const double p10[33] = B2A_P10_VALUES;
    return p10[k + 16];
#endif
}

// one LM trial step of OpenCV's CvLevMarq::step(): p = prev - (JtJ with diagonal * (1 + 10^lg))^-1 JtErr
// LDL^T solve of the 6x6 symmetric positive definite LM system (upper part of M is not read);
// false if a pivot is not positive.  (CvLevMarq solves the same system by SVD; only the solution matters.)
B2A_HD bool solve_spd6(const double *M, double *d)
{
    double L[36], D[6], Dinv[6];
    B2A_UNROLL
    for (int j = 0; j < 6; ++j) {
        double v = M[j * 6 + j];
        B2A_UNROLL
        for (int k = 0; k < j; ++k) v -= L[j * 6 + k] * L[j * 6 + k] * D[k];
        if (!(v > 0)) return false;
        D[j] = v;
        const double inv = 1. / v;
        Dinv[j] = inv;
        B2A_UNROLL
        for (int i = j + 1; i < 6; ++i) {
            double t = M[i * 6 + j];
            B2A_UNROLL
            for (int k = 0; k < j; ++k) t -= L[i * 6 + k] * L[j * 6 + k] * D[k];
            L[i * 6 + j] = t * inv;
        }
        // forward substitution and the division by D folded in
        double y = d[j];
        B2A_UNROLL
        for (int k = 0; k < j; ++k) y -= L[j * 6 + k] * d[k];
        d[j] = y;
    }
    B2A_UNROLL
    for (int j = 0; j < 6; ++j) d[j] = d[j] * Dinv[j];                 // the pivots' reciprocals are already there (one rounding more than a division; the LM iteration absorbs it)
    B2A_UNROLL
    for (int j = 5; j >= 0; --j) {
        double x = d[j];
        B2A_UNROLL
        for (int k = j + 1; k < 6; ++k) x -= L[k * 6 + j] * d[k];
        d[j] = x;
    }
    return true;
}

// one LM trial step of OpenCV's CvLevMarq::step(): p = prev - (JtJ with diagonal * (1 + 10^lg))^-1 JtErr
B2A_HD void lm_step(const double *shN /* [6][7]: JtJ row | JtErr */, int lambdaLg10, const double *prev, double *p)
{
    double M[36], d[6];
    // CvLevMarq: lambda = exp(lambdaLg10 * log(10)), lambdaLg10 in [-16, 16]
    const int li = lambdaLg10 < -16 ? -16 : (lambdaLg10 > 16 ? 16 : lambdaLg10);
    const double lambda = pow10_table(li);
    B2A_UNROLL
    for (int a = 0; a < 6; ++a) {
        B2A_UNROLL
        for (int c = 0; c < 6; ++c) M[a * 6 + c] = shN[a * 7 + c];
        M[a * 6 + a] *= 1. + lambda;
        d[a] = shN[a * 7 + 6];
    }
    if (!solve_spd6(M, d)) {                        // not positive definite (degenerate corners): pivoted elimination
        B2A_UNROLL
        for (int a = 0; a < 6; ++a) d[a] = shN[a * 7 + 6];
        if (!solve_linear<6>(M, d)) { B2A_UNROLL for (int a = 0; a < 6; ++a) d[a] = 0; }
    }
    B2A_UNROLL
    for (int a = 0; a < 6; ++a) p[a] = prev[a] - d[a];
}

// A group of lanes that solves one marker together: the 8 rows of the reprojection system and the
// 6 rows of the normal equations are spread over the lanes; everything else (Rodrigues, the 6x6
// solve, the schedule) is computed redundantly by every lane from the same shared values, so the
// control flow is identical on all lanes.  `sh` = POSE_SH doubles visible to the whole group.
struct OneLane {
    B2A_HD int lane() const { return 0; }
    B2A_HD int nlanes() const { return 1; }
    B2A_HD void sync() const {}
    B2A_HD void mark() const {}
};
constexpr int POSE_SH = 8 * 7 + 6 * 7;

// rows of the reprojection system at p into shJ (lane r computes row r); returns the L2 norm of the
// residual vector (the same value on every lane)
template <class LG>
B2A_HD double pose_rows(const LG &lg, const Camera &cam, double h, const float *corners, const double *p, double *shJ)
{
    lg.sync();                                   // the previous contents have been consumed by every lane
    {
        double R[9], dR[27];
        rodrigues_with_jacobian(p, R, dR);
        for (int r = lg.lane(); r < 8; r += lg.nlanes()) pose_row(cam, h, corners, p, R, dR, r, shJ + r * 7);
    }
    lg.sync();
    double e = 0;
    B2A_UNROLL
    for (int i = 0; i < 4; ++i) { const double a = shJ[(2 * i) * 7 + 6], b = shJ[(2 * i + 1) * 7 + 6]; e += a * a + b * b; }
    return sqrt(e);
}

// one marker: corners (4x2, image pixels) -> rvec, tvec.  Follows cv::solvePnP(SOLVEPNP_ITERATIVE)
// for 4 coplanar points: homography initialisation, then OpenCV's Levenberg-Marquardt schedule
// (CvLevMarq: lambda = 10^-3 start, x10 on a worse step (up to 10^16), /10 after every accepted
// iteration, at most 20 iterations, stop when the relative parameter change < FLT_EPSILON) on the
// rotation *vector* and translation, so that the same local minimum and rvec branch are reached.
template <class LG>
B2A_HD void solve_marker_pose(const LG &lg, const Camera &cam, float marker_length, const float *corners, double *sh, double *rvec, double *tvec)
{
    const float hf = marker_length / 2.f;                         // Vec3f(-L/2.f, L/2.f, 0) ...
    const double h = (double)hf;
    // ---- initialisation: homography obj.xy -> undistorted normalised points (every lane) ----
    double p[6], prev[6];
    {
        // homography obj.xy -> undistorted normalised points, h33 = 1.  The object points are the corners
        // of a square, so the 8x8 system has the closed form of the unit-square -> quadrilateral map
        // (s = (X + h) / 2h, t = (h - Y) / 2h sends corners 0..3 to (0,0), (1,0), (1,1), (0,1)).
        double qx[4], qy[4];
        B2A_UNROLL
        for (int i = 0; i < 4; ++i) undistort_point(cam, (double)corners[2 * i], (double)corners[2 * i + 1], qx[i], qy[i]);
        const double dx1 = qx[1] - qx[2], dx2 = qx[3] - qx[2], sx = qx[0] - qx[1] + qx[2] - qx[3];
        const double dy1 = qy[1] - qy[2], dy2 = qy[3] - qy[2], sy = qy[0] - qy[1] + qy[2] - qy[3];
        const double den = dx1 * dy2 - dx2 * dy1;
        const double g = (sx * dy2 - dx2 * sy) / den, k = (dx1 * sy - sx * dy1) / den;
        const double a = qx[1] - qx[0] + g * qx[1], b = qx[3] - qx[0] + k * qx[3], c = qx[0];
        const double d = qy[1] - qy[0] + g * qy[1], e = qy[3] - qy[0] + k * qy[3], f = qy[0];
        const double i2h = 1. / (2. * h);
        // H = Hs * [[i2h, 0, 1/2], [0, -i2h, 1/2], [0, 0, 1]], then scaled to h33 = 1
        const double w = 1. / (0.5 * (g + k) + 1.);
        const double hm[9] = {a * i2h * w, -b * i2h * w, (0.5 * (a + b) + c) * w,
                              d * i2h * w, -e * i2h * w, (0.5 * (d + e) + f) * w,
                              g * i2h * w, -k * i2h * w, 1.0};
        const double n1 = sqrt(hm[0] * hm[0] + hm[3] * hm[3] + hm[6] * hm[6]);
        const double n2 = sqrt(hm[1] * hm[1] + hm[4] * hm[4] + hm[7] * hm[7]);
        const double a1[3] = {hm[0] / n1, hm[3] / n1, hm[6] / n1}, a2[3] = {hm[1] / n2, hm[4] / n2, hm[7] / n2};
        const double a3[3] = {a1[1] * a2[2] - a1[2] * a2[1], a1[2] * a2[0] - a1[0] * a2[2], a1[0] * a2[1] - a1[1] * a2[0]};
        double R[9];
        for (int i = 0; i < 3; ++i) { R[3 * i] = a1[i]; R[3 * i + 1] = a2[i]; R[3 * i + 2] = a3[i]; }
        nearest_rotation(R);
        R_to_rodrigues(R, p);
        const double sc = 2. / (n1 + n2);
        p[3] = hm[2] * sc; p[4] = hm[5] * sc; p[5] = hm[8] * sc;
    }
    // ---- Levenberg-Marquardt, CvLevMarq schedule ----
    double *shJ = sh, *shN = sh + 56;
    int lambdaLg10 = -3, iters = 0;
    double prevErr = 0;
    const int max_iter = 20;
    lg.mark();
    // rows (Jacobian + residual) of the current estimate; every trial step below leaves the rows of the
    // accepted estimate in shJ, so Rodrigues and the projection are evaluated once per trial
    double err = pose_rows(lg, cam, h, corners, p, shJ);
    for (;;) {
        for (int a = lg.lane(); a < 6; a += lg.nlanes()) {
            double s = 0;
            B2A_UNROLL
            for (int i = 0; i < 8; ++i) s += shJ[i * 7 + a] * shJ[i * 7 + 6];
            shN[a * 7 + 6] = s;
            B2A_UNROLL
            for (int c = 0; c < 6; ++c) {
                double q = 0;
                B2A_UNROLL
                for (int i = 0; i < 8; ++i) q += shJ[i * 7 + a] * shJ[i * 7 + c];
                shN[a * 7 + c] = q;
            }
        }
        lg.sync();
        B2A_UNROLL
        for (int a = 0; a < 6; ++a) prev[a] = p[a];
        if (iters == 0) prevErr = err;
        lm_step(shN, lambdaLg10, prev, p);
        err = pose_rows(lg, cam, h, corners, p, shJ);
        while (err > prevErr && ++lambdaLg10 <= 16) {
            lm_step(shN, lambdaLg10, prev, p);
            err = pose_rows(lg, cam, h, corners, p, shJ);
        }
        lambdaLg10 = lambdaLg10 - 1 > -16 ? lambdaLg10 - 1 : -16;
        double dn = 0, pn = 0;
        B2A_UNROLL
        for (int a = 0; a < 6; ++a) { dn += (p[a] - prev[a]) * (p[a] - prev[a]); pn += prev[a] * prev[a]; }
        lg.mark();
        if (++iters >= max_iter || sqrt(dn) < (double)FLT_EPSILON * sqrt(pn)) break;
        prevErr = err;
    }
    if (lg.lane() == 0) {
        rvec[0] = p[0]; rvec[1] = p[1]; rvec[2] = p[2];
        tvec[0] = p[3]; tvec[1] = p[4]; tvec[2] = p[5];
    }
}

// ---- observation mapping (reference src/aruco_slam.cpp:325-374, 437-471) ----
struct ObsParams {
    double R_x, R_y, R_theta, marker_length, r2c_tx, r2c_ty;
    float useful_distance_threshold;
};
struct Observation {
    int32_t aruco_id, aruco_index;
    double x, y, theta;
    double cov[9];
};
B2A_HD void norm_angle(double &a)
{   // aruco_slam.cpp:412-421, single wrap
    const double PI = 3.14159265358979323846, TWO_PI = 2.0 * PI;
    if (a >= PI) a -= TWO_PI;
    if (a < -PI) a += TWO_PI;
}
// returns false when the marker is gated out (range :327-333 or covariance norm :367-368)
B2A_HD bool make_observation(const Camera &cam, const ObsParams &op, const float *corners, int id,
                             const double *rvec, const double *tvec, Observation &o)
{
    const double tn = sqrt(tvec[0] * tvec[0] + tvec[1] * tvec[1] + tvec[2] * tvec[2]);
    const float dist = (float)tn;
    if (dist > op.useful_distance_threshold) return false;
    double R[9];
    rodrigues_to_R(rvec, R);
    const double x = tvec[2] + op.r2c_tx, y = -tvec[0] + op.r2c_ty;
    double theta = atan2(-R[2], R[8]);
    norm_angle(theta);
    const float hf = (float)op.marker_length / 2.f;               // objectPoints_ (Point3f), aruco_slam.h:189
    const double h = (double)hf;
    const double obj[12] = {-h, h, 0, h, h, 0, h, -h, 0, -h, -h, 0};
    double total = 0;
    for (int j = 0; j < 4; ++j) {
        double u, v;
        project_point(cam, R, tvec, obj + 3 * j, u, v);
        const double dx = (double)corners[2 * j] - (double)(float)u, dy = (double)corners[2 * j + 1] - (double)(float)v;
        const double err = sqrt(dx * dx + dy * dy);
        total += err * err;
    }
    const double rms = total / 4.0;                                // "rmserror" is a mean of squares (:465)
    const double gx = (double)corners[0] - (double)corners[4], gy = (double)corners[1] - (double)corners[5];
    const double oe = (rms / sqrt(gx * gx + gy * gy)) * (tn / op.marker_length);
    for (int i = 0; i < 9; ++i) o.cov[i] = 0;
    o.cov[0] = oe * op.R_x + 1e-2; o.cov[4] = oe * op.R_y + 1e-2; o.cov[8] = oe * op.R_theta + 1e-3;
    const double fro = sqrt(o.cov[0] * o.cov[0] + o.cov[4] * o.cov[4] + o.cov[8] * o.cov[8]);
    if (fro > 1) return false;
    o.aruco_id = id; o.aruco_index = -1; o.x = x; o.y = y; o.theta = theta;
    return true;
}

}  // namespace b2a

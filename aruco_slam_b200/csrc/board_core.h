// board_core.h -- host-side geometry of refineDetectedMarkers (b2a_refine_detected_markers): where the board's undetected
// markers should appear in the image.  Replaces the two _projectUndetectedMarkers of cv2 4.13's aruco_detector.cpp (part of the
// cv::aruco surface of reference src/aruco_slam.cpp:313):
//   no camera   findHomography(board xy -> detected corners, method 0) + perspectiveTransform
//   camera      solvePnP(matched board corners, detected corners, ITERATIVE) + projectPoints: start from the plane homography when
//               the corners are (nearly) coplanar -- cv2's test, W[2] / W[1] < 1e-3 on the scatter's singular values -- else from
//               the DLT over at least 6 points
// A few dozen points per call: plain double-precision host code (the per-candidate bit extraction is the GPU part).
// Both fits are least-squares problems; cv2 reaches their minima with its normalised DLT + LM and its LM on (rvec, tvec), this
// file with the same DLT and Gauss-Newton, so the projections agree to ~1e-3 px.
#pragma once
#include <cmath>
#include <vector>

#include "pose_core.h"

namespace b2a {

// cyclic Jacobi on a symmetric n x n matrix (row-major, destroyed); eigenvalues ascending in w, eigenvectors in the columns of V
inline void sym_eig_jacobi(int n, double *A, double *V, double *w)
{
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i * n + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0;
        for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
        if (off < 1e-300) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p * n + q];
                if (std::fabs(apq) < 1e-300) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
                const double c = 1 / std::sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
    for (int i = 0; i < n; ++i)                         // selection sort, ascending
        for (int j = i + 1; j < n; ++j)
            if (w[j] < w[i]) {
                std::swap(w[i], w[j]);
                for (int k = 0; k < n; ++k) std::swap(V[k * n + i], V[k * n + j]);
            }
}

// findHomography(src, dst, 0): src, dst n x 2; H row-major with H[8] = 1
inline bool homography_ls(const double *src, const double *dst, int n, double *H, int refine_iters)
{
    if (n < 4) return false;
    double cM[2] = {0, 0}, cm[2] = {0, 0}, sM[2] = {0, 0}, sm[2] = {0, 0};
    for (int i = 0; i < n; ++i) for (int k = 0; k < 2; ++k) { cM[k] += src[2 * i + k]; cm[k] += dst[2 * i + k]; }
    for (int k = 0; k < 2; ++k) { cM[k] /= n; cm[k] /= n; }
    for (int i = 0; i < n; ++i) for (int k = 0; k < 2; ++k) { sM[k] += std::fabs(src[2 * i + k] - cM[k]); sm[k] += std::fabs(dst[2 * i + k] - cm[k]); }
    for (int k = 0; k < 2; ++k) { if (sM[k] < 1e-300 || sm[k] < 1e-300) return false; sM[k] = n / sM[k]; sm[k] = n / sm[k]; }
    double LtL[81] = {0};
    for (int i = 0; i < n; ++i) {
        const double X = (src[2 * i] - cM[0]) * sM[0], Y = (src[2 * i + 1] - cM[1]) * sM[1];
        const double x = (dst[2 * i] - cm[0]) * sm[0], y = (dst[2 * i + 1] - cm[1]) * sm[1];
        const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x}, Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
        for (int a = 0; a < 9; ++a) for (int b = 0; b < 9; ++b) LtL[a * 9 + b] += Lx[a] * Lx[b] + Ly[a] * Ly[b];
    }
    double V[81], w[9];
    sym_eig_jacobi(9, LtL, V, w);
    double H0[9];
    for (int k = 0; k < 9; ++k) H0[k] = V[k * 9 + 0];
    // H = inv(norm of dst) * H0 * (norm of src)
    const double A[9] = {1 / sm[0], 0, cm[0], 0, 1 / sm[1], cm[1], 0, 0, 1}, B[9] = {sM[0], 0, -cM[0] * sM[0], 0, sM[1], -cM[1] * sM[1], 0, 0, 1};
    double T[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { T[r * 3 + c] = 0; for (int k = 0; k < 3; ++k) T[r * 3 + c] += A[r * 3 + k] * H0[k * 3 + c]; }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { H[r * 3 + c] = 0; for (int k = 0; k < 3; ++k) H[r * 3 + c] += T[r * 3 + k] * B[k * 3 + c]; }
    if (std::fabs(H[8]) < 1e-300) return false;
    for (int k = 0; k < 9; ++k) H[k] /= H[8];
    if (n > 4)
        for (int it = 0; it < refine_iters; ++it) {            // Gauss-Newton on the reprojection error over h0 .. h7
            double JtJ[64] = {0}, Jtr[8] = {0};
            for (int i = 0; i < n; ++i) {
                const double X = src[2 * i], Y = src[2 * i + 1];
                const double ww = 1.0 / (H[6] * X + H[7] * Y + 1.0);
                const double xi = (H[0] * X + H[1] * Y + H[2]) * ww, yi = (H[3] * X + H[4] * Y + H[5]) * ww;
                const double Jx[8] = {X * ww, Y * ww, ww, 0, 0, 0, -X * ww * xi, -Y * ww * xi}, Jy[8] = {0, 0, 0, X * ww, Y * ww, ww, -X * ww * yi, -Y * ww * yi};
                const double rx = xi - dst[2 * i], ry = yi - dst[2 * i + 1];
                for (int a = 0; a < 8; ++a) { for (int b = 0; b < 8; ++b) JtJ[a * 8 + b] += Jx[a] * Jx[b] + Jy[a] * Jy[b]; Jtr[a] += Jx[a] * rx + Jy[a] * ry; }
            }
            if (!solve_linear<8>(JtJ, Jtr)) break;
            for (int a = 0; a < 8; ++a) H[a] -= Jtr[a];
        }
    return true;
}

inline void homography_apply(const double *H, double X, double Y, double &x, double &y)
{
    const double w = H[6] * X + H[7] * Y + H[8];
    x = (H[0] * X + H[1] * Y + H[2]) / w; y = (H[3] * X + H[4] * Y + H[5]) / w;
}

// solvePnP(ITERATIVE) for object points obj (n x 3) seen at img (n x 2).  Returns 0 ok, 1 degenerate, 2 points in general position
// but fewer than 6 of them (cv2 throws: "DLT algorithm needs at least 6 points")
inline int board_pose(const Camera &cam, const double *obj, const double *img, int n, double *rvec, double *tvec)
{
    if (n < 4) return 1;
    double mean[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) mean[k] += obj[3 * i + k];
    for (int k = 0; k < 3; ++k) mean[k] /= n;
    double C[9] = {0};
    for (int i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) C[a * 3 + b] += (obj[3 * i + a] - mean[a]) * (obj[3 * i + b] - mean[b]);
    double V[9], w[3];
    sym_eig_jacobi(3, C, V, w);
    if (w[2] <= 0) return 1;
    const bool planar = w[0] / w[1] < 1e-3;
    if (!planar && n < 6) return 2;
    double p[6];
    if (!planar) {
        // DLT: the 3 x 4 projection is the eigenvector of L^T L with the smallest eigenvalue (normalised image points), its left
        // block becomes a rotation (SVD), the translation is scaled alike
        double LtL[144] = {0};
        for (int i = 0; i < n; ++i) {
            double x, y;
            undistort_point(cam, img[2 * i], img[2 * i + 1], x, y);
            const double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
            const double Lx[12] = {X, Y, Z, 1, 0, 0, 0, 0, -x * X, -x * Y, -x * Z, -x}, Ly[12] = {0, 0, 0, 0, X, Y, Z, 1, -y * X, -y * Y, -y * Z, -y};
            for (int a = 0; a < 12; ++a) for (int b = 0; b < 12; ++b) LtL[a * 12 + b] += Lx[a] * Lx[b] + Ly[a] * Ly[b];
        }
        double V12[144], w12[12], RR[12];
        sym_eig_jacobi(12, LtL, V12, w12);
        for (int k = 0; k < 12; ++k) RR[k] = V12[k * 12 + 0];
        const double dRR = RR[0] * (RR[5] * RR[10] - RR[6] * RR[9]) - RR[1] * (RR[4] * RR[10] - RR[6] * RR[8]) + RR[2] * (RR[4] * RR[9] - RR[5] * RR[8]);
        if (dRR < 0) for (int k = 0; k < 12; ++k) RR[k] = -RR[k];
        double R[9], nrm = 0;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { R[r * 3 + c] = RR[r * 4 + c]; nrm += RR[r * 4 + c] * RR[r * 4 + c]; }
        if (nrm <= 0) return 1;
        nearest_rotation(R);
        R_to_rodrigues(R, p);
        const double sc = std::sqrt(3.0) / std::sqrt(nrm);                    // |R|_F / |RR[:, :3]|_F
        for (int k = 0; k < 3; ++k) p[3 + k] = RR[k * 4 + 3] * sc;
    } else {
    // plane frame: e1, e2 = the two in-plane axes (largest eigenvalues), e3 = normal, right-handed
    double E[9];
    for (int k = 0; k < 3; ++k) { E[k * 3 + 0] = V[k * 3 + 2]; E[k * 3 + 1] = V[k * 3 + 1]; E[k * 3 + 2] = V[k * 3 + 0]; }
    const double det = E[0] * (E[4] * E[8] - E[5] * E[7]) - E[1] * (E[3] * E[8] - E[5] * E[6]) + E[2] * (E[3] * E[7] - E[4] * E[6]);
    if (det < 0) for (int k = 0; k < 3; ++k) E[k * 3 + 2] = -E[k * 3 + 2];
    std::vector<double> uv((size_t)2 * n), mn((size_t)2 * n);
    for (int i = 0; i < n; ++i) {
        for (int c = 0; c < 2; ++c) { double s = 0; for (int k = 0; k < 3; ++k) s += (obj[3 * i + k] - mean[k]) * E[k * 3 + c]; uv[2 * i + c] = s; }
        undistort_point(cam, img[2 * i], img[2 * i + 1], mn[2 * i], mn[2 * i + 1]);
    }
    double H[9];
    if (!homography_ls(uv.data(), mn.data(), n, H, 0)) return 1;
    const double n1 = std::sqrt(H[0] * H[0] + H[3] * H[3] + H[6] * H[6]), n2 = std::sqrt(H[1] * H[1] + H[4] * H[4] + H[7] * H[7]);
    double lam = 2.0 / (n1 + n2);
    if (H[8] * lam < 0) lam = -lam;
    double Rh[9];                                         // columns r1, r2, r1 x r2
    const double r1[3] = {H[0] * lam, H[3] * lam, H[6] * lam}, r2[3] = {H[1] * lam, H[4] * lam, H[7] * lam};
    const double r3[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
    for (int k = 0; k < 3; ++k) { Rh[k * 3 + 0] = r1[k]; Rh[k * 3 + 1] = r2[k]; Rh[k * 3 + 2] = r3[k]; }
    nearest_rotation(Rh);
    double R[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { R[r * 3 + c] = 0; for (int k = 0; k < 3; ++k) R[r * 3 + c] += Rh[r * 3 + k] * E[c * 3 + k]; }   // Rh * E^T
    R_to_rodrigues(R, p);
    for (int k = 0; k < 3; ++k) p[3 + k] = H[2 + 3 * k] * lam - (R[k * 3] * mean[0] + R[k * 3 + 1] * mean[1] + R[k * 3 + 2] * mean[2]);
    }
    // Gauss-Newton on the reprojection error, numeric Jacobian
    auto residuals = [&](const double *q, double *r) {
        double Rq[9];
        rodrigues_to_R(q, Rq);
        for (int i = 0; i < n; ++i) {
            double u, v;
            project_point(cam, Rq, q + 3, obj + 3 * i, u, v);
            r[2 * i] = u - img[2 * i]; r[2 * i + 1] = v - img[2 * i + 1];
        }
    };
    std::vector<double> r0((size_t)2 * n), rp((size_t)2 * n), rm((size_t)2 * n), J((size_t)12 * n);
    for (int it = 0; it < 30; ++it) {
        residuals(p, r0.data());
        for (int k = 0; k < 6; ++k) {
            double qp[6], qm[6];
            for (int a = 0; a < 6; ++a) { qp[a] = p[a]; qm[a] = p[a]; }
            qp[k] += 1e-7; qm[k] -= 1e-7;
            residuals(qp, rp.data()); residuals(qm, rm.data());
            for (int i = 0; i < 2 * n; ++i) J[(size_t)i * 6 + k] = (rp[i] - rm[i]) / 2e-7;
        }
        double JtJ[36] = {0}, Jtr[6] = {0};
        for (int i = 0; i < 2 * n; ++i)
            for (int a = 0; a < 6; ++a) { for (int b = 0; b < 6; ++b) JtJ[a * 6 + b] += J[(size_t)i * 6 + a] * J[(size_t)i * 6 + b]; Jtr[a] += J[(size_t)i * 6 + a] * r0[i]; }
        if (!solve_linear<6>(JtJ, Jtr)) return 1;
        double sn = 0;
        for (int a = 0; a < 6; ++a) { p[a] -= Jtr[a]; sn += Jtr[a] * Jtr[a]; }
        if (sn < 1e-24) break;
    }
    for (int k = 0; k < 3; ++k) { rvec[k] = p[k]; tvec[k] = p[3 + k]; }
    return 0;
}

}  // namespace b2a

// ekf_kernels.cuh -- EKF landmark update with the covariance resident in HBM (FP64).
// Replaces the Eigen arithmetic of the reference:
//   k_ekf_predict   ArucoSlam::addEncoder            src/aruco_slam.cpp:21-74
//   k_ekf_gain      known-landmark branch, gain      src/aruco_slam.cpp:108-146
//   k_ekf_rank3     mu_ += K ze; Sigma = (I-K Gx)Sigma  src/aruco_slam.cpp:202-204
//   k_ekf_augment   new-landmark branch              src/aruco_slam.cpp:208-256
// The reference forms dense N x N products (Hx Sigma Hx^T, (I - K Gx) Sigma); Gx is nonzero in
// six columns only, so the same results are  Sigma -= K (Gx Sigma)  (rank 3; differs <= 4e-16,
// SURVEY probe P11) and a 3-row / 3-column update for the prediction.  The rank-3 pass is the
// HBM-bound kernel: 16 N^2 bytes per observation.
// Sigma is row-major with leading dimension LD (capacity), mu has LD entries.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include "pose_core.h"

namespace b2a {

struct EkfObs {            // one observation, passed by value
    int index;             // landmark index (>= 0: known, -1: new)
    double z[3];
    double Rk[9];
};

__device__ __forceinline__ void inv3x3(const double *a, double *t)
{
    double d = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    d = 1. / d;
    t[0] = (a[4] * a[8] - a[5] * a[7]) * d; t[1] = (a[2] * a[7] - a[1] * a[8]) * d; t[2] = (a[1] * a[5] - a[2] * a[4]) * d;
    t[3] = (a[5] * a[6] - a[3] * a[8]) * d; t[4] = (a[0] * a[8] - a[2] * a[6]) * d; t[5] = (a[2] * a[3] - a[0] * a[5]) * d;
    t[6] = (a[3] * a[7] - a[4] * a[6]) * d; t[7] = (a[1] * a[6] - a[0] * a[7]) * d; t[8] = (a[0] * a[4] - a[1] * a[3]) * d;
}

// geometry of the known-landmark update from the frame-start snapshot mu_s (:119-143)
__device__ __forceinline__ void ekf_linearise(const double *__restrict__ mu_s, const EkfObs &ob, double *Gxm /*3x6*/, double *ze)
{
    const int L = 3 + 3 * ob.index;
    const double mx = mu_s[L], my = mu_s[L + 1], mth = mu_s[L + 2];
    const double x = mu_s[0], y = mu_s[1], th = mu_s[2];
    const double s = sin(th), c = cos(th);
    const double gdx = mx - x, gdy = my - y;
    double gdt = mth - th;
    norm_angle(gdt);
    const double zh[3] = {gdx * c + gdy * s, -gdx * s + gdy * c, gdt};
    ze[0] = ob.z[0] - zh[0]; ze[1] = ob.z[1] - zh[1]; ze[2] = ob.z[2] - zh[2];
    norm_angle(ze[2]);
    const double G[18] = {-c, -s, -gdx * s + gdy * c, c, s, 0,
                          s, -c, -gdx * c - gdy * s, -s, c, 0,
                          0, 0, -1, 0, 0, 1};
    for (int i = 0; i < 18; ++i) Gxm[i] = G[i];
}

// K (N x 3, row-major) and GS = Gx Sigma (3 x N, row-major with pitch LD) from the current Sigma
__global__ void k_ekf_gain(const double *__restrict__ sigma, const double *__restrict__ mu_s, int N, int LD, EkfObs ob,
                           double *__restrict__ Kout, double *__restrict__ GS)
{
    __shared__ double s_G[18], s_Si[9];
    const int L = 3 + 3 * ob.index;
    if (threadIdx.x == 0) {
        double G[18], ze[3];
        ekf_linearise(mu_s, ob, G, ze);
        const int cols[6] = {0, 1, 2, L, L + 1, L + 2};
        double T[18];   // Gxm * Sigma_cc  (3 x 6)
        for (int r = 0; r < 3; ++r) for (int j = 0; j < 6; ++j) {
            double a = 0;
            for (int k = 0; k < 6; ++k) a += G[6 * r + k] * sigma[(size_t)cols[k] * LD + cols[j]];
            T[6 * r + j] = a;
        }
        // S = Gx Sigma Gx^T + Rk, evaluated as (Sigma Gx^T) restricted to the six rows, like the
        // reference's K = Sigma Gx^T (Gx Sigma Gx^T + Rk)^-1
        double SGc[18]; // (Sigma Gx^T)[cols[j]][r]
        for (int j = 0; j < 6; ++j) for (int r = 0; r < 3; ++r) {
            double a = 0;
            for (int k = 0; k < 6; ++k) a += sigma[(size_t)cols[j] * LD + cols[k]] * G[6 * r + k];
            SGc[3 * j + r] = a;
        }
        double S[9];
        for (int r = 0; r < 3; ++r) for (int c2 = 0; c2 < 3; ++c2) {
            double a = 0;
            for (int j = 0; j < 6; ++j) a += G[6 * r + j] * SGc[3 * j + c2];
            S[3 * r + c2] = a + ob.Rk[3 * r + c2];
        }
        (void)T;
        inv3x3(S, s_Si);
        for (int i = 0; i < 18; ++i) s_G[i] = G[i];
    }
    __syncthreads();
    const int cols[6] = {0, 1, 2, L, L + 1, L + 2};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        double sg[3] = {0, 0, 0}, gs[3] = {0, 0, 0};
        for (int k = 0; k < 6; ++k) {
            const double col = sigma[(size_t)i * LD + cols[k]];       // Sigma[i][cols[k]]
            const double row = sigma[(size_t)cols[k] * LD + i];       // Sigma[cols[k]][i]
            for (int r = 0; r < 3; ++r) { sg[r] += col * s_G[6 * r + k]; gs[r] += s_G[6 * r + k] * row; }
        }
        for (int c2 = 0; c2 < 3; ++c2) Kout[(size_t)i * 3 + c2] = sg[0] * s_Si[c2] + sg[1] * s_Si[3 + c2] + sg[2] * s_Si[6 + c2];
        for (int r = 0; r < 3; ++r) GS[(size_t)r * LD + i] = gs[r];
    }
}

// mu += K ze  and  Sigma -= K GS : the streaming pass (read + write 8 N^2 bytes each)
constexpr int EK_TX = 32, EK_TY = 8, EK_ROWS = 32;   // tile: 64 columns (2 per thread) x 32 rows
__global__ void __launch_bounds__(EK_TX * EK_TY)
k_ekf_rank3(double *__restrict__ sigma, double *__restrict__ mu, const double *__restrict__ mu_s, int N, int LD, EkfObs ob,
            const double *__restrict__ K, const double *__restrict__ GS)
{
    const int j = (blockIdx.x * EK_TX + threadIdx.x) * 2;
    const int i0 = blockIdx.y * EK_ROWS;
    if (blockIdx.x == 0 && threadIdx.x == 0) {           // mean update for this block's rows
        double G[18], ze[3];
        ekf_linearise(mu_s, ob, G, ze);
        for (int i = i0 + threadIdx.y; i < i0 + EK_ROWS && i < N; i += EK_TY)
            mu[i] += K[(size_t)i * 3] * ze[0] + K[(size_t)i * 3 + 1] * ze[1] + K[(size_t)i * 3 + 2] * ze[2];
    }
    if (j >= N) return;
    const bool two = (j + 1 < N);
    const double g0a = GS[j], g1a = GS[(size_t)LD + j], g2a = GS[2 * (size_t)LD + j];
    const double g0b = two ? GS[j + 1] : 0, g1b = two ? GS[(size_t)LD + j + 1] : 0, g2b = two ? GS[2 * (size_t)LD + j + 1] : 0;
#pragma unroll
    for (int r = 0; r < EK_ROWS / EK_TY; ++r) {
        const int i = i0 + threadIdx.y + r * EK_TY;
        if (i >= N) break;
        const double k0 = K[(size_t)i * 3], k1 = K[(size_t)i * 3 + 1], k2 = K[(size_t)i * 3 + 2];
        double2 *p = reinterpret_cast<double2 *>(sigma + (size_t)i * LD + j);      // LD even, j even: 16-byte aligned
        double2 v = *p;
        v.x -= k0 * g0a + k1 * g1a + k2 * g2a;
        if (two) v.y -= k0 * g0b + k1 * g1b + k2 * g2b;
        *p = v;
    }
}

// All known-landmark corrections of one frame in ONE cooperative launch (the per-observation launches of
// k_ekf_gain + k_ekf_rank3 spend as long in launch gaps as in the 16 N^2-byte streaming pass).  Every CTA owns a
// block of rows of Sigma.  Per observation:
//   phase A  every CTA forms GS = Gx Sigma (3 x N, needs six *old* rows of Sigma) in shared memory and the gain rows
//            K[i] of its own rows;                                   -- grid barrier (nobody may still read old rows)
//   phase B  mu[i] += K[i] ze and Sigma[i][:] -= K[i] GS for its own rows (the streaming pass);   -- grid barrier
// Same expressions, in the same order, as k_ekf_gain / k_ekf_rank3.  Sigma ping-pongs between two buffers (observation o
// reads buffer o % 2 and writes the other), so nobody overwrites rows that another CTA still reads and one grid barrier
// per observation is enough.
__global__ void __launch_bounds__(512)
k_ekf_frame(double *__restrict__ sigma_a, double *__restrict__ sigma_b, double *__restrict__ mu, const double *__restrict__ mu_s, int N, int LD,
            const EkfObs *__restrict__ obs, int n_obs)
{
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    extern __shared__ double s_gs[];                 // [3][LD]  GS of the current observation
    __shared__ double s_G[18], s_Si[9], s_ze[3], s_blk[36];
    __shared__ double s_K[64][3];                    // gain rows of this CTA's rows (rows_per_cta <= 64)
    const int rows_per_cta = (N + gridDim.x - 1) / gridDim.x;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
    for (int o = 0; o < n_obs; ++o) {
        const EkfObs ob = obs[o];
        const int L = 3 + 3 * ob.index;
        const int cols[6] = {0, 1, 2, L, L + 1, L + 2};
        const double *__restrict__ sigma = (o & 1) ? sigma_b : sigma_a;        // read
        double *__restrict__ sigma_out = (o & 1) ? sigma_a : sigma_b;          // written
        if (threadIdx.x < 36) s_blk[threadIdx.x] = sigma[(size_t)cols[threadIdx.x / 6] * LD + cols[threadIdx.x % 6]];      // the 6 x 6 block, loads in parallel
        __syncthreads();
        if (threadIdx.x == 0) {
            double G[18], ze[3];
            ekf_linearise(mu_s, ob, G, ze);
            double SGc[18];                          // (Sigma Gx^T)[cols[j]][r]
            for (int j = 0; j < 6; ++j) for (int r = 0; r < 3; ++r) {
                double a = 0;
                for (int k = 0; k < 6; ++k) a += s_blk[6 * j + k] * G[6 * r + k];
                SGc[3 * j + r] = a;
            }
            double S[9];
            for (int r = 0; r < 3; ++r) for (int c2 = 0; c2 < 3; ++c2) {
                double a = 0;
                for (int j = 0; j < 6; ++j) a += G[6 * r + j] * SGc[3 * j + c2];
                S[3 * r + c2] = a + ob.Rk[3 * r + c2];
            }
            inv3x3(S, s_Si);
            for (int i = 0; i < 18; ++i) s_G[i] = G[i];
            for (int i = 0; i < 3; ++i) s_ze[i] = ze[i];
        }
        __syncthreads();
        // phase A: GS for every column (each CTA its own copy), K for the CTA's rows
        for (int j = threadIdx.x; j < N; j += blockDim.x) {
            double gs[3] = {0, 0, 0};
            for (int k = 0; k < 6; ++k) {
                const double row = sigma[(size_t)cols[k] * LD + j];
                for (int r = 0; r < 3; ++r) gs[r] += s_G[6 * r + k] * row;
            }
            s_gs[j] = gs[0]; s_gs[LD + j] = gs[1]; s_gs[2 * LD + j] = gs[2];
        }
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
            double sg[3] = {0, 0, 0};
            for (int k = 0; k < 6; ++k) {
                const double col = sigma[(size_t)i * LD + cols[k]];
                for (int r = 0; r < 3; ++r) sg[r] += col * s_G[6 * r + k];
            }
            for (int c2 = 0; c2 < 3; ++c2) s_K[i - r0][c2] = sg[0] * s_Si[c2] + sg[1] * s_Si[3 + c2] + sg[2] * s_Si[6 + c2];
        }
        __syncthreads();
        // phase B: the streaming pass over the CTA's rows
        for (int i = r0 + (threadIdx.x >> 7); i < r1; i += (blockDim.x >> 7)) {      // 128 threads per row, two columns each per step
            const double k0 = s_K[i - r0][0], k1 = s_K[i - r0][1], k2 = s_K[i - r0][2];
            if ((threadIdx.x & 127) == 0) mu[i] += k0 * s_ze[0] + k1 * s_ze[1] + k2 * s_ze[2];
            const double *row = sigma + (size_t)i * LD;
            double *row_out = sigma_out + (size_t)i * LD;
            // up to 8 column pairs per thread (N <= 2048): all loads first, then the updates
            double2 v[8];
            const int j0 = 2 * (threadIdx.x & 127);
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int j = j0 + 256 * u; if (j < N) v[u] = *reinterpret_cast<const double2 *>(row + j); }     // LD even, j even: 16-byte aligned
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + 256 * u;
                if (j >= N) break;
                v[u].x -= k0 * s_gs[j] + k1 * s_gs[LD + j] + k2 * s_gs[2 * LD + j];
                if (j + 1 < N) v[u].y -= k0 * s_gs[j + 1] + k1 * s_gs[LD + j + 1] + k2 * s_gs[2 * LD + j + 1];
                *reinterpret_cast<double2 *>(row_out + j) = v[u];
            }
            for (int j = j0 + 2048; j < N; j += 256) {                                 // larger states: the plain loop
                double2 w = *reinterpret_cast<const double2 *>(row + j);
                w.x -= k0 * s_gs[j] + k1 * s_gs[LD + j] + k2 * s_gs[2 * LD + j];
                if (j + 1 < N) w.y -= k0 * s_gs[j + 1] + k1 * s_gs[LD + j + 1] + k2 * s_gs[2 * LD + j + 1];
                *reinterpret_cast<double2 *>(row_out + j) = w;
            }
        }
        grid.sync();
    }
}

// new landmark (:208-256): appends (x,y,theta) and the 3 new rows / columns; N is the old dimension
__global__ void k_ekf_augment(double *__restrict__ sigma, double *__restrict__ mu, const double *__restrict__ mu_s, int N, int LD, EkfObs ob)
{
    __shared__ double s_GG[9], s_Smm[9];
    if (threadIdx.x == 0) {
        const float sinth = (float)sin(mu_s[2]), costh = (float)cos(mu_s[2]);        // float, as the reference (:210-211)
        const double map_x = mu_s[0] + costh * ob.z[0] - sinth * ob.z[1];
        const double map_y = mu_s[1] + sinth * ob.z[0] + costh * ob.z[1];
        double map_th = mu_s[2] + ob.z[2];
        norm_angle(map_th);
        const double dx = map_x - mu_s[0], dy = map_y - mu_s[1];
        const double Gsk[9] = {-costh, -sinth, -sinth * dx + costh * dy, sinth, -costh, -dx * costh - dy * sinth, 0, 0, -1};
        const double Gmi[9] = {costh, sinth, 0, -sinth, costh, 0, 0, 0, 1};
        double A[9], Bm[9], C[9];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += Gsk[3 * i + k] * sigma[(size_t)k * LD + j]; A[3 * i + j] = a; }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += A[3 * i + k] * Gsk[3 * j + k]; Bm[3 * i + j] = a + ob.Rk[3 * i + j]; }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += Gmi[3 * i + k] * Bm[3 * j + k]; C[3 * i + j] = a; }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += C[3 * i + k] * Gmi[3 * j + k]; s_Smm[3 * i + j] = a; }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += (-Gmi[3 * i + k]) * Gsk[3 * k + j]; s_GG[3 * i + j] = a; }
        mu[N] = map_x; mu[N + 1] = map_y; mu[N + 2] = map_th;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const double s0 = sigma[j], s1 = sigma[(size_t)LD + j], s2 = sigma[2 * (size_t)LD + j];    // Sigma[0:3][j]
        for (int i = 0; i < 3; ++i) {
            const double v = s_GG[3 * i] * s0 + s_GG[3 * i + 1] * s1 + s_GG[3 * i + 2] * s2;
            sigma[(size_t)(N + i) * LD + j] = v;
            sigma[(size_t)j * LD + N + i] = v;
        }
    }
    if (threadIdx.x < 9) sigma[(size_t)(N + threadIdx.x / 3) * LD + N + threadIdx.x % 3] = s_Smm[threadIdx.x];
}

// prediction (:21-74): only rows / columns 0..2 of Sigma change
struct EkfMotion { double Hxi[9]; double Qk[9]; double dmu[3]; };
__global__ void k_ekf_predict(double *__restrict__ sigma, double *__restrict__ mu, int N, int LD, double wl, double wr, double dt,
                              double kl, double kr, double b, double Q_k, double *__restrict__ scratch /* 3*LD */)
{
    __shared__ double s_H[9], s_Q[9];
    if (threadIdx.x == 0) {
        const double dsl = kl * (dt * wl), dsr = kr * (dt * wr);
        const double dth = (dsr - dsl) / (2 * b), ds = 0.5 * (dsr + dsl);
        const double tmp = mu[2] + 0.5 * dth;
        const double c = cos(tmp), s = sin(tmp);
        mu[0] += ds * c; mu[1] += ds * s;
        double th = mu[2] + dth;
        norm_angle(th);
        mu[2] = th;
        const double H[9] = {1, 0, -ds * s, 0, 1, ds * c, 0, 0, 1};
        const double f = 0.5 * kl * dt;                                   // kl for both wheels (:62)
        const double wkh[6] = {c * f, c * f, s * f, s * f, (1 / b) * f, (-1 / b) * f};
        const double su0 = Q_k * fabs(wl), su1 = Q_k * fabs(wr);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s_Q[3 * i + j] = wkh[2 * i] * su0 * wkh[2 * j] + wkh[2 * i + 1] * su1 * wkh[2 * j + 1];
        for (int i = 0; i < 9; ++i) s_H[i] = H[i];
    }
    __syncthreads();
    // new rows 0..2 over columns >= 3: H * Sigma[0:3][j]; new columns 0..2 over rows >= 3: Sigma[i][0:3] * H^T
    for (int j = 3 + threadIdx.x; j < N; j += blockDim.x) {
        const double r0 = sigma[j], r1 = sigma[(size_t)LD + j], r2 = sigma[2 * (size_t)LD + j];
        const double c0 = sigma[(size_t)j * LD], c1 = sigma[(size_t)j * LD + 1], c2 = sigma[(size_t)j * LD + 2];
        for (int i = 0; i < 3; ++i) {
            scratch[(size_t)i * LD + j] = s_H[3 * i] * r0 + s_H[3 * i + 1] * r1 + s_H[3 * i + 2] * r2;
            scratch[(size_t)(3 + i) * LD + j] = c0 * s_H[3 * i] + c1 * s_H[3 * i + 1] + c2 * s_H[3 * i + 2];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double A[9], T[9];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += s_H[3 * i + k] * sigma[(size_t)k * LD + j]; A[3 * i + j] = a; }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += A[3 * i + k] * s_H[3 * j + k]; T[3 * i + j] = a + s_Q[3 * i + j]; }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sigma[(size_t)i * LD + j] = T[3 * i + j];
    }
    for (int j = 3 + threadIdx.x; j < N; j += blockDim.x)
        for (int i = 0; i < 3; ++i) {
            sigma[(size_t)i * LD + j] = scratch[(size_t)i * LD + j];
            sigma[(size_t)j * LD + i] = scratch[(size_t)(3 + i) * LD + j];
        }
}

}  // namespace b2a

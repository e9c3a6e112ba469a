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
#ifdef B2A_EKF_CLOCKS
#include <cstdio>
#endif
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

// ---------------------------------------------------------------------------------------------------------------
// Panel form: all M known-landmark corrections of a frame as ONE rank-3M update of Sigma (reference loop :92-207, dense
// product :204).  Every Jacobian G_m (3 x N, six non-zero columns) and innovation ze_m comes from the frame-start snapshot
// (:88), so with  U0 = Sigma0 [G_0^T .. G_{M-1}^T]  (N x 3M),  V0 = [G_0; ..; G_{M-1}] Sigma0  (3M x N)  and the block LU
// factorisation  G Sigma0 G^T + blockdiag(R_m) = L Uh  (3 x 3 blocks, L unit lower, Uh upper with the innovation
// covariances S_m on its diagonal; elimination step m is exactly "apply correction m") the sequential gains are the block
// columns of  K = U0 Uh^-1,  the rows G_m Sigma_m are the block rows of  V = L^-1 V0,  and
//     Sigma_M = Sigma0 - K V,      mu_M = mu_0 + K [ze_0; ..; ze_{M-1}]          (the reference's mean: sequential gains, fixed innovations).
// Sigma is read and written once per frame (16 N^2 bytes) instead of once per observation; the N x 3M x N contraction runs
// on the FP64 tensor-core path (mma.sync m8n8k4 f64).  Same result as the sequential form up to rounding (tests: 1e-9).
//   k_ekf_panel_gather   G_m, ze_m, U0, V0                       (N / 128 CTAs)
//   k_ekf_panel_factor   C = G U0, block LU in shared memory     (1 CTA)
//   k_ekf_panel_solve    K = U0 Uh^-1 (per row), mu += K ze, V = L^-1 V0 (per column)
//   k_ekf_panel_gemm     Sigma -= K V                            (64 x 64 tiles, DMMA)
// ---------------------------------------------------------------------------------------------------------------
constexpr int EP_MAX_OBS = 32;                    // corrections per panel
constexpr int EP_K = 3 * EP_MAX_OBS;              // 96: padded inner dimension (multiple of 4)
struct EkfPanel {
    const EkfObs *obs; int M;                     // device array of the panel's corrections
    double *ze;                                   // [EP_K] innovations (zero padded)
    double *U0, *Kall;                            // [N][EP_K]
    double *V0;                                   // [N][EP_K]: the TRANSPOSE of V0 (row j = column j of the 3M x N matrix), one contiguous row per index
    double *Vall;                                 // [EP_K][LD]: V as the contraction reads it
    double *fac;                                  // [EP_K][EP_K]: G Sigma0 G^T on entry of k_ekf_panel_factor, the LU factors after it:
                                                  // upper block triangle = Uh with S_m^-1 in the diagonal blocks, strict lower = L
    double *facT;                                 // its transpose (the row solve walks columns of Uh)
};

// grid (ceil(N / 128), M + 1).  y < M: column block y of U0 and row block y of V0 (one thread per state index);
// y == M: the zero padding, the innovations and the 3 x 3 blocks of G Sigma0 G^T (one thread per block pair)
__global__ void __launch_bounds__(128)
k_ekf_panel_gather(const double *__restrict__ sigma, const double *__restrict__ mu_s, int N, int LD, EkfPanel p)
{
    const int M = p.M;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ((int)blockIdx.y < M) {
        __shared__ double s_G[18];
        __shared__ int s_L;
        const int m = blockIdx.y;
        if (threadIdx.x == 0) {
            double G[18], ze[3];
            const EkfObs ob = p.obs[m];
            ekf_linearise(mu_s, ob, G, ze);
            for (int k = 0; k < 18; ++k) s_G[k] = G[k];
            s_L = 3 + 3 * ob.index;
        }
        __syncthreads();
        if (i >= N) return;
        const int L = s_L;
        const double *row = sigma + (size_t)i * LD;
        const double c0 = row[0], c1 = row[1], c2 = row[2], c3 = row[L], c4 = row[L + 1], c5 = row[L + 2];          // Sigma[i][cols]
        const double r0 = sigma[i], r1 = sigma[(size_t)LD + i], r2 = sigma[2 * (size_t)LD + i];                       // Sigma[cols][i]
        const double r3 = sigma[(size_t)L * LD + i], r4 = sigma[(size_t)(L + 1) * LD + i], r5 = sigma[(size_t)(L + 2) * LD + i];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double *g = s_G + 6 * r;
            p.U0[(size_t)i * EP_K + 3 * m + r] = c0 * g[0] + c1 * g[1] + c2 * g[2] + c3 * g[3] + c4 * g[4] + c5 * g[5];
            p.V0[(size_t)i * EP_K + 3 * m + r] = g[0] * r0 + g[1] * r1 + g[2] * r2 + g[3] * r3 + g[4] * r4 + g[5] * r5;
        }
        return;
    }
    if (i < N) for (int c = 3 * M; c < EP_K; ++c) { p.U0[(size_t)i * EP_K + c] = 0.0; p.V0[(size_t)i * EP_K + c] = 0.0; }
    if (i < EP_K * EP_K && (i / EP_K >= 3 * M || i % EP_K >= 3 * M)) p.fac[i] = 0.0;
    if (i < M * M) {
        const int a = i / M, b = i - a * M;
        double Ga[18], Gb[18], ze[3], zb[3];
        const EkfObs oa = p.obs[a], ob = p.obs[b];
        ekf_linearise(mu_s, oa, Ga, ze);
        ekf_linearise(mu_s, ob, Gb, zb);
        if (b == 0) for (int r = 0; r < 3; ++r) p.ze[3 * a + r] = ze[r];
        const int La = 3 + 3 * oa.index, Lb = 3 + 3 * ob.index;
        double T[18];                              // Sigma[cols_a][cols_b] G_b^T   (6 x 3)
        for (int k = 0; k < 6; ++k) {
            const double *row = sigma + (size_t)(k < 3 ? k : La + k - 3) * LD;
            const double v0 = row[0], v1 = row[1], v2 = row[2], v3 = row[Lb], v4 = row[Lb + 1], v5 = row[Lb + 2];
            for (int c = 0; c < 3; ++c) T[3 * k + c] = v0 * Gb[6 * c] + v1 * Gb[6 * c + 1] + v2 * Gb[6 * c + 2] + v3 * Gb[6 * c + 3] + v4 * Gb[6 * c + 4] + v5 * Gb[6 * c + 5];
        }
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
            double v = 0;
            for (int k = 0; k < 6; ++k) v += Ga[6 * r + k] * T[3 * k + c];
            p.fac[(size_t)(3 * a + r) * EP_K + 3 * b + c] = v;
        }
    }
    if (i >= 3 * M && i < EP_K) p.ze[i] = 0.0;
}

// one CTA: block LU without pivoting of C + blockdiag(R_m) (symmetric positive definite up to rounding); elimination step m is
// correction m of the reference's loop.  Thread (A, B) keeps the 3 x 3 block D_AB in registers for the whole factorisation:
//   pivot      thread (m, m): S_m = D_mm + R_m (:146), inverted in closed form; threads (m, B > m) publish their finished Uh_mB
//   multiplier threads (A > m, m): L_Am = D_Am S_m^-1, published
//   update     threads (A > m, B > m): D_AB -= L_Am Uh_mB   (27 FMAs on registers)
// Two barriers per step; the published blocks are double buffered by the step's parity.  (Measured alternatives, all slower on the
// one SM this runs on: rows of a shared-memory matrix per warp, 51-85 us; 2 x 2 blocks per thread with one barrier per step, 47-58 us;
// this form 42 us for 30 corrections.)
__global__ void __launch_bounds__(EP_MAX_OBS * EP_MAX_OBS)
k_ekf_panel_factor(EkfPanel p)
{
    __shared__ double s_L[2][EP_MAX_OBS][9], s_U[2][EP_MAX_OBS][9], s_Si[9];
    const int M = p.M, t = threadIdx.x;
    const int A = t / M, B = t - A * M;
    const bool live = t < M * M;
    double D[9], Rk[9];
    if (live) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) D[3 * r + c] = p.fac[(size_t)(3 * A + r) * EP_K + 3 * B + c];
        if (A == B) {
#pragma unroll
            for (int k = 0; k < 9; ++k) Rk[k] = p.obs[A].Rk[k];
        }
    }
    for (int m = 0; m < M; ++m) {
        const int par = m & 1;
        if (live && A == m) {
            if (B == m) {
                double S[9], Si[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) S[k] = D[k] + Rk[k];
                inv3x3(S, Si);
#pragma unroll
                for (int k = 0; k < 9; ++k) { D[k] = Si[k]; s_Si[k] = Si[k]; }                     // the diagonal block keeps S_m^-1
            } else if (B > m) {
#pragma unroll
                for (int k = 0; k < 9; ++k) s_U[par][B][k] = D[k];
            }
        }
        __syncthreads();
        if (live && B == m && A > m) {
            double L[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) L[3 * r + c] = D[3 * r] * s_Si[c] + D[3 * r + 1] * s_Si[3 + c] + D[3 * r + 2] * s_Si[6 + c];
#pragma unroll
            for (int k = 0; k < 9; ++k) { D[k] = L[k]; s_L[par][A][k] = L[k]; }
        }
        __syncthreads();
        if (live && A > m && B > m) {
            double L[9], U[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) { L[k] = s_L[par][A][k]; U[k] = s_U[par][B][k]; }
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) D[3 * r + c] -= L[3 * r] * U[c] + L[3 * r + 1] * U[3 + c] + L[3 * r + 2] * U[6 + c];
        }
    }
    if (live) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                p.fac[(size_t)(3 * A + r) * EP_K + 3 * B + c] = D[3 * r + c];
                p.facT[(size_t)(3 * B + c) * EP_K + 3 * A + r] = D[3 * r + c];
            }
    }
}

// column-oriented triangular solves (no reductions).  A warp solves EP_SV state indices at once; lane l keeps block l (entries
// 3l .. 3l + 2) of each of their vectors.
//   rows     x = U0[i][:]:  K_m = x_m S_m^-1, then x_l -= K_m Uh[m][l] for the later blocks;  mu += K ze (:203)
//   columns  x = V0[:][j]:  V_m = x_m,        then x_l -= L[l][m] V_m
// The coefficients of a step are three rows of fac (rows) or of its transpose (columns).  A CTA stages that matrix in shared memory as
// three planes (entry 3l + c of a row at plane c, word l), so a lane's three coefficients per row are conflict-free reads and the
// four vectors of a warp share them (read through L1 instead, the 24-byte lane stride made every load seven cache lines wide).
constexpr int EP_SW = 8, EP_SV = 4;               // warps per CTA, state indices per warp
constexpr size_t EP_SSMEM = (size_t)(3 * EP_K * EP_MAX_OBS + 9 * EP_MAX_OBS + EP_K) * sizeof(double);
__global__ void __launch_bounds__(32 * EP_SW)
k_ekf_panel_solve(double *__restrict__ mu, int N, int LD, EkfPanel p, int row_ctas)
{
    extern __shared__ double s_F[];               // [3][EP_K][32] coefficient planes, [32][9] S^-1 blocks, [EP_K] ze
    double *s_Si = s_F + 3 * EP_K * EP_MAX_OBS, *s_ze = s_Si + 9 * EP_MAX_OBS;
    const int M = p.M, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool rows = (int)blockIdx.x < row_ctas;
    {
        const double *__restrict__ fm = rows ? p.fac : p.facT;
        for (int t = threadIdx.x; t < 3 * M * EP_K; t += blockDim.x) {      // 8-byte asynchronous copies: all in flight at once
            const int r = t / EP_K, c = t - r * EP_K;          // coalesced read of row r
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((uint32_t)__cvta_generic_to_shared(s_F + ((c % 3) * EP_K + r) * EP_MAX_OBS + c / 3)),
                         "l"(fm + (size_t)r * EP_K + c) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        for (int t = threadIdx.x; t < 9 * M; t += blockDim.x) { const int m = t / 9, k = t - 9 * m; s_Si[t] = p.fac[(size_t)(3 * m + k / 3) * EP_K + 3 * m + k % 3]; }
        for (int t = threadIdx.x; t < EP_K; t += blockDim.x) s_ze[t] = p.ze[t];
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int i0 = (((int)blockIdx.x - (rows ? 0 : row_ctas)) * EP_SW + warp) * EP_SV;
    if (i0 >= N) return;
    const double *__restrict__ src = rows ? p.U0 : p.V0;
    double x[EP_SV][3], dmu[EP_SV];
#pragma unroll
    for (int v = 0; v < EP_SV; ++v) {
        const int i = min(i0 + v, N - 1);
        const double *in = src + (size_t)i * EP_K + 3 * lane;
        x[v][0] = in[0]; x[v][1] = in[1]; x[v][2] = in[2];
        dmu[v] = 0.0;
    }
    for (int m = 0; m < M; ++m) {
        double f[3][3];                            // this lane's coefficients of the step: rows 3m + r, entries 3 lane + c
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) f[r][c] = s_F[(c * EP_K + 3 * m + r) * EP_MAX_OBS + lane];
        double Si[9];
        if (rows) {
#pragma unroll
            for (int k = 0; k < 9; ++k) Si[k] = s_Si[9 * m + k];
        }
#pragma unroll
        for (int v = 0; v < EP_SV; ++v) {
            double k0 = __shfl_sync(0xFFFFFFFFu, x[v][0], m), k1 = __shfl_sync(0xFFFFFFFFu, x[v][1], m), k2 = __shfl_sync(0xFFFFFFFFu, x[v][2], m);
            if (rows) {                            // times S_m^-1 from the right
                const double a0 = k0, a1 = k1, a2 = k2;
                k0 = a0 * Si[0] + a1 * Si[3] + a2 * Si[6];
                k1 = a0 * Si[1] + a1 * Si[4] + a2 * Si[7];
                k2 = a0 * Si[2] + a1 * Si[5] + a2 * Si[8];
                dmu[v] += k0 * s_ze[3 * m] + k1 * s_ze[3 * m + 1] + k2 * s_ze[3 * m + 2];
            }
            if (lane > m) {
                x[v][0] -= k0 * f[0][0] + k1 * f[1][0] + k2 * f[2][0];
                x[v][1] -= k0 * f[0][1] + k1 * f[1][1] + k2 * f[2][1];
                x[v][2] -= k0 * f[0][2] + k1 * f[1][2] + k2 * f[2][2];
            } else if (lane == m) { x[v][0] = k0; x[v][1] = k1; x[v][2] = k2; }
        }
    }
#pragma unroll
    for (int v = 0; v < EP_SV; ++v) {
        const int i = i0 + v;
        if (i >= N) break;
        const double o0 = lane < M ? x[v][0] : 0.0, o1 = lane < M ? x[v][1] : 0.0, o2 = lane < M ? x[v][2] : 0.0;
        if (rows) {
            double *out = p.Kall + (size_t)i * EP_K + 3 * lane;
            out[0] = o0; out[1] = o1; out[2] = o2;
            if (lane == 0) mu[i] += dmu[v];
        } else {
            double *out = p.Vall + (size_t)(3 * lane) * LD + i;
            out[0] = o0; out[LD] = o1; out[2 * (size_t)LD] = o2;
        }
    }
}

// Sigma -= K V on the FP64 tensor-core path: a CTA of 16 warps owns a 128 x 128 tile of Sigma (144 tiles at N = 1503: one per SM),
// a warp a 32 x 32 part (4 x 4 mma tiles); the tile's K rows and V columns come to shared memory by 16-byte asynchronous copies
// and stay for the whole contraction
__device__ __forceinline__ void dmma_8x8x4(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void *dst_shared, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dst_shared)), "l"(src) : "memory");
}
constexpr int EG_T = 128, EG_AP = EP_K + 4, EG_BP = EG_T + 4;     // pitches = 4 (mod 16) doubles: the fragment loads of a half-warp cover 16 distinct 8-byte banks
constexpr size_t EG_SMEM = ((size_t)EG_T * EG_AP + (size_t)EP_K * EG_BP) * sizeof(double);
__global__ void __launch_bounds__(512)
k_ekf_panel_gemm(double *__restrict__ sigma, int N, int LD, EkfPanel p, int kdim)
{
    extern __shared__ __align__(16) double s_ab[];
    double *sA = s_ab, *sB = s_ab + EG_T * EG_AP;
    const int i0 = blockIdx.y * EG_T, j0 = blockIdx.x * EG_T, tid = threadIdx.x;
    const int kd2 = kdim / 2;
    for (int t = tid; t < EG_T * kd2; t += blockDim.x) {                    // K rows of the tile (rows beyond N: zeros)
        const int r = t / kd2, c = 2 * (t - r * kd2);
        if (i0 + r < N) cp_async16(sA + r * EG_AP + c, p.Kall + (size_t)(i0 + r) * EP_K + c);
        else { sA[r * EG_AP + c] = 0.0; sA[r * EG_AP + c + 1] = 0.0; }
    }
    for (int t = tid; t < kdim * (EG_T / 2); t += blockDim.x) {             // V columns of the tile (columns beyond the array: zeros)
        const int r = t / (EG_T / 2), c = 2 * (t - r * (EG_T / 2));
        if (j0 + c + 1 < LD) cp_async16(sB + r * EG_BP + c, p.Vall + (size_t)r * LD + j0 + c);
        else { sB[r * EG_BP + c] = 0.0; sB[r * EG_BP + c + 1] = 0.0; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int warp = tid >> 5, lane = tid & 31, gr = lane >> 2, gc = lane & 3;
    const int wi = (warp >> 2) * 32, wj = (warp & 3) * 32;
    {   // the tile of Sigma itself is wanted at the end: start it towards L2 now
        const int r = tid >> 2, q = tid & 3;                                // 128 rows x 4 quarter rows of 256 bytes
        if (i0 + r < N && j0 + 32 * q < LD) {
            const char *line = reinterpret_cast<const char *>(sigma + (size_t)(i0 + r) * LD + j0) + 256 * q;
            asm volatile("prefetch.global.L2 [%0];" :: "l"(line));
            if (j0 + 32 * q + 16 < LD) asm volatile("prefetch.global.L2 [%0];" :: "l"(line + 128));
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    for (int k0 = 0; k0 < kdim; k0 += 4) {
        double fa[4], fb[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) fa[a] = sA[(wi + 8 * a + gr) * EG_AP + k0 + gc];
#pragma unroll
        for (int b = 0; b < 4; ++b) fb[b] = sB[(k0 + gc) * EG_BP + wj + 8 * b + gr];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) dmma_8x8x4(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + wi + 8 * a + gr;
        double2 v[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {                                       // four loads in flight, then the four updates
            const int j = j0 + wj + 8 * b + 2 * gc;
            v[b] = (i < N && j < N) ? *reinterpret_cast<const double2 *>(sigma + (size_t)i * LD + j) : make_double2(0.0, 0.0);      // LD even, j even: 16-byte aligned
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = j0 + wj + 8 * b + 2 * gc;
            if (i >= N || j >= N) continue;
            v[b].x -= acc[a][b][0];
            if (j + 1 < N) v[b].y -= acc[a][b][1];
            *reinterpret_cast<double2 *>(sigma + (size_t)i * LD + j) = v[b];
        }
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

// detect_kernels.cuh -- hand-written sm_100a kernels of the detect path (one batch of frames
// per launch).  Replaces the inside of cv::aruco::detectMarkers (reference
// src/aruco_slam.cpp:313); stage letters follow SURVEY.md Appendix A.
//
//   k_bgr2gray        A1   bgr8 -> gray (15-bit fixed point)
//   k_threshold       A2   gray -> nScales bit-packed adaptive-threshold masks, one pass, shared-memory
//                          staged row-prefix tile, ballot-packed output
//   k_anchors         A3a  word-parallel enumeration of the border graph's anchor states and start candidates
//   k_segments        A3a  every anchor walks to the next anchor (short, independent walks)
//   k_skip, k_cycles  A3a  hop over each border's (super) anchors: leader (first point), border length;
//                          borders without anchors from their start candidates
//   k_sort_scan       A3a  per (frame,scale): order kept borders like cv2.findContours, offsets
//   k_assign, k_emit  A3a  position of every segment inside its border; emit the border points
//   k_approx          A3b  warp-cooperative approxPolyDP + quad gates
//   k_group           A4-5 per frame: grouping / selection / hierarchy (frame_logic.h)
//   k_identify        A7   per candidate: homography, NN warp, Otsu, bits, dictionary match
//   k_finalize        A6   per frame: depth-ordered acceptance, output
//   k_subpix          A8   optional cornerSubPix
//   k_refine_contour  A8   optional CORNER_REFINE_CONTOUR (refine_core.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "core.h"
#include "frame_logic.h"
#include "refine_core.h"
#include "pyr_core.h"

namespace b2a {

constexpr int MAX_SCALES = 8;
constexpr int R_MAX = 15;           // largest supported threshold window radius (window 31)
constexpr int SORT_CAP = 4096;      // largest survivors-per-(frame,scale) the sort kernel handles

struct DetGeom {
    int W, H, B, nScales;
    int radius[MAX_SCALES];
    int Cfloor;                     // floor(adaptiveThreshConstant)
    int WW, PWW;                    // mask words per row, padded pitch (WW + 2)
    long long mask_plane;           // words per (frame,scale) mask plane = PWW * (H + 2)
    int KS;                         // key stride (W + 1)
    int minPerim, maxPerim, maxWH;
    double approxRate, minCornerDistRate;
    int surv_cap, pts_cap;
    int count_all;                  // debug tap only: count every border (also the ones the perimeter gate drops) and the one-pixel regions
};

// ---------------------------------------------------------------------------------------------
// A1
// ---------------------------------------------------------------------------------------------
__global__ void k_bgr2gray(const uint8_t *__restrict__ bgr, size_t in_pitch, size_t in_frame,
                           uint8_t *__restrict__ gray, size_t out_pitch, size_t out_frame, int W, int H, int B)
{
    const long long total = (long long)B * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const long long t = i / W;
        const int y = (int)(t % H), b = (int)(t / H);
        const uint8_t *p = bgr + (size_t)b * in_frame + (size_t)y * in_pitch + (size_t)x * 3;
        const int v = (3735 * (int)p[0] + 19235 * (int)p[1] + 9798 * (int)p[2] + 16384) >> 15;
        gray[(size_t)b * out_frame + (size_t)y * out_pitch + x] = (uint8_t)v;
    }
}

// ---------------------------------------------------------------------------------------------
// A2: adaptive threshold, all scales in one pass.
//   tile = TW x TH output pixels; shared memory holds the replicate-clamped input tile (halo
//   R_MAX + 1) and its per-row exclusive prefix sums E (u16).  Thread (column c, half h) slides
//   the three vertical window sums down its half of the rows; a box sum is
//   sum_rows (E[row][x + r + 1] - E[row][x - r]).  mask bit:  g - mean <= -C  with
//   mean = (2 S + k^2) div (2 k^2)   <=>   2 S >= (2 (g + C) - 1) k^2      (exact integers).
//   The 32 lanes of a warp hold 32 neighbouring columns: one ballot = one packed mask word.
// ---------------------------------------------------------------------------------------------
constexpr int TH_TW = 128, TH_TH = 64, TH_HALO = 16;
constexpr int TH_COLS = TH_TW + 2 * TH_HALO;          // 160
constexpr int TH_ROWS = TH_TH + 2 * R_MAX;            // 94
constexpr int TH_EPITCH = TH_COLS + 2;                // 162 (161 used)
constexpr int TH_THREADS = 256;

__global__ void __launch_bounds__(TH_THREADS)
k_threshold(const uint8_t *__restrict__ gray, size_t pitch, size_t frame_stride, uint32_t *__restrict__ masks, DetGeom g)
{
    __shared__ __align__(16) uint8_t s_in[TH_ROWS][TH_COLS];
    __shared__ uint16_t s_E[TH_ROWS + 1][TH_EPITCH];   // +1: the last slide reads one row past the tile (unused)

    const int b = blockIdx.z;
    const int x0 = blockIdx.x * TH_TW, y0 = blockIdx.y * TH_TH;
    const uint8_t *src = gray + (size_t)b * frame_stride;
    const int tid = threadIdx.x;

    // ---- load the clamped tile ----
    const bool fast = (x0 - TH_HALO >= 0) && (x0 + TH_TW + TH_HALO <= g.W) && ((pitch & 15) == 0) &&
                      ((((size_t)src) & 15) == 0);
    if (fast) {
        for (int t = tid; t < TH_ROWS * (TH_COLS / 16); t += TH_THREADS) {
            const int r = t / (TH_COLS / 16), c16 = t - r * (TH_COLS / 16);
            int gy = y0 - R_MAX + r; gy = gy < 0 ? 0 : (gy >= g.H ? g.H - 1 : gy);
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)gy * pitch + (x0 - TH_HALO)) + c16);
            *reinterpret_cast<uint4 *>(&s_in[r][c16 * 16]) = v;
        }
    } else {
        for (int t = tid; t < TH_ROWS * TH_COLS; t += TH_THREADS) {
            const int r = t / TH_COLS, c = t - r * TH_COLS;
            int gy = y0 - R_MAX + r; gy = gy < 0 ? 0 : (gy >= g.H ? g.H - 1 : gy);
            int gx = x0 - TH_HALO + c; gx = gx < 0 ? 0 : (gx >= g.W ? g.W - 1 : gx);
            s_in[r][c] = __ldg(src + (size_t)gy * pitch + gx);
        }
    }
    __syncthreads();

    // ---- per-row exclusive prefix sums (one warp per row, 5 bytes per lane) ----
    const int lane = tid & 31, warp = tid >> 5;
    for (int r = warp; r < TH_ROWS; r += TH_THREADS / 32) {
        unsigned v[5], s = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) { v[i] = s_in[r][lane * 5 + i]; s += v[i]; }
        unsigned incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += o; }
        unsigned run = incl - s;
#pragma unroll
        for (int i = 0; i < 5; ++i) { s_E[r][lane * 5 + i] = (uint16_t)run; run += v[i]; }
        if (lane == 31) s_E[r][TH_COLS] = (uint16_t)run;
    }
    __syncthreads();

    // ---- vertical sliding sums, compare, ballot ----
    const int c = tid & (TH_TW - 1), half = tid >> 7;             // 128 columns x 2 halves
    const int xo = TH_HALO + c;                                   // tile column of the output pixel
    const int j0 = half * (TH_TH / 2);                            // first output row of this thread
    const int gx = x0 + c;
    const bool col_ok = gx < g.W;
    unsigned V[MAX_SCALES];
    int rhs_mul[MAX_SCALES];
#pragma unroll
    for (int s = 0; s < MAX_SCALES; ++s) {
        if (s < g.nScales) {
            const int r = g.radius[s];
            unsigned acc = 0;
            for (int dy = -r; dy <= r; ++dy) acc += (unsigned)s_E[R_MAX + j0 + dy][xo + r + 1] - (unsigned)s_E[R_MAX + j0 + dy][xo - r];
            V[s] = acc;
            rhs_mul[s] = (2 * r + 1) * (2 * r + 1);
        }
    }
    const int wi = (x0 >> 5) + (c >> 5) + 1;                      // padded word index of this warp's 32 columns
    for (int j = j0; j < j0 + TH_TH / 2; ++j) {
        const int gy = y0 + j;
        const int gv = s_in[R_MAX + j][xo];
        const bool ok = col_ok && gy < g.H;
#pragma unroll
        for (int s = 0; s < MAX_SCALES; ++s) {
            if (s < g.nScales) {
                const int r = g.radius[s];
                const long long rhs = (long long)(2 * (gv + g.Cfloor) - 1) * rhs_mul[s];
                const bool bit = ok && ((long long)(2u * V[s]) >= rhs);
                const unsigned word = __ballot_sync(0xFFFFFFFFu, bit);
                if (lane == 0 && gy < g.H && wi <= g.WW)
                    masks[((size_t)b * g.nScales + s) * g.mask_plane + (size_t)(gy + 1) * g.PWW + wi] = word;
                // slide to row j + 1
                V[s] += ((unsigned)s_E[R_MAX + j + 1 + r][xo + r + 1] - (unsigned)s_E[R_MAX + j + 1 + r][xo - r])
                      - ((unsigned)s_E[R_MAX + j - r][xo + r + 1] - (unsigned)s_E[R_MAX + j - r][xo - r]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// A2, specialised on a compile-time list of three window radii (the default 3 / 13 / 23 windows):
// same tile scheme, 128 x 128 outputs per CTA (64 rows per thread), every shared-memory address is
// "row pointer + immediate", the test runs in 32-bit integers and the row loop is unrolled by 8.
//   2 S >= (2 (g + C) - 1) k^2   with S <= 23^2 * 255 and |C| <= 2^11 (checked by the host) fits int32.
// ---------------------------------------------------------------------------------------------
constexpr int T3_TW = 128, T3_TH = 128, T3_HALO = 16;
constexpr int T3_COLS = T3_TW + 2 * T3_HALO;          // 160
constexpr int T3_ROWS = T3_TH + 2 * R_MAX;            // 158
constexpr int T3_EPITCH = T3_COLS + 2;                // 162
constexpr int T3_THREADS = 256;
constexpr size_t T3_SMEM = (size_t)T3_ROWS * T3_COLS + (size_t)(T3_ROWS + 1) * T3_EPITCH * sizeof(uint16_t);

template <int R>
__device__ __forceinline__ int t3_hsum(const uint16_t *row)       // horizontal window sum at the row's centre column pointer
{
    return (int)row[R + 1] - (int)row[-R];
}

template <int R0, int R1, int R2>
__global__ void __launch_bounds__(T3_THREADS)
k_threshold3(const uint8_t *__restrict__ gray, size_t pitch, size_t frame_stride, uint32_t *__restrict__ masks, DetGeom g)
{
    extern __shared__ __align__(16) uint8_t t3_smem[];
    uint8_t (*s_in)[T3_COLS] = reinterpret_cast<uint8_t (*)[T3_COLS]>(t3_smem);
    uint16_t (*s_E)[T3_EPITCH] = reinterpret_cast<uint16_t (*)[T3_EPITCH]>(t3_smem + (size_t)T3_ROWS * T3_COLS);

    const int b = blockIdx.z;
    const int x0 = blockIdx.x * T3_TW, y0 = blockIdx.y * T3_TH;
    const uint8_t *src = gray + (size_t)b * frame_stride;
    const int tid = threadIdx.x;

    // ---- load the clamped tile ----
    const bool fast = (x0 - T3_HALO >= 0) && (x0 + T3_TW + T3_HALO <= g.W) && ((pitch & 15) == 0) && ((((size_t)src) & 15) == 0);
    if (fast) {
        for (int t = tid; t < T3_ROWS * (T3_COLS / 16); t += T3_THREADS) {
            const int r = t / (T3_COLS / 16), c16 = t - r * (T3_COLS / 16);
            int gy = y0 - R_MAX + r; gy = gy < 0 ? 0 : (gy >= g.H ? g.H - 1 : gy);
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)gy * pitch + (x0 - T3_HALO)) + c16);
            *reinterpret_cast<uint4 *>(&s_in[r][c16 * 16]) = v;
        }
    } else {
        for (int t = tid; t < T3_ROWS * T3_COLS; t += T3_THREADS) {
            const int r = t / T3_COLS, c = t - r * T3_COLS;
            int gy = y0 - R_MAX + r; gy = gy < 0 ? 0 : (gy >= g.H ? g.H - 1 : gy);
            int gx = x0 - T3_HALO + c; gx = gx < 0 ? 0 : (gx >= g.W ? g.W - 1 : gx);
            s_in[r][c] = __ldg(src + (size_t)gy * pitch + gx);
        }
    }
    __syncthreads();

    // ---- per-row exclusive prefix sums (one warp per row, 5 bytes per lane) ----
    const int lane = tid & 31, warp = tid >> 5;
    for (int r = warp; r < T3_ROWS; r += T3_THREADS / 32) {
        unsigned v[5], sum = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) { v[i] = s_in[r][lane * 5 + i]; sum += v[i]; }
        unsigned incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += o; }
        unsigned run = incl - sum;
#pragma unroll
        for (int i = 0; i < 5; ++i) { s_E[r][lane * 5 + i] = (uint16_t)run; run += v[i]; }
        if (lane == 31) s_E[r][T3_COLS] = (uint16_t)run;
    }
    __syncthreads();

    // ---- vertical sliding sums, compare, ballot ----
    const int c = tid & (T3_TW - 1), half = tid >> 7;             // 128 columns x 2 halves of 64 rows
    const int xo = T3_HALO + c;
    const int j0 = half * (T3_TH / 2);
    const bool col_ok = x0 + c < g.W;
    const uint16_t *e = &s_E[R_MAX + j0][xo];                     // row pointer of output row j0, centred on this column
    const uint8_t *pin = &s_in[R_MAX + j0][xo];
    int V0 = 0, V1 = 0, V2 = 0;
#pragma unroll
    for (int dy = -R0; dy <= R0; ++dy) V0 += t3_hsum<R0>(e + dy * T3_EPITCH);
#pragma unroll
    for (int dy = -R1; dy <= R1; ++dy) V1 += t3_hsum<R1>(e + dy * T3_EPITCH);
#pragma unroll
    for (int dy = -R2; dy <= R2; ++dy) V2 += t3_hsum<R2>(e + dy * T3_EPITCH);
    constexpr int K0 = (2 * R0 + 1) * (2 * R0 + 1), K1 = (2 * R1 + 1) * (2 * R1 + 1), K2 = (2 * R2 + 1) * (2 * R2 + 1);
    const int c2 = 2 * g.Cfloor - 1;
    const int wi = (x0 >> 5) + (c >> 5) + 1;                      // padded word index of this warp's 32 columns
    // the 24 mask words of 8 rows x 3 scales are parked on lanes 0..23 (lane = scale * 8 + row) and leave in one store
    const int my_u = lane & 7, my_s = lane >> 3;
    uint32_t *outp = masks + ((size_t)b * 3 + (my_s < 3 ? my_s : 0)) * g.mask_plane + (size_t)(y0 + j0 + my_u + 1) * g.PWW + wi;
    const bool lane_stores = (lane < 24) && (wi <= g.WW);
    for (int jb = 0; jb < T3_TH / 2; jb += 8) {
        unsigned mine = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int t = 2 * (int)pin[u * T3_COLS] + c2;         // 2 (g + C) - 1
            const bool ok = col_ok && (y0 + j0 + jb + u < g.H);
            const unsigned w0 = __ballot_sync(0xFFFFFFFFu, ok && (2 * V0 - t * K0 >= 0));
            const unsigned w1 = __ballot_sync(0xFFFFFFFFu, ok && (2 * V1 - t * K1 >= 0));
            const unsigned w2 = __ballot_sync(0xFFFFFFFFu, ok && (2 * V2 - t * K2 >= 0));
            if (lane == u) mine = w0;
            if (lane == 8 + u) mine = w1;
            if (lane == 16 + u) mine = w2;
            // slide to the next row
            V0 += t3_hsum<R0>(e + (u + 1 + R0) * T3_EPITCH) - t3_hsum<R0>(e + (u - R0) * T3_EPITCH);
            V1 += t3_hsum<R1>(e + (u + 1 + R1) * T3_EPITCH) - t3_hsum<R1>(e + (u - R1) * T3_EPITCH);
            V2 += t3_hsum<R2>(e + (u + 1 + R2) * T3_EPITCH) - t3_hsum<R2>(e + (u - R2) * T3_EPITCH);
        }
        if (lane_stores && y0 + j0 + jb + my_u < g.H) *outp = mine;
        e += 8 * T3_EPITCH; pin += 8 * T3_COLS;
        outp += (size_t)8 * g.PWW;
    }
}

// ---------------------------------------------------------------------------------------------
// A2, marching form for three windows (radii <= 11): the instruction count per pixel, not the memory system,
// bounds this stage, so the kernel is organised around the fewest instructions per pixel (~20 instead of ~50):
//   work item  = a strip of TM_WT output columns x Hs rows of one frame; a CTA marches down it TM_RC rows at a time.
//   V phase    thread t owns 4 neighbouring columns (one 32-bit load per row, straight from global memory) and keeps
//              the running column prefix C (two registers of packed u16 pairs: even / odd bytes) of the last 24 rows
//              in a register ring: every vertical window sum is ONE packed subtraction  V_r[m] = C[m+r] - C[m-r-1]
//              (no borrow between the lanes: C is monotone and < 2^16 for Hs + 22 <= 257 rows).  The three V planes
//              of the chunk go to shared memory as u16; the raw row goes to a small ring for the centre pixel.
//   H phase    thread (row, 64-column segment) slides the three box sums along its row: dp2a adds the entering and
//              subtracts the leaving u16 element (extract + accumulate in one instruction), every 16-byte piece of a
//              plane row is loaded once and kept in registers while a window still needs it; the test
//              S - k^2 g - c_k >= 0 (c_k = k^2 C - (k^2 - 1) / 2, exact integers, same as 2 S >= (2 (g + C) - 1) k^2)
//              is one IMAD, and its sign bit enters the mask word through one funnel shift.
// ---------------------------------------------------------------------------------------------
constexpr int TM_WT = 320;                            // output columns of a strip (10 mask words)
constexpr int TM_HALO = 12;                           // >= largest radius + 1, multiple of 4
constexpr int TM_VCOLS = TM_WT + 2 * TM_HALO;         // 344 columns of V per strip row
constexpr int TM_VG = TM_VCOLS / 4;                   // 86 four-column groups = V-phase threads
constexpr int TM_RING = 24;                           // length of the prefix ring (largest window + 1)
constexpr int TM_THREADS = 128;
constexpr int TM_VPITCH = TM_VG * 8;                  // 688 bytes per plane row: 43 x 16 (odd => 16-byte loads of 8 rows hit 8 bank groups)
constexpr int TM_GPITCH = TM_WT + 16;                 // 336 = 21 x 16
constexpr int TM_MAX_HS = 216;                        // (Hs + 22) * 255 < 2^16
// RC = rows per chunk (24 or 12): the three u16 planes of a chunk are what shared memory holds, so 12-row chunks halve the
// footprint (37 KB instead of 74 KB) and more CTAs share an SM; the H phase then covers a chunk with 12 rows x 10 segments of
// 32 columns instead of 24 x 5 of 64.  The raw-row ring of the centre pixels holds two chunks.
// BGR: the kernel reads bgr8 frames, converts on the fly (S0 of SURVEY 8(a)) and also leaves the gray plane for the later stages;
// a staged row word is then three words (four pixels x 3 bytes).
template <int RC, bool BGR = false> struct TmCfg {
    static_assert(RC == 24 || RC == 12, "chunk rows");
    static constexpr int SW = BGR ? 3 : 1;                // staging words per (row, V thread)
    static constexpr int SEG = 64 * RC / 24;              // output columns of one H-phase thread
    static constexpr int NSEG = TM_WT / SEG;
    static constexpr int GROWS = 2 * RC;
    static constexpr size_t SMEM = (size_t)3 * RC * TM_VPITCH + (size_t)GROWS * TM_GPITCH + (size_t)RC * TM_VG * 4 * SW;
    static_assert(RC * NSEG <= TM_THREADS, "thread roles");
};
constexpr int TM_RC = 24;                             // the chunk height the host sizes work items by (a multiple of both variants)
static_assert((TM_VPITCH / 16) % 2 == 1 && (TM_GPITCH / 16) % 2 == 1 && TM_VPITCH % 16 == 0, "bank mapping");
static_assert(TM_VG <= TM_THREADS, "thread roles");

__device__ __forceinline__ int tm_dp2a(uint32_t a, int sel, int c)      // c + s16(a.lo) * s8(sel.b0) + s16(a.hi) * s8(sel.b1)
{
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(sel), "r"(c));
    return d;
}
// element e (0..7) of a 16-byte piece = two groups of four columns stored as (c0 | c2 << 16), (c1 | c3 << 16)
__device__ __forceinline__ uint32_t tm_word(const uint4 &c, int e)
{
    const int wi = 2 * (e >> 2) + (e & 1);
    return wi == 0 ? c.x : wi == 1 ? c.y : wi == 2 ? c.z : c.w;
}
__device__ __forceinline__ int tm_half(int e) { return (e >> 1) & 1; }

// the box sums of one window along a row segment: pieces ch[] of the plane row are loaded on demand by the caller
template <int R>
struct TmWin {
    static constexpr int K2 = (2 * R + 1) * (2 * R + 1);
    static constexpr int first_piece = (TM_HALO - R) / 8;
    static constexpr int __host__ __device__ newest_piece(int blk) { return (8 * blk + 8 + TM_HALO + R) / 8; }     // piece of the last entering element of block blk
    uint4 ch[12];
    int S;
    __device__ __forceinline__ void load(const uint8_t *row, int from, int to)
    {
#pragma unroll
        for (int c = 0; c < 12; ++c)
            if (c >= from && c <= to) ch[c] = *reinterpret_cast<const uint4 *>(row + 16 * c);
    }
    __device__ __forceinline__ void init(int bias)                 // S at segment column 0: columns 12 - R .. 12 + R of the row
    {
        S = bias;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int wi = 0; wi < 4; ++wi) {
                const int qa = 8 * c + 4 * (wi >> 1) + (wi & 1), qb = qa + 2;
                const int sel = ((qa >= TM_HALO - R && qa <= TM_HALO + R) ? 1 : 0) | ((qb >= TM_HALO - R && qb <= TM_HALO + R) ? 0x100 : 0);
                if (sel) S = tm_dp2a(wi == 0 ? ch[c].x : wi == 1 ? ch[c].y : wi == 2 ? ch[c].z : ch[c].w, sel, S);
            }
    }
    __device__ __forceinline__ void slide(int i)                   // S(i) -> S(i + 1)
    {
        const int qn = i + TM_HALO + 1 + R, qo = i + TM_HALO - R;
        S = tm_dp2a(tm_word(ch[qn >> 3], qn & 7), tm_half(qn & 7) ? 0x0100 : 0x0001, S);
        S = tm_dp2a(tm_word(ch[qo >> 3], qo & 7), tm_half(qo & 7) ? 0xFF00 : 0x00FF, S);
    }
};

// one 4-byte asynchronous copy global -> shared (LDGSTS): the V phase prefetches the next chunk's rows with it
__device__ __forceinline__ void tm_cp_async4(uint32_t dst_shared, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst_shared), "l"(src) : "memory");
}
__device__ __forceinline__ void tm_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tm_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// four bgr8 pixels (three words) -> four gray bytes, g = (3735 B + 19235 G + 9798 R + 16384) >> 15 (cv2's fixed point, bit-exact):
// every pixel is two dp2a with doubled weights, so that the result sits in byte 2 of the sum and three byte permutes pack the word
__device__ __forceinline__ uint32_t tm_dp2a_lo_u(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t tm_dp2a_hi_u(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t tm_bgr_to_gray4(uint32_t w0, uint32_t w1, uint32_t w2)
{
    constexpr uint32_t WB = 2 * 3735, WG = 2 * 19235, WR = 2 * 9798, RND = 2 * 16384;
    // w0 = B0 G0 R0 B1, w1 = G1 R1 B2 G2, w2 = R2 B3 G3 R3 (byte 0 first)
    const uint32_t s0 = tm_dp2a_hi_u(WR, w0, tm_dp2a_lo_u(WB | (WG << 16), w0, RND));
    const uint32_t s1 = tm_dp2a_lo_u(WG | (WR << 16), w1, tm_dp2a_hi_u(WB << 16, w0, RND));
    const uint32_t s2 = tm_dp2a_lo_u(WR, w2, tm_dp2a_hi_u(WB | (WG << 16), w1, RND));
    const uint32_t s3 = tm_dp2a_hi_u(WG | (WR << 16), w2, tm_dp2a_lo_u(WB << 16, w2, RND));
    return __byte_perm(__byte_perm(s0, s1, 0x4462), __byte_perm(s2, s3, 0x4462), 0x5410);
}

// shared-memory views and per-thread constants of one CTA
struct TmCtx {
    uint8_t *pl0, *pl1, *pl2, *stash, *my_stash, *my_stage;
    uint32_t my_stage_s;
    int tid, wmax;
    bool vthread, stash_ok;
    uint8_t *gdst;            // BGR only: this thread's word of image row Y0 + 11 in the gray plane (null: column outside the frame)
    uint32_t gpitch;
};

// V phase of one chunk: the chunk's RC raw rows (already in the staging words) extend the prefix ring and leave the three
// plane rows.  PAR = parity of the chunk: the ring position of a 12-row chunk ((first output row) mod 24) and the half of the
// raw-row ring a chunk writes alternate, and the ring lives in registers, so the two parities are two instantiations.
// BGR: grows = how many of the chunk's rows are rows of this work item (their gray words go to the gray plane), m0 = first output row of the chunk
template <int R0, int R1, int R2, int RC, int PAR, bool BGR>
__device__ __forceinline__ void tm_v_chunk(const TmCtx &c, uint32_t (&Ce)[TM_RING], uint32_t (&Co)[TM_RING], uint32_t &ce, uint32_t &co,
                                           uint32_t sel_e, uint32_t sel_o, int m0, int grows)
{
    constexpr int G = TmCfg<RC>::GROWS, OFF = (RC * PAR) % TM_RING, SW = TmCfg<RC, BGR>::SW;
    uint32_t w[RC];
    tm_cp_async_wait_all();
#pragma unroll
    for (int u = 0; u < RC; ++u) {
        const volatile uint32_t *sp = reinterpret_cast<const volatile uint32_t *>(c.my_stage + u * (TM_VG * 4 * SW));
        if (BGR) {
            w[u] = tm_bgr_to_gray4(sp[0], sp[1], sp[2]);
            if (c.gdst && u < grows) *reinterpret_cast<uint32_t *>(c.gdst + (size_t)(m0 + u) * c.gpitch) = w[u];
        } else w[u] = sp[0];
    }
#pragma unroll
    for (int u = 0; u < RC; ++u) {
        ce += __byte_perm(w[u], 0, sel_e); co += __byte_perm(w[u], 0, sel_o);
        Ce[(OFF + u + 23) % TM_RING] = ce; Co[(OFF + u + 23) % TM_RING] = co;             // C[m + 22]: raw index m + 22 -> slot (m + 23) % 24
        if (c.stash_ok) *reinterpret_cast<uint32_t *>(c.my_stash + ((RC * PAR + u + 22 + 26) % G) * TM_GPITCH) = w[u];   // raw index m0 + 22 + u
        // V_r[m] = C[m + r + 11] - C[m - r + 10]   (raw index = output row + 11)
        *reinterpret_cast<uint2 *>(c.pl0 + u * TM_VPITCH + 8 * c.tid) =
            make_uint2(Ce[(OFF + u + R0 + 12) % TM_RING] - Ce[(OFF + u + 11 - R0) % TM_RING], Co[(OFF + u + R0 + 12) % TM_RING] - Co[(OFF + u + 11 - R0) % TM_RING]);
        *reinterpret_cast<uint2 *>(c.pl1 + u * TM_VPITCH + 8 * c.tid) =
            make_uint2(Ce[(OFF + u + R1 + 12) % TM_RING] - Ce[(OFF + u + 11 - R1) % TM_RING], Co[(OFF + u + R1 + 12) % TM_RING] - Co[(OFF + u + 11 - R1) % TM_RING]);
        *reinterpret_cast<uint2 *>(c.pl2 + u * TM_VPITCH + 8 * c.tid) =
            make_uint2(Ce[(OFF + u + R2 + 12) % TM_RING] - Ce[(OFF + u + 11 - R2) % TM_RING], Co[(OFF + u + R2 + 12) % TM_RING] - Co[(OFF + u + 11 - R2) % TM_RING]);
    }
}

// requires pitch, frame_stride and the base pointer to be multiples of 4 (the host checks and falls back to k_threshold3)
// BGR: `gray` points at bgr8 frames (pitch, frame_stride in bytes of that layout, W a multiple of 4) and the gray plane is written to gray_out
template <int R0, int R1, int R2, int RC, int MINB, bool BGR = false>
__global__ void __launch_bounds__(TM_THREADS, MINB)
k_threshold_march(const uint8_t *__restrict__ gray, uint32_t pitch, size_t frame_stride, uint32_t *__restrict__ masks, DetGeom g,
                  int Hs, int n_sy, int n_sx, int n_items, uint8_t *__restrict__ gray_out = nullptr, uint32_t gray_pitch = 0, size_t gray_frame = 0)
{
    static_assert(R0 <= 11 && R1 <= 11 && R2 <= 11 && R0 >= 1 && R1 >= 1 && R2 >= 1, "radii");
    using Cfg = TmCfg<RC, BGR>;
    constexpr int SW = Cfg::SW;
    constexpr int SEG = Cfg::SEG;
    extern __shared__ __align__(16) uint8_t tm_smem[];
    TmCtx c;
    c.pl0 = tm_smem; c.pl1 = c.pl0 + RC * TM_VPITCH; c.pl2 = c.pl1 + RC * TM_VPITCH;
    c.stash = c.pl2 + RC * TM_VPITCH;                               // raw rows of the output columns, slot = (raw index + 26) % GROWS
    uint8_t *stage = c.stash + Cfg::GROWS * TM_GPITCH;              // [RC][86] words: the chunk's raw rows, each word private to its V thread
    const int tid = threadIdx.x;
    c.tid = tid;
    c.vthread = tid < TM_VG;
    c.stash_ok = tid >= TM_HALO / 4 && tid < TM_VG - TM_HALO / 4;
    c.my_stash = c.stash + 4 * (tid - TM_HALO / 4);
    c.my_stage = stage + 4 * SW * tid;
    c.my_stage_s = (uint32_t)__cvta_generic_to_shared(c.my_stage);
    c.gpitch = gray_pitch;
    const int rho = tid % RC, sigma = tid / RC;
    const bool hthread = tid < RC * Cfg::NSEG;
    const int bias0 = -(TmWin<R0>::K2 * g.Cfloor - (TmWin<R0>::K2 - 1) / 2);
    const int bias1 = -(TmWin<R1>::K2 * g.Cfloor - (TmWin<R1>::K2 - 1) / 2);
    const int bias2 = -(TmWin<R2>::K2 * g.Cfloor - (TmWin<R2>::K2 - 1) / 2);
    c.wmax = (g.W - 1) >> 2;                                        // last 4-column word that starts inside a row

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int sy = item % n_sy;
        const int t2 = item / n_sy;
        const int sx = t2 % n_sx, b = t2 / n_sx;
        const int X0 = sx * TM_WT, Y0 = sy * Hs;
        const int rows = min(Hs, g.H - Y0);
        const int ytop = Y0 - (TM_HALO - 1);                        // image row of raw index 0
        // this thread's word of every row: replicate border = clamp the word index, then pick the bytes with the two
        // unpack permutes (even columns -> Ce lanes, odd columns -> Co lanes); a selector nibble of 4 reads a zero byte
        const int wcol = (X0 - TM_HALO) / 4 + tid;
        const int wc = min(max(wcol, 0), c.wmax);
        uint32_t sel_e, sel_o;
        {
            int bj[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bj[j] = min(max(4 * wcol + j, 0), g.W - 1) - 4 * wc;
            sel_e = (uint32_t)bj[0] | 0x40u | ((uint32_t)bj[2] << 8) | 0x4000u;
            sel_o = (uint32_t)bj[1] | 0x40u | ((uint32_t)bj[3] << 8) | 0x4000u;
        }
        const uint8_t *colp = gray + (size_t)b * frame_stride + 4 * SW * (size_t)wc;
        // the gray plane gets this item's own rows and columns (the V phase runs 11 rows ahead of the output rows)
        c.gdst = (BGR && c.stash_ok && wcol <= c.wmax) ? gray_out + (size_t)b * gray_frame + (size_t)(Y0 + 11) * gray_pitch + 4 * (size_t)wcol : nullptr;

        // prefix ring: slot (j + 1) % 24 holds the column prefix through raw index j; slot 0 starts as C[-1] = 0
        uint32_t Ce[TM_RING], Co[TM_RING];
#pragma unroll
        for (int k = 0; k < TM_RING; ++k) { Ce[k] = 0; Co[k] = 0; }
        uint32_t ce = 0, co = 0;
        // rows raw index r_first .. r_first + RC - 1 start towards the staging words
        auto prefetch = [&](int r_first) {
            if (r_first >= 0 && r_first + RC - 1 <= g.H - 1) {                                // no clamping inside the frame
                const uint8_t *p = colp + (size_t)((uint32_t)r_first * (uint64_t)pitch);
#pragma unroll
                for (int u = 0; u < RC; ++u)
#pragma unroll
                    for (int k = 0; k < SW; ++k) tm_cp_async4(c.my_stage_s + u * (TM_VG * 4 * SW) + 4 * k, p + (size_t)((uint32_t)u * (uint64_t)pitch) + 4 * k);
            } else {
#pragma unroll
                for (int u = 0; u < RC; ++u) {
                    const int gy = min(max(r_first + u, 0), g.H - 1);
#pragma unroll
                    for (int k = 0; k < SW; ++k) tm_cp_async4(c.my_stage_s + u * (TM_VG * 4 * SW) + 4 * k, colp + (size_t)((uint32_t)gy * (uint64_t)pitch) + 4 * k);
                }
            }
            tm_cp_async_commit();
        };
        if (c.vthread) {
            // chunk 0's rows (raw index 22 .. 22 + RC - 1) start towards shared memory, then the 22 warm-up rows come through registers
            prefetch(ytop + 22);
            uint32_t w[22];
            if (BGR) {
#pragma unroll
                for (int i0 = 0; i0 < 22; i0 += 11) {                                           // two rounds of 33 loads in flight
                    uint32_t t[33];
#pragma unroll
                    for (int i = 0; i < 11; ++i) {
                        const int gy = min(max(ytop + i0 + i, 0), g.H - 1);
                        const uint32_t *q = reinterpret_cast<const uint32_t *>(colp + (size_t)((uint32_t)gy * (uint64_t)pitch));
                        t[3 * i] = __ldg(q); t[3 * i + 1] = __ldg(q + 1); t[3 * i + 2] = __ldg(q + 2);
                    }
#pragma unroll
                    for (int i = 0; i < 11; ++i) {
                        w[i0 + i] = tm_bgr_to_gray4(t[3 * i], t[3 * i + 1], t[3 * i + 2]);
                        // raw index i0 + i is image row Y0 - 11 + i0 + i: rows 11 .. 21 are this item's first output rows
                        if (c.gdst && i0 + i >= 11 && i0 + i - 11 < rows) *reinterpret_cast<uint32_t *>(c.gdst + (ptrdiff_t)(i0 + i - 22) * (ptrdiff_t)gray_pitch) = w[i0 + i];
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 22; ++i) {
                    const int gy = min(max(ytop + i, 0), g.H - 1);
                    w[i] = __ldg(reinterpret_cast<const uint32_t *>(colp + (size_t)((uint32_t)gy * (uint64_t)pitch)));
                }
            }
#pragma unroll
            for (int i = 0; i < 22; ++i) {
                ce += __byte_perm(w[i], 0, sel_e); co += __byte_perm(w[i], 0, sel_o);
                Ce[i + 1] = ce; Co[i + 1] = co;
                if (c.stash_ok) *reinterpret_cast<uint32_t *>(c.my_stash + ((i + 26) % Cfg::GROWS) * TM_GPITCH) = w[i];
            }
        }
        for (int m0 = 0, par = 0; m0 < rows; m0 += RC, par ^= 1) {
            if (c.vthread) {
                // chunk row u is image row Y0 + 11 + m0 + u: the item's own rows end at Y0 + rows - 1
                if (par == 0) tm_v_chunk<R0, R1, R2, RC, 0, BGR>(c, Ce, Co, ce, co, sel_e, sel_o, m0, rows - 11 - m0);
                else tm_v_chunk<R0, R1, R2, RC, 1, BGR>(c, Ce, Co, ce, co, sel_e, sel_o, m0, rows - 11 - m0);
                if (m0 + RC < rows) prefetch(ytop + m0 + RC + 22);                            // the next chunk's rows fly during this chunk's H phase
            }
            __syncthreads();
            const int y = Y0 + m0 + rho;
            if (hthread && y < g.H && X0 + SEG * sigma < g.W) {
                const uint8_t *row0 = c.pl0 + rho * TM_VPITCH + 2 * SEG * sigma;
                const uint8_t *row1 = c.pl1 + rho * TM_VPITCH + 2 * SEG * sigma;
                const uint8_t *row2 = c.pl2 + rho * TM_VPITCH + 2 * SEG * sigma;
                const uint8_t *grow = c.stash + ((m0 + rho + 11 + 26) % Cfg::GROWS) * TM_GPITCH + SEG * sigma;   // the centre pixels: raw index = row + 11
                TmWin<R0> w0; TmWin<R1> w1; TmWin<R2> w2;
                w0.load(row0, TmWin<R0>::first_piece, TmWin<R0>::newest_piece(0) - 1);
                w1.load(row1, TmWin<R1>::first_piece, TmWin<R1>::newest_piece(0) - 1);
                w2.load(row2, TmWin<R2>::first_piece, TmWin<R2>::newest_piece(0) - 1);
                w0.init(bias0); w1.init(bias1); w2.init(bias2);
                uint32_t m_0 = 0, m_1 = 0, m_2 = 0;
                uint32_t *mrow = masks + (size_t)b * 3 * g.mask_plane + (size_t)(y + 1) * g.PWW + 1 + X0 / 32 + (SEG / 32) * sigma;
                uint4 gq;
#pragma unroll
                for (int blk = 0; blk < SEG / 8; ++blk) {
                    w0.load(row0, blk == 0 ? TmWin<R0>::newest_piece(0) : TmWin<R0>::newest_piece(blk - 1) + 1, TmWin<R0>::newest_piece(blk));
                    w1.load(row1, blk == 0 ? TmWin<R1>::newest_piece(0) : TmWin<R1>::newest_piece(blk - 1) + 1, TmWin<R1>::newest_piece(blk));
                    w2.load(row2, blk == 0 ? TmWin<R2>::newest_piece(0) : TmWin<R2>::newest_piece(blk - 1) + 1, TmWin<R2>::newest_piece(blk));
                    if ((blk & 1) == 0) gq = *reinterpret_cast<const uint4 *>(grow + 8 * blk);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int i = 8 * blk + j;
                        const int e = i & 15;
                        const uint32_t gw = (e >> 2) == 0 ? gq.x : (e >> 2) == 1 ? gq.y : (e >> 2) == 2 ? gq.z : gq.w;
                        const int gv = (int)__byte_perm(gw, 0, 0x4440 + (e & 3));
                        const int d0 = w0.S - TmWin<R0>::K2 * gv, d1 = w1.S - TmWin<R1>::K2 * gv, d2 = w2.S - TmWin<R2>::K2 * gv;
                        m_0 = __funnelshift_l((uint32_t)d0, m_0, 1);                 // (m << 1) | sign(d)
                        m_1 = __funnelshift_l((uint32_t)d1, m_1, 1);
                        m_2 = __funnelshift_l((uint32_t)d2, m_2, 1);
                        if (i + 1 < SEG) { w0.slide(i); w1.slide(i); w2.slide(i); }
                    }
                    if ((blk & 3) == 3) {                                            // 32 columns done: first column ends up in bit 0, set = test passed
                        const int wd = blk >> 2;
                        const int left = g.W - (X0 + SEG * sigma + 32 * wd);
                        if (left > 0) {
                            const uint32_t valid = left >= 32 ? 0xFFFFFFFFu : ((1u << left) - 1u);
                            mrow[wd] = ~__brev(m_0) & valid;
                            mrow[g.mask_plane + wd] = ~__brev(m_1) & valid;
                            mrow[2 * g.mask_plane + wd] = ~__brev(m_2) & valid;
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
}

// expand packed masks to 0/255 bytes (debug / parity tap)
__global__ void k_unpack_masks(const uint32_t *__restrict__ masks, uint8_t *__restrict__ out, DetGeom g)
{
    const long long total = (long long)g.B * g.nScales * g.H * g.W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % g.W);
        long long t = i / g.W;
        const int y = (int)(t % g.H);
        t /= g.H;                                             // frame * nScales + scale
        const uint32_t w = masks[(size_t)t * g.mask_plane + (size_t)(y + 1) * g.PWW + (x >> 5) + 1];
        out[i] = ((w >> (x & 31)) & 1u) ? 255 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// A3a step 1: anchors and start candidates of the border graph (core.h), one mask word (32 pixels) per thread.
//   ast[i]    = (x | y << 16, (frame*nScales+scale) << 4 | super << 3 | s_in)
//   starts[j] = (x | y << 16, (frame*nScales+scale) << 4 | s_in)
//   amap[(fs*H + y)*WW + wx] = index of the word's first anchor (anchors of one word are consecutive)
// ---------------------------------------------------------------------------------------------
struct BorderGraph {
    uint2 *ast;          // anchor states
    Seg *seg;            // (next, prev, len | SEG_SUPER, minkey)
    uint32_t *minoff;    // offset of the segment's min-key state
    Seg *sseg;           // super anchors only: (next super, previous super, length up to it, min key)
    uint32_t *ssoff;     // super anchors only: offset of the min-key state from the super anchor
    int2 *emit;          // (position of the segment's first state in its border, 1 + slot of the border in `sorted`; 0 = not kept)
    uint32_t *codes;     // per anchor SEG_CODE_WORDS words: the segment's first successor directions, 3 bits each
    uint32_t *amap;      // per mask word of every (frame,scale): first anchor index
    unsigned *n_anchors; // this sub-batch's counter
    unsigned cap;
    uint2 *starts;       // start candidates
    unsigned *n_starts;
    unsigned starts_cap;
};

__device__ __forceinline__ void load_walk_tables(const WalkTables *__restrict__ g, uint16_t *s_succ, uint16_t *s_pred)
{
    const uint4 *src = reinterpret_cast<const uint4 *>(g);
    if (s_succ) for (int i = threadIdx.x; i < 512; i += blockDim.x) reinterpret_cast<uint4 *>(s_succ)[i] = __ldg(src + i);
    if (s_pred) for (int i = threadIdx.x; i < 512; i += blockDim.x) reinterpret_cast<uint4 *>(s_pred)[i] = __ldg(src + 512 + i);
    __syncthreads();
}

// warp-aggregated append of cnt items per lane: returns the lane's first slot
__device__ __forceinline__ unsigned warp_append(unsigned *counter, int cnt, int lane)
{
    int incl = cnt;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) { const int o = __shfl_up_sync(0xFFFFFFFFu, incl, dd); if (lane >= dd) incl += o; }
    const int tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
    unsigned base = 0;
    if (tot) {
        if (lane == 0) base = atomicAdd(counter, (unsigned)tot);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
    }
    return base + (unsigned)(incl - cnt);
}

// The word logic costs a few hundred instructions but only ~10 % of the mask words touch a border, so a
// CTA first filters 256 words per iteration with three loads each (empty words and words in the
// interior of a region are dropped) and queues the others in shared memory; whenever 256 are queued
// every thread takes one, so the logic always runs on full warps, and the two global list counters see
// one atomic per 256 words (same-address atomics serialise in L2: one per warp was 60 % of this kernel).
constexpr int ANCHOR_SCAN_U = 4;      // mask words a thread filters per iteration (independent loads in flight, a quarter of the barriers)
struct AnchorBlock {
    unsigned queue[256 + 256 * ANCHOR_SCAN_U];
    int nq;
    int wsum[2][16];
    unsigned base[2];
};

// exclusive positions of this thread's cnt0 anchors and cnt1 start candidates in the CTA-wide appends to the two list counters:
// one round (two barriers, the two atomics issued together) for both lists
__device__ __forceinline__ void block_append2(AnchorBlock &sb, unsigned *counter0, unsigned *counter1, int cnt0, int cnt1, unsigned &pos0, unsigned &pos1)
{
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    int incl = cnt0 | (cnt1 << 16);                     // a word holds at most 128 states of either kind: the two scans share the shuffles
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) { const int o = __shfl_up_sync(0xFFFFFFFFu, incl, dd); if (lane >= dd) incl += o; }
    if (lane == 31) sb.wsum[0][wp] = incl;
    __syncthreads();
    if (threadIdx.x < 2) {
        const int sh = 16 * threadIdx.x;
        int tot = 0;
        for (int w = 0; w < 8; ++w) { const int v = (sb.wsum[0][w] >> sh) & 0xFFFF; sb.wsum[1][2 * w + threadIdx.x] = tot; tot += v; }
        sb.base[threadIdx.x] = tot ? atomicAdd(threadIdx.x ? counter1 : counter0, (unsigned)tot) : 0u;
    }
    __syncthreads();
    const int excl = incl - (cnt0 | (cnt1 << 16));
    pos0 = sb.base[0] + (unsigned)sb.wsum[1][2 * wp] + (unsigned)(excl & 0xFFFF);
    pos1 = sb.base[1] + (unsigned)sb.wsum[1][2 * wp + 1] + (unsigned)(excl >> 16);
}

__device__ __forceinline__ void anchors_of_word(AnchorBlock &sb, const uint32_t *__restrict__ masks, const BorderGraph &bg, int *__restrict__ iso_count,
                                                int Rm, int Rm2, const DetGeom &g, unsigned widx, bool valid)
{
    uint32_t A[4] = {0, 0, 0, 0}, SU[4] = {0, 0, 0, 0}, U[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
    int fs = 0, y = 0, wx = 0;
    if (valid) {
        const unsigned words_per_plane = (unsigned)g.H * (unsigned)g.WW;
        fs = (int)(widx / words_per_plane);
        const unsigned rem = widx - (unsigned)fs * words_per_plane;
        y = (int)(rem / (unsigned)g.WW); wx = (int)(rem - (unsigned)y * (unsigned)g.WW);
        const uint32_t *row = masks + (size_t)fs * g.mask_plane + (size_t)(y + 1) * g.PWW + wx + 1;
        const uint32_t m = __ldg(row), ml = __ldg(row - 1), mr = __ldg(row + 1);
        const uint32_t u = __ldg(row - g.PWW), ul = __ldg(row - g.PWW - 1), ur = __ldg(row - g.PWW + 1);
        const uint32_t d = __ldg(row + g.PWW), dl = __ldg(row + g.PWW - 1), dr = __ldg(row + g.PWW + 1);
        uint32_t iso;
        anchor_words(m, ml, mr, u, ul, ur, d, dl, dr, !(y & Rm), grid_cols(wx, Rm), !(y & Rm2), grid_cols(wx, Rm2), A, SU, U, hi, iso);
        if (iso && g.count_all) atomicAdd(&iso_count[fs], __popc(iso));       // one hot address per mask: never in the product call
    }
    // anchors: pixel by pixel, canonical directions E, N, W, S inside a pixel
    const int cnt = __popc(A[0]) + __popc(A[1]) + __popc(A[2]) + __popc(A[3]);
    const int ucnt = __popc(U[0]) + __popc(U[1]) + __popc(U[2]) + __popc(U[3]);
    unsigned pos, upos;
    block_append2(sb, bg.n_anchors, bg.n_starts, cnt, ucnt, pos, upos);
    if (cnt && pos < bg.cap) bg.amap[((size_t)fs * g.H + y) * g.WW + wx] = pos;
    for (uint32_t px = A[0] | A[1] | A[2] | A[3]; px; px &= px - 1) {
        const int b = __ffs(px) - 1, x = wx * 32 + b;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!((A[k] >> b) & 1u)) continue;
            if (pos < bg.cap) {
                const unsigned sup = (SU[k] >> b) & 1u;
                bg.ast[pos] = make_uint2((unsigned)x | ((unsigned)y << 16), ((unsigned)fs << 4) | (sup << 3) | (unsigned)state_dir(k, hi[k], b));
                bg.seg[pos] = Seg{A_NONE, A_NONE, 0u, A_NONE};
                bg.emit[pos] = make_int2(0, 0);
                if (sup) bg.sseg[pos] = Seg{A_NONE, A_NONE, 0u, A_NONE};
            }
            ++pos;
        }
    }
    // start candidates
#pragma unroll
    for (int k = 0; k < 4; ++k)
        for (uint32_t px = U[k]; px; px &= px - 1) {
            const int b = __ffs(px) - 1;
            if (upos < bg.starts_cap) bg.starts[upos] = make_uint2((unsigned)(wx * 32 + b) | ((unsigned)y << 16), ((unsigned)fs << 4) | (unsigned)state_dir(k, hi[k], b));
            ++upos;
        }
}

__global__ void __launch_bounds__(256)
k_anchors(const uint32_t *__restrict__ masks, BorderGraph bg, int *__restrict__ iso_count, int Rm, int Rm2, DetGeom g)
{
    __shared__ AnchorBlock sb;
    const unsigned words_per_plane = (unsigned)g.H * (unsigned)g.WW;
    const unsigned total = (unsigned)(g.B * g.nScales) * words_per_plane;
    const int lane = threadIdx.x & 31;
    const unsigned stride = gridDim.x * blockDim.x;
    if (threadIdx.x == 0) sb.nq = 0;
    __syncthreads();
    // word index = (fs * H + y) * WW + wx; every thread of the CTA runs the same number of iterations
    constexpr int U = ANCHOR_SCAN_U;
    for (unsigned i0 = blockIdx.x * blockDim.x * U; i0 < total; i0 += stride * U) {
        const uint32_t *row[U];
        uint32_t m[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const unsigned i = i0 + k * blockDim.x + threadIdx.x;
            m[k] = 0; row[k] = masks;
            if (i < total) {
                const unsigned fs = i / words_per_plane, rem = i - fs * words_per_plane;
                const unsigned y = rem / (unsigned)g.WW, wx = rem - y * (unsigned)g.WW;
                row[k] = masks + (size_t)fs * g.mask_plane + (size_t)(y + 1) * g.PWW + wx + 1;
                m[k] = __ldg(row[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            bool keep = false;
            if (m[k]) {
                // interior of a region: the word, the words above and below are full and so are the six flanking bits
                const uint32_t *r = row[k];
                const uint32_t u = __ldg(r - g.PWW), d = __ldg(r + g.PWW);
                keep = (m[k] & u & d) != 0xFFFFFFFFu || !(__ldg(r - 1) >> 31) || !(__ldg(r + 1) & 1u) ||
                       !(__ldg(r - g.PWW - 1) >> 31) || !(__ldg(r - g.PWW + 1) & 1u) || !(__ldg(r + g.PWW - 1) >> 31) || !(__ldg(r + g.PWW + 1) & 1u);
            }
            const unsigned km = __ballot_sync(0xFFFFFFFFu, keep);
            int wbase = 0;
            if (lane == 0 && km) wbase = atomicAdd(&sb.nq, __popc(km));
            wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
            if (keep) sb.queue[wbase + __popc(km & ((1u << lane) - 1u))] = i0 + k * blockDim.x + threadIdx.x;   // nq < 256 before the iteration: never past the queue
        }
        __syncthreads();
        for (int nq = sb.nq; nq >= 256; nq -= 256) {                              // the same count in every thread
            const unsigned widx = sb.queue[nq - 256 + threadIdx.x];                  // take the last 256: the rest stays in place
            anchors_of_word(sb, masks, bg, iso_count, Rm, Rm2, g, widx, true);
        }
        __syncthreads();
        if (threadIdx.x == 0) sb.nq &= 255;
        __syncthreads();
    }
    const int nq = sb.nq;
    if (nq > 0) anchors_of_word(sb, masks, bg, iso_count, Rm, Rm2, g, (int)threadIdx.x < nq ? sb.queue[threadIdx.x] : 0u, (int)threadIdx.x < nq);
}

// ---------------------------------------------------------------------------------------------
// A3a step 2: one thread per anchor walks its segment to the next anchor and links the two.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_segments(const uint32_t *__restrict__ masks, BorderGraph bg, int max_len, const WalkTables *__restrict__ tables, int Rm, DetGeom g)
{
    __shared__ __align__(16) uint16_t s_succ[4096];
    unsigned n = *bg.n_anchors;
    if (n > bg.cap) n = bg.cap;
    if (blockIdx.x * blockDim.x >= n) return;
    load_walk_tables(tables, s_succ, nullptr);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 a = bg.ast[i];
        const int fs = (int)(a.y >> 4);
        int x = (int)(a.x & 0xFFFFu), y = (int)(a.x >> 16), s = (int)(a.y & 7u);
        CachedMaskView rd(masks + (size_t)fs * g.mask_plane, g.PWW);
        uint32_t len, minkey, moff;
        seg_walk(rd, s_succ, g.KS, Rm, max_len, x, y, s, len, minkey, moff, bg.codes + (size_t)i * SEG_CODE_WORDS);
        uint32_t j = A_NONE;
        if (len != SEG_OVERFLOW) {
            const MaskView mv{masks + (size_t)fs * g.mask_plane, g.PWW};
            const int r = anchor_rank_in_word(mv, x, y, s, Rm);
            j = bg.amap[((size_t)fs * g.H + y) * g.WW + (x >> 5)] + (uint32_t)r;
            if (r < 0 || j >= n) j = A_NONE;                          // only after an anchor-list overflow (status 3)
        }
        bg.seg[i].next = j; bg.seg[i].len = len | ((a.y & 8u) ? SEG_SUPER : 0u); bg.seg[i].minkey = minkey;
        bg.minoff[i] = moff;
        if (j != A_NONE) bg.seg[j].prev = i;
    }
}

// ---------------------------------------------------------------------------------------------
// A3a step 3: super anchors skip to the next super anchor; then one thread per anchor hops over its
// border's (super) anchors and one thread per start candidate walks the borders without anchors; the
// leaders report.   surv[fs*surv_cap + slot] = (key, length, leader anchor | x + (y << 16), kind)
//   kind 0: leader is a plain anchor, 1: a super anchor, 2 | s_in << 8: a border without anchors from its first state
// ---------------------------------------------------------------------------------------------
struct SegLoad {
    const Seg *seg;
    __device__ __forceinline__ Seg operator()(uint32_t i) const
    {
        const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(seg) + i);
        return Seg{v.x, v.y, v.z, v.w};
    }
};
struct MinoffLoad {
    const uint32_t *m;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return __ldcg(m + i); }
};

__global__ void __launch_bounds__(256)
k_skip(BorderGraph bg, int max_len)
{
    unsigned n = *bg.n_anchors;
    if (n > bg.cap) n = bg.cap;
    const SegLoad seg_at{bg.seg};
    const MinoffLoad minoff_at{bg.minoff};
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (!(bg.ast[i].y & 8u)) continue;
        uint32_t snext, slen, smin, soff;
        super_skip(seg_at, minoff_at, i, max_len, snext, slen, smin, soff);
        bg.sseg[i].next = snext; bg.sseg[i].len = slen; bg.sseg[i].minkey = smin;
        bg.ssoff[i] = soff;
        if (snext != A_NONE) bg.sseg[snext].prev = i;
    }
}

__device__ __forceinline__ void report_border(uint4 *__restrict__ surv, int *__restrict__ surv_count, int *__restrict__ contour_count,
                                              int fs, uint32_t key, uint32_t len, uint32_t who, uint32_t kind, const DetGeom &g)
{
    if (g.count_all) atomicAdd(&contour_count[fs], 1);                          // one hot address per mask: never in the product call
    if ((int)len >= g.minPerim && (int)len <= g.maxPerim) {
        const int slot = atomicAdd(&surv_count[fs], 1);
        if (slot < g.surv_cap) surv[(size_t)fs * g.surv_cap + slot] = make_uint4(key, len, who, kind);
    }
}

__global__ void __launch_bounds__(256)
k_cycles(const uint32_t *__restrict__ masks, BorderGraph bg, uint4 *__restrict__ surv, int *__restrict__ surv_count, int *__restrict__ contour_count,
         int max_len, const WalkTables *__restrict__ tables, int Rm, DetGeom g)
{
    __shared__ __align__(16) uint16_t s_succ[4096];
    __shared__ __align__(16) uint16_t s_pred[4096];
    unsigned n = *bg.n_anchors, ns = *bg.n_starts;
    if (n > bg.cap) n = bg.cap;
    if (ns > bg.starts_cap) ns = bg.starts_cap;
    load_walk_tables(tables, s_succ, s_pred);
    const SegLoad seg_at{bg.seg}, sseg_at{bg.sseg};
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    for (unsigned i = tid; i < n; i += nthr) {
        const uint2 a = bg.ast[i];
        const bool sup = (a.y & 8u) != 0;
        const uint32_t len = sup ? cycle_leader(sseg_at, NeverStop{}, i, max_len) : cycle_leader(seg_at, StopAtSuper{}, i, max_len);
        if (len) report_border(surv, surv_count, contour_count, (int)(a.y >> 4), sup ? bg.sseg[i].minkey : bg.seg[i].minkey, len, i, sup ? 1u : 0u, g);
    }
    // start candidates: a short probe decides almost all of them (the next candidate up the edge has a smaller
    // key); the few that are still undecided are queued and walked by the first threads of the CTA together
    __shared__ unsigned s_q[256];
    __shared__ int s_nq;
    for (unsigned base = blockIdx.x * blockDim.x; base < ns; base += nthr) {
        if (threadIdx.x == 0) s_nq = 0;
        __syncthreads();
        const unsigned i = base + threadIdx.x;
        if (i < ns) {
            const uint2 c = bg.starts[i];
            const int fs = (int)(c.y >> 4), x = (int)(c.x & 0xFFFFu), y = (int)(c.x >> 16), s0 = (int)(c.y & 7u);
            MaskView rd{masks + (size_t)fs * g.mask_plane, g.PWW};
            const unsigned e0 = s_succ[rd.win9(x, y) | ((unsigned)s0 << 9)];
            const uint32_t key0 = key_of(x, y, e0, g.KS);
            const int len = direct_walk(rd, rd, s_succ, s_pred, g.KS, Rm, x, y, s0, key0, max_len < 4 ? max_len : 4);
            if (len > 0) report_border(surv, surv_count, contour_count, fs, key0, (uint32_t)len, c.x, 2u | ((unsigned)s0 << 8), g);
            else if (len < 0 && max_len > 4) s_q[atomicAdd(&s_nq, 1)] = i;
        }
        __syncthreads();
        if ((int)threadIdx.x < s_nq) {
            const uint2 c = bg.starts[s_q[threadIdx.x]];
            const int fs = (int)(c.y >> 4), x = (int)(c.x & 0xFFFFu), y = (int)(c.x >> 16), s0 = (int)(c.y & 7u);
            MaskView rd{masks + (size_t)fs * g.mask_plane, g.PWW};
            const unsigned e0 = s_succ[rd.win9(x, y) | ((unsigned)s0 << 9)];
            const uint32_t key0 = key_of(x, y, e0, g.KS);
            // (the cached window was measured slower here: 0.178 against 0.145 ms -- the two walkers' twelve independent loads per step hide more latency than the cache saves instructions)
            const int len = direct_walk(rd, rd, s_succ, s_pred, g.KS, Rm, x, y, s0, key0, max_len);
            if (len > 0) report_border(surv, surv_count, contour_count, fs, key0, (uint32_t)len, c.x, 2u | ((unsigned)s0 << 8), g);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// A3a step 3: per (frame,scale) sort the kept borders by start key, descending (= the order of
// cv2.findContours' list) and assign point offsets.  One CTA per (frame,scale).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_sort_scan(const uint4 *__restrict__ surv, int *__restrict__ surv_count, uint4 *__restrict__ sorted,
            int *__restrict__ pts_off, int *__restrict__ status, const unsigned *__restrict__ n_anchors, unsigned anchors_cap,
            const unsigned *__restrict__ n_starts, unsigned starts_cap, DetGeom g)
{
    __shared__ uint32_t s_key[SORT_CAP];
    __shared__ uint16_t s_idx[SORT_CAP];
    __shared__ int s_scan[SORT_CAP];
    const int fs = blockIdx.x;
    int n = surv_count[fs];
    if (n > g.surv_cap) { n = g.surv_cap; if (threadIdx.x == 0) status[fs / g.nScales] = 3; }
    if (threadIdx.x == 0 && (*n_anchors > anchors_cap || *n_starts > starts_cap)) status[fs / g.nScales] = 3;   // a list overflowed: borders are missing
    int N = 32; while (N < n) N <<= 1;
    const uint4 *in = surv + (size_t)fs * g.surv_cap;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        // ascending sort of (0xFFFFFFFE - key) == descending key; pads (0xFFFFFFFF) strictly last even
        // for key 0 (an outer border that starts at pixel (0,0)); keys are <= 2 * 32767 * 32768 + 1
        s_key[i] = (i < n) ? 0xFFFFFFFEu - in[i].x : 0xFFFFFFFFu;
        s_idx[i] = (uint16_t)i;
    }
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < N; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    const uint32_t a = s_key[i], b = s_key[ixj];
                    if ((a > b) == up) { s_key[i] = b; s_key[ixj] = a; const uint16_t t = s_idx[i]; s_idx[i] = s_idx[ixj]; s_idx[ixj] = t; }
                }
            }
            __syncthreads();
        }
    uint4 *out = sorted + (size_t)fs * g.surv_cap;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        int len = 0;
        if (i < n) { const uint4 e = in[s_idx[i]]; out[i] = e; len = (int)e.y; }
        s_scan[i] = len;
    }
    __syncthreads();
    for (int d = 1; d < N; d <<= 1) {                 // inclusive Hillis-Steele scan
        int v[SORT_CAP / 1024];
        int k = 0;
        for (int i = threadIdx.x; i < N; i += blockDim.x, ++k) v[k] = (i >= d) ? s_scan[i - d] : 0;
        __syncthreads();
        k = 0;
        for (int i = threadIdx.x; i < N; i += blockDim.x, ++k) s_scan[i] += v[k];
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int len = (int)out[i].y;
        const int off = s_scan[i] - len;
        const bool fits = off + len <= g.pts_cap;
        pts_off[(size_t)fs * g.surv_cap + i] = fits ? off : -1;
        if (!fits) status[fs / g.nScales] = 3;
    }
    if (threadIdx.x == 0) surv_count[fs] = n;
}

struct EmitSet {
    int2 *emit; int slot1;
    __device__ __forceinline__ void operator()(uint32_t a, int pos) const { emit[a] = make_int2(pos, slot1); }
};

// A3a step 5: the leader of every kept border tells the border's (super) anchors where their segments go;
// a border without anchors is written out directly from its first state
__global__ void __launch_bounds__(128)
k_assign(const uint32_t *__restrict__ masks, BorderGraph bg, const uint4 *__restrict__ sorted, const int *__restrict__ surv_count,
         const int *__restrict__ pts_off, uint32_t *__restrict__ pts, const WalkTables *__restrict__ tables, DetGeom g)
{
    __shared__ __align__(16) uint16_t s_succ[4096];
    const int fs = blockIdx.y;
    const int n = surv_count[fs];
    if ((int)(blockIdx.x * blockDim.x) >= n) return;
    load_walk_tables(tables, s_succ, nullptr);
    const SegLoad seg_at{bg.seg};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int off = pts_off[(size_t)fs * g.surv_cap + i];
        if (off < 0) continue;
        const uint4 e = sorted[(size_t)fs * g.surv_cap + i];
        const int slot1 = fs * g.surv_cap + i + 1, len = (int)e.y;
        const unsigned kind = e.w & 3u;
        if (kind == 1u) cycle_assign(SegLoad{bg.sseg}, EmitSet{bg.emit, slot1}, e.z, len, (int)bg.ssoff[e.z]);
        else if (kind == 0u) cycle_assign(seg_at, EmitSet{bg.emit, slot1}, e.z, len, (int)bg.minoff[e.z]);
        else {
            MaskView rd{masks + (size_t)fs * g.mask_plane, g.PWW};
            seg_emit(rd, s_succ, (int)(e.z & 0xFFFFu), (int)(e.z >> 16), (int)(e.w >> 8), len, 0, len, pts + (size_t)fs * g.pts_cap + off);
        }
    }
}

// ... and every super anchor of a kept border passes the positions on to the plain anchors behind it
__global__ void __launch_bounds__(256)
k_assign_sub(BorderGraph bg)
{
    unsigned n = *bg.n_anchors;
    if (n > bg.cap) n = bg.cap;
    const SegLoad seg_at{bg.seg};
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (!(bg.ast[i].y & 8u)) continue;
        const int2 e = bg.emit[i];
        if (e.y == 0) continue;
        super_assign(seg_at, EmitSet{bg.emit, e.y}, i, e.x);
    }
}

// A3a step 6: the points of every kept border.  The lanes of a warp first read the records of 32 anchors, then the warp takes
// the segments one at a time with one lane per point: the positions are an exclusive warp scan of the stored direction codes
// ((dx + 1) and (dy + 1) packed in one word), so the 4-byte points of a segment leave as one coalesced store per 32 points
// instead of one 32-byte sector per point and lane.  (A segment longer than its 80 stored codes is re-walked by its own lane.)
__global__ void __launch_bounds__(256)
k_emit(const uint32_t *__restrict__ masks, BorderGraph bg, const uint4 *__restrict__ sorted, const int *__restrict__ pts_off,
       uint32_t *__restrict__ pts, const WalkTables *__restrict__ tables, DetGeom g)
{
    __shared__ __align__(16) uint16_t s_succ[4096];
    __shared__ __align__(16) uint32_t s_codes[8][32][SEG_CODE_WORDS];       // the direction codes of the 32 anchors a warp is working on
    constexpr unsigned FULL = 0xFFFFFFFFu;
    unsigned n = *bg.n_anchors;
    if (n > bg.cap) n = bg.cap;
    if (blockIdx.x * blockDim.x >= n) return;
    load_walk_tables(tables, s_succ, nullptr);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // code word and shift of the points lane, lane + 32, lane + 64 of a segment
    int cw_idx[3], cw_sh[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) { const int t = 32 * r + lane; cw_idx[r] = t / 10; cw_sh[r] = 3 * (t - 10 * cw_idx[r]); }
    for (unsigned base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += gridDim.x * blockDim.x) {
        const unsigned i = base + lane;
        int pos = 0, lens = 0;                 // lens = border length | segment length << 24
        uint32_t xy = 0, obase = 0;
        bool coop = false;
        __syncwarp();                           // the previous 32 anchors' codes have been consumed
        if (i < n) {
            const int2 e = bg.emit[i];
            if (e.y != 0) {
                const uint2 a = bg.ast[i];
                const int fs = (int)(a.y >> 4);
                const int len = (int)__ldg(&sorted[e.y - 1].y), off = __ldg(&pts_off[e.y - 1]);
                const int slen = (int)(bg.seg[i].len & SEG_LEN);
                pos = e.x; xy = a.x;
                obase = (uint32_t)fs * (uint32_t)g.pts_cap + (uint32_t)off;
                coop = slen <= SEG_CODE_WORDS * 10;
                if (coop) {
                    lens = len | (slen << 24);
                    const uint4 *cp = reinterpret_cast<const uint4 *>(bg.codes + (size_t)i * SEG_CODE_WORDS);
                    uint4 *sp = reinterpret_cast<uint4 *>(&s_codes[warp][lane][0]);
                    sp[0] = __ldg(cp);
                    if (slen > 40) sp[1] = __ldg(cp + 1);
                } else {
                    MaskView rd{masks + (size_t)fs * g.mask_plane, g.PWW};
                    seg_emit(rd, s_succ, (int)(a.x & 0xFFFFu), (int)(a.x >> 16), (int)(a.y & 7u), slen, pos, len, pts + obase);
                }
            }
        }
        __syncwarp();
        unsigned todo = __ballot_sync(FULL, coop);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const int jpos = __shfl_sync(FULL, pos, j), jl = __shfl_sync(FULL, lens, j);
            const uint32_t jxy = __shfl_sync(FULL, xy, j);
            uint32_t *out = pts + __shfl_sync(FULL, obase, j);
            const int jlen = jl & 0xFFFFFF, jslen = jl >> 24;
            int cx = (int)(jxy & 0xFFFFu), cy = (int)(jxy >> 16);                          // point 32 * round of the segment
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                if (32 * r >= jslen) break;
                const int t = 32 * r + lane;
                unsigned v = 0;
                if (t < jslen) {
                    const int so = (int)((s_codes[warp][j][cw_idx[r]] >> cw_sh[r]) & 7u);
                    v = (unsigned)(dir_dx(so) + 1) | ((unsigned)(dir_dy(so) + 1) << 16);
                }
                unsigned incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += u; }
                if (t < jslen) {
                    const unsigned excl = incl - v;                                      // sums of (d + 1) over the lane earlier points of this round
                    const int px = cx + (int)(excl & 0xFFFFu) - lane, py = cy + (int)(excl >> 16) - lane;
                    const int q = jpos + t;
                    out[q < 0 ? q + jlen : q] = (uint32_t)px | ((uint32_t)py << 16);
                }
                if (32 * (r + 1) < jslen) {
                    const unsigned tot = __shfl_sync(FULL, incl, 31);
                    cx += (int)(tot & 0xFFFFu) - 32; cy += (int)(tot >> 16) - 32;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// A3b: one warp per kept border
// ---------------------------------------------------------------------------------------------
struct WarpLanes {
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int nlanes() const { return 32; }
    // largest d (d >= 0), smallest pos among equals, on every lane: three warp reductions (REDUX) instead of five shuffle rounds
    __device__ __forceinline__ void argmax_first(long long &d, int &pos) const
    {
        const unsigned hi = (unsigned)((unsigned long long)d >> 32), lo = (unsigned)d;
        const unsigned mh = __reduce_max_sync(0xFFFFFFFFu, hi);
        const unsigned ml = __reduce_max_sync(0xFFFFFFFFu, hi == mh ? lo : 0u);
        const unsigned mp = __reduce_min_sync(0xFFFFFFFFu, (hi == mh && lo == ml) ? (unsigned)pos : 0xFFFFFFFFu);
        d = (long long)(((unsigned long long)mh << 32) | ml);
        pos = (int)mp;
    }
};

// borders of at least this many points are left to k_approx_long (a whole CTA per border)
#ifndef B2A_APPROX_LONG
#define B2A_APPROX_LONG 2048
#endif
constexpr int APPROX_LONG = B2A_APPROX_LONG;
#ifndef B2A_APPROX_CTAS
#define B2A_APPROX_CTAS 16
#endif
constexpr int APPROX_LONG_CTAS = B2A_APPROX_CTAS;          // CTAs per (frame,scale) of k_approx_long

__device__ __forceinline__ void approx_store(bool ok, int len, const int *ox, const int *oy, size_t slot, uint8_t *quad_ok, int32_t *quad_xy, int32_t *quad_len)
{
    quad_ok[slot] = ok ? 1 : 0;
    quad_len[slot] = len;
    if (ok) for (int k = 0; k < 4; ++k) { quad_xy[slot * 8 + 2 * k] = ox[k]; quad_xy[slot * 8 + 2 * k + 1] = oy[k]; }
}

// the same algorithm with the lanes of a whole CTA on one (long) border: the sweeps over the points
// are the parallel part, the recursion is shared control flow
struct BlockLanes {
    long long *s_d; int *s_p;           // [warps] reduction scratch
    __device__ __forceinline__ int lane() const { return threadIdx.x; }
    __device__ __forceinline__ int nlanes() const { return blockDim.x; }
    __device__ __forceinline__ void argmax_first(long long &d, int &pos) const
    {
        WarpLanes{}.argmax_first(d, pos);
        const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        __syncthreads();                                     // previous round's scratch has been read by everyone
        if ((threadIdx.x & 31) == 0) { s_d[w] = d; s_p[w] = pos; }
        __syncthreads();
        d = s_d[0]; pos = s_p[0];
        for (int k = 1; k < nw; ++k) { const long long od = s_d[k]; const int op = s_p[k]; if (od > d || (od == d && op < pos)) { d = od; pos = op; } }
    }
};

// One launch, two roles: the first APPROX_LONG_CTAS CTAs of every (frame,scale) take the long borders
// (a whole CTA per border), the other CTAs take the rest (one warp per border), so a frame's few long
// borders are worked on while the many short ones are.
__global__ void __launch_bounds__(256)
k_approx(const uint4 *__restrict__ sorted, const int *__restrict__ surv_count, const int *__restrict__ pts_off,
         const uint32_t *__restrict__ pts, uint8_t *__restrict__ quad_ok, int32_t *__restrict__ quad_xy,
         int32_t *__restrict__ quad_len, DetGeom g)
{
    __shared__ long long s_d[8];
    __shared__ int s_p[8];
    __shared__ int s_list[SORT_CAP / APPROX_LONG_CTAS], s_n;     // every APPROX_LONG_CTAS-th border index can belong to this CTA
    const int fs = blockIdx.y;
    const int n = surv_count[fs];
    if ((int)blockIdx.x >= APPROX_LONG_CTAS) {
        const int warps_per_block = blockDim.x >> 5, nblk = (int)gridDim.x - APPROX_LONG_CTAS;
        WarpLanes lg;
        for (int i = ((int)blockIdx.x - APPROX_LONG_CTAS) * warps_per_block + (threadIdx.x >> 5); i < n; i += nblk * warps_per_block) {
            const size_t slot = (size_t)fs * g.surv_cap + i;
            const int off = pts_off[slot];
            const int len = (int)sorted[slot].y;
            if (len >= APPROX_LONG && off >= 0) continue;
            bool ok = false;
            int ox[8], oy[8];
            if (off >= 0) {
                const int m = approx_closed(lg, pts + (size_t)fs * g.pts_cap + off, len, (double)len * g.approxRate, ox, oy);
                ok = (m == 4) && quad_passes(ox, oy, m, len, g.maxWH, g.minCornerDistRate);
            }
            if (lg.lane() == 0) approx_store(ok, len, ox, oy, slot, quad_ok, quad_xy, quad_len);
        }
        return;
    }
    BlockLanes lg{s_d, s_p};
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    // this CTA's long borders (every APPROX_LONG_CTAS-th border index), found by all threads at once
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if ((i % APPROX_LONG_CTAS) != (int)blockIdx.x) continue;
        const size_t slot = (size_t)fs * g.surv_cap + i;
        if ((int)__ldg(&sorted[slot].y) >= APPROX_LONG && pts_off[slot] >= 0) s_list[atomicAdd(&s_n, 1)] = i;
    }
    __syncthreads();
    const int nlong = s_n;
    for (int k = 0; k < nlong; ++k) {
        const int i = s_list[k];
        const size_t slot = (size_t)fs * g.surv_cap + i;
        const int len = (int)sorted[slot].y, off = pts_off[slot];
        int ox[8], oy[8];
        const int m = approx_closed(lg, pts + (size_t)fs * g.pts_cap + off, len, (double)len * g.approxRate, ox, oy);
        const bool ok = (m == 4) && quad_passes(ox, oy, m, len, g.maxWH, g.minCornerDistRate);
        if (threadIdx.x == 0) approx_store(ok, len, ox, oy, slot, quad_ok, quad_xy, quad_len);
    }
}

// ---------------------------------------------------------------------------------------------
// per-frame stages
// ---------------------------------------------------------------------------------------------
struct BlockCtx {
    int *s_warp;   // [33] shared scratch
    long long *marks = nullptr;   // optional (debug): clock64 of thread 0 after every sync()
    mutable int nmark = 0;
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int nthreads() const { return blockDim.x; }
    __device__ __forceinline__ void sync() const
    {
        __syncthreads();
        if (marks && threadIdx.x == 0 && nmark < 32) marks[nmark++] = clock64();
    }
    __device__ __forceinline__ void mark() const { if (marks && threadIdx.x == 0 && nmark < 32) marks[nmark++] = -clock64(); }
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int lanes() const { return 32; }
    __device__ __forceinline__ int warp() const { return threadIdx.x >> 5; }
    __device__ __forceinline__ int warps() const { return blockDim.x >> 5; }
    __device__ __forceinline__ uint32_t ballot(bool p) const { return __ballot_sync(0xFFFFFFFFu, p); }
    __device__ __forceinline__ uint32_t warp_or(uint32_t v) const { return __reduce_or_sync(0xFFFFFFFFu, v); }
    __device__ __forceinline__ void atomic_or(uint32_t *p, uint32_t v) const { atomicOr(p, v); }
    __device__ __forceinline__ int block_sum(int v) const        // same value on every thread
    {
        v = __reduce_add_sync(0xFFFFFFFFu, v);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        __syncthreads();
        if (lane == 0) s_warp[warp] = v;
        __syncthreads();
        int tot = 0;
        for (int w = 0; w < nw; ++w) tot += s_warp[w];
        __syncthreads();
        return tot;
    }
    __device__ __forceinline__ int exclusive_scan(int flag, int &total) const
    {
        const unsigned m = __ballot_sync(0xFFFFFFFFu, flag != 0);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = 0, tot = 0;
        for (int w = 0; w < nw; ++w) { const int v = s_warp[w]; tot += v; if (w < warp) before += v; }
        __syncthreads();
        total = tot;
        return before + __popc(m & ((1u << lane) - 1u));
    }
};

struct FrameArrays {             // device base pointers; frame f uses offset f * (per-frame size)
    FrameScratch fs0;            // pointers of frame 0
    FrameOutputs fo0;
    const int32_t *surv_count;   // [B*nScales]
    const uint8_t *quad_ok;      // [B*nScales*surv_cap]
    const int32_t *quad_xy;
    const int32_t *quad_len;
    long long *marks;            // debug: per frame 32 sync timestamps of k_group (null = off)
};

__device__ __forceinline__ FrameScratch frame_scratch(const FrameScratch &b, int f, int mc)
{
    FrameScratch s = b;
    const size_t o = (size_t)f * mc;
    s.cq += o * 8; s.clen += o; s.tq += o * 8; s.tper += o; s.cent += o * 3; s.gid += o; s.sel += o;
    s.gstart += (size_t)f * (mc + 1); s.gfill += o; s.members += o; s.closeIdx += o; s.closeCnt += o;
    s.S += o; s.parent += o; s.depth += o; s.selGroup += o;
    s.closeM += o * 2 * (size_t)((mc + 31) / 32);
    s.wq += o * 8; s.wres += o; s.closeStart += o; s.closeNum += o;
    s.counters += (size_t)f * 8;
    if (s.tlen) { s.tlen += o; s.wlen += o; }
    return s;
}

// A4-A5 in three launches: k_group_a (one CTA per frame: candidates in T order), k_close (many CTAs per frame:
// the O(n^2) closeness matrix), k_group (one CTA per frame: grouping, selection, hierarchy, work list).
// dynamic shared memory of k_group_a / k_group: 8 per-candidate arrays (gid, sel, gstart, gfill, members,
// closeIdx, closeCnt, tper), then (k_group) the T-order quads, the two bit matrices and the selected-candidate arrays
__device__ __forceinline__ void group_setup(const FrameArrays &fa, const FrameParams &fp, int f, uint32_t *s_dyn, FrameScratch &fs, ScaleQuads &sq)
{
    fs = frame_scratch(fa.fs0, f, fp.max_cand);
    const int mc1 = fp.max_cand + 1;
    int32_t *base = reinterpret_cast<int32_t *>(s_dyn);
    fs.gid = base; fs.sel = base + mc1; fs.gstart = base + 2 * mc1; fs.gfill = base + 3 * mc1;
    fs.members = base + 4 * mc1; fs.closeIdx = base + 5 * mc1; fs.closeCnt = base + 6 * mc1;
    fs.tper = reinterpret_cast<float *>(base + 7 * mc1);
    sq.count = fa.surv_count + (size_t)f * fp.nScales;
    sq.quad_ok = fa.quad_ok + (size_t)f * fp.nScales * fp.surv_cap;
    sq.quad_xy = fa.quad_xy + (size_t)f * fp.nScales * fp.surv_cap * 8;
    sq.len = fa.quad_len + (size_t)f * fp.nScales * fp.surv_cap;
}

__global__ void __launch_bounds__(1024)
k_group_a(FrameArrays fa, FrameParams fp)
{
    extern __shared__ uint32_t s_dyn[];
    __shared__ int s_warp[33];
    BlockCtx ctx{s_warp};
    FrameScratch fs; ScaleQuads sq;
    group_setup(fa, fp, blockIdx.x, s_dyn, fs, sq);
    frame_group(ctx, fp, sq, fs, nullptr, 0, 1);
}

constexpr int CLOSE_CTAS = 16;
__global__ void __launch_bounds__(256)
k_close(FrameArrays fa, FrameParams fp)
{
    __shared__ int s_warp[33];
    BlockCtx ctx{s_warp};
    const FrameScratch fs = frame_scratch(fa.fs0, blockIdx.y, fp.max_cand);
    const int n = fs.counters[FC_NCAND], wpr = (n + 31) >> 5;
    const int nwarps = (int)gridDim.x * (blockDim.x >> 5);
    for (int t = (int)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n * wpr; t += nwarps)
        close_task(ctx, fp, fs.tq, fs.cent, n, wpr, t, fs.closeM, fs.closeM + (size_t)n * wpr);
}

__global__ void __launch_bounds__(1024)
k_group(FrameArrays fa, FrameParams fp, int smem_words)
{
    extern __shared__ uint32_t s_dyn[];
    __shared__ int s_warp[33];
    const int f = blockIdx.x;
    BlockCtx ctx{s_warp};
    if (fa.marks) { ctx.marks = fa.marks + (size_t)f * 32; if (threadIdx.x == 0) ctx.marks[ctx.nmark++] = clock64(); }
    FrameScratch fs; ScaleQuads sq;
    group_setup(fa, fp, f, s_dyn, fs, sq);
    const int mc1 = fp.max_cand + 1;
    frame_group(ctx, fp, sq, fs, s_dyn + 8 * mc1, smem_words - 8 * mc1, 2);
}

// dynamic shared memory of k_finalize: copies of depth, parent, closeStart, closeNum (n_sel entries),
// wres (n_work entries) and the valid / was / chosen scratch, so the single-lane A6 loop runs on-chip
__global__ void __launch_bounds__(128)
k_finalize(FrameArrays fa, FrameParams fp)
{
    extern __shared__ uint32_t s_dyn[];
    __shared__ int s_warp[33];
    const int f = blockIdx.x;
    BlockCtx ctx{s_warp};
    FrameScratch fs = frame_scratch(fa.fs0, f, fp.max_cand);
    FrameOutputs fo = fa.fo0;
    fo.n_accepted += f; fo.n_rejected += f; fo.status += f;
    fo.corners += (size_t)f * fp.max_markers * 8; fo.ids += (size_t)f * fp.max_markers; fo.rejected += (size_t)f * fp.max_markers * 8;
    const int mc = fp.max_cand;
    int32_t *base = reinterpret_cast<int32_t *>(s_dyn);
    const int nS = fs.counters[FC_NSEL], nW = fs.counters[FC_NWORK];
    for (int i = threadIdx.x; i < nS; i += blockDim.x) {
        base[i] = fs.depth[i]; base[mc + i] = fs.parent[i]; base[2 * mc + i] = fs.closeStart[i]; base[3 * mc + i] = fs.closeNum[i];
    }
    for (int i = threadIdx.x; i < nW; i += blockDim.x) base[4 * mc + i] = fs.wres[i];
    __syncthreads();
    fs.depth = base; fs.parent = base + mc; fs.closeStart = base + 2 * mc; fs.closeNum = base + 3 * mc; fs.wres = base + 4 * mc;
    fs.gid = base + 5 * mc; fs.sel = base + 6 * mc; fs.gfill = base + 7 * mc;
    frame_finalize(ctx, fp, fs, fo);
}

// ---------------------------------------------------------------------------------------------
// A7: one CTA per identification work item
// ---------------------------------------------------------------------------------------------
struct IdentParams {
    int markerSize, borderBits, cellSize, cellMargin;
    int nMarkers, maxCorr, maxBorderErr, detectInverted;
    double minOtsuStdDev;
    int W, H;
    size_t pitch, frame_stride;
    int max_cand;
    long long *marks;            // debug: clock64 after each phase of work item 0 of frame 0 (null = off)
    unsigned long long *codes;   // optional [frames][max_cand]: every work item's inner bits as extracted (before the inverted-marker choice)
    PyrLevels pyr;               // ArUco3: the image pyramid a work item picks its level from (n = 0: off)
};

// A7 step 1: the inverse perspective map of every work item, one thread each (cv2's 8x8 LU lives in
// ~150 registers; keeping it out of k_identify lets that kernel run many more warps per SM)
__global__ void __launch_bounds__(64)
k_homography(FrameArrays fa, double *__restrict__ wM, int S, int max_cand, PyrLevels pyr)
{
    const int f = blockIdx.y;
    const int nw = fa.fs0.counters[(size_t)f * 8 + FC_NWORK];
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < nw; w += gridDim.x * blockDim.x) {
        double M[9];
        const float *q = fa.fs0.wq + ((size_t)f * max_cand + w) * 8;
        float sq[8];
        if (pyr.n > 0) {                 // ArUco3: the quad in the coordinates of the pyramid level it is read in
            const int lvl = pyr_opt_level(pyr.W, pyr.n, pyr.segW, fa.fs0.wlen[(size_t)f * max_cand + w], pyr.minPerimeter);
            pyr_scale_quad(q, f_div((float)pyr.W[lvl], (float)pyr.segW), sq);
            q = sq;
        }
        perspective_inverse(q, S, M);
#pragma unroll
        for (int i = 0; i < 9; ++i) wM[((size_t)f * max_cand + w) * 9 + i] = M[i];
    }
}

constexpr int ID_WARPS = 4;         // work items in flight per CTA (one warp each)
constexpr int ID_THREADS = ID_WARPS * 32;
constexpr int ID_MAX_S = 9 * 8;     // (7 + 2) cells * up to 8 px
inline size_t identify_smem_bytes(int S) { return (size_t)ID_WARPS * (3 * 256 * sizeof(double) + (size_t)((S * S + 15) & ~15) + 96); }

// A7 step 2.  One WARP per work item: the sequential pieces (the Otsu recurrence, the border
// check) run on lane 0 while the other warps of the SM work on other candidates; sampling, the
// between-class variances, the cell votes and the dictionary scan are spread over the 32 lanes.
// PYR (ArUco3): every work item reads the pyramid level its contour length picks (pyr_opt_level) instead of `gray`
template <bool PYR>
__global__ void __launch_bounds__(ID_THREADS, 7)
k_identify(const uint8_t *__restrict__ gray, const unsigned long long *__restrict__ dict, const double *__restrict__ wM, FrameArrays fa, IdentParams ip)
{
    // dynamic shared memory, per warp: q1 / mu1 / y [256] f64 (the histogram [256] i32 lives in y until Otsu starts),
    // patch [S*S rounded up to 16] u8, bits [96] u8
    extern __shared__ __align__(16) uint8_t id_smem[];
    const int patch_bytes = (((ip.markerSize + 2 * ip.borderBits) * ip.cellSize) * ((ip.markerSize + 2 * ip.borderBits) * ip.cellSize) + 15) & ~15;
    const size_t per_warp = 3 * 256 * sizeof(double) + (size_t)patch_bytes + 96;
    uint8_t *wbase = id_smem + (size_t)(threadIdx.x >> 5) * per_warp;
    double *const q1w = reinterpret_cast<double *>(wbase), *const mu1w = q1w + 256, *const yw = mu1w + 256;
    int *const histw = reinterpret_cast<int *>(yw);
    uint8_t *const patchw = reinterpret_cast<uint8_t *>(yw + 256), *const bitsw = patchw + patch_bytes;
    const int f = blockIdx.y;
    const int *counters = fa.fs0.counters + (size_t)f * 8;
    const int nw = counters[FC_NWORK];
    const int nb = ip.markerSize + 2 * ip.borderBits;
    const int S = nb * ip.cellSize;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const uint8_t *img = gray + (size_t)f * ip.frame_stride;
    int iW = ip.W, iH = ip.H;
    size_t ipitch = ip.pitch;
    const unsigned FULL = 0xFFFFFFFFu;
    for (int w = blockIdx.x * ID_WARPS + wp; w < nw; w += gridDim.x * ID_WARPS) {
        for (int i = lane; i < 256; i += 32) histw[i] = 0;
        __syncwarp();
        if (PYR) {
            const int lvl = pyr_opt_level(ip.pyr.W, ip.pyr.n, ip.pyr.segW, fa.fs0.wlen[(size_t)f * ip.max_cand + w], ip.pyr.minPerimeter);
            img = ip.pyr.base[lvl] + (size_t)f * ip.pyr.frame_stride[lvl];
            iW = ip.pyr.W[lvl]; iH = ip.pyr.H[lvl]; ipitch = ip.pyr.pitch[lvl];
        }
        double M[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) M[i] = __ldg(wM + ((size_t)f * ip.max_cand + w) * 9 + i);
        const int m0 = ip.cellSize / 2;
        long long *mk = (ip.marks && f == 0 && w == 0 && lane == 0) ? ip.marks : nullptr;
        int nmk = 0;
        long long t_start = 0;
        if (ip.marks && lane == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start)); }
#define ID_MARK() do { if (mk) mk[nmk++] = clock64(); } while (0)
        ID_MARK();
        int ls = 0, lq = 0;
        // sampling: no dependence between iterations, so the (uncached) gathers of several pixels are in flight together
        // 4 pixels per lane at a time: four independent FP64 coordinate chains, then four gathers in flight
        for (int base = 0; base < S * S; base += 128) {
            long long off[4];
            int pp[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                pp[k] = base + k * 32 + lane;
                const int y = pp[k] / S, x = pp[k] - y * S;
                off[k] = pp[k] < S * S ? warp_source(iW, iH, ipitch, M, x, y) : -1;
            }
            unsigned v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = off[k] >= 0 ? (unsigned)__ldg(img + off[k]) : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (pp[k] >= S * S) continue;
                const int y = pp[k] / S, x = pp[k] - y * S;
                patchw[pp[k]] = (uint8_t)v[k];
                if (x >= m0 && x < S - m0 && y >= m0 && y < S - m0) { ls += (int)v[k]; lq += (int)(v[k] * v[k]); }
            }
        }
        __syncwarp();
        ID_MARK();
        // warp-aggregated histogram: a marker patch has two dominant levels, so plain shared atomics serialise
        for (int base = 0; base < S * S; base += 32) {
            const int p = base + lane;
            const unsigned v = (p < S * S) ? patchw[p] : 256u;            // sentinel for the lanes past the patch
            const unsigned peers = __match_any_sync(FULL, v);
            if (v < 256u && lane == __ffs(peers) - 1) histw[v] += __popc(peers);
            __syncwarp();
        }
        ID_MARK();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { ls += __shfl_xor_sync(FULL, ls, o); lq += __shfl_xor_sync(FULL, lq, o); }
        __syncwarp();
        const int mode = ident_mode(ls, lq, S, m0, ip.minOtsuStdDev);          // same value on every lane
        int thr = 0;
        if (mode == 2) {
            // histogram range and first moment on all lanes (8 consecutive bins each), chain on lane 0
            int isum = 0, nzmask = 0, hv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) hv[k] = histw[lane * 8 + k];
            __syncwarp();                                                       // the histogram shares its storage with y
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = lane * 8 + k;
                isum += i * hv[k]; nzmask |= (hv[k] != 0) << k;
                double p_i, ip_i;
                otsu_bin_inputs(i, hv[k], S * S, p_i, ip_i);
                q1w[i] = p_i; mu1w[i] = ip_i; yw[i] = -1.0;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) isum += __shfl_xor_sync(FULL, isum, o);
            const unsigned nzl = __ballot_sync(FULL, nzmask != 0);              // lanes that own a non-empty bin (never 0 here)
            const int l_lo = __ffs(nzl) - 1, l_hi = 31 - __clz(nzl);
            const int lo = l_lo * 8 + __ffs(__shfl_sync(FULL, nzmask, l_lo)) - 1;
            const int hi = l_hi * 8 + 31 - __clz(__shfl_sync(FULL, nzmask, l_hi));
            const double mu = otsu_mu(isum, S * S);
            __syncwarp();
            ID_MARK();
            if (((S * S) & (S * S - 1)) == 0) {
                // the patch has a power-of-two number of pixels (S = 32 for the 6x6 dictionaries): 1 / n and every p_i = h_i / n are
                // exact, so are their running sums, and the one-lane pass below equals the integer prefix scaled once
                int run = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) run += hv[k];
                int incl = run;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
                int acc = incl - run;
                const double scale = d_div(1.0, (double)(S * S));
#pragma unroll
                for (int k = 0; k < 8; ++k) { acc += hv[k]; q1w[lane * 8 + k] = d_mul((double)acc, scale); }
            } else if (lane == 0) otsu_prefix(lo, hi, q1w);                // running sums q1 (one add per bin)
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k) { const int i = lane * 8 + k; if (i >= lo && i <= hi) yw[i] = otsu_bin(q1w[i]); }
            __syncwarp();
            if (lane == 0) otsu_chain(lo, hi, q1w, yw, mu1w);    // the mu1 recurrence, division-free
            __syncwarp();
            ID_MARK();
            double best = 0; int bi = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {                                        // lane owns 8 consecutive bins: order kept
                const int i = lane * 8 + k;
                const double sg = otsu_sigma(mu, q1w[i], yw[i], mu1w[i]);
                if (sg > best) { best = sg; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(FULL, best, o);
                const int oi = __shfl_xor_sync(FULL, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            thr = best > 0 ? bi : 0;
        }
        ID_MARK();
        unsigned long long code = 0;
        int berr = 0;
        for (int cidx = lane; cidx < nb * nb; cidx += 32) {
            const int cy = cidx / nb, cx = cidx - cy * nb;
            const unsigned bit = (unsigned)((mode < 2) ? mode : ident_cell_bit(patchw, S, ip.cellSize, ip.cellMargin, cy, cx, thr));
            ident_cell_accumulate(cidx, bit, ip.markerSize, ip.borderBits, berr, code);
        }
        berr = __reduce_add_sync(FULL, berr);
        code = ((unsigned long long)__reduce_or_sync(FULL, (unsigned)(code >> 32)) << 32) | __reduce_or_sync(FULL, (unsigned)code);
        if (ip.codes && lane == 0) ip.codes[(size_t)f * ip.max_cand + w] = code;
        if (ip.detectInverted) ident_choose_inverted(ip.markerSize, ip.borderBits, berr, code);
        const int ok = berr <= ip.maxBorderErr;
        int best_m = 0x7FFFFFFF;
        if (ok) {
            for (int m = lane; m < ip.nMarkers; m += 32) {
                int rot;
                if (ident_marker_distance(dict + (size_t)m * 4, code, ip.markerSize, rot) <= ip.maxCorr) { best_m = m; break; }   // later m of this lane are larger
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best_m = min(best_m, __shfl_xor_sync(FULL, best_m, o));
        }
        if (lane == 0) {
            int res = 0;
            if (ok && best_m != 0x7FFFFFFF) {
                int rot;
                ident_marker_distance(dict + (size_t)best_m * 4, code, ip.markerSize, rot);
                res = (int)(0x80000000u | ((unsigned)best_m << 8) | (unsigned)rot);
            }
            fa.fs0.wres[(size_t)f * ip.max_cand + w] = res;
        }
        ID_MARK();
        if (ip.marks && lane == 0) {
            long long t_end; unsigned smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            long long *rec = ip.marks + 16 + ((size_t)f * 128 + (w & 127)) * 3;
            rec[0] = t_start; rec[1] = t_end; rec[2] = smid;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// A8: cornerSubPix on the accepted corners (one thread per corner)
// ---------------------------------------------------------------------------------------------
struct SubpixParams {
    int W, H; size_t pitch, frame_stride;
    int max_markers, markerSize, borderBits, maxWin, maxIter;
    double relWin, eps;
    // ArUco3's refinement up the pyramid (findCornerInPyrImage): the corner is first multiplied by mul0, then by mul1 (two float
    // products, as cv2 scales to the closest level and then doubles per level), the window is fixedWin (0 = the relative rule of
    // CORNER_REFINE_SUBPIX) and refine = 0 only scales.  With fixedWin a thread touches its own corner only: in place is fine.
    int fixedWin, refine;
    float mul0, mul1;
};

__device__ __forceinline__ float subpix_sample(const uint8_t *img, int W, int H, size_t pitch, int ipx, int ipy, float a11, float a12, float a21, float a22, int x, int y)
{
    int x0 = ipx + x, x1 = x0 + 1, y0 = ipy + y, y1 = y0 + 1;
    x0 = x0 < 0 ? 0 : (x0 >= W ? W - 1 : x0); x1 = x1 < 0 ? 0 : (x1 >= W ? W - 1 : x1);
    y0 = y0 < 0 ? 0 : (y0 >= H ? H - 1 : y0); y1 = y1 < 0 ? 0 : (y1 >= H ? H - 1 : y1);
    return f_add(f_add(f_add(f_mul((float)img[y0 * pitch + x0], a11), f_mul((float)img[y0 * pitch + x1], a12)),
                       f_mul((float)img[y1 * pitch + x0], a21)), f_mul((float)img[y1 * pitch + x1], a22));
}

__global__ void k_subpix(const uint8_t *__restrict__ gray, const int32_t *__restrict__ n_acc, const float *__restrict__ corners,
                         float *__restrict__ corners_out, int B, SubpixParams sp)
{
    const int total = B * sp.max_markers * 4;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int f = t / (sp.max_markers * 4), r = t - f * sp.max_markers * 4, mk = r >> 2;
        if (mk >= n_acc[f]) continue;
        int win = sp.fixedWin;
        if (win == 0) {
            const float *c = corners + ((size_t)f * sp.max_markers + mk) * 8;
            const float per = quad_perimeter(c);
            const int nm = sp.markerSize + 2 * sp.borderBits;
            win = (int)lroundf((float)sp.relWin * f_div(per, 4.f * (float)nm));
            win = win < 1 ? 1 : (win > sp.maxWin ? sp.maxWin : win);
        }
        const uint8_t *img = gray + (size_t)f * sp.frame_stride;
        const float *pin = corners + ((size_t)f * sp.max_markers + mk) * 8 + (r & 3) * 2;
        float *pt = corners_out + ((size_t)f * sp.max_markers + mk) * 8 + (r & 3) * 2;
        const float cTx = f_mul(f_mul(pin[0], sp.mul0), sp.mul1), cTy = f_mul(f_mul(pin[1], sp.mul0), sp.mul1);
        if (!sp.refine) { pt[0] = cTx; pt[1] = cTy; continue; }
        float cIx = cTx, cIy = cTy;
        const int ww = 2 * win + 1;
        const double eps2 = sp.eps * sp.eps;
        int iter = 0; double err = 0;
        do {
            const float fx = cIx - (float)(ww + 1) * 0.5f, fy = cIy - (float)(ww + 1) * 0.5f;   // centre - (ww+2-1)/2
            const int ipx = (int)floorf(fx), ipy = (int)floorf(fy);
            const float a = fx - (float)ipx, b = fy - (float)ipy;
            const float a11 = f_mul(1.f - a, 1.f - b), a12 = f_mul(a, 1.f - b), a21 = f_mul(1.f - a, b), a22 = f_mul(a, b);
            double A = 0, Bm = 0, C = 0, bb1 = 0, bb2 = 0;
            for (int i = 0; i < ww; ++i) {
                const float yy = (float)(i - win) / (float)win;
                const float my = (float)exp(-(double)(yy * yy));
                const double py = i - win;
                for (int j = 0; j < ww; ++j) {
                    const float xx = (float)(j - win) / (float)win;
                    const float mx = (float)exp(-(double)(xx * xx));
                    const double m = (double)f_mul(my, mx);
                    // subimage index (j+1, i+1) in the (ww+2)^2 patch
                    const double tgx = (double)subpix_sample(img, sp.W, sp.H, sp.pitch, ipx, ipy, a11, a12, a21, a22, j + 2, i + 1)
                                     - (double)subpix_sample(img, sp.W, sp.H, sp.pitch, ipx, ipy, a11, a12, a21, a22, j, i + 1);
                    const double tgy = (double)subpix_sample(img, sp.W, sp.H, sp.pitch, ipx, ipy, a11, a12, a21, a22, j + 1, i + 2)
                                     - (double)subpix_sample(img, sp.W, sp.H, sp.pitch, ipx, ipy, a11, a12, a21, a22, j + 1, i);
                    const double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
                    const double px = j - win;
                    A += gxx; Bm += gxy; C += gyy;
                    bb1 += gxx * px + gxy * py;
                    bb2 += gxy * px + gyy * py;
                }
            }
            const double det = A * C - Bm * Bm;
            if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
            const double scale = 1.0 / det;
            const float nx = (float)((double)cIx + (C * scale * bb1 - Bm * scale * bb2));
            const float ny = (float)((double)cIy + (-Bm * scale * bb1 + A * scale * bb2));
            err = (double)((nx - cIx) * (nx - cIx) + (ny - cIy) * (ny - cIy));
            cIx = nx; cIy = ny;
            if (cIx < 0 || cIx >= (float)sp.W || cIy < 0 || cIy >= (float)sp.H) break;
        } while (++iter < sp.maxIter && err > eps2);
        if (fabsf(cIx - cTx) > (float)win || fabsf(cIy - cTy) > (float)win) { cIx = cTx; cIy = cTy; }
        pt[0] = cIx; pt[1] = cIy;
    }
}

// The same refinement with one WARP per corner (the form the pipeline launches; k_subpix stays for windows too large for shared
// memory): the lanes build the bilinearly interpolated (ww + 2)^2 patch around the corner once per iteration in shared memory
// (cv2's getRectSubPix; the thread form recomputes every patch value up to four times), split the ww^2 gradient terms -- each term
// computed exactly as above -- and add their partial sums with a butterfly, so every lane holds the same normal equations and takes
// the same step; the Gaussian window weights are computed once per corner.  Only the ORDER of the double-precision sums differs
// from the sequential loop (results agree to the last bits of the float corner).
__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
inline size_t subpix_warp_smem_bytes(int wmax) { return (size_t)4 * ((size_t)(2 * wmax + 3) * (2 * wmax + 3) + 2 * wmax + 1) * sizeof(float); }

__global__ void __launch_bounds__(128)
k_subpix_warp(const uint8_t *__restrict__ gray, const int32_t *__restrict__ n_acc, const float *__restrict__ corners,
              float *__restrict__ corners_out, int B, SubpixParams sp, int wmax)
{
    extern __shared__ float spw_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int pw_max = 2 * wmax + 3;
    float *patch = spw_smem + (size_t)wib * (pw_max * pw_max + 2 * wmax + 1);
    float *mask = patch + pw_max * pw_max;
    const int total = B * sp.max_markers * 4;
    for (int t = blockIdx.x * 4 + wib; t < total; t += gridDim.x * 4) {
        const int f = t / (sp.max_markers * 4), r = t - f * sp.max_markers * 4, mk = r >> 2;
        if (mk >= n_acc[f]) continue;
        int win = sp.fixedWin;
        if (win == 0) {
            const float *c = corners + ((size_t)f * sp.max_markers + mk) * 8;
            const float per = quad_perimeter(c);
            const int nm = sp.markerSize + 2 * sp.borderBits;
            win = (int)lroundf((float)sp.relWin * f_div(per, 4.f * (float)nm));
            win = win < 1 ? 1 : (win > sp.maxWin ? sp.maxWin : win);
        }
        if (win > wmax) win = wmax;                                       // cannot happen: wmax is the largest window of the launch
        const uint8_t *img = gray + (size_t)f * sp.frame_stride;
        const float *pin = corners + ((size_t)f * sp.max_markers + mk) * 8 + (r & 3) * 2;
        float *pt = corners_out + ((size_t)f * sp.max_markers + mk) * 8 + (r & 3) * 2;
        const float cTx = f_mul(f_mul(pin[0], sp.mul0), sp.mul1), cTy = f_mul(f_mul(pin[1], sp.mul0), sp.mul1);
        if (!sp.refine) { if (lane == 0) { pt[0] = cTx; pt[1] = cTy; } continue; }
        float cIx = cTx, cIy = cTy;
        const int ww = 2 * win + 1, pw = ww + 2;
        __syncwarp();
        for (int i = lane; i < ww; i += 32) {
            const float yy = (float)(i - win) / (float)win;
            mask[i] = (float)exp(-(double)(yy * yy));
        }
        const double eps2 = sp.eps * sp.eps;
        int iter = 0; double err = 0;
        do {
            const float fx = cIx - (float)(ww + 1) * 0.5f, fy = cIy - (float)(ww + 1) * 0.5f;
            const int ipx = (int)floorf(fx), ipy = (int)floorf(fy);
            const float a = fx - (float)ipx, b = fy - (float)ipy;
            const float a11 = f_mul(1.f - a, 1.f - b), a12 = f_mul(a, 1.f - b), a21 = f_mul(1.f - a, b), a22 = f_mul(a, b);
            __syncwarp();
            for (int p = lane; p < pw * pw; p += 32) {
                const int y = p / pw, x = p - y * pw;
                patch[p] = subpix_sample(img, sp.W, sp.H, sp.pitch, ipx, ipy, a11, a12, a21, a22, x, y);
            }
            __syncwarp();
            double A = 0, Bm = 0, C = 0, bb1 = 0, bb2 = 0;
            for (int p = lane; p < ww * ww; p += 32) {
                const int i = p / ww, j = p - i * ww;
                const double m = (double)f_mul(mask[i], mask[j]);
                const double tgx = (double)patch[(i + 1) * pw + j + 2] - (double)patch[(i + 1) * pw + j];
                const double tgy = (double)patch[(i + 2) * pw + j + 1] - (double)patch[i * pw + j + 1];
                const double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
                const double px = j - win, py = i - win;
                A += gxx; Bm += gxy; C += gyy;
                bb1 += gxx * px + gxy * py;
                bb2 += gxy * px + gyy * py;
            }
            A = warp_sum_f64(A); Bm = warp_sum_f64(Bm); C = warp_sum_f64(C); bb1 = warp_sum_f64(bb1); bb2 = warp_sum_f64(bb2);
            const double det = A * C - Bm * Bm;
            if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
            const double scale = 1.0 / det;
            const float nx = (float)((double)cIx + (C * scale * bb1 - Bm * scale * bb2));
            const float ny = (float)((double)cIy + (-Bm * scale * bb1 + A * scale * bb2));
            err = (double)((nx - cIx) * (nx - cIx) + (ny - cIy) * (ny - cIy));
            cIx = nx; cIy = ny;
            if (cIx < 0 || cIx >= (float)sp.W || cIy < 0 || cIy >= (float)sp.H) break;
        } while (++iter < sp.maxIter && err > eps2);
        if (fabsf(cIx - cTx) > (float)win || fabsf(cIy - cTy) > (float)win) { cIx = cTx; cIy = cTy; }
        if (lane == 0) { pt[0] = cIx; pt[1] = cIy; }
    }
}

// ---------------------------------------------------------------------------------------------
// A8 (optional): CORNER_REFINE_CONTOUR, one warp per accepted marker (refine_core.h).  The marker's contour is still in the
// point arrays of the contour stage; the warp finds it by its quad, sums the four sides and lane 0 writes the crossings.
// ---------------------------------------------------------------------------------------------
struct RefineLanes {
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int nlanes() const { return 32; }
    __device__ __forceinline__ uint32_t ballot(bool p) const { return __ballot_sync(0xFFFFFFFFu, p); }
    __device__ __forceinline__ int bcast(int v, int l) const { return __shfl_sync(0xFFFFFFFFu, v, l); }
    __device__ __forceinline__ long long sum(long long v) const
    {
#pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
        return v;
    }
    __device__ __forceinline__ int imin(int v) const { return __reduce_min_sync(0xFFFFFFFFu, v); }
    __device__ __forceinline__ int imax(int v) const { return __reduce_max_sync(0xFFFFFFFFu, v); }
};

__global__ void __launch_bounds__(128)
k_refine_contour(const float *__restrict__ corners, const int32_t *__restrict__ n_acc, float *__restrict__ out,
                 const int *__restrict__ surv_count, const int *__restrict__ pts_off, const uint32_t *__restrict__ pts,
                 const uint8_t *__restrict__ quad_ok, const int32_t *__restrict__ quad_xy, const int32_t *__restrict__ quad_len,
                 DetGeom g, int nb, int max_markers)
{
    const int w = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int f = w / max_markers, m = w - f * max_markers;
    if (f >= nb || m >= n_acc[f]) return;
    const RefineLanes lg;
    const float *cin = corners + ((size_t)f * max_markers + m) * 8;
    float *co = out + ((size_t)f * max_markers + m) * 8;
    const size_t fs0 = (size_t)f * g.nScales;
    const int hit = refine_find_border(lg, surv_count + fs0, quad_ok + fs0 * g.surv_cap, quad_xy + fs0 * g.surv_cap * 8, g.nScales, g.surv_cap, cin);
    float r[8];
    bool done = false;
    if (hit >= 0) {
        const size_t slot = fs0 * g.surv_cap + hit, fs = fs0 + hit / g.surv_cap;
        const int off = pts_off[slot];
        if (off >= 0) { refine_marker_lines(lg, pts + fs * (size_t)g.pts_cap + off, quad_len[slot], cin, r); done = true; }
    }
    if (lg.lane() == 0)
        for (int k = 0; k < 8; ++k) co[k] = done ? r[k] : cin[k];
}

// ---------------------------------------------------------------------------------------------
// ArUco3 (useAruco3Detection): the image pyramid (cv::buildPyramid = repeated cv::pyrDown) and the reduced segmentation image
// (cv::resize, INTER_LINEAR) of every frame of a sub-batch; one thread per destination pixel (pyr_core.h)
// ---------------------------------------------------------------------------------------------
// a thread makes four neighbouring destination pixels (one word store when all four exist).  Source window inside the image: word
// loads + dp4a when the rows are word aligned, byte loads otherwise; window across the border: reflected indices (images of at
// least 3 x 3 pixels), pixel by pixel below that
__device__ __forceinline__ void pyr_down_group(const uint8_t *sb, int W, int H, size_t spitch, bool aligned, uint8_t *drow, int dW, int x0, int y)
{
    uint32_t word;
    if (pyr_down_is_interior4(W, H, dW, x0, y)) {
        const bool words = aligned && 2 * x0 - 4 >= 0 && 2 * x0 + 11 < W;
        word = words ? pyr_down_interior4<true>(sb, spitch, x0, y) : pyr_down_interior4<false>(sb, spitch, x0, y);
    } else if (W >= 3 && H >= 3) word = pyr_down_border4(sb, W, H, spitch, dW, x0, y);
    else {
        word = 0;
        for (int q = 0; q < 4 && x0 + q < dW; ++q) word |= (uint32_t)pyr_down_pixel(sb, W, H, spitch, x0 + q, y) << (8 * q);
    }
    if (x0 + 3 < dW) *reinterpret_cast<uint32_t *>(drow + x0) = word;
    else for (int q = 0; x0 + q < dW; ++q) drow[x0 + q] = (uint8_t)(word >> (8 * q));
}

// one level of every frame: block = 64 groups x 4 rows, blockIdx.x walks the rows of all frames (row index = frame * dH + y).
// Rows whose 5 source rows lie inside the image: the interior groups by all threads (no divergence: every lane takes the same
// path), then the few border groups of the block's four rows together on the lanes of one warp -- with the border groups inside
// the row loop a quarter of all warps ran both paths (profiles/r2_aruco3_ncu.md).  Top and bottom rows: the general path.
__global__ void __launch_bounds__(256)
k_pyr_down(const uint8_t *__restrict__ src, int W, int H, size_t spitch, size_t sframe, uint8_t *__restrict__ dst, size_t dpitch, size_t dframe, int nb)
{
    const int dW = (W + 1) / 2, dH = (H + 1) / 2, gW = (dW + 3) / 4;
    const bool aligned = ((spitch | sframe | (size_t)src) & 3) == 0;
    const int gR = pyr_interior_gx_end(W, dW), yB = pyr_interior_y_end(H), nbm = 1 + (gW - gR);
    const unsigned rows = (unsigned)nb * (unsigned)dH;
    for (unsigned R0 = blockIdx.x * 4u; R0 < rows; R0 += gridDim.x * 4u) {
        const unsigned R = R0 + threadIdx.y;
        if (R < rows) {
            const unsigned b = R / (unsigned)dH;
            const int y = (int)(R - b * (unsigned)dH);
            const uint8_t *sb = src + (size_t)b * sframe;
            uint8_t *drow = dst + (size_t)b * dframe + (size_t)y * dpitch;
            if (y >= 1 && y < yB) {
                for (int gx = 1 + threadIdx.x; gx < gR; gx += 64) {
                    const int x0 = 4 * gx;
                    const bool words = aligned && 2 * x0 + 11 < W;
                    *reinterpret_cast<uint32_t *>(drow + x0) = words ? pyr_down_interior4<true>(sb, spitch, x0, y) : pyr_down_interior4<false>(sb, spitch, x0, y);
                }
            } else for (int gx = threadIdx.x; gx < gW; gx += 64) pyr_down_group(sb, W, H, spitch, aligned, drow, dW, 4 * gx, y);
        }
        if (threadIdx.y == 0 && threadIdx.x < 32)
            for (int i = threadIdx.x; i < 4 * nbm; i += 32) {
                const int q = i / nbm, j = i - q * nbm;
                const unsigned Rq = R0 + (unsigned)q;
                if (Rq >= rows) continue;
                const unsigned b = Rq / (unsigned)dH;
                const int y = (int)(Rq - b * (unsigned)dH);
                if (!(y >= 1 && y < yB)) continue;                       // that row went through the general path above
                const int gx = j == 0 ? 0 : gR + j - 1;
                pyr_down_group(src + (size_t)b * sframe, W, H, spitch, aligned, dst + (size_t)b * dframe + (size_t)y * dpitch, dW, 4 * gx, y);
            }
    }
}

// levels first .. n - 1 of one frame per CTA, one after the other (a level depends on the whole level before it, and the last
// levels of a frame are a few thousand pixels: launches and their gaps would cost more than the work)
__global__ void __launch_bounds__(1024)
k_pyr_chain(PyrLevels pl, int first)
{
    const int b = blockIdx.x;
    for (int l = first; l < pl.n; ++l) {
        const uint8_t *sb = pl.base[l - 1] + (size_t)b * pl.frame_stride[l - 1];
        uint8_t *db = const_cast<uint8_t *>(pl.base[l]) + (size_t)b * pl.frame_stride[l];
        const int W = pl.W[l - 1], H = pl.H[l - 1], dW = pl.W[l], dH = pl.H[l], gW = (dW + 3) / 4;
        const size_t spitch = pl.pitch[l - 1], dpitch = pl.pitch[l];
        const bool aligned = ((spitch | (size_t)sb) & 3) == 0;
        for (int t = threadIdx.x; t < gW * dH; t += blockDim.x) {
            const int y = t / gW, gx = t - y * gW;
            pyr_down_group(sb, W, H, spitch, aligned, db + (size_t)y * dpitch, dW, 4 * gx, y);
        }
        __syncthreads();                 // the level is complete (and visible to this CTA) before the next one reads it
    }
}

// the two tables of one resize (source index; 11-bit weights c0 | c1 << 16 per destination column / row): dW + dH entries per call
__global__ void k_resize_tabs(int W, int H, int dW, int dH, int2 *__restrict__ tab)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dW + dH; i += gridDim.x * blockDim.x) {
        int s, c0, c1;
        if (i < dW) resize_tab(i, dW, W, s, c0, c1); else resize_tab(i - dW, dH, H, s, c0, c1);
        tab[i] = make_int2(s, c0 | (c1 << 16));
    }
}

// a thread makes four neighbouring destination pixels of a row (one word store; dpitch is a multiple of 4); block = 64 groups x 4
// rows, blockIdx.x walks the rows of all frames; the row's two source rows and weights are set up once per thread and row
__global__ void __launch_bounds__(256)
k_resize_linear(const uint8_t *__restrict__ src, int W, int H, size_t spitch, size_t sframe, uint8_t *__restrict__ dst, int dW, int dH, size_t dpitch, size_t dframe, int nb,
                const int2 *__restrict__ tab)
{
    const int gW = (dW + 3) / 4;
    const bool area = (W == 2 * dW && H == 2 * dH);
    const unsigned rows = (unsigned)nb * (unsigned)dH;
    for (unsigned R = blockIdx.x * 4u + threadIdx.y; R < rows; R += gridDim.x * 4u) {
        const unsigned b = R / (unsigned)dH;
        const int y = (int)(R - b * (unsigned)dH);
        const uint8_t *sb = src + (size_t)b * sframe;
        uint8_t *drow = dst + (size_t)b * dframe + (size_t)y * dpitch;
        const int2 ty = area ? make_int2(0, 0) : __ldg(tab + dW + y);
        const int sy1 = ty.x + 1 < H ? ty.x + 1 : H - 1, b0 = ty.y & 0xFFFF, b1 = ty.y >> 16;
        const uint8_t *s0 = sb + (size_t)ty.x * spitch, *s1 = sb + (size_t)sy1 * spitch;
        for (int gx = threadIdx.x; gx < gW; gx += 64) {
            const int x0 = 4 * gx;
            uint32_t word = 0;
            for (int q = 0; q < 4 && x0 + q < dW; ++q) {
                uint32_t v;
                if (area) v = resize_pixel(sb, W, H, spitch, dW, dH, x0 + q, y);
                else { const int2 tx = __ldg(tab + x0 + q); v = resize_pixel_rows(s0, s1, W, tx.x, tx.y & 0xFFFF, tx.y >> 16, b0, b1); }
                word |= v << (8 * q);
            }
            if (x0 + 3 < dW) *reinterpret_cast<uint32_t *>(drow + x0) = word;
            else for (int q = 0; x0 + q < dW; ++q) drow[x0 + q] = (uint8_t)(word >> (8 * q));
        }
    }
}

}  // namespace b2a

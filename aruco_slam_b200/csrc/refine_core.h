// refine_core.h -- CORNER_REFINE_CONTOUR of one accepted marker, host- and device-callable (hostemu runs it on one lane).
// Replaces cv::aruco's _refineCandidateLines / _interpolate2Dline / _getCrossPoint, which detectMarkers (reference
// src/aruco_slam.cpp:313) runs on every accepted marker when DetectorParameters::cornerRefinementMethod is
// CORNER_REFINE_CONTOUR: each side of the marker becomes the least-squares line through the contour points between two
// corners, the corners move to the crossings of neighbouring lines.
//
// Arithmetic follows cv2 4.13: the fit is cv::solve(A, B, DECOMP_NORMAL) on CV_32F matrices, i.e. A^T A and A^T B rounded to
// float (the sums are exact here: integer coordinates, 64-bit accumulators), the 2 x 2 system through the float LU with partial
// pivoting of hal::LU32f, the crossing through the closed form of Matx22f::solve.  Every float operation is a single rounded
// operation (no contraction into FMAs).
#pragma once
#include "core.h"
#include <float.h>
#include <math.h>

namespace b2a {

#ifdef __CUDA_ARCH__
B2A_HD float rf_mul(float a, float b) { return __fmul_rn(a, b); }
B2A_HD float rf_add(float a, float b) { return __fadd_rn(a, b); }
B2A_HD float rf_sub(float a, float b) { return __fsub_rn(a, b); }
B2A_HD float rf_div(float a, float b) { return __fdiv_rn(a, b); }
#else
B2A_HD float rf_mul(float a, float b) { volatile float r = a * b; return r; }
B2A_HD float rf_add(float a, float b) { volatile float r = a + b; return r; }
B2A_HD float rf_sub(float a, float b) { volatile float r = a - b; return r; }
B2A_HD float rf_div(float a, float b) { volatile float r = a / b; return r; }
#endif

constexpr int REFINE_MAX_EVENTS = 16;      // corner hits along one contour (4 unless the border passes a corner pixel twice)

struct SideSums {
    long long n, sx, sy, sxx, syy, sxy;
    int minx, maxx, miny, maxy;
};

// hal::LU32f on a 2 x 2 system with one right-hand side (LUImpl unrolled for m = 2, eps = FLT_EPSILON * 10)
B2A_HD bool refine_lu2(float a00, float a01, float a10, float a11, float b0, float b1, float &x0, float &x1)
{
    if (fabsf(a10) > fabsf(a00)) {
        float t = a00; a00 = a10; a10 = t;
        t = a01; a01 = a11; a11 = t;
        t = b0; b0 = b1; b1 = t;
    }
    if (fabsf(a00) < FLT_EPSILON * 10) return false;
    const float d = rf_div(-1.f, a00), alpha = rf_mul(a10, d);
    a11 = rf_add(a11, rf_mul(alpha, a01));
    b1 = rf_add(b1, rf_mul(alpha, b0));
    if (fabsf(a11) < FLT_EPSILON * 10) return false;
    x1 = rf_div(b1, a11);
    x0 = rf_div(rf_sub(b0, rf_mul(a01, x1)), a00);
    return true;
}

// _interpolate2Dline: (a, -1, b) for y = a x + b when the side is wider than tall, else (-1, a, b) for x = a y + b
B2A_HD void refine_side_line(const SideSums &g, float line[3])
{
    float x0 = 0.f, x1 = 0.f;
    if (g.maxx - g.minx > g.maxy - g.miny) {
        if (!refine_lu2((float)(double)g.sxx, (float)(double)g.sx, (float)(double)g.sx, (float)(double)g.n, (float)(double)g.sxy, (float)(double)g.sy, x0, x1)) x0 = x1 = 0.f;
        line[0] = x0; line[1] = -1.f; line[2] = x1;
    } else {
        if (!refine_lu2((float)(double)g.syy, (float)(double)g.sy, (float)(double)g.sy, (float)(double)g.n, (float)(double)g.sxy, (float)(double)g.sx, x0, x1)) x0 = x1 = 0.f;
        line[0] = -1.f; line[1] = x0; line[2] = x1;
    }
}

// _getCrossPoint: Matx22f(l1.x, l1.y, l2.x, l2.y).solve(Vec2f(-l1.z, -l2.z)), zero when singular
B2A_HD void refine_cross(const float l1[3], const float l2[3], float &ox, float &oy)
{
    float d = rf_sub(rf_mul(l1[0], l2[1]), rf_mul(l1[1], l2[0]));
    ox = oy = 0.f;
    if (d == 0.f) return;
    d = rf_div(1.f, d);
    const float b0 = -l1[2], b1 = -l2[2];
    ox = rf_mul(rf_sub(rf_mul(b0, l2[1]), rf_mul(b1, l1[1])), d);
    oy = rf_mul(rf_sub(rf_mul(b1, l1[0]), rf_mul(b0, l2[0])), d);
}

// One marker.  P: the contour's points (x | y << 16) in cv2.findContours order; cin: the four corners (each a contour point,
// final order); cout: refined corners.  Returns false and copies cin when a corner is not on the contour or a side has fewer
// than two points (cv2 raises there) or the border passes corner pixels more than REFINE_MAX_EVENTS times.
// LG: lanes working together -- lane(), nlanes(), ballot(bool), bcast(int, lane), sum(long long), imin(int), imax(int).
template <class LG>
B2A_HD bool refine_marker_lines(const LG &lg, const uint32_t *P, int count, const float *cin, float *cout)
{
    int cx[4], cy[4];
    for (int j = 0; j < 4; ++j) { cx[j] = (int)cin[2 * j]; cy[j] = (int)cin[2 * j + 1]; }
    bool ok = true;
    for (int j = 0; j < 4; ++j) if ((float)cx[j] != cin[2 * j] || (float)cy[j] != cin[2 * j + 1]) ok = false;
    // the corner hits in contour order: a point belongs to the side of the last corner met before it
    int evp[REFINE_MAX_EVENTS], evj[REFINE_MAX_EVENTS], nev = 0;
    for (int base = 0; ok && base < count; base += lg.nlanes()) {
        const int i = base + lg.lane();
        int j = -1;
        if (i < count) {
            const int x = px_of(P[i]), y = py_of(P[i]);
            for (int k = 0; k < 4; ++k) if (x == cx[k] && y == cy[k]) j = k;
        }
        for (uint32_t m = lg.ballot(j >= 0); m; m &= m - 1) {
            const int l = ffs32(m) - 1;
            const int jj = lg.bcast(j, l);
            if (nev < REFINE_MAX_EVENTS) { evp[nev] = base + l; evj[nev] = jj; }
            ++nev;
        }
    }
    int cornerIndex[4] = {-1, -1, -1, -1};
    if (nev > REFINE_MAX_EVENTS) ok = false;
    for (int k = 0; ok && k < nev; ++k) cornerIndex[evj[k]] = evp[k];
    for (int j = 0; j < 4; ++j) if (cornerIndex[j] < 0) ok = false;
    SideSums g[4];
    for (int j = 0; j < 4; ++j) { g[j] = SideSums{0, 0, 0, 0, 0, 0, 0x7FFFFFFF, -0x7FFFFFFF, 0x7FFFFFFF, -0x7FFFFFFF}; }
    for (int k = 0; ok && k < nev; ++k) {
        const int s = evp[k], e = k + 1 < nev ? evp[k + 1] : count + evp[0];      // the last side wraps to the first corner hit
        long long sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
        int mnx = 0x7FFFFFFF, mxx = -0x7FFFFFFF, mny = 0x7FFFFFFF, mxy = -0x7FFFFFFF;
        for (int t = s + lg.lane(); t < e; t += lg.nlanes()) {
            const uint32_t p = P[t < count ? t : t - count];
            const int x = px_of(p), y = py_of(p);
            sx += x; sy += y; sxx += (long long)x * x; syy += (long long)y * y; sxy += (long long)x * y;
            mnx = x < mnx ? x : mnx; mxx = x > mxx ? x : mxx; mny = y < mny ? y : mny; mxy = y > mxy ? y : mxy;
        }
        SideSums &q = g[evj[k]];
        q.n += e - s; q.sx += lg.sum(sx); q.sy += lg.sum(sy); q.sxx += lg.sum(sxx); q.syy += lg.sum(syy); q.sxy += lg.sum(sxy);
        mnx = lg.imin(mnx); mxx = lg.imax(mxx); mny = lg.imin(mny); mxy = lg.imax(mxy);
        q.minx = mnx < q.minx ? mnx : q.minx; q.maxx = mxx > q.maxx ? mxx : q.maxx;
        q.miny = mny < q.miny ? mny : q.miny; q.maxy = mxy > q.maxy ? mxy : q.maxy;
    }
    for (int j = 0; j < 4; ++j) if (g[j].n < 2) ok = false;
    if (!ok) {
        for (int k = 0; k < 8; ++k) cout[k] = cin[k];
        return false;
    }
    // direction of the contour relative to the corner order
    int inc = 1;
    if (cornerIndex[0] > cornerIndex[1] && cornerIndex[3] > cornerIndex[0]) inc = -1;
    if (cornerIndex[2] > cornerIndex[3] && cornerIndex[1] > cornerIndex[2]) inc = -1;
    float lines[4][3];
    for (int j = 0; j < 4; ++j) refine_side_line(g[j], lines[j]);
    for (int j = 0; j < 4; ++j) refine_cross(lines[j], lines[inc < 0 ? (j + 1) & 3 : (j + 3) & 3], cout[2 * j], cout[2 * j + 1]);
    return true;
}

// The contour of an accepted marker: the first kept border, in candidate order (threshold scales in order, borders in
// cv2.findContours order), whose quad is the marker's four corners.  (Equal quads of several scales fall into one group of
// filterTooCloseCandidates with equal perimeters, and the stable sort keeps the first of them in front.)
// Returns scale * surv_cap + slot or -1; every lane gets the same value.
template <class LG>
B2A_HD int refine_find_border(const LG &lg, const int32_t *surv_count, const uint8_t *quad_ok, const int32_t *quad_xy, int nScales, int surv_cap,
                              const float *cin)
{
    int cx[4], cy[4];
    for (int j = 0; j < 4; ++j) { cx[j] = (int)cin[2 * j]; cy[j] = (int)cin[2 * j + 1]; }
    for (int s = 0; s < nScales; ++s) {
        int n = surv_count[s];
        if (n > surv_cap) n = surv_cap;
        for (int base = 0; base < n; base += lg.nlanes()) {
            const int i = base + lg.lane();
            bool hit = false;
            if (i < n && quad_ok[(size_t)s * surv_cap + i]) {
                const int32_t *q = quad_xy + ((size_t)s * surv_cap + i) * 8;
                unsigned seen = 0;
                for (int a = 0; a < 4; ++a)
                    for (int b = 0; b < 4; ++b) if (q[2 * a] == cx[b] && q[2 * a + 1] == cy[b]) seen |= 1u << b;
                hit = seen == 15u;
            }
            const uint32_t m = lg.ballot(hit);
            if (m) return s * surv_cap + base + ffs32(m) - 1;
        }
    }
    return -1;
}

}  // namespace b2a

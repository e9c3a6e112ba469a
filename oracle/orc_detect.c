/* orc_detect.c -- CPU restatement of cv::aruco::detectMarkers (OpenCV 4.13.0,
 * default DetectorParameters), the call at reference src/aruco_slam.cpp:313.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Sequential, single-threaded, written
 * for clarity; every stage follows SURVEY.md Appendix A (A1..A8, A3a(i), A3b) and
 * is pinned against the cv2 4.13.0 wheel by tests/golden (tools/make_golden.py).
 */
#include "oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

void orc_free(void *p) { free(p); }

void orc_default_params(orc_params *p)
{
    p->adaptiveThreshWinSizeMin = 3;
    p->adaptiveThreshWinSizeMax = 23;
    p->adaptiveThreshWinSizeStep = 10;
    p->adaptiveThreshConstant = 7.0;
    p->minMarkerPerimeterRate = 0.03;
    p->maxMarkerPerimeterRate = 4.0;
    p->polygonalApproxAccuracyRate = 0.03;
    p->minCornerDistanceRate = 0.05;
    p->minDistanceToBorder = 3;
    p->minMarkerDistanceRate = 0.125;
    p->minGroupDistance = 0.21f;
    p->markerBorderBits = 1;
    p->perspectiveRemovePixelPerCell = 4;
    p->perspectiveRemoveIgnoredMarginPerCell = 0.13;
    p->maxErroneousBitsInBorderRate = 0.35;
    p->minOtsuStdDev = 5.0;
    p->errorCorrectionRate = 0.6;
    p->cornerRefinementMethod = 0;
    p->cornerRefinementWinSize = 5;
    p->relativeCornerRefinmentWinSize = 0.3;
    p->cornerRefinementMaxIterations = 30;
    p->cornerRefinementMinAccuracy = 0.1;
    p->detectInvertedMarker = 0;
    p->useAruco3Detection = 0;
    p->minSideLengthCanonicalImg = 32;
    p->minMarkerLengthRatioOriginalImg = 0.f;
}

/* A1: cvtColor(BGR2GRAY) 8-bit: 15-bit fixed point, SURVEY App. A1 / probe P3 */
void orc_bgr2gray(const uint8_t *bgr, int W, int H, uint8_t *gray)
{
    long n = (long)W * H;
    for (long i = 0; i < n; i++) {
        int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        gray[i] = (uint8_t)((3735 * b + 19235 * g + 9798 * r + 16384) >> 15);
    }
}

/* A2: adaptiveThreshold(MEAN_C, THRESH_BINARY_INV, k, C): box mean with replicate
 * border rounded to nearest, mask = 255 iff g - mean <= -floor(C).  Probe P4. */
void orc_adaptive_threshold(const uint8_t *gray, int W, int H, int k, double C, uint8_t *mask)
{
    int r = k / 2;
    int idelta = (int)floor(C);
    /* integral image of the replicate-padded frame */
    int PW = W + 2 * r, PH = H + 2 * r;
    uint32_t *I = (uint32_t *)calloc((size_t)(PW + 1) * (PH + 1), sizeof(uint32_t));
    for (int y = 0; y < PH; y++) {
        int sy = y - r; if (sy < 0) sy = 0; if (sy >= H) sy = H - 1;
        uint32_t row = 0;
        for (int x = 0; x < PW; x++) {
            int sx = x - r; if (sx < 0) sx = 0; if (sx >= W) sx = W - 1;
            row += gray[(size_t)sy * W + sx];
            I[(size_t)(y + 1) * (PW + 1) + x + 1] = I[(size_t)y * (PW + 1) + x + 1] + row;
        }
    }
    int kk = k * k;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            uint32_t S = I[(size_t)(y + k) * (PW + 1) + x + k] - I[(size_t)y * (PW + 1) + x + k]
                       - I[(size_t)(y + k) * (PW + 1) + x] + I[(size_t)y * (PW + 1) + x];
            int mean = (int)((2 * S + kk) / (2 * kk));
            int g = gray[(size_t)y * W + x];
            mask[(size_t)y * W + x] = (g - mean <= -idelta) ? 255 : 0;
        }
    free(I);
}

/* ------------------------------------------------------------------------- */
/* A3a(i): findContours(RETR_LIST, CHAIN_APPROX_NONE), Suzuki-Abe sequential  */
/* ------------------------------------------------------------------------- */
typedef struct { int32_t *v; size_t n, cap; } ivec;
static void iv_push(ivec *a, int32_t x)
{
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 1024; a->v = (int32_t *)realloc(a->v, a->cap * sizeof(int32_t)); }
    a->v[a->n++] = x;
}

int orc_find_contours(const uint8_t *mask, int W, int H, int32_t **pts_out, int32_t **offs_out)
{
    const int PW = W + 2;
    int32_t *f = (int32_t *)calloc((size_t)PW * (H + 2), sizeof(int32_t));
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            f[(size_t)(y + 1) * PW + x + 1] = mask[(size_t)y * W + x] ? 1 : 0;
    /* d = 0..7: E, NE, N, NW, W, SW, S, SE (y down); 16 entries so ++s may run past 7 */
    int delta[16];
    const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1}, dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    for (int i = 0; i < 16; i++) delta[i] = dy[i & 7] * PW + dx[i & 7];

    ivec pts = {0, 0, 0}, starts = {0, 0, 0};   /* discovery order */
    int nbd = 1;
    for (int y = 1; y <= H; y++) {
        int32_t prev = 0;
        for (int x = 1; x <= W; x++) {
            int32_t *px = f + (size_t)y * PW + x;
            int32_t p = *px;
            if (p != prev) {
                int is_hole = 0, start = 0;
                int32_t *o = px;
                if (prev == 0 && p == 1) start = 1;
                else if (p == 0 && prev >= 1) { start = 1; is_hole = 1; o = px - 1; }
                if (start) {
                    nbd++;
                    iv_push(&starts, (int32_t)(pts.n / 2));
                    int s_end = is_hole ? 0 : 4, s = s_end;
                    int32_t *i1;
                    do { s = (s - 1) & 7; i1 = o + delta[s]; } while (*i1 == 0 && s != s_end);
                    if (s == s_end) {                       /* isolated pixel */
                        *o = -nbd;
                        long off = o - f;
                        iv_push(&pts, (int32_t)(off % PW) - 1); iv_push(&pts, (int32_t)(off / PW) - 1);
                    } else {
                        int32_t *c = o, *n4;
                        for (;;) {
                            int s_e = s;
                            for (;;) { n4 = c + delta[++s]; if (*n4 != 0) break; }
                            s &= 7;
                            if ((unsigned)(s - 1) < (unsigned)s_e) *c = -nbd;
                            else if (*c == 1) *c = nbd;
                            long off = c - f;
                            iv_push(&pts, (int32_t)(off % PW) - 1); iv_push(&pts, (int32_t)(off / PW) - 1);
                            if (n4 == o && c == i1) break;
                            c = n4;
                            s = (s + 4) & 7;
                        }
                    }
                    p = *px;                                 /* marked value */
                }
            }
            prev = p;
        }
    }
    free(f);
    /* returned list = reverse discovery order */
    int nc = (int)starts.n;
    int32_t *offs = (int32_t *)malloc((size_t)(nc + 1) * sizeof(int32_t));
    int32_t *out = (int32_t *)malloc((pts.n ? pts.n : 1) * sizeof(int32_t));
    int32_t pos = 0;
    for (int i = 0; i < nc; i++) {
        int j = nc - 1 - i;
        int32_t b = starts.v[j], e = (j + 1 < nc) ? starts.v[j + 1] : (int32_t)(pts.n / 2);
        offs[i] = pos;
        memcpy(out + 2 * (size_t)pos, pts.v + 2 * (size_t)b, (size_t)(e - b) * 2 * sizeof(int32_t));
        pos += e - b;
    }
    offs[nc] = pos;
    free(pts.v); free(starts.v);
    *pts_out = out; *offs_out = offs;
    return nc;
}

/* ------------------------------------------------------------------------- */
/* A3b: approxPolyDP(closed) on integer points, 4.13 point-to-segment variant  */
/* ------------------------------------------------------------------------- */
int orc_approx_poly_dp(const int32_t *P, int count, double eps, int32_t *out)
{
    if (count == 0) return 0;
    double eps2 = eps * eps;
    typedef struct { int s, e; } rng;
    rng *stack = (rng *)malloc(sizeof(rng) * (size_t)(count + 8));
    int top = 0, m = 0;
    int right = 0, pos = 0, le;
    double sx = 0, sy = 0, maxd = 0;
    for (int it = 0; it < 3; it++) {
        pos = (pos + right) % count;
        sx = P[2 * pos]; sy = P[2 * pos + 1];
        pos = (pos + 1) % count;
        maxd = 0;
        for (int j = 1; j < count; j++) {
            double ddx = P[2 * pos] - sx, ddy = P[2 * pos + 1] - sy;
            pos = (pos + 1) % count;
            double d = ddx * ddx + ddy * ddy;
            if (d > maxd) { maxd = d; right = j; }
        }
    }
    le = maxd <= eps2;
    if (le) {
        out[0] = (int32_t)sx; out[1] = (int32_t)sy; m = 1;
    } else {
        int a = pos % count, b = (right + a) % count;
        stack[top].s = b; stack[top].e = a; top++;
        stack[top].s = a; stack[top].e = b; top++;
    }
    while (top > 0) {
        rng sl = stack[--top];
        double ex = P[2 * sl.e], ey = P[2 * sl.e + 1];
        int p2 = sl.s, split = 0;
        double stx = P[2 * p2], sty = P[2 * p2 + 1];
        p2 = (p2 + 1) % count;
        if (p2 != sl.e) {
            double ddx = ex - stx, ddy = ey - sty, L2 = ddx * ddx + ddy * ddy;
            double md = 0;
            while (p2 != sl.e) {
                double px = P[2 * p2] - stx, py = P[2 * p2 + 1] - sty;
                p2 = (p2 + 1) % count;
                double dot = px * ddx + py * ddy, d;
                if (dot < 0) d = (px * px + py * py) * L2;
                else if (dot > L2) { double qx = px - ddx, qy = py - ddy; d = (qx * qx + qy * qy) * L2; }
                else { double cr = py * ddx - px * ddy; d = cr * cr; }
                if (d > md) { md = d; split = (p2 + count - 1) % count; }
            }
            le = md <= eps2 * L2;
        } else le = 1;
        if (le) { out[2 * m] = (int32_t)stx; out[2 * m + 1] = (int32_t)sty; m++; }
        else {
            stack[top].s = split; stack[top].e = sl.e; top++;
            stack[top].s = sl.s; stack[top].e = split; top++;
        }
    }
    free(stack);
    /* clean-up pass: drop vertices that are (nearly) collinear with their neighbours */
    int new_count = m;
    if (m > 0) {
        int32_t *o = out;
        int r = 0, w = 0;
        double stx = o[2 * (m - 1)], sty = o[2 * (m - 1) + 1];
        double ptx = o[0], pty = o[1];
        r = 1 % m;
        /* work on a copy for reads (writes trail reads, as in the in-place original) */
        for (int i = 0; i < m && new_count > 2; i++) {
            double ex = o[2 * r], ey = o[2 * r + 1];
            r = (r + 1) % m;
            double ddx = ex - stx, ddy = ey - sty;
            double dist = fabs((ptx - stx) * ddy - (pty - sty) * ddx);
            double sip = (ptx - stx) * (ex - ptx) + (pty - sty) * (ey - pty);
            if (dist * dist <= 0.5 * eps2 * (ddx * ddx + ddy * ddy) && ddx != 0 && ddy != 0 && sip >= 0) {
                new_count--;
                o[2 * w] = (int32_t)ex; o[2 * w + 1] = (int32_t)ey; w = (w + 1) % m;
                stx = ex; sty = ey;
                ptx = o[2 * r]; pty = o[2 * r + 1];
                r = (r + 1) % m;
                i++;
                continue;
            }
            o[2 * w] = (int32_t)ptx; o[2 * w + 1] = (int32_t)pty; w = (w + 1) % m;
            stx = ptx; sty = pty;
            ptx = ex; pty = ey;
        }
    }
    return new_count;
}

int orc_is_contour_convex(const int32_t *q, int n)
{
    if (n < 3) return 0;   /* not used by the detector for n != 4 */
    long long px = q[2 * (n - 2)], py = q[2 * (n - 2) + 1];
    long long cx = q[2 * (n - 1)], cy = q[2 * (n - 1) + 1];
    long long dx0 = cx - px, dy0 = cy - py;
    int o = 0;
    for (int i = 0; i < n; i++) {
        px = cx; py = cy;
        cx = q[2 * i]; cy = q[2 * i + 1];
        long long ddx = cx - px, ddy = cy - py;
        long long a = ddx * dy0, b = ddy * dx0;
        o |= (b > a) ? 1 : ((b < a) ? 2 : 3);
        if (o == 3) return 0;
        dx0 = ddx; dy0 = ddy;
    }
    return 1;
}

int orc_point_polygon_test(const float *poly, int n, float ptx, float pty)
{
    int counter = 0;
    float vx = poly[2 * (n - 1)], vy = poly[2 * (n - 1) + 1];
    for (int i = 0; i < n; i++) {
        float v0x = vx, v0y = vy;
        vx = poly[2 * i]; vy = poly[2 * i + 1];
        if ((v0y <= pty && vy <= pty) || (v0y > pty && vy > pty) || (v0x < ptx && vx < ptx)) {
            if (pty == vy && (ptx == vx || (pty == v0y && ((v0x <= ptx && ptx <= vx) || (vx <= ptx && ptx <= v0x)))))
                return 0;
            continue;
        }
        double dist = (double)(pty - v0y) * (vx - v0x) - (double)(ptx - v0x) * (vy - v0y);
        if (dist == 0) return 0;
        if (vy < v0y) dist = -dist;
        counter += dist > 0;
    }
    return (counter % 2 == 0) ? -1 : 1;
}

/* getPerspectiveTransform: 8x8 system, LU with partial pivoting in double, no FMA */
void orc_get_perspective_transform(const float *src, const float *dst, double *Hout)
{
    double A[8][8], b[8];
    memset(A, 0, sizeof(A));
    for (int i = 0; i < 4; i++) {
        double sx = src[2 * i], sy = src[2 * i + 1], ddx = dst[2 * i], ddy = dst[2 * i + 1];
        A[i][0] = A[i + 4][3] = sx;
        A[i][1] = A[i + 4][4] = sy;
        A[i][2] = A[i + 4][5] = 1;
        A[i][6] = -sx * ddx; A[i][7] = -sy * ddx;
        A[i + 4][6] = -sx * ddy; A[i + 4][7] = -sy * ddy;
        b[i] = ddx; b[i + 4] = ddy;
    }
    const int m = 8;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++) if (fabs(A[j][i]) > fabs(A[k][i])) k = j;
        if (k != i) {
            for (int j = i; j < m; j++) { double t = A[i][j]; A[i][j] = A[k][j]; A[k][j] = t; }
            double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        double d = -1 / A[i][i];
        for (int j = i + 1; j < m; j++) {
            double alpha = A[j][i] * d;
            for (int kk = i + 1; kk < m; kk++) A[j][kk] += alpha * A[i][kk];
            b[j] += alpha * b[i];
        }
    }
    for (int i = m - 1; i >= 0; i--) {
        double s = b[i];
        for (int k = i + 1; k < m; k++) s -= A[i][k] * b[k];
        b[i] = s / A[i][i];
    }
    for (int i = 0; i < 8; i++) Hout[i] = b[i];
    Hout[8] = 1.0;
}

static void inv3x3(const double *a, double *t)
{
    double d = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    d = 1. / d;
    t[0] = (a[4] * a[8] - a[5] * a[7]) * d;
    t[1] = (a[2] * a[7] - a[1] * a[8]) * d;
    t[2] = (a[1] * a[5] - a[2] * a[4]) * d;
    t[3] = (a[5] * a[6] - a[3] * a[8]) * d;
    t[4] = (a[0] * a[8] - a[2] * a[6]) * d;
    t[5] = (a[2] * a[3] - a[0] * a[5]) * d;
    t[6] = (a[3] * a[7] - a[4] * a[6]) * d;
    t[7] = (a[1] * a[6] - a[0] * a[7]) * d;
    t[8] = (a[0] * a[4] - a[1] * a[3]) * d;
}

void orc_warp_nearest(const uint8_t *gray, int W, int H, const double *H9, int S, uint8_t *patch)
{
    double M[9];
    inv3x3(H9, M);
    for (int y = 0; y < S; y++) {
        double X0 = M[1] * y + M[2], Y0 = M[4] * y + M[5], W0 = M[7] * y + M[8];
        for (int x = 0; x < S; x++) {
            double w = W0 + M[6] * x;
            w = w ? 1. / w : 0;
            double fx = (X0 + M[0] * x) * w, fy = (Y0 + M[3] * x) * w;
            if (fx < -2147483648.0) fx = -2147483648.0; if (fx > 2147483647.0) fx = 2147483647.0;
            if (fy < -2147483648.0) fy = -2147483648.0; if (fy > 2147483647.0) fy = 2147483647.0;
            long X = lrint(fx), Y = lrint(fy);
            patch[y * S + x] = (X >= 0 && X < W && Y >= 0 && Y < H) ? gray[(size_t)Y * W + X] : 0;
        }
    }
}

int orc_otsu(const uint8_t *img, int n)
{
    int h[256] = {0};
    for (int i = 0; i < n; i++) h[img[i]]++;
    double mu = 0, scale = 1. / n;
    for (int i = 0; i < 256; i++) mu += i * (double)h[i];
    mu *= scale;
    double mu1 = 0, q1 = 0, max_sigma = 0, max_val = 0;
    for (int i = 0; i < 256; i++) {
        double p_i = h[i] * scale, q2, mu2, sigma;
        mu1 *= q1;
        q1 += p_i;
        q2 = 1. - q1;
        if (fmin(q1, q2) < FLT_EPSILON || fmax(q1, q2) > 1. - FLT_EPSILON) continue;
        mu1 = (mu1 + i * p_i) / q1;
        mu2 = (mu - q1 * mu1) / q2;
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    return (int)max_val;
}

/* A7 */
int orc_identify_one(const uint8_t *gray, int W, int H, const float *corners, const orc_dict *d,
                     const orc_params *p, int *id_out, int *rot_out, uint8_t *bits_out)
{
    int bb = p->markerBorderBits, cell = p->perspectiveRemovePixelPerCell;
    int nb = d->markerSize + 2 * bb, S = nb * cell;
    int margin = (int)(p->perspectiveRemoveIgnoredMarginPerCell * cell);
    float dst[8] = {0, 0, (float)S - 1, 0, (float)S - 1, (float)S - 1, 0, (float)S - 1};
    double Hm[9];
    orc_get_perspective_transform(corners, dst, Hm);
    uint8_t *patch = (uint8_t *)malloc((size_t)S * S);
    uint8_t *bits = (uint8_t *)calloc((size_t)nb * nb, 1);
    orc_warp_nearest(gray, W, H, Hm, S, patch);
    /* meanStdDev of the inner region (cell/2 margin) */
    int m0 = cell / 2;
    long long s = 0, sq = 0; int cnt = 0;
    for (int y = m0; y < S - m0; y++) for (int x = m0; x < S - m0; x++) { int v = patch[y * S + x]; s += v; sq += v * v; cnt++; }
    double scale = 1. / cnt, mean = s * scale, var = sq * scale - mean * mean;
    if (var < 0) var = 0;
    double sd = sqrt(var);
    if (sd < p->minOtsuStdDev) {
        memset(bits, mean > 127 ? 1 : 0, (size_t)nb * nb);
    } else {
        int t = orc_otsu(patch, S * S);
        int cw = cell - 2 * margin;
        for (int y = 0; y < nb; y++) for (int x = 0; x < nb; x++) {
            int nz = 0;
            for (int yy = 0; yy < cw; yy++) for (int xx = 0; xx < cw; xx++)
                nz += patch[(y * cell + margin + yy) * S + x * cell + margin + xx] > t;
            if (nz > (cw * cw) / 2) bits[y * nb + x] = 1;
        }
    }
    free(patch);
    if (bits_out) memcpy(bits_out, bits, (size_t)nb * nb);
    int maxErr = (int)(d->markerSize * d->markerSize * p->maxErroneousBitsInBorderRate);
    int err = 0;
    for (int y = 0; y < nb; y++) for (int k = 0; k < bb; k++) { err += bits[y * nb + k] != 0; err += bits[y * nb + nb - 1 - k] != 0; }
    for (int x = bb; x < nb - bb; x++) for (int k = 0; k < bb; k++) { err += bits[k * nb + x] != 0; err += bits[(nb - 1 - k) * nb + x] != 0; }
    if (p->detectInvertedMarker) {
        /* white marker: the inverted bit matrix is taken when its border has fewer errors (cv2 _identifyOneCandidate) */
        int nborder = nb * nb - d->markerSize * d->markerSize, inv = nborder - err;
        if (inv < err) { err = inv; for (int i = 0; i < nb * nb; i++) bits[i] = (uint8_t)!bits[i]; }
    }
    if (err > maxErr) { free(bits); return 0; }
    /* pack inner bits row-major MSB first, last partial byte right-aligned */
    uint8_t code[16] = {0};
    int ms = d->markerSize, nbits = ms * ms, nby = d->nBytes;
    for (int i = 0; i < nbits; i++) {
        int y = i / ms, x = i % ms, byte = i / 8, shift;
        if (byte == nby - 1 && (nbits % 8)) shift = (nbits % 8) - 1 - (i % 8); else shift = 7 - (i % 8);
        code[byte] |= (uint8_t)(bits[(y + bb) * nb + x + bb] << shift);
    }
    free(bits);
    int maxCorr = (int)((double)d->maxCorrectionBits * p->errorCorrectionRate);
    for (int m = 0; m < d->nMarkers; m++) {
        int best = ms * ms + 1, brot = -1;
        for (int r = 0; r < 4; r++) {
            const uint8_t *t = d->table + ((size_t)m * 4 + r) * nby;
            int hd = 0;
            for (int k = 0; k < nby; k++) hd += __builtin_popcount((unsigned)(t[k] ^ code[k]));
            if (hd < best) { best = hd; brot = r; }
        }
        if (best <= maxCorr) { *id_out = m; *rot_out = brot; return 1; }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* cornerSubPix (A8), restated from OpenCV's published algorithm               */
/* ------------------------------------------------------------------------- */
static void get_rect_subpix_u8_f32(const uint8_t *src, int W, int H, float cx, float cy, int ww, int wh, float *dst)
{
    /* getRectSubPix 8u->32f: bilinear, replicate border */
    cx -= (ww - 1) * 0.5f; cy -= (wh - 1) * 0.5f;
    int ipx = (int)floorf(cx), ipy = (int)floorf(cy);
    float a = cx - ipx, b = cy - ipy;
    float a11 = (1.f - a) * (1.f - b), a12 = a * (1.f - b), a21 = (1.f - a) * b, a22 = a * b;
    for (int y = 0; y < wh; y++) {
        int y0 = ipy + y, y1 = y0 + 1;
        if (y0 < 0) y0 = 0; if (y0 >= H) y0 = H - 1; if (y1 < 0) y1 = 0; if (y1 >= H) y1 = H - 1;
        for (int x = 0; x < ww; x++) {
            int x0 = ipx + x, x1 = x0 + 1;
            if (x0 < 0) x0 = 0; if (x0 >= W) x0 = W - 1; if (x1 < 0) x1 = 0; if (x1 >= W) x1 = W - 1;
            dst[y * ww + x] = src[(size_t)y0 * W + x0] * a11 + src[(size_t)y0 * W + x1] * a12
                            + src[(size_t)y1 * W + x0] * a21 + src[(size_t)y1 * W + x1] * a22;
        }
    }
}

void orc_corner_subpix(const uint8_t *gray, int W, int H, float *corners, int n, int win, int maxIter, double eps)
{
    int ww = win * 2 + 1, wh = ww;
    float *mask = (float *)malloc(sizeof(float) * (size_t)ww * wh);
    float *sub = (float *)malloc(sizeof(float) * (size_t)(ww + 2) * (wh + 2));
    float *mx = (float *)malloc(sizeof(float) * (size_t)ww);
    double eps2 = eps * eps;
    for (int i = 0; i < ww; i++) { float y = (float)(i - win) / win; mx[i] = (float)exp(-y * y); }
    for (int i = 0; i < wh; i++) for (int j = 0; j < ww; j++) mask[i * ww + j] = mx[i] * mx[j];
    for (int pt = 0; pt < n; pt++) {
        float cTx = corners[2 * pt], cTy = corners[2 * pt + 1], cIx = cTx, cIy = cTy;
        int iter = 0; double err = 0;
        do {
            double a = 0, b = 0, c = 0, bb1 = 0, bb2 = 0;
            get_rect_subpix_u8_f32(gray, W, H, cIx, cIy, ww + 2, wh + 2, sub);
            const float *sp = sub + (ww + 2) + 1;
            for (int i = 0, k = 0; i < wh; i++, sp += ww + 2) {
                double py = i - win;
                for (int j = 0; j < ww; j++, k++) {
                    double m = mask[k];
                    double tgx = sp[j + 1] - sp[j - 1];
                    double tgy = sp[j + ww + 2] - sp[j - ww - 2];
                    double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
                    double px = j - win;
                    a += gxx; b += gxy; c += gyy;
                    bb1 += gxx * px + gxy * py;
                    bb2 += gxy * px + gyy * py;
                }
            }
            double det = a * c - b * b;
            if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
            double scale = 1.0 / det;
            float nx = (float)(cIx + (c * scale * bb1 - b * scale * bb2));
            float ny = (float)(cIy + (-b * scale * bb1 + a * scale * bb2));
            err = (nx - cIx) * (nx - cIx) + (ny - cIy) * (ny - cIy);
            cIx = nx; cIy = ny;
            if (cIx < 0 || cIx >= W || cIy < 0 || cIy >= H) break;
        } while (++iter < maxIter && err > eps2);
        if (fabsf(cIx - cTx) > win || fabsf(cIy - cTy) > win) { cIx = cTx; cIy = cTy; }
        corners[2 * pt] = cIx; corners[2 * pt + 1] = cIy;
    }
    free(mask); free(sub); free(mx);
}

/* ------------------------------------------------------------------------- */
/* full detector                                                             */
/* ------------------------------------------------------------------------- */

/* ---- CORNER_REFINE_CONTOUR: cv::aruco _refineCandidateLines / _interpolate2Dline / _getCrossPoint (cv2 4.13
 *      aruco_detector.cpp).  Every side of the marker becomes the least-squares line through the contour points between two
 *      corners; the corners move to the crossings.  cv2 fits with cv::solve(A, B, DECOMP_NORMAL) on CV_32F matrices: A^T A
 *      and A^T B are rounded to float (exact sums here: the coordinates are integers), the 2 x 2 system goes through the
 *      float LU of hal::LU32f, the crossing through Matx22f::solve's closed form -- all restated below in float.  cv2
 *      itself computes A^T B with its BLAS once a side has 100 points or more, so there its sums carry float rounding
 *      that depends on the BLAS build: parity with cv2 is exact below 100 points per side and within 0.05 px above
 *      (tests/golden/contour_refine_*.npz). ---- */
static int lu2_solve_f32(float A[2][2], float b[2], float x[2])
{
    for (int i = 0; i < 2; i++) {
        int k = i;
        for (int j = i + 1; j < 2; j++) if (fabsf(A[j][i]) > fabsf(A[k][i])) k = j;
        if (fabsf(A[k][i]) < FLT_EPSILON * 10) return 0;
        if (k != i) {
            for (int c = 0; c < 2; c++) { float t = A[i][c]; A[i][c] = A[k][c]; A[k][c] = t; }
            float t = b[i]; b[i] = b[k]; b[k] = t;
        }
        float d = -1 / A[i][i];
        for (int j = i + 1; j < 2; j++) {
            float alpha = A[j][i] * d;
            for (int c = i + 1; c < 2; c++) A[j][c] += alpha * A[i][c];
            b[j] += alpha * b[i];
        }
    }
    for (int i = 1; i >= 0; i--) {
        float s = b[i];
        for (int c = i + 1; c < 2; c++) s -= A[i][c] * x[c];
        x[i] = s / A[i][i];
    }
    return 1;
}

typedef struct { double n, sx, sy, sxx, syy, sxy; int minx, maxx, miny, maxy; } side_sums;

static void interpolate_2d_line(const side_sums *g, float line[3])
{
    float A[2][2], b[2], x[2] = {0.f, 0.f};
    if (g->maxx - g->minx > g->maxy - g->miny) {          /* y = a x + b  ->  (a, -1, b) */
        A[0][0] = (float)g->sxx; A[0][1] = A[1][0] = (float)g->sx; A[1][1] = (float)g->n;
        b[0] = (float)g->sxy; b[1] = (float)g->sy;
        if (!lu2_solve_f32(A, b, x)) x[0] = x[1] = 0.f;
        line[0] = x[0]; line[1] = -1.f; line[2] = x[1];
    } else {                                              /* x = a y + b  ->  (-1, a, b) */
        A[0][0] = (float)g->syy; A[0][1] = A[1][0] = (float)g->sy; A[1][1] = (float)g->n;
        b[0] = (float)g->sxy; b[1] = (float)g->sx;
        if (!lu2_solve_f32(A, b, x)) x[0] = x[1] = 0.f;
        line[0] = -1.f; line[1] = x[0]; line[2] = x[1];
    }
}

static void cross_point(const float l1[3], const float l2[3], float out[2])
{
    float d = l1[0] * l2[1] - l1[1] * l2[0];
    out[0] = out[1] = 0.f;
    if (d == 0) return;
    d = 1 / d;
    const float b0 = -l1[2], b1 = -l2[2];
    out[0] = (b0 * l2[1] - b1 * l1[1]) * d;
    out[1] = (b1 * l1[0] - b0 * l2[0]) * d;
}

/* corners: 4 x (x, y), each one a point of the contour; returns 0 (corners untouched) when a corner is not on the contour or a
 * side has fewer than two points (cv2 raises there) */
int orc_refine_candidate_lines(const int32_t *contour, int n, float *corners)
{
    side_sums g[5];
    int cornerIndex[4] = {-1, -1, -1, -1}, group = 4;
    memset(g, 0, sizeof(g));
    for (int k = 0; k < 5; k++) { g[k].minx = g[k].miny = INT32_MAX; g[k].maxx = g[k].maxy = INT32_MIN; }
    for (int i = 0; i < n; i++) {
        const int x = contour[2 * i], y = contour[2 * i + 1];
        for (int j = 0; j < 4; j++) if (corners[2 * j] == (float)x && corners[2 * j + 1] == (float)y) { cornerIndex[j] = i; group = j; }
        side_sums *s = &g[group];
        s->n += 1; s->sx += x; s->sy += y; s->sxx += (double)x * x; s->syy += (double)y * y; s->sxy += (double)x * y;
        if (x < s->minx) s->minx = x; if (x > s->maxx) s->maxx = x; if (y < s->miny) s->miny = y; if (y > s->maxy) s->maxy = y;
    }
    for (int j = 0; j < 4; j++) if (cornerIndex[j] == -1) return 0;
    if (g[4].n > 0) {                                     /* the points before the first corner belong to the last side */
        side_sums *s = &g[group], *e = &g[4];
        s->n += e->n; s->sx += e->sx; s->sy += e->sy; s->sxx += e->sxx; s->syy += e->syy; s->sxy += e->sxy;
        if (e->minx < s->minx) s->minx = e->minx; if (e->maxx > s->maxx) s->maxx = e->maxx;
        if (e->miny < s->miny) s->miny = e->miny; if (e->maxy > s->maxy) s->maxy = e->maxy;
    }
    for (int j = 0; j < 4; j++) if (g[j].n < 2) return 0;
    int inc = 1;
    if (cornerIndex[0] > cornerIndex[1] && cornerIndex[3] > cornerIndex[0]) inc = -1;
    if (cornerIndex[2] > cornerIndex[3] && cornerIndex[1] > cornerIndex[2]) inc = -1;
    float lines[4][3];
    for (int j = 0; j < 4; j++) interpolate_2d_line(&g[j], lines[j]);
    for (int j = 0; j < 4; j++) cross_point(lines[j], lines[inc < 0 ? (j + 1) % 4 : (j + 3) % 4], corners + 2 * j);
    return 1;
}

typedef struct {
    float c[8];
    float perimeter;
    int   len;          /* contour length */
    int   parent, depth;
    int   n_close; int *close;   /* indices into the sorted candidate array T */
    int32_t *contour;            /* the candidate's contour points (x, y), kept only for CORNER_REFINE_CONTOUR */
} cand_t;

static float perimeter_f(const float *c)
{
    float p = 0.f;
    for (int i = 0; i < 4; i++) {
        float ddx = c[2 * i] - c[2 * ((i + 1) & 3)], ddy = c[2 * i + 1] - c[2 * ((i + 1) & 3) + 1];
        p += sqrtf(ddx * ddx + ddy * ddy);
    }
    return p;
}

static float avg_distance(const float *a, const float *b)
{
    float minsq = FLT_MAX;
    for (int fc = 0; fc < 4; fc++) {
        float dsq = 0;
        for (int c = 0; c < 4; c++) {
            int mc = (c + fc) % 4;
            float ddx = a[2 * mc] - b[2 * c], ddy = a[2 * mc + 1] - b[2 * c + 1];
            dsq += ddx * ddx + ddy * ddy;
        }
        dsq /= 4.f;
        if (dsq < minsq) minsq = dsq;
    }
    return sqrtf(minsq);
}

static float avg_module_size(const float *c, int markerSize, int borderBits)
{
    float a = perimeter_f(c);
    int nm = markerSize + borderBits * 2;
    a /= (4.f * nm);
    return a;
}

static void stable_sort_desc(cand_t *a, int n)
{   /* insertion-stable merge sort by perimeter descending */
    if (n < 2) return;
    cand_t *tmp = (cand_t *)malloc(sizeof(cand_t) * (size_t)n);
    for (int w = 1; w < n; w *= 2) {
        for (int lo = 0; lo < n; lo += 2 * w) {
            int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int i = lo, j = mid, k = lo;
            while (i < mid && j < hi) { if (a[j].perimeter > a[i].perimeter) tmp[k++] = a[j++]; else tmp[k++] = a[i++]; }
            while (i < mid) tmp[k++] = a[i++];
            while (j < hi) tmp[k++] = a[j++];
        }
        memcpy(a, tmp, sizeof(cand_t) * (size_t)n);
    }
    free(tmp);
}

static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

/* _findOptPyrImageForCanonicalImg: the level whose scaled contour length exceeds min_perimeter by the least (level 0 when none does) */
static int find_opt_pyr_level(const int *pyrW, int nLevels, int scaled_width, int cur_perimeter, int min_perimeter)
{
    int opt = 0;
    float dist = FLT_MAX;
    for (int i = 0; i < nLevels; i++) {
        const float scale = (float)pyrW[i] / (float)scaled_width;
        const float perimeter_scaled = (float)cur_perimeter * scale;
        const float new_dist = perimeter_scaled - (float)min_perimeter;
        if (new_dist < dist && new_dist > 0.f) { dist = new_dist; opt = i; }
    }
    return opt;
}

/* _identifyOneCandidate with its `scale` argument: the corners are scaled (in float) to the image they are read in */
static int identify_scaled(const uint8_t *im, int W, int H, const float *c, float scale, const orc_dict *d, const orc_params *p, int *id, int *rot)
{
    float sc[8];
    for (int j = 0; j < 8; j++) sc[j] = c[j] * scale;
    return orc_identify_one(im, W, H, sc, d, p, id, rot, NULL);
}

int orc_detect(const uint8_t *img, int W, int H, int channels, const orc_dict *d, const orc_params *p, orc_detections *out)
{
    memset(out, 0, sizeof(*out));
    size_t P = (size_t)W * H;
    uint8_t *gray = (uint8_t *)malloc(P);
    if (channels == 3) orc_bgr2gray(img, W, H, gray); else memcpy(gray, img, P);

    /* ArUco3 (useAruco3Detection): the image pyramid of the full-size gray image, and the smaller "segmentation image" the
     * candidates are searched in.  Off: one level, factor 1 (detectMarkers zeroes the two ArUco3 parameters itself). */
    const int a3 = p->useAruco3Detection != 0;
    const int minSide = a3 ? p->minSideLengthCanonicalImg : 0;
    const int refineMethod = a3 ? 1 : p->cornerRefinementMethod;      /* "always turn on corner refinement in case of Aruco3, due to upsampling" */
    float fxfy = 1.f;
    int numLevels = 0, closestIdx = 0;
    if (a3) {
        fxfy = (float)minSide / ((float)minSide + (float)(W > H ? W : H) * p->minMarkerLengthRatioOriginalImg);
        const float img_area = (float)(H * W), min_area_marker = (float)(minSide * minSide);
        numLevels = (int)(log2f(img_area / min_area_marker) / 2.f);
        const float scale_img_area = img_area * fxfy * fxfy;
        closestIdx = (int)lrintf(log2f(img_area / scale_img_area) / 2.f);
    }
    uint8_t *pyr[32]; int pyrW[32], pyrH[32];
    pyr[0] = gray; pyrW[0] = W; pyrH[0] = H;
    if (numLevels > 31) numLevels = 31;
    if (numLevels < 0) numLevels = 0;
    for (int l = 1; l <= numLevels; l++) {
        pyrW[l] = (pyrW[l - 1] + 1) / 2; pyrH[l] = (pyrH[l - 1] + 1) / 2;
        pyr[l] = (uint8_t *)malloc((size_t)pyrW[l] * pyrH[l]);
        orc_pyr_down(pyr[l - 1], pyrW[l - 1], pyrH[l - 1], pyr[l]);
    }
    const int W0 = W;
    if (fxfy != 1.f) {
        const int sW = (int)lrintf(fxfy * (float)W), sH = (int)lrintf(fxfy * (float)H);
        uint8_t *seg = (uint8_t *)malloc((size_t)sW * sH);
        orc_resize_linear(gray, W, H, seg, sW, sH);
        gray = seg; W = sW; H = sH; P = (size_t)W * H;
    }

    int nScales = (p->adaptiveThreshWinSizeMax - p->adaptiveThreshWinSizeMin) / p->adaptiveThreshWinSizeStep + 1;
    out->n_scales = nScales;
    out->n_contours = (int32_t *)calloc((size_t)nScales, sizeof(int32_t));
    int maxWH = W > H ? W : H;
    unsigned minPerim = (unsigned)(p->minMarkerPerimeterRate * maxWH);
    unsigned maxPerim = (unsigned)(p->maxMarkerPerimeterRate * maxWH);
    if (minSide) minPerim = 4u * (unsigned)minSide;          /* _findMarkerContours: "for aruco3 we want to filter contours with min size" */

    cand_t *T = NULL; int nT = 0, capT = 0;
    uint8_t *mask = (uint8_t *)malloc(P);
    for (int sc = 0; sc < nScales; sc++) {
        int k = p->adaptiveThreshWinSizeMin + sc * p->adaptiveThreshWinSizeStep;
        orc_adaptive_threshold(gray, W, H, k, p->adaptiveThreshConstant, mask);
        int32_t *pts, *offs;
        int nc = orc_find_contours(mask, W, H, &pts, &offs);
        out->n_contours[sc] = nc;
        for (int ci = 0; ci < nc; ci++) {
            int n = offs[ci + 1] - offs[ci];
            if ((unsigned)n < minPerim || (unsigned)n > maxPerim) continue;
            int32_t *ap = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)n);
            int na = orc_approx_poly_dp(pts + 2 * (size_t)offs[ci], n, (double)n * p->polygonalApproxAccuracyRate, ap);
            if (na == 4 && orc_is_contour_convex(ap, 4)) {
                double minDistSq = (double)maxWH * maxWH;
                for (int j = 0; j < 4; j++) {
                    double ddx = ap[2 * j] - ap[2 * ((j + 1) % 4)], ddy = ap[2 * j + 1] - ap[2 * ((j + 1) % 4) + 1];
                    double dd = ddx * ddx + ddy * ddy;
                    if (dd < minDistSq) minDistSq = dd;
                }
                double mcd = (double)n * p->minCornerDistanceRate;
                if (!(minDistSq < mcd * mcd)) {
                    if (nT == capT) { capT = capT ? capT * 2 : 256; T = (cand_t *)realloc(T, sizeof(cand_t) * (size_t)capT); }
                    cand_t *c = &T[nT++];
                    memset(c, 0, sizeof(*c));
                    for (int j = 0; j < 8; j++) c->c[j] = (float)ap[j];
                    c->len = n; c->parent = -1; c->depth = 0;
                    if (refineMethod == 2) {
                        c->contour = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)n);
                        memcpy(c->contour, pts + 2 * (size_t)offs[ci], sizeof(int32_t) * 2 * (size_t)n);
                    }
                    /* A4: clockwise */
                    double dx1 = c->c[2] - c->c[0], dy1 = c->c[3] - c->c[1];
                    double dx2 = c->c[4] - c->c[0], dy2 = c->c[5] - c->c[1];
                    if (dx1 * dy2 - dy1 * dx2 < 0.0) {
                        float tx = c->c[2], ty = c->c[3];
                        c->c[2] = c->c[6]; c->c[3] = c->c[7]; c->c[6] = tx; c->c[7] = ty;
                    }
                    c->perimeter = perimeter_f(c->c);
                }
            }
            free(ap);
        }
        free(pts); free(offs);
    }
    free(mask);
    out->n_cand = nT;
    out->cand = (float *)malloc(sizeof(float) * 8 * (size_t)(nT ? nT : 1));
    out->cand_len = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nT ? nT : 1));
    for (int i = 0; i < nT; i++) { memcpy(out->cand + 8 * i, T[i].c, sizeof(float) * 8); out->cand_len[i] = T[i].len; }

    /* A5: filterTooCloseCandidates */
    stable_sort_desc(T, nT);
    int *gid = (int *)malloc(sizeof(int) * (size_t)(nT ? nT : 1));
    char *sel = (char *)malloc((size_t)(nT ? nT : 1));
    for (int i = 0; i < nT; i++) { gid[i] = -1; sel[i] = 1; }
    int **groups = NULL; int *gsz = NULL; int ng = 0, capg = 0;
    float rate = (float)p->minMarkerDistanceRate;
    for (int i = 0; i < nT; i++)
        for (int j = i + 1; j < nT; j++) {
            float md = avg_distance(T[i].c, T[j].c);
            if (md < T[j].perimeter * rate) {
                sel[i] = 0; sel[j] = 0;
                int target = -1, add = -1;
                if (gid[i] < 0 && gid[j] < 0) {
                    if (ng == capg) { capg = capg ? capg * 2 : 64; groups = (int **)realloc(groups, sizeof(int *) * (size_t)capg); gsz = (int *)realloc(gsz, sizeof(int) * (size_t)capg); }
                    groups[ng] = (int *)malloc(sizeof(int) * (size_t)nT); gsz[ng] = 0;
                    gid[i] = gid[j] = ng;
                    groups[ng][gsz[ng]++] = i; groups[ng][gsz[ng]++] = j;
                    ng++;
                } else if (gid[i] > -1 && gid[j] == -1) { target = gid[i]; add = j; }
                else if (gid[j] > -1 && gid[i] == -1) { target = gid[j]; add = i; }
                if (target >= 0) { gid[add] = target; groups[target][gsz[target]++] = add; }
            }
        }
    for (int g = 0; g < ng; g++) {
        qsort(groups[g], (size_t)gsz[g], sizeof(int), cmp_int);
        if (p->detectInvertedMarker) for (int a = 0, b = gsz[g] - 1; a < b; a++, b--) { int t = groups[g][a]; groups[g][a] = groups[g][b]; groups[g][b] = t; }
        int cur = groups[g][0];
        sel[cur] = 1;
        T[cur].close = (int *)malloc(sizeof(int) * (size_t)gsz[g]);
        for (int a = 1; a < gsz[g]; a++) {
            int id = groups[g][a];
            float dist = avg_distance(T[id].c, T[cur].c);
            float ms = avg_module_size(T[id].c, d->markerSize, p->markerBorderBits);
            if (dist > p->minGroupDistance * ms) { cur = id; T[groups[g][0]].close[T[groups[g][0]].n_close++] = id; }
        }
    }
    /* selected, minus the ones too near the image border (4.13 placement) */
    int nS = 0;
    int *S = (int *)malloc(sizeof(int) * (size_t)(nT ? nT : 1));
    int mdb = p->minDistanceToBorder;
    for (int i = 0; i < nT; i++) {
        if (!sel[i]) continue;
        int near = 0;
        for (int j = 0; j < 4; j++) {
            float x = T[i].c[2 * j], y = T[i].c[2 * j + 1];
            if (x < mdb || y < mdb || x > W - 1 - mdb || y > H - 1 - mdb) near = 1;
        }
        if (!near) S[nS++] = i;
    }
    for (int i = nS - 1; i >= 0; i--)
        for (int j = i - 1; j >= 0; j--) {
            const float *a = T[S[i]].c, *b = T[S[j]].c;
            if (orc_point_polygon_test(b, 4, a[0], a[1]) >= 0 && orc_point_polygon_test(b, 4, a[2], a[3]) >= 0 &&
                orc_point_polygon_test(b, 4, a[4], a[5]) >= 0 && orc_point_polygon_test(b, 4, a[6], a[7]) >= 0) {
                T[S[i]].parent = j;
                if (T[S[j]].depth < T[S[i]].depth + 1) T[S[j]].depth = T[S[i]].depth + 1;
                break;
            }
        }
    out->n_sel = nS;
    out->sel = (float *)malloc(sizeof(float) * 8 * (size_t)(nS ? nS : 1));
    out->sel_info = (int32_t *)calloc(5 * (size_t)(nS ? nS : 1), sizeof(int32_t));
    for (int i = 0; i < nS; i++) memcpy(out->sel + 8 * i, T[S[i]].c, sizeof(float) * 8);

    /* A6: identifyCandidates, depth 0 first */
    char *valid = (char *)calloc((size_t)(nS ? nS : 1), 1), *was = (char *)calloc((size_t)(nS ? nS : 1), 1);
    int *ids = (int *)malloc(sizeof(int) * (size_t)(nS ? nS : 1)), *rots = (int *)calloc((size_t)(nS ? nS : 1), sizeof(int));
    float *fc = (float *)malloc(sizeof(float) * 8 * (size_t)(nS ? nS : 1));   /* final corners per selected */
    int *fcand = (int *)malloc(sizeof(int) * (size_t)(nS ? nS : 1));          /* the candidate (index into T) they come from */
    for (int i = 0; i < nS; i++) { ids[i] = -1; memcpy(fc + 8 * i, T[S[i]].c, sizeof(float) * 8); fcand[i] = S[i]; }
    int maxDepth = 0;
    for (int i = 0; i < nS; i++) if (T[S[i]].depth > maxDepth) maxDepth = T[S[i]].depth;
    int counter = 0;
    for (int depth = 0; counter < nS && depth <= maxDepth; depth++) {
        for (int v = 0; v < nS; v++) {
            if (T[S[v]].depth != depth) continue;
            was[v] = 1;
            /* ArUco3, "equation (4)": the candidate is read in the pyramid level where its contour is closest to (and longer
             * than) the canonical perimeter; the close candidates of its group are read in the same level */
            const uint8_t *im = gray; int iW = W, iH = H; float scale = 1.f;
            if (a3) {
                int lvl = find_opt_pyr_level(pyrW, numLevels + 1, W, T[S[v]].len, 4 * minSide);
                im = pyr[lvl]; iW = pyrW[lvl]; iH = pyrH[lvl]; scale = (float)iW / (float)W;
            }
            valid[v] = (char)identify_scaled(im, iW, iH, T[S[v]].c, scale, d, p, &ids[v], &rots[v]);
            if (!valid[v]) {
                for (int k = 0; k < T[S[v]].n_close; k++) {
                    const float *cc = T[T[S[v]].close[k]].c;
                    if (identify_scaled(im, iW, iH, cc, scale, d, p, &ids[v], &rots[v])) {
                        valid[v] = 1; memcpy(fc + 8 * v, cc, sizeof(float) * 8); fcand[v] = T[S[v]].close[k]; break;
                    }
                }
            }
        }
        for (int v = 0; v < nS; v++) {
            if (T[S[v]].depth != depth) continue;
            if (valid[v]) {
                int par = T[S[v]].parent;
                while (par != -1) { if (!was[par]) { was[par] = 1; counter++; } par = T[S[par]].parent; }
            }
            counter++;
        }
    }
    /* outputs in S order */
    out->corners = (float *)malloc(sizeof(float) * 8 * (size_t)(nS ? nS : 1));
    out->ids = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nS ? nS : 1));
    out->rejected = (float *)malloc(sizeof(float) * 8 * (size_t)(nS ? nS : 1));
    for (int v = 0; v < nS; v++) {
        out->sel_info[5 * v + 0] = T[S[v]].parent; out->sel_info[5 * v + 1] = T[S[v]].depth;
        out->sel_info[5 * v + 2] = valid[v]; out->sel_info[5 * v + 3] = ids[v]; out->sel_info[5 * v + 4] = rots[v];
        if (valid[v]) {
            float *o = out->corners + 8 * out->n_acc;
            int r = rots[v];
            /* std::rotate(begin, begin + 4 - rot, end): out[j] = in[(j + 4 - rot) % 4] */
            for (int j = 0; j < 4; j++) { o[2 * j] = fc[8 * v + 2 * ((j + 4 - r) % 4)]; o[2 * j + 1] = fc[8 * v + 2 * ((j + 4 - r) % 4) + 1]; }
            /* optional refinement with the contour's side lines (after the rotation, as detectMarkers does) */
            if (refineMethod == 2) orc_refine_candidate_lines(T[fcand[v]].contour, T[fcand[v]].len, o);
            out->ids[out->n_acc++] = ids[v];
        } else {
            memcpy(out->rejected + 8 * out->n_rej, fc + 8 * v, sizeof(float) * 8);
            out->n_rej++;
        }
    }
    /* ArUco3 forces CORNER_REFINE_SUBPIX and refines up the pyramid (findCornerInPyrImage): corners to the level closest to the
     * segmentation image, then level by level x 2 with a 3-pixel (5 above 1080 px) window down to the full-size image */
    if (a3) {
        const float scale_init = (float)pyrW[closestIdx <= numLevels ? closestIdx : numLevels] / (float)W;
        for (int i = 0; i < out->n_acc; i++) {
            float *c = out->corners + 8 * i;
            if (scale_init != 1.f) for (int j = 0; j < 8; j++) c[j] *= scale_init;
            for (int idx = closestIdx - 1; idx >= 0; --idx) {
                for (int j = 0; j < 8; j++) c[j] *= 2.f;
                const int mx = pyrW[idx] > pyrH[idx] ? pyrW[idx] : pyrH[idx];
                orc_corner_subpix(pyr[idx], pyrW[idx], pyrH[idx], c, 4, mx > 1080 ? 5 : 3, p->cornerRefinementMaxIterations, p->cornerRefinementMinAccuracy);
            }
        }
    }
    /* A8: optional sub-pixel refinement of accepted markers */
    if (!a3 && refineMethod == 1) {
        for (int i = 0; i < out->n_acc; i++) {
            float *c = out->corners + 8 * i;
            float per = perimeter_f(c);
            int nm = d->markerSize + 2 * p->markerBorderBits;
            int win = (int)lroundf((float)p->relativeCornerRefinmentWinSize * (per / (4.f * nm)));
            if (win < 1) win = 1;
            if (win > p->cornerRefinementWinSize) win = p->cornerRefinementWinSize;
            orc_corner_subpix(gray, W, H, c, 4, win, p->cornerRefinementMaxIterations, p->cornerRefinementMinAccuracy);
        }
    }
    for (int i = 0; i < nT; i++) { free(T[i].close); free(T[i].contour); }
    free(fcand);
    for (int g = 0; g < ng; g++) free(groups[g]);
    free(groups); free(gsz); free(gid); free(sel); free(S); free(valid); free(was); free(ids); free(rots); free(fc);
    free(T);
    if (gray != pyr[0]) free((void *)gray);
    for (int l = 0; l <= numLevels; l++) free(pyr[l]);
    (void)W0;
    return out->n_acc;
}

void orc_free_detections(orc_detections *o)
{
    free(o->corners); free(o->ids); free(o->rejected); free(o->cand); free(o->cand_len);
    free(o->sel); free(o->sel_info); free(o->n_contours);
    memset(o, 0, sizeof(*o));
}

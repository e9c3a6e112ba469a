// draw_core.h -- cv::aruco::drawDetectedMarkers (reference src/aruco_slam.cpp:319, result kept as markered_img_ :318 and
// returned by getMarkedImg, include/aruco_slam/aruco_slam.h:152) as device / host shared logic.
// Per marker, in this order (OpenCV 4.13 aruco_utils / drawing):
//   1. four cv::line(p_j, p_(j+1), borderColor, 1): 8-connected Bresenham, always walked from the left end point
//      (LineIterator with leftToRight): major-axis step every iteration, minor-axis step when the running error is negative;
//   2. cv::rectangle(p_0 - (3,3), p_0 + (3,3), cornerColor, 1, LINE_AA): four anti-aliased 7-pixel lines (left, top, right,
//      bottom), each a 3 x 8 alpha kernel blended TWICE per pixel, d += ((c - d) a + 127) >> 8;
//   3. cv::putText("id=N", centre, FONT_HERSHEY_SIMPLEX, 0.5, textColor, 2): a fixed bitmap for "id=" and one per digit
//      (10 px apart); textColor / cornerColor are borderColor with channels (0,1) / (1,2) swapped.
// The alpha kernel and the glyph bitmaps are tables recorded from the cv2 4.13.0 wheel (tools/make_overlay_tables.py, which also
// proves that the composition reproduces every string "id=0" .. "id=1023").  Exact when the corners lie inside the image (what
// detectMarkers returns) and are integer valued (CORNER_REFINE_NONE, the reference's setting): OpenCV clips lines that leave the
// image before walking them and shifts the rectangle by the corner's fraction.
#pragma once
#include <math.h>
#include <stdint.h>
#ifndef B2A_HD
#ifdef __CUDACC__
#define B2A_HD __host__ __device__ __forceinline__
#else
#define B2A_HD inline
#endif
#endif

namespace b2a {

struct OverlayTables {
    int aa[24];
    int pre_x0, pre_y0, pre_rows; uint32_t pre[16];
    int dig_x0, dig_y0, dig_rows; uint16_t dig[10][16];
};

inline OverlayTables make_overlay_tables()
{
    OverlayTables t{};
#define B2A_OVERLAY_AA_KERNEL(...) { const int v[] = {__VA_ARGS__}; for (int i = 0; i < 24; ++i) t.aa[i] = v[i]; }
#define B2A_OVERLAY_PREFIX(x0, y0, rows, ...) { t.pre_x0 = x0; t.pre_y0 = y0; t.pre_rows = rows; const uint32_t v[] = {__VA_ARGS__}; for (int i = 0; i < rows; ++i) t.pre[i] = v[i]; }
#define B2A_OVERLAY_DIGITS(x0, y0, rows, ...) { t.dig_x0 = x0; t.dig_y0 = y0; t.dig_rows = rows; const uint16_t v[] = {__VA_ARGS__}; for (int d = 0; d < 10; ++d) for (int i = 0; i < rows; ++i) t.dig[d][i] = v[d * rows + i]; }
#include "../data/overlay_tables.inc"
#undef B2A_OVERLAY_AA_KERNEL
#undef B2A_OVERLAY_PREFIX
#undef B2A_OVERLAY_DIGITS
    return t;
}

struct OverlayImage { uint8_t *data; int W, H, channels; size_t pitch; };

B2A_HD int overlay_round(float v) { return (int)rintf(v); }                 // cv::Point(Point2f): cvRound, ties to even

B2A_HD void overlay_put(const OverlayImage &im, int x, int y, const uint8_t *col)
{
    if ((unsigned)x >= (unsigned)im.W || (unsigned)y >= (unsigned)im.H) return;
    uint8_t *p = im.data + (size_t)y * im.pitch + (size_t)x * im.channels;
    for (int c = 0; c < im.channels; ++c) p[c] = col[c];
}

// cv::line(img, p1, p2, col, 1, LINE_8): the whole segment, walked by one caller
B2A_HD void overlay_line(const OverlayImage &im, int x1, int y1, int x2, int y2, const uint8_t *col)
{
    int dx = x2 - x1, dy = y2 - y1;
    if (dx < 0) { const int tx = x1, ty = y1; x1 = x2; y1 = y2; x2 = tx; y2 = ty; dx = -dx; dy = -dy; }
    const int sy = dy < 0 ? -1 : 1;
    if (dy < 0) dy = -dy;
    const bool steep = dy > dx;
    if (steep) { const int t = dx; dx = dy; dy = t; }
    int err = dx - 2 * dy, x = x1, y = y1;
    for (int i = 0; i <= dx; ++i) {
        overlay_put(im, x, y, col);
        const bool minor = err < 0;
        err += -2 * dy + (minor ? 2 * dx : 0);
        if (steep) { y += sy; if (minor) x += 1; } else { x += 1; if (minor) y += sy; }
    }
}

B2A_HD int overlay_blend2(int d, int c, int a)
{
    d += ((c - d) * a + 127) >> 8;
    d += ((c - d) * a + 127) >> 8;
    return d;
}

// pixel (sx, sy) of the 11 x 11 neighbourhood (offsets -5 .. +5 from the corner) under the anti-aliased 7 x 7 rectangle:
// the four lines in drawing order (left, top, right, bottom)
B2A_HD void overlay_stamp_pixel(const OverlayImage &im, const OverlayTables &t, int cx, int cy, int ox, int oy, const uint8_t *col)
{
    const int x = cx + ox, y = cy + oy;
    if ((unsigned)x >= (unsigned)im.W || (unsigned)y >= (unsigned)im.H) return;
    int a4[4] = {0, 0, 0, 0};
    // vertical lines at ox = -3 (left) and +3 (right): across = ox -+ 3 in -1..1, along = oy + 3 in 0..7
    if (oy + 3 >= 0 && oy + 3 < 8) {
        if (ox + 3 >= -1 && ox + 3 <= 1) a4[0] = t.aa[(ox + 3 + 1) * 8 + oy + 3];
        if (ox - 3 >= -1 && ox - 3 <= 1) a4[2] = t.aa[(ox - 3 + 1) * 8 + oy + 3];
    }
    if (ox + 3 >= 0 && ox + 3 < 8) {
        if (oy + 3 >= -1 && oy + 3 <= 1) a4[1] = t.aa[(oy + 3 + 1) * 8 + ox + 3];
        if (oy - 3 >= -1 && oy - 3 <= 1) a4[3] = t.aa[(oy - 3 + 1) * 8 + ox + 3];
    }
    uint8_t *p = im.data + (size_t)y * im.pitch + (size_t)x * im.channels;
    for (int c = 0; c < im.channels; ++c) {
        int d = p[c];
        for (int k = 0; k < 4; ++k) if (a4[k]) d = overlay_blend2(d, col[c], a4[k]);
        p[c] = (uint8_t)d;
    }
}

// is pixel (tx, ty) relative to the text origin set in "id=<id>"?  (0 <= id <= 9999: up to four digits)
B2A_HD bool overlay_text_bit(const OverlayTables &t, int id, int tx, int ty)
{
    const int py = ty - t.pre_y0, px = tx - t.pre_x0;
    if (py >= 0 && py < t.pre_rows && px >= 0 && px < 32 && ((t.pre[py] >> px) & 1u)) return true;
    const int dy = ty - t.dig_y0;
    if (dy < 0 || dy >= t.dig_rows) return false;
    int nd = 1;
    for (int v = id; v >= 10; v /= 10) ++nd;
    int pw = 1;
    for (int k = 1; k < nd; ++k) pw *= 10;
    for (int k = 0; k < nd; ++k, pw /= 10) {
        const int digit = (id / pw) % 10, dx = tx - (t.dig_x0 + 10 * k);
        if (dx >= 0 && dx < 16 && ((t.dig[digit][dy] >> dx) & 1u)) return true;
    }
    return false;
}

B2A_HD void overlay_colours(const uint8_t *border, uint8_t *text, uint8_t *corner)
{
    text[0] = border[1]; text[1] = border[0]; text[2] = border[2];          // swap(textColor[0], textColor[1])
    corner[0] = border[0]; corner[1] = border[2]; corner[2] = border[1];    // swap(cornerColor[1], cornerColor[2])
}

// text origin: the mean of the four corners (float accumulation, division in double, back to float), rounded
B2A_HD void overlay_text_origin(const float *q, int &ox, int &oy)
{
    float sx = 0.f, sy = 0.f;
    for (int j = 0; j < 4; ++j) { sx += q[2 * j]; sy += q[2 * j + 1]; }
    ox = overlay_round((float)((double)sx / 4.)); oy = overlay_round((float)((double)sy / 4.));
}

// the whole call, one thread (the host emulation of the CPU test tier)
inline void overlay_draw_sequential(const OverlayImage &im, const OverlayTables &t, const float *corners, const int32_t *ids, int n, const uint8_t *border)
{
    uint8_t text[3], corner[3];
    overlay_colours(border, text, corner);
    for (int i = 0; i < n; ++i) {
        const float *q = corners + 8 * i;
        int px[4], py[4];
        for (int j = 0; j < 4; ++j) { px[j] = overlay_round(q[2 * j]); py[j] = overlay_round(q[2 * j + 1]); }
        for (int j = 0; j < 4; ++j) overlay_line(im, px[j], py[j], px[(j + 1) & 3], py[(j + 1) & 3], border);
        for (int oy = -5; oy <= 5; ++oy) for (int ox = -5; ox <= 5; ++ox) overlay_stamp_pixel(im, t, px[0], py[0], ox, oy, corner);
        if (ids) {
            int ox, oy;
            overlay_text_origin(q, ox, oy);
            for (int ty = -13; ty <= 2; ++ty) for (int tx = 0; tx < 72; ++tx)
                if (overlay_text_bit(t, ids[i], tx, ty)) overlay_put(im, ox + tx, oy + ty, text);
        }
    }
}

}  // namespace b2a

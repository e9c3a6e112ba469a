// pyr_core.h -- the arithmetic of the ArUco3 mode of cv::aruco::detectMarkers (DetectorParameters::useAruco3Detection, OpenCV 4.13.0
// aruco_detector.cpp "Step 0 / Step 1" and identifyCandidates' "equation (4)"; inside the detectMarkers(image, dictionary, params)
// surface of reference src/aruco_slam.cpp:313), per output pixel / per candidate, for the device (k_pyr_down, k_resize_linear,
// k_homography, k_identify, the refinement chain in b2a_api.cu) and for the host emulation of the CPU tests:
//   pyr_down_pixel      cv::pyrDown, 8-bit: [1 4 6 4 1]^2 / 256, BORDER_REFLECT_101, (sum + 128) >> 8
//   resize_tab          the source index and the two 11-bit weights of one destination coordinate of cv::resize(INTER_LINEAR)
//   resize_pixel        one destination pixel of that resize (the exact 2 x 2 reduction is the area average, as resize() makes it)
//   Aruco3Plan          factor of the segmentation image, pyramid depth, the level the refinement starts from
//   pyr_opt_level       _findOptPyrImageForCanonicalImg
#pragma once
#include "core.h"

namespace b2a {

constexpr int PYR_MAX_LEVELS = 12;       // level 0 = the full-size gray image

B2A_HD int pyr_reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}
// the same for -2 <= p <= len + 1 and len >= 3 (every index pyrDown forms on an image of at least 3 pixels): one reflection, no loop
B2A_HD int pyr_reflect101_near(int p, int len)
{
    p = p < 0 ? -p : p;
    return p >= len ? 2 * (len - 1) - p : p;
}

// destination pixel (dx, dy) of pyrDown of a W x H image
B2A_HD uint8_t pyr_down_pixel(const uint8_t *src, int W, int H, size_t pitch, int dx, int dy)
{
    int xs[5];
    B2A_UNROLL
    for (int k = 0; k < 5; ++k) xs[k] = pyr_reflect101(2 * dx + k - 2, W);
    int v = 0;
    B2A_UNROLL
    for (int k = 0; k < 5; ++k) {
        const uint8_t *s = src + (size_t)pyr_reflect101(2 * dy + k - 2, H) * pitch;
        const int row = (int)s[xs[0]] + (int)s[xs[4]] + 4 * ((int)s[xs[1]] + (int)s[xs[3]]) + 6 * (int)s[xs[2]];
        v += (k == 0 || k == 4) ? row : (k == 2 ? 6 * row : 4 * row);
    }
    return (uint8_t)((v + 128) >> 8);
}

B2A_HD int pyr_round_half_even(float v)
{
#if defined(__CUDA_ARCH__)
    return __float2int_rn(v);
#else
    return (int)lrintf(v);
#endif
}

// resize.cpp, INTER_LINEAR table set-up: fx = (float)((d + 0.5) * scale - 0.5) with scale = 1 / (dlen / slen) in double,
// weights saturate_cast<short>(w * 2048) (round half to even)
B2A_HD void resize_tab(int d, int dlen, int slen, int &s, int &c0, int &c1)
{
    const double scale = d_div(1.0, d_div((double)dlen, (double)slen));
    float f = (float)d_sub(d_mul((double)d + 0.5, scale), 0.5);
    s = (int)floorf(f);
    f = f - (float)s;
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= slen - 1) { f = 0.f; s = slen - 1; }
    c0 = pyr_round_half_even(f_mul(1.f - f, 2048.f));
    c1 = pyr_round_half_even(f_mul(f, 2048.f));
}

// one destination pixel from its two source rows s0, s1 (weights b0, b1) and its source column sx (weights a0, a1)
B2A_HD uint8_t resize_pixel_rows(const uint8_t *s0, const uint8_t *s1, int W, int sx, int a0, int a1, int b0, int b1)
{
    const int sx1 = sx + 1 < W ? sx + 1 : W - 1;
    const int r0 = (int)s0[sx] * a0 + (int)s0[sx1] * a1, r1 = (int)s1[sx] * a0 + (int)s1[sx1] * a1;
    const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
    return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
}
// the same from the two table entries (source column sx with weights a0, a1; source row sy with b0, b1)
B2A_HD uint8_t resize_pixel_tab(const uint8_t *src, int W, int H, size_t pitch, int sx, int a0, int a1, int sy, int b0, int b1)
{
    const int sy1 = sy + 1 < H ? sy + 1 : H - 1;
    return resize_pixel_rows(src + (size_t)sy * pitch, src + (size_t)sy1 * pitch, W, sx, a0, a1, b0, b1);
}

// destination pixel (dx, dy) of cv::resize(src W x H -> dW x dH, INTER_LINEAR)
B2A_HD uint8_t resize_pixel(const uint8_t *src, int W, int H, size_t pitch, int dW, int dH, int dx, int dy)
{
    if (W == 2 * dW && H == 2 * dH) {
        const uint8_t *s = src + (size_t)(2 * dy) * pitch + 2 * dx;
        return (uint8_t)(((int)s[0] + (int)s[1] + (int)s[pitch] + (int)s[pitch + 1] + 2) >> 2);
    }
    int sx, a0, a1, sy, b0, b1;
    resize_tab(dx, dW, W, sx, a0, a1);
    resize_tab(dy, dH, H, sy, b0, b1);
    return resize_pixel_tab(src, W, H, pitch, sx, a0, a1, sy, b0, b1);
}

// sum of the four bytes of a weighted by the four bytes of b (dp4a.u32.u32)
B2A_HD uint32_t pyr_dp4a(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __dp4a(a, b, c);
#else
    for (int k = 0; k < 4; ++k) c += ((a >> (8 * k)) & 255u) * ((b >> (8 * k)) & 255u);
    return c;
#endif
}

// four neighbouring destination pixels (dx0 .. dx0 + 3, dy; dx0 a multiple of 4) of pyrDown whose 5 x 11 source window lies inside
// the image: the same sums as pyr_down_pixel without the border reflection, packed into one word (byte 0 = dx0).
// ALIGNED (rows of src start on word boundaries): a row's 11 bytes come as four words starting at column 2 dx0 - 4 and every
// horizontal [1 4 6 4 1] is two dp4a; otherwise byte loads.
template <bool ALIGNED>
B2A_HD uint32_t pyr_down_interior4(const uint8_t *src, size_t pitch, int dx0, int dy)
{
    uint32_t v[4] = {0, 0, 0, 0};
    B2A_UNROLL
    for (int k = 0; k < 5; ++k) {
        const uint32_t wk = (k == 0 || k == 4) ? 1u : (k == 2 ? 6u : 4u);
        const uint8_t *row = src + (size_t)(2 * dy + k - 2) * pitch;
        if (ALIGNED) {
            const uint32_t *w = reinterpret_cast<const uint32_t *>(row + (2 * dx0 - 4));
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
            // window byte n = source column 2 dx0 - 4 + n; pixel q needs bytes 2 q + 2 .. 2 q + 6 with weights 1 4 6 4 1
            const uint32_t h0 = pyr_dp4a(w0, 0x04010000u, pyr_dp4a(w1, 0x00010406u, 0u));
            const uint32_t h1 = pyr_dp4a(w1, 0x04060401u, pyr_dp4a(w2, 0x00000001u, 0u));
            const uint32_t h2 = pyr_dp4a(w1, 0x04010000u, pyr_dp4a(w2, 0x00010406u, 0u));
            const uint32_t h3 = pyr_dp4a(w2, 0x04060401u, pyr_dp4a(w3, 0x00000001u, 0u));
            v[0] += wk * h0; v[1] += wk * h1; v[2] += wk * h2; v[3] += wk * h3;
        } else {
            const uint8_t *s = row + (2 * dx0 - 2);
            uint32_t t[11];
            B2A_UNROLL
            for (int i = 0; i < 11; ++i) t[i] = s[i];
            B2A_UNROLL
            for (int q = 0; q < 4; ++q) v[q] += wk * (t[2 * q] + t[2 * q + 4] + 4 * (t[2 * q + 1] + t[2 * q + 3]) + 6 * t[2 * q + 2]);
        }
    }
    return ((v[0] + 128) >> 8) | (((v[1] + 128) >> 8) << 8) | (((v[2] + 128) >> 8) << 16) | (((v[3] + 128) >> 8) << 24);
}
// up to four neighbouring destination pixels (dx0 .. min(dx0 + 3, dW - 1), dy) anywhere in an image of at least 3 x 3 pixels: the
// 11 source columns and 5 source rows are reflected once each, then the sums of pyr_down_pixel; bytes of missing pixels are 0
B2A_HD uint32_t pyr_down_border4(const uint8_t *src, int W, int H, size_t pitch, int dW, int dx0, int dy)
{
    int xs[11];
    B2A_UNROLL
    for (int i = 0; i < 11; ++i) { const int p = 2 * dx0 - 2 + i; xs[i] = pyr_reflect101_near(p > W + 1 ? W + 1 : p, W); }     // columns past W + 1 belong to pixels >= dW
    uint32_t v[4] = {0, 0, 0, 0};
    B2A_UNROLL
    for (int k = 0; k < 5; ++k) {
        const uint32_t wk = (k == 0 || k == 4) ? 1u : (k == 2 ? 6u : 4u);
        const uint8_t *s = src + (size_t)pyr_reflect101_near(2 * dy + k - 2, H) * pitch;
        uint32_t t[11];
        B2A_UNROLL
        for (int i = 0; i < 11; ++i) t[i] = s[xs[i]];
        B2A_UNROLL
        for (int q = 0; q < 4; ++q) v[q] += wk * (t[2 * q] + t[2 * q + 4] + 4 * (t[2 * q + 1] + t[2 * q + 3]) + 6 * t[2 * q + 2]);
    }
    uint32_t word = 0;
    B2A_UNROLL
    for (int q = 0; q < 4; ++q) if (dx0 + q < dW) word |= ((v[q] + 128) >> 8) << (8 * q);
    return word;
}
B2A_HD bool pyr_down_is_interior4(int W, int H, int dW, int dx0, int dy)
{
    return dx0 + 3 < dW && 2 * dx0 - 2 >= 0 && 2 * (dx0 + 3) + 2 < W && 2 * dy - 2 >= 0 && 2 * dy + 2 < H;
}
// the same test as ranges: groups gx (dx0 = 4 gx) with 1 <= gx < pyr_interior_gx_end and rows 1 <= dy < pyr_interior_y_end are interior
B2A_HD int pyr_interior_gx_end(int W, int dW)
{
    const int a = dW >= 4 ? (dW - 4) / 4 : -1, b = W >= 9 ? (W - 9) / 8 : -1;
    const int m = a < b ? a : b;
    return m + 1 > 1 ? m + 1 : 1;
}
B2A_HD int pyr_interior_y_end(int H)
{
    const int m = H >= 3 ? (H - 3) / 2 : -1;
    return m + 1 > 1 ? m + 1 : 1;
}

// what detectMarkers derives from the image size and the two ArUco3 parameters before it looks at a pixel
struct Aruco3Plan {
    float fxfy;           // size factor of the segmentation image ("equation (2)")
    int segW, segH;       // its size (the image itself when fxfy == 1)
    int numLevels;        // pyrDown steps: the pyramid has numLevels + 1 images
    int closestIdx;       // the level the refinement chain starts from
    int W[PYR_MAX_LEVELS], H[PYR_MAX_LEVELS];
};

// returns false when the pyramid would need more than PYR_MAX_LEVELS images or closestIdx lies past its last level (cv2 then
// indexes its pyramid vector out of range)
inline bool aruco3_plan(int W, int H, int minSide, float ratio, Aruco3Plan &p)
{
    p.fxfy = (float)minSide / ((float)minSide + (float)(W > H ? W : H) * ratio);
    const float img_area = (float)(H * W), min_area_marker = (float)(minSide * minSide);
    p.numLevels = (int)(log2f(img_area / min_area_marker) / 2.f);
    const float scale_img_area = img_area * p.fxfy * p.fxfy;
    p.closestIdx = (int)lrintf(log2f(img_area / scale_img_area) / 2.f);
    if (p.numLevels < 0) p.numLevels = 0;
    if (p.numLevels + 1 > PYR_MAX_LEVELS || p.closestIdx < 0 || p.closestIdx > p.numLevels) return false;
    p.W[0] = W; p.H[0] = H;
    for (int l = 1; l <= p.numLevels; ++l) { p.W[l] = (p.W[l - 1] + 1) / 2; p.H[l] = (p.H[l - 1] + 1) / 2; }
    p.segW = W; p.segH = H;
    if (p.fxfy != 1.f) { p.segW = (int)lrintf(p.fxfy * (float)W); p.segH = (int)lrintf(p.fxfy * (float)H); }
    return p.segW >= 1 && p.segH >= 1;
}

// the pyramid as the identification kernels see it: level 0 is the full-size gray image
struct PyrLevels {
    int n;                                   // 0 = ArUco3 off (candidates are read in the image they were found in)
    int segW;                                // width of the segmentation image
    int minPerimeter;                        // 4 * minSideLengthCanonicalImg
    int W[PYR_MAX_LEVELS], H[PYR_MAX_LEVELS];
    const uint8_t *base[PYR_MAX_LEVELS];     // frame 0 of the sub-batch
    size_t pitch[PYR_MAX_LEVELS], frame_stride[PYR_MAX_LEVELS];
};

// _findOptPyrImageForCanonicalImg: the level whose scaled contour length exceeds min_perimeter by the least (level 0 when none does)
B2A_HD int pyr_opt_level(const int *levelW, int n, int scaled_width, int cur_perimeter, int min_perimeter)
{
    int opt = 0;
    float dist = FLT_MAX;
    for (int i = 0; i < n; ++i) {
        const float scale = f_div((float)levelW[i], (float)scaled_width);
        const float new_dist = f_mul((float)cur_perimeter, scale) - (float)min_perimeter;
        if (new_dist < dist && new_dist > 0.f) { dist = new_dist; opt = i; }
    }
    return opt;
}

// _identifyOneCandidate's `scale`: the quad in the coordinates of the level it is read in
B2A_HD void pyr_scale_quad(const float *c, float scale, float *out)
{
    B2A_UNROLL
    for (int k = 0; k < 8; ++k) out[k] = f_mul(c[k], scale);
}

}  // namespace b2a

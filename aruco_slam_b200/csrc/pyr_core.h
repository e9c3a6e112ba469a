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
    const int sx1 = sx + 1 < W ? sx + 1 : W - 1, sy1 = sy + 1 < H ? sy + 1 : H - 1;
    const uint8_t *s0 = src + (size_t)sy * pitch, *s1 = src + (size_t)sy1 * pitch;
    const int r0 = (int)s0[sx] * a0 + (int)s0[sx1] * a1, r1 = (int)s1[sx] * a0 + (int)s1[sx1] * a1;
    const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
    return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
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

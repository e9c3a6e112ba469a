/* orc_pyramid.c -- CPU restatement of the two image operations cv::aruco::detectMarkers adds when
 * DetectorParameters::useAruco3Detection is set (OpenCV 4.13.0 aruco_detector.cpp, "Step 1: create image pyramid" and
 * "resize to segmentation image"): cv::pyrDown (buildPyramid) and cv::resize(INTER_LINEAR) on 8-bit single-channel images.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Pinned against the cv2 4.13.0 wheel by tests/golden/aruco3.npz
 * (tools/make_golden_aruco3.py) and by random images in tests/test_oracle_aruco3.py's fixtures.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* cv::borderInterpolate(p, len, BORDER_REFLECT_101) */
static int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

/* cv::pyrDown, 8-bit: separable [1 4 6 4 1] / 16 in both directions, integer sums, (sum + 128) >> 8, destination
 * ((W + 1) / 2, (H + 1) / 2), BORDER_REFLECT_101 */
void orc_pyr_down(const uint8_t *src, int W, int H, uint8_t *dst)
{
    const int dW = (W + 1) / 2, dH = (H + 1) / 2;
    int *rows = (int *)malloc(sizeof(int) * (size_t)dW * 5);
    for (int dy = 0; dy < dH; dy++) {
        for (int k = 0; k < 5; k++) {
            const uint8_t *s = src + (size_t)reflect101(2 * dy + k - 2, H) * W;
            int *r = rows + (size_t)k * dW;
            for (int dx = 0; dx < dW; dx++) {
                int x0 = reflect101(2 * dx - 2, W), x1 = reflect101(2 * dx - 1, W), x2 = 2 * dx, x3 = reflect101(2 * dx + 1, W), x4 = reflect101(2 * dx + 2, W);
                r[dx] = s[x0] + s[x4] + 4 * (s[x1] + s[x3]) + 6 * s[x2];
            }
        }
        for (int dx = 0; dx < dW; dx++) {
            int v = rows[dx] + rows[4 * dW + dx] + 4 * (rows[dW + dx] + rows[3 * dW + dx]) + 6 * rows[2 * dW + dx];
            dst[(size_t)dy * dW + dx] = (uint8_t)((v + 128) >> 8);
        }
    }
    free(rows);
}

/* saturate_cast<short>(float): round half to even */
static short sat_short(float v)
{
    long r = lrintf(v);
    return (short)(r < -32768 ? -32768 : r > 32767 ? 32767 : r);
}

/* the source index and the two 11-bit weights of one destination coordinate (resize.cpp, the INTER_LINEAR branch of the table set-up) */
static void linear_tab(int dlen, int slen, int *ofs, short *coef)
{
    const double scale = 1.0 / ((double)dlen / (double)slen);
    for (int d = 0; d < dlen; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= slen - 1) { f = 0.f; s = slen - 1; }
        ofs[d] = s;
        coef[2 * d] = sat_short((1.f - f) * 2048.f);
        coef[2 * d + 1] = sat_short(f * 2048.f);
    }
}

/* cv::resize(src, dst, Size(dW, dH), 0, 0, INTER_LINEAR), 8-bit single channel.  An exact 2 x 2 reduction is the area average
 * (resize() switches INTER_LINEAR to INTER_AREA when both scale factors are exactly 2); everything else is the fixed-point
 * bilinear: horizontal pass into 32-bit rows with 11-bit weights, vertical pass
 * ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2. */
void orc_resize_linear(const uint8_t *src, int W, int H, uint8_t *dst, int dW, int dH)
{
    if (dW == W && dH == H) { memcpy(dst, src, (size_t)W * H); return; }
    if (W == 2 * dW && H == 2 * dH) {
        for (int y = 0; y < dH; y++)
            for (int x = 0; x < dW; x++) {
                const uint8_t *s = src + (size_t)(2 * y) * W + 2 * x;
                dst[(size_t)y * dW + x] = (uint8_t)((s[0] + s[1] + s[W] + s[W + 1] + 2) >> 2);
            }
        return;
    }
    int *xofs = (int *)malloc(sizeof(int) * (size_t)dW), *yofs = (int *)malloc(sizeof(int) * (size_t)dH);
    short *alpha = (short *)malloc(sizeof(short) * 2 * (size_t)dW), *beta = (short *)malloc(sizeof(short) * 2 * (size_t)dH);
    linear_tab(dW, W, xofs, alpha);
    linear_tab(dH, H, yofs, beta);
    int *r0 = (int *)malloc(sizeof(int) * (size_t)dW), *r1 = (int *)malloc(sizeof(int) * (size_t)dW);
    for (int y = 0; y < dH; y++) {
        int sy0 = yofs[y], sy1 = sy0 + 1 < H ? sy0 + 1 : H - 1;
        const uint8_t *s0 = src + (size_t)sy0 * W, *s1 = src + (size_t)sy1 * W;
        for (int x = 0; x < dW; x++) {
            int sx = xofs[x], sx1 = sx + 1 < W ? sx + 1 : W - 1;
            r0[x] = s0[sx] * alpha[2 * x] + s0[sx1] * alpha[2 * x + 1];
            r1[x] = s1[sx] * alpha[2 * x] + s1[sx1] * alpha[2 * x + 1];
        }
        int b0 = beta[2 * y], b1 = beta[2 * y + 1];
        for (int x = 0; x < dW; x++) {
            int v = (((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2;
            dst[(size_t)y * dW + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    free(xofs); free(yofs); free(alpha); free(beta); free(r0); free(r1);
}

#!/usr/bin/env python
"""Write aruco_slam_b200/data/overlay_tables.inc and tests/golden/draw_*.npz from the cv2 4.13.0 wheel (authoring container only).

cv::aruco::drawDetectedMarkers (reference src/aruco_slam.cpp:319) = per marker four cv::line calls (thickness 1, 8-connected), a
7 x 7 anti-aliased cv::rectangle on the first corner and cv::putText("id=N", FONT_HERSHEY_SIMPLEX, 0.5, thickness 2).  OpenCV's
rasterisers are not in /root/reference; what is recorded here from the wheel is
  * the alpha kernel of a 7-pixel anti-aliased line (3 x 8 values; the rectangle is four of them, blended twice per pixel as
    OpenCV's LineAA does: d += ((c - d) a + 127) >> 8, twice),
  * the Hershey-simplex bitmaps of the prefix "id=" and of the ten digits at the sub-pixel phase they always have after that prefix
    (composition of these reproduces every string "id=0" .. "id=1023" bit for bit -- checked below),
and golden overlays of detected frames.  Run: python tools/make_overlay_tables.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from aruco_slam_b200 import synth  # noqa: E402

F = cv2.FONT_HERSHEY_SIMPLEX
ORG = (40, 40)


def blend2(d, c, a):
    d = d + (((c - d) * a + 127) >> 8)
    return d + (((c - d) * a + 127) >> 8)


def render(s):
    im = np.zeros((80, 220), np.uint8)
    cv2.putText(im, s, ORG, F, 0.5, 255, 2)
    return im > 0


def main():
    inv = {}
    for a in range(256):
        inv.setdefault(blend2(0, 255, a), a)
    im = np.zeros((30, 30), np.uint8)
    cv2.line(im, (10, 10), (16, 10), 255, 1, cv2.LINE_AA)
    ker = np.array([[inv[int(v)] if v else 0 for v in im[9 + r, 10:18]] for r in range(3)])
    imv = np.zeros((30, 30), np.uint8)
    cv2.line(imv, (10, 10), (10, 16), 255, 1, cv2.LINE_AA)
    assert np.array_equal(np.array([[inv[int(v)] if v else 0 for v in imv[10:18, 9 + r]] for r in range(3)]), ker), "vertical = transposed horizontal"
    assert not im[:9].any() and not im[12:].any() and not im[:, :10].any() and not im[:, 18:].any()
    pre = render("id=")
    ys, xs = np.nonzero(pre)
    px0, px1, py0, py1 = xs.min() - ORG[0], xs.max() - ORG[0], ys.min() - ORG[1], ys.max() - ORG[1]
    digs = {d: render("id=" + d) & ~pre for d in "0123456789"}
    dx0 = min(np.nonzero(digs[d])[1].min() for d in digs) - ORG[0]
    dx1 = max(np.nonzero(digs[d])[1].max() for d in digs) - ORG[0]
    dy0 = min(np.nonzero(digs[d])[0].min() for d in digs) - ORG[1]
    dy1 = max(np.nonzero(digs[d])[0].max() for d in digs) - ORG[1]
    for n in range(1024):                                  # the composition is exact for every id of every predefined dictionary
        comp = pre.copy()
        for k, ch in enumerate(str(n)):
            comp |= np.roll(digs[ch], 10 * k, axis=1)
        assert np.array_equal(comp, render("id=%d" % n)), n
    assert px1 - px0 < 32 and dx1 - dx0 < 16

    def rows(bm, x0, x1, y0, y1):
        out = []
        for y in range(y0, y1 + 1):
            v = 0
            for x in range(x0, x1 + 1):
                if bm[ORG[1] + y, ORG[0] + x]:
                    v |= 1 << (x - x0)
            out.append(v)
        return out

    with open(os.path.join(ROOT, "aruco_slam_b200", "data", "overlay_tables.inc"), "w") as f:
        f.write("// written by tools/make_overlay_tables.py from cv2 %s -- do not edit\n" % cv2.__version__)
        f.write("// alpha of a 7-pixel anti-aliased line: rows -1, 0, +1 across the line, 8 positions along it (the last is the one-pixel tail)\n")
        f.write("B2A_OVERLAY_AA_KERNEL(%s)\n" % ", ".join(str(int(v)) for v in ker.ravel()))
        f.write("// \"id=\" (Hershey simplex, scale 0.5, thickness 2): x0, y0 relative to the text origin, rows, then one 32-bit mask per row (bit k = x0 + k)\n")
        f.write("B2A_OVERLAY_PREFIX(%d, %d, %d, %s)\n" % (px0, py0, py1 - py0 + 1, ", ".join("0x%xu" % v for v in rows(pre, px0, px1, py0, py1))))
        f.write("// digits at the phase they have after that prefix: x0 (first digit; each further digit +10), y0, rows, then per digit one 16-bit mask per row\n")
        f.write("B2A_OVERLAY_DIGITS(%d, %d, %d, %s)\n" % (dx0, dy0, dy1 - dy0 + 1,
                                                          ", ".join("0x%x" % v for d in "0123456789" for v in rows(digs[d], dx0, dx1, dy0, dy1))))
    print("tables written: aa kernel", ker.tolist(), "prefix box", (px0, px1, py0, py1), "digit box", (dx0, dx1, dy0, dy1))

    # ---- golden overlays ----
    out = os.path.join(ROOT, "tests", "golden")
    prov = "cv2 %s cv2.aruco.drawDetectedMarkers, tools/make_overlay_tables.py" % cv2.__version__

    def detect(img, dict_id):
        det = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(dict_id), cv2.aruco.DetectorParameters())
        c, ids, _ = det.detectMarkers(img)
        return c, ids

    # colour inputs are regenerated in the tests from the stored gray frame (synth.gray_to_bgr(gray, seed)); outputs are stored as the
    # pixels that differ from the input (flat index + value)
    cases = {}
    cases["vga_gray"] = (synth.render_config("C1", 2).image, -1, 0)
    cases["vga_bgr"] = (synth.render_config("C1", 5).image, 3, 0)
    cases["540p_orig_bgr"] = (synth.render_frame(960, 540, 14, 16, seed=9, side_range=(50.0, 110.0)).image, 1, 16)      # ids up to 1023, overlapping labels
    near = synth.background(200, 160)
    dic16 = synth.getPredefinedDictionary(16)
    synth.paste_marker(near, dic16, 1017, np.array([[9.0, 9.0], [70.0, 11.0], [68.0, 72.0], [8.0, 69.0]]))
    synth.paste_marker(near, dic16, 5, np.array([[120.0, 85.0], [188.0, 90.0], [186.0, 150.0], [118.0, 148.0]]))
    cases["near_border_gray"] = (np.clip(np.rint(near), 0, 255).astype(np.uint8), -1, 16)

    def sparse(drawn, img):
        idx = np.flatnonzero(drawn.ravel() != img.ravel())
        return idx.astype(np.int32), drawn.ravel()[idx].copy()

    for name, (gray, bgr_seed, dict_id) in cases.items():
        img = gray if bgr_seed < 0 else synth.gray_to_bgr(gray, bgr_seed)
        c, ids = detect(img, dict_id)
        assert ids is not None and len(ids) > 0, name
        kw = dict(provenance=np.array(prov), gray=gray, bgr_seed=bgr_seed, corners=np.array(c, np.float32).reshape(-1, 4, 2), ids=ids.ravel().astype(np.int32),
                  colour=np.array([200, 30, 90], np.uint8))
        for key, args in (("ids", (c, ids)), ("no_ids", (c,)), ("colour", (c, ids, (200, 30, 90)))):
            d = img.copy()
            cv2.aruco.drawDetectedMarkers(d, *args)
            kw["idx_" + key], kw["val_" + key] = sparse(d, img)
        np.savez_compressed(os.path.join(out, "draw_%s.npz" % name), **kw)
        print("draw_%-18s %d markers, ids %s, %d changed values" % (name, len(ids), ids.ravel().tolist()[:12], len(kw["idx_ids"])))


if __name__ == "__main__":
    main()

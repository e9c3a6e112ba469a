#!/usr/bin/env python
"""Write tests/golden/aruco3.npz from the cv2 4.13.0 wheel (authoring container only): `ArucoDetector.detectMarkers` with
`useAruco3Detection = True` on the frames the `detect_*.npz` fixtures already hold, for several settings of the two ArUco3
parameters (keys `<fixture>/<case>/corners | ids | rejected`; the accepted corners are in full-size image coordinates, the
rejected candidates stay in the coordinates of the reduced segmentation image -- cv2 does not scale them back), and the two
image operations the mode adds, on small random images: `cv2.pyrDown` (keys `pyr/<i>/src | dst`) and
`cv2.resize(..., INTER_LINEAR)` (keys `resize/<i>/src | dst`, one case an exact 2 x 2 reduction, which cv2 turns into the area
average).

Run:  python tools/make_golden_aruco3.py        (needs cv2)
"""
import glob
import os
import sys

import numpy as np
import cv2

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
A = cv2.aruco
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
PROV = "cv2 %s (opencv-python-headless), tools/make_golden_aruco3.py" % cv2.__version__

# (name, minMarkerLengthRatioOriginalImg, minSideLengthCanonicalImg, extra parameters)
CASES = [("default", 0.0, 32, {}), ("r010", 0.01, 32, {}), ("r020", 0.02, 32, {}), ("r035_s16", 0.035, 16, {}), ("r030_s64", 0.03, 64, {}),
         ("half", None, 32, {}),                                 # the ratio that makes the factor exactly 0.5
         ("r015_inv", 0.015, 32, {"detectInvertedMarker": True}),
         ("r020_contour", 0.02, 32, {"cornerRefinementMethod": 2})]    # ArUco3 forces SUBPIX whatever the method says


def main():
    rng = np.random.default_rng(33)
    kw = {}
    shapes = [(37, 53), (64, 64), (5, 9), (2, 3), (1, 7), (121, 80)]
    for i, (h, w) in enumerate(shapes):
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        kw["pyr/%d/src" % i] = src
        kw["pyr/%d/dst" % i] = cv2.pyrDown(src)
    sizes = [((90, 120), (43, 57)), ((90, 120), (45, 60)), ((90, 120), (81, 119)), ((33, 47), (9, 13)), ((60, 80), (20, 27)), ((64, 48), (3, 2)), ((50, 50), (50, 49))]
    for i, ((h, w), (dh, dw)) in enumerate(sizes):
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        kw["resize/%d/src" % i] = src
        kw["resize/%d/dst" % i] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    names = []
    for path in sorted(glob.glob(os.path.join(OUT, "detect_*.npz")) + glob.glob(os.path.join(OUT, "inverted_vga_4x4_half.npz"))):
        name = os.path.basename(path)[:-4]
        g = np.load(path)
        frame = g["frame"]
        names.append(name)
        for case, ratio, side, extra in CASES:
            prm = A.DetectorParameters()
            prm.useAruco3Detection = True
            prm.minSideLengthCanonicalImg = side
            prm.minMarkerLengthRatioOriginalImg = side / max(frame.shape[:2]) if ratio is None else ratio
            for k, v in extra.items():
                setattr(prm, k, v)
            c, ids, rej = A.ArucoDetector(A.getPredefinedDictionary(int(g["dict_id"])), prm).detectMarkers(frame)
            key = name + "/" + case
            kw[key + "/corners"] = np.array(c, np.float32).reshape(-1, 4, 2)
            kw[key + "/ids"] = np.zeros(0, np.int32) if ids is None else ids.ravel().astype(np.int32)
            kw[key + "/rejected"] = np.array(rej, np.float32).reshape(-1, 4, 2)
            kw[key + "/ratio"] = np.float32(prm.minMarkerLengthRatioOriginalImg)
            print("%-32s %-13s %3d markers, %3d rejected" % (name, case, len(kw[key + "/ids"]), len(kw[key + "/rejected"])))
    path = os.path.join(OUT, "aruco3.npz")
    np.savez_compressed(path, provenance=np.array(PROV), fixtures=np.array(names), cases=np.array([c[0] for c in CASES]),
                        sides=np.array([c[2] for c in CASES], np.int32), n_pyr=np.int32(len(shapes)), n_resize=np.int32(len(sizes)), **kw)
    print("aruco3.npz %.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()

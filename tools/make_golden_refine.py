#!/usr/bin/env python
"""Write tests/golden/contour_refine.npz from the cv2 4.13.0 wheel (authoring container only): what
`ArucoDetector.detectMarkers` returns with `cornerRefinementMethod = CORNER_REFINE_CONTOUR` on the frames the
`detect_*.npz` fixtures already hold (keys `<fixture>/corners`, `<fixture>/ids`, `<fixture>/rejected`).

cv2 fits the side lines in float32 normal equations and computes A^T B with its BLAS once a side has 100 points or
more, so its corners carry up to a few hundredths of a pixel of build-dependent rounding there; the tests ask for exact
equality where every side of a marker is shorter than 90 px and for 0.05 px elsewhere.

Run:  python tools/make_golden_refine.py        (needs cv2)
"""
import glob
import os
import sys

import numpy as np
import cv2

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
A = cv2.aruco
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
PROV = "cv2 %s (opencv-python-headless), tools/make_golden_refine.py" % cv2.__version__


def main():
    kw = {}
    names = []
    for path in sorted(glob.glob(os.path.join(OUT, "detect_*.npz"))):
        name = os.path.basename(path)[:-4]
        g = np.load(path)
        prm = A.DetectorParameters()
        prm.cornerRefinementMethod = A.CORNER_REFINE_CONTOUR
        c, ids, rej = A.ArucoDetector(A.getPredefinedDictionary(int(g["dict_id"])), prm).detectMarkers(g["frame"])
        kw[name + "/corners"] = np.array(c, np.float32).reshape(-1, 4, 2)
        kw[name + "/ids"] = np.zeros(0, np.int32) if ids is None else ids.ravel().astype(np.int32)
        kw[name + "/rejected"] = np.array(rej, np.float32).reshape(-1, 4, 2)
        names.append(name)
        print("%-32s %3d markers" % (name, len(kw[name + "/ids"])))
    path = os.path.join(OUT, "contour_refine.npz")
    np.savez_compressed(path, provenance=np.array(PROV), fixtures=np.array(names), **kw)
    print("contour_refine.npz %.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()

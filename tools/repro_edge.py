"""debug helper: front end on ragged noise frames, one shape at a time"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aruco_slam_b200 import aruco, dictionaries as D
rng = np.random.default_rng(3)
dic = D.getPredefinedDictionary(0)
for (H, W) in ((31, 33), (64, 129), (95, 257), (130, 64), (7, 300)):
    img = rng.integers(0, 256, (H, W)).astype(np.uint8)
    det = aruco.ArucoDetector(dic, max_shape=(H, W), max_batch=1, device=0)
    try:
        det.debug_threshold(img)
        print(H, W, "ok", flush=True)
    except Exception as e:
        print(H, W, "FAIL", e, flush=True)
        break
    det.close()
img = rng.integers(0, 256, (120, 200)).astype(np.uint8)
p = aruco.DetectorParameters(adaptiveThreshWinSizeMin=5, adaptiveThreshWinSizeMax=29, adaptiveThreshWinSizeStep=8, adaptiveThreshConstant=3.0)
det = aruco.ArucoDetector(dic, p, max_shape=img.shape, max_batch=1, device=0)
try:
    det.debug_threshold(img)
    print("4 scales ok", flush=True)
except Exception as e:
    print("4 scales FAIL", e, flush=True)

"""Short driver for ncu: a few detect+pose calls on the C2 batch (32 x 1080p, host frames).
usage: python tools/profile_step.py [calls] [batch]"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aruco_slam_b200 import aruco, synth, dictionaries as D

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
K = np.array([[1400.0, 0, 960], [0, 1400.0, 540], [0, 0, 1]])
Dc = np.array([0.05, -0.1, 0.001, -0.002, 0.02])
frames = synth.render_batch("C2", B, base_seed=0)
det = aruco.ArucoDetector(D.getPredefinedDictionary(D.DICT_6X6_250), max_shape=frames.shape[1:], max_batch=B, device=0)
for _ in range(calls):
    r = det.detect_pose_batch(frames, 0.27, K, Dc)
print("markers", sum(len(x) for x in r.ids), "launches", det.last_launch_count(), {k: round(v, 3) for k, v in det.last_stage_times().items()})
det.close()

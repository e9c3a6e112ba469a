import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from aruco_slam_b200 import aruco, synth, dictionaries as D
import torch
B=32
K = np.array([[1400.0, 0, 960], [0, 1400.0, 540], [0, 0, 1]]); Dc = np.array([0.05, -0.1, 0.001, -0.002, 0.02])
frames = synth.render_batch("C2", B, base_seed=0)
fh = torch.from_numpy(frames).pin_memory()
det = aruco.ArucoDetector(D.getPredefinedDictionary(D.DICT_6X6_250), max_shape=frames.shape[1:], max_batch=B, device=0)
cam = det.make_camera(0.27, K, Dc) if hasattr(det,'make_camera') else None
for _ in range(4):
    r = det.detect_pose_batch(fh.numpy(), 0.27, K, Dc)
t0=time.perf_counter()
for _ in range(10):
    r = det.detect_pose_batch(fh.numpy(), 0.27, K, Dc)
print("ms per call (python wall)", (time.perf_counter()-t0)*100)
det.close()

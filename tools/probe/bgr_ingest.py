"""python tools/probe/bgr_ingest.py -- device-resident detect+pose on 32 bgr8 1080p frames (the reference's node delivers bgr8,
src/aruco_slam_node.cpp:93): ms per batch with the conversion fused into the threshold kernel (default) or as its own pass
(B2A_FUSE_BGR=0), CUDA events around 20 calls after 3 warm-ups."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from aruco_slam_b200 import aruco as A, dictionaries as D, synth  # noqa: E402

B = 32
gray = synth.render_batch("C2", B)
bgr = np.stack([synth.gray_to_bgr(g, i) for i, g in enumerate(gray)])
dev = torch.from_numpy(bgr).cuda()
det = A.ArucoDetector(D.getPredefinedDictionary(10), A.DetectorParameters(), max_shape=gray.shape[1:], max_batch=B)
fr = A.ArucoDetector.frames_device(dev.data_ptr(), B, 1080, 1920, channels=3)
K = np.array([[1400.0, 0, 960], [0, 1400.0, 540], [0, 0, 1]])
cam = A._camera(K, np.zeros(5), 0.1)
for _ in range(3):
    det.detect_raw(fr, cam)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    r = det.detect_raw(fr, cam)
e1.record()
torch.cuda.synchronize()
print("B2A_FUSE_BGR=%s  %.4f ms per batch of %d bgr8 1080p frames (calls are synchronous; includes the result copy)" % (os.environ.get("B2A_FUSE_BGR", "1"), e0.elapsed_time(e1) / 20, B))

"""python tools/probe/c4_host_time.py [inflight] -- where the host thread of the C4 loop (one 1080p stream, full SLAM loop) spends its
time: mean microseconds per frame inside submitImageFrames / addEncoder / waitImage, device-resident frames."""
import collections
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from aruco_slam_b200 import slam, aruco, synth  # noqa: E402

inflight = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n, W, H = 120, 1920, 1080
frames, enc, truth = synth.c4_stream(0, n)
d_frames = torch.from_numpy(frames).cuda()
s = slam.ArucoSlam(synth.C4_DICT, synth.C4_MARKER_LENGTH, image_shape=(H, W), device=0, r2c_tx=synth.C4_R2C[0], r2c_ty=synth.C4_R2C[1], max_landmarks=96)
s.setCameraParameters(synth.C4_K, synth.C4_D)
s.addEncoder(0, 0, None)
s.detector.set_inflight(inflight)
descs = [aruco.ArucoDetector.frames_device(d_frames[f].data_ptr(), 1, H, W) for f in range(n)]
tks = [s.detector.submit_raw(descs[0], s._cam) for _ in range(inflight)]
for t_ in tks:
    s.detector.wait_raw(t_)
s.detector.set_inflight(inflight)
warm = 20
for f in range(warm):
    tk = s.submitImageFrames(descs[f]); s.addEncoder(*enc[f]); s.waitImage(tk)
s.synchronize()
torch.cuda.synchronize()
acc = collections.Counter()
q, nxt = collections.deque(), warm
t_all = time.perf_counter()
for f in range(warm, n):
    while nxt < n and len(q) < inflight:
        t0 = time.perf_counter(); q.append(s.submitImageFrames(descs[nxt])); acc["submit"] += time.perf_counter() - t0
        nxt += 1
    t0 = time.perf_counter(); s.addEncoder(*enc[f]); acc["encoder"] += time.perf_counter() - t0
    t0 = time.perf_counter(); s.waitImage(q.popleft()); acc["wait"] += time.perf_counter() - t0
s.synchronize()
tot = time.perf_counter() - t_all
k = n - warm
print("inflight %d: %.1f us per frame wall (%.0f frames/s); host time per frame: submit %.1f, addEncoder %.1f, waitImage %.1f us; launches per frame %d" %
      (inflight, 1e6 * tot / k, k / tot, 1e6 * acc["submit"] / k, 1e6 * acc["encoder"] / k, 1e6 * acc["wait"] / k, s.detector.last_launch_count()))

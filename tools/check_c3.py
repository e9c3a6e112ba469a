"""debug helper: C3 (4K, noise + blur) batch through the GPU path, every frame against the CPU oracle"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aruco_slam_b200 import aruco, synth, dictionaries as D
from oracle import oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
frames = np.stack([synth.render_config("C3", seed0 + i).image for i in range(B)])
dic = D.getPredefinedDictionary(D.DICT_6X6_250)
det = aruco.ArucoDetector(dic, max_shape=frames.shape[1:], max_batch=B, device=0)
_H, _W = frames.shape[1:]
_f = 1400.0 * _W / 1920.0                                  # the bench's camera: ~70 degree field of view centred on the frame (the distortion model is only sane inside it)
K = np.array([[_f, 0, _W / 2.0], [0, _f, _H / 2.0], [0, 0, 1]]); Dc = np.array([0.05, -0.1, 0.001, -0.002, 0.02])
r = det.detect_pose_batch(frames, 0.27, K, Dc)
bad = []
for b in range(B):
    oc, oi, orj = O.detect(frames[b], dic)
    ok = np.array_equal(r.ids[b], oi) and np.array_equal(r.corners[b], oc) and np.array_equal(r.rejected[b], orj)
    if ok and len(oi):
        orv, otv = O.estimate_pose_single_markers(oc, 0.27, K, Dc)
        dt = np.abs(r.tvecs[b] - otv).max(axis=1)
        dr = np.array([synth.rvec_distance(x, y) for x, y in zip(r.rvecs[b], orv)])
        if dt.max() > 1e-4 or dr.max() > 1e-4:
            k = int(np.argmax(np.maximum(dt, dr)))
            print("frame", b, "pose diff: max dt %.3g dr %.3g at marker %d id %d corners %s gpu r %s t %s oracle r %s t %s" % (dt.max(), dr.max(), k, oi[k], oc[k].tolist(), r.rvecs[b][k], r.tvecs[b][k], orv[k], otv[k]), flush=True)
    if not ok:
        bad.append(b)
        print("frame", b, "seed", seed0 + b, "ids", len(r.ids[b]), "vs", len(oi), "rej", len(r.rejected[b]), "vs", len(orj),
              "missing", sorted(set(oi.tolist()) - set(r.ids[b].tolist())), "extra", sorted(set(r.ids[b].tolist()) - set(oi.tolist())), flush=True)
print("B", B, "bad frames", bad)
det.close()

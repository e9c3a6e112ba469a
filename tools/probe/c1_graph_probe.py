import os, sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from aruco_slam_b200 import aruco as A, dictionaries as D, synth
frs = [synth.render_config("C1", s).image for s in range(3)]
det = A.ArucoDetector(D.getPredefinedDictionary(0), A.DetectorParameters(), max_shape=frs[0].shape, max_batch=1)
K = np.array([[600.0, 0, 320], [0, 600.0, 240], [0, 0, 1]])
for g in (True, False):
    det.set_graph(g)
    for f in frs:
        r = det.detect_pose_batch(f, 0.05, K, np.zeros(5))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(200):
        r = det.detect_pose_batch(frs[i % 3], 0.05, K, np.zeros(5))
    t1 = time.perf_counter()
    print("graph", g, "wall per call %.1f us" % (1e6 * (t1 - t0) / 200), "ids", r.ids[0].tolist(), "launches", det.last_launch_count())
    d = torch.from_numpy(frs[0]).cuda()
    fr = A.ArucoDetector.frames_device(d.data_ptr(), 1, 480, 640)
    cam = A._camera(K, np.zeros(5), 0.05)
    st = torch.cuda.ExternalStream(det.stream)
    tot = 0
    for i in range(50):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); det.detect_raw(fr, cam); e1.record(st); e1.synchronize(); tot += e0.elapsed_time(e1)
    print("   events on the detector stream: %.1f us per call" % (1e3 * tot / 50))

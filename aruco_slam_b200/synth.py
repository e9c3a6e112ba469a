"""Deterministic synthetic rendered-marker frames.

numpy only (no cv2), so the same frames can be produced on the GPU box.  Frame
`f` of a batch uses `seed = base_seed + f`.  The reference has no test images
(SURVEY section 4); these frames are the benchmark and parity inputs for the
`detectMarkers` / `estimatePoseSingleMarkers` path of reference
src/aruco_slam.cpp:313-314.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .dictionaries import Dictionary, getPredefinedDictionary

_CELL = 24  # source pixels per marker cell


@dataclass
class Frame:
    image: np.ndarray                 # (H, W) uint8 gray
    ids: np.ndarray                   # (n,) int32 ground-truth marker ids
    corners: np.ndarray               # (n, 4, 2) float64 ground-truth corners (TL,TR,BR,BL)
    meta: dict = field(default_factory=dict)


def _homography(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """3x3 H with dst ~ H src, from 4 point pairs."""
    A = np.zeros((8, 8))
    b = np.zeros(8)
    for i in range(4):
        x, y = src[i]
        u, v = dst[i]
        A[2 * i] = [x, y, 1, 0, 0, 0, -u * x, -u * y]
        A[2 * i + 1] = [0, 0, 0, x, y, 1, -v * x, -v * y]
        b[2 * i], b[2 * i + 1] = u, v
    h = np.linalg.solve(A, b)
    return np.append(h, 1.0).reshape(3, 3)


def _gauss_blur(img: np.ndarray, sigma: float) -> np.ndarray:
    r = int(np.ceil(3 * sigma))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    out = img.astype(np.float64)
    for axis in (0, 1):
        pad = [(0, 0), (0, 0)]
        pad[axis] = (r, r)
        p = np.pad(out, pad, mode="reflect")
        acc = np.zeros_like(out)
        for i in range(2 * r + 1):
            sl = [slice(None), slice(None)]
            sl[axis] = slice(i, i + out.shape[axis])
            acc += k[i] * p[tuple(sl)]
        out = acc
    return out


def background(W: int, H: int) -> np.ndarray:
    x = np.arange(W, dtype=np.float64)[None, :]
    y = np.arange(H, dtype=np.float64)[:, None]
    return 110.0 + 30.0 * np.sin(x / 97.0) * np.cos(y / 71.0)


def paste_marker(img: np.ndarray, dic: Dictionary, marker_id: int, quad: np.ndarray,
                 quiet: float = 0.2) -> None:
    """Warp marker `marker_id` (plus a white quiet zone of `quiet`*side) so that
    its four outer corners land on `quad` (TL,TR,BR,BL; pixel-centre coords)."""
    H, W = img.shape
    m = dic.marker_image(marker_id, _CELL)
    M = m.shape[0]
    q = int(round(quiet * M))
    src = np.full((M + 2 * q, M + 2 * q), 255, np.uint8)
    src[q:q + M, q:q + M] = m
    S = src.shape[0]
    # continuous coords: source pixel i covers [i-0.5, i+0.5]
    c0, c1 = q - 0.5, q + M - 0.5
    src_c = np.array([[c0, c0], [c1, c0], [c1, c1], [c0, c1]])
    Hm = _homography(src_c, quad)            # src -> dst
    Hi = np.linalg.inv(Hm)
    ext = np.array([[-0.5, -0.5], [S - 0.5, -0.5], [S - 0.5, S - 0.5], [-0.5, S - 0.5]])
    e = np.c_[ext, np.ones(4)] @ Hm.T
    e = e[:, :2] / e[:, 2:3]
    x0 = max(int(np.floor(e[:, 0].min())), 0)
    x1 = min(int(np.ceil(e[:, 0].max())) + 1, W)
    y0 = max(int(np.floor(e[:, 1].min())), 0)
    y1 = min(int(np.ceil(e[:, 1].max())) + 1, H)
    if x1 <= x0 or y1 <= y0:
        return
    xs, ys = np.meshgrid(np.arange(x0, x1, dtype=np.float64), np.arange(y0, y1, dtype=np.float64))
    w = Hi[2, 0] * xs + Hi[2, 1] * ys + Hi[2, 2]
    u = (Hi[0, 0] * xs + Hi[0, 1] * ys + Hi[0, 2]) / w
    v = (Hi[1, 0] * xs + Hi[1, 1] * ys + Hi[1, 2]) / w
    inside = (u >= -0.5) & (u <= S - 0.5) & (v >= -0.5) & (v <= S - 0.5)
    uc = np.clip(u, 0, S - 1)
    vc = np.clip(v, 0, S - 1)
    u0 = np.floor(uc).astype(np.int64)
    v0 = np.floor(vc).astype(np.int64)
    u1 = np.minimum(u0 + 1, S - 1)
    v1 = np.minimum(v0 + 1, S - 1)
    fu = uc - u0
    fv = vc - v0
    sf = src.astype(np.float64)
    val = (sf[v0, u0] * (1 - fu) * (1 - fv) + sf[v0, u1] * fu * (1 - fv)
           + sf[v1, u0] * (1 - fu) * fv + sf[v1, u1] * fu * fv)
    roi = img[y0:y1, x0:x1]
    roi[inside] = val[inside]


def render_frame(W: int, H: int, n_markers: int, dict_id: int, seed: int,
                 noise_sigma: float = 0.0, blur_sigma: float = 0.0,
                 side_range=(60.0, 160.0), jitter: float = 0.12) -> Frame:
    dic = getPredefinedDictionary(dict_id)
    rng = np.random.default_rng(seed)
    img = background(W, H)
    ids = rng.choice(dic.n_markers, size=min(n_markers, dic.n_markers), replace=False)
    placed = []          # (cx, cy, side)
    out_ids, out_corners = [], []
    for mid in ids:
        s = rng.uniform(*side_range)
        rad = 0.9 * s
        ok = False
        for _ in range(1000):
            cx = rng.uniform(rad, W - rad) if W > 2 * rad else W / 2
            cy = rng.uniform(rad, H - rad) if H > 2 * rad else H / 2
            if all((cx - px) ** 2 + (cy - py) ** 2 > (0.8 * (s + ps)) ** 2 for px, py, ps in placed):
                ok = True
                break
        if not ok:
            continue
        th = rng.uniform(0, 2 * np.pi)
        base = np.array([[-0.5, -0.5], [0.5, -0.5], [0.5, 0.5], [-0.5, 0.5]]) * s
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        quad = base @ R.T + np.array([cx, cy]) + rng.uniform(-jitter, jitter, (4, 2)) * s
        paste_marker(img, dic, int(mid), quad)
        placed.append((cx, cy, s))
        out_ids.append(int(mid))
        out_corners.append(quad)
    if blur_sigma > 0:
        img = _gauss_blur(img, blur_sigma)
    if noise_sigma > 0:
        img = img + rng.normal(0.0, noise_sigma, img.shape)
    u8 = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return Frame(u8, np.array(out_ids, np.int32), np.array(out_corners, np.float64).reshape(-1, 4, 2),
                 dict(W=W, H=H, dict_id=dict_id, seed=seed, noise_sigma=noise_sigma, blur_sigma=blur_sigma))


# BASELINE.json configs: name -> kwargs
CONFIGS = {
    "C1": dict(W=640, H=480, n_markers=4, dict_id=0, side_range=(50.0, 90.0)),
    "C2": dict(W=1920, H=1080, n_markers=30, dict_id=10, side_range=(60.0, 160.0)),
    "C3": dict(W=3840, H=2160, n_markers=100, dict_id=10, side_range=(80.0, 220.0),
               noise_sigma=4.0, blur_sigma=1.0),
}


def render_config(name: str, seed: int) -> Frame:
    return render_frame(seed=seed, **CONFIGS[name])


def render_batch(name: str, batch: int, base_seed: int = 0) -> np.ndarray:
    """(batch, H, W) uint8 frames with seeds base_seed .. base_seed+batch-1."""
    return np.stack([render_config(name, base_seed + f).image for f in range(batch)])


def gray_to_bgr(gray: np.ndarray, seed: int = 0) -> np.ndarray:
    """A colour frame whose channels differ (exercises the BGR->gray ingest,
    reference src/aruco_slam_node.cpp:93 delivers bgr8)."""
    rng = np.random.default_rng(seed)
    d = rng.integers(-6, 7, size=gray.shape + (3,))
    return np.clip(gray[..., None].astype(np.int64) + d, 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------
# 3-D scenes (config C4 style): markers on planes, pinhole + Brown distortion
# ----------------------------------------------------------------------------

def rodrigues(r: np.ndarray) -> np.ndarray:
    th = float(np.linalg.norm(r))
    if th < 1e-12:
        return np.eye(3)
    k = r / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def rvec_distance(r1, r2) -> float:
    """Largest element difference of the rotation matrices of two rotation vectors (radians to first
    order).  solvePnP can return |rvec| > pi near a half turn; r and r (1 - 2 pi/|r|) are the same rotation."""
    return float(np.abs(rodrigues(np.asarray(r1, float).ravel()) - rodrigues(np.asarray(r2, float).ravel())).max())


def project(obj: np.ndarray, rvec, tvec, K: np.ndarray, D: np.ndarray) -> np.ndarray:
    """Pinhole + (k1,k2,p1,p2,k3) projection of (n,3) points."""
    P = obj @ rodrigues(np.asarray(rvec, float)).T + np.asarray(tvec, float)
    x = P[:, 0] / P[:, 2]
    y = P[:, 1] / P[:, 2]
    k1, k2, p1, p2, k3 = (list(D) + [0] * 5)[:5]
    r2 = x * x + y * y
    rad = 1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3
    xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return np.stack([K[0, 0] * xd + K[0, 2], K[1, 1] * yd + K[1, 2]], axis=1)


def marker_object_points(L: float) -> np.ndarray:
    """estimatePoseSingleMarkers object points (reference aruco_slam.h:189)."""
    h = L / 2.0
    return np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]], float)


def render_scene(W: int, H: int, dict_id: int, K: np.ndarray, D: np.ndarray, marker_length: float,
                 poses, seed: int = 0, noise_sigma: float = 0.0, blur_sigma: float = 0.0) -> Frame:
    """Render markers given as [(id, rvec, tvec)] camera-frame poses; the corner
    positions are the projected object points, the interior is warped by the
    homography through them (exact for D=0, first-order otherwise)."""
    dic = getPredefinedDictionary(dict_id)
    rng = np.random.default_rng(seed)
    img = background(W, H)
    obj = marker_object_points(marker_length)
    ids, cs = [], []
    for mid, rvec, tvec in poses:
        quad = project(obj, rvec, tvec, K, D)
        paste_marker(img, dic, int(mid), quad)
        ids.append(int(mid))
        cs.append(quad)
    if blur_sigma > 0:
        img = _gauss_blur(img, blur_sigma)
    if noise_sigma > 0:
        img = img + rng.normal(0.0, noise_sigma, img.shape)
    u8 = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return Frame(u8, np.array(ids, np.int32), np.array(cs, float).reshape(-1, 4, 2),
                 dict(W=W, H=H, dict_id=dict_id, seed=seed))


# ----------------------------------------------------------------------------
# World scenes (config C4): a landmark map in the reference's map.txt layout, a differential-drive robot with a
# forward-looking camera, encoder readings.  Frames: robot X forward / Y left / Z up; camera optical X right /
# Y down / Z forward (the mapping the reference's observation model assumes, src/aruco_slam.cpp:359-361:
# x_obs = t_z + r2c.x, y_obs = -t_x + r2c.y, theta_obs = heading of the marker's normal).
# ----------------------------------------------------------------------------

R_CAM_TO_ROBOT = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])     # columns: camera axes in the robot frame


def rvec_from_matrix(R: np.ndarray) -> np.ndarray:
    """rotation vector of a rotation matrix (inverse of `rodrigues`)."""
    c = (np.trace(R) - 1.0) / 2.0
    th = float(np.arccos(np.clip(c, -1.0, 1.0)))
    if th < 1e-12:
        return np.zeros(3)
    if np.pi - th < 1e-6:                       # half turn: axis from the symmetric part
        A = (R + np.eye(3)) / 2.0
        k = int(np.argmax(np.diag(A)))
        ax = A[:, k] / np.sqrt(A[k, k])
        return th * ax
    ax = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2.0 * np.sin(th))
    return th * ax


def rpy_matrix(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """fixed-axes X-Y-Z rotation (tf2 setRPY)."""
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def quaternion_from_matrix(R: np.ndarray) -> np.ndarray:
    """(x, y, z, w) of a rotation matrix."""
    w = np.sqrt(max(0.0, 1.0 + R[0, 0] + R[1, 1] + R[2, 2])) / 2.0
    if w > 1e-6:
        return np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])
    r = rvec_from_matrix(R)
    th = np.linalg.norm(r)
    return np.append(np.sin(th / 2) * r / th, np.cos(th / 2))


REFERENCE_MAP = (        # /root/reference/map/map.txt:2-8  (id, length, x, y, z, roll, pitch, yaw)
    (0, 0.27, 5.10375, 0.0, 0.3, 0.0, -1.5708, 0.0),
    (1, 0.27, 5.10375, -1.5, 0.3, 0.0, -1.5708, 0.0),
    (2, 0.27, 5.10375, -3.0, 0.3, 0.0, -1.5708, 0.0),
    (3, 0.27, 4.0, 0.6025, 0.3, 1.5708, -0.0, 0.0),
    (4, 0.27, 2.0, 0.6025, 0.3, 1.5708, -0.0, 0.0),
    (5, 0.27, 4.0, -4.09375, 0.3, -1.5708, -0.0, 0.0),
    (6, 0.27, 2.0, -4.09375, 0.3, -1.5708, -0.0, 0.0),
)


def marker_world_frame(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """columns = the marker's x (right), y (up), z (normal, out of the printed face) axes in the world.  The map stores the
    orientation of a thin cube whose local z is the face normal; the printed side is rendered upright (y = world up)."""
    n = rpy_matrix(roll, pitch, yaw)[:, 2].copy()
    n[2] = 0.0
    n /= np.linalg.norm(n)
    up = np.array([0.0, 0.0, 1.0])
    x = np.cross(up, n)
    return np.stack([x, up, n], axis=1)


def scene_poses(map_markers, robot, r2c_t=(0.0, 0.0, 0.3), K=None, D=None, W=0, H=0, margin=12.0, max_incidence_deg=72.0):
    """camera-frame (id, rvec, tvec) of the map markers the camera sees from robot = (x, y, theta)."""
    rx, ry, rth = robot
    c, s = np.cos(rth), np.sin(rth)
    Rwr = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]])                 # world -> robot
    out = []
    for (mid, L, mx, my, mz, roll, pitch, yaw) in map_markers:
        Rmw = marker_world_frame(roll, pitch, yaw)
        R = R_CAM_TO_ROBOT.T @ Rwr @ Rmw
        t = R_CAM_TO_ROBOT.T @ (Rwr @ (np.array([mx, my, mz]) - np.array([rx, ry, 0.0])) - np.asarray(r2c_t, float))
        if t[2] < 0.35:
            continue
        nrm = R[:, 2]
        cosi = -float(nrm @ t) / float(np.linalg.norm(t))
        if cosi < np.cos(np.radians(max_incidence_deg)):
            continue
        rvec = rvec_from_matrix(R)
        if K is not None:
            q = project(marker_object_points(L), rvec, t, K, np.zeros(5) if D is None else D)
            if q[:, 0].min() < margin or q[:, 1].min() < margin or q[:, 0].max() > W - 1 - margin or q[:, 1].max() > H - 1 - margin:
                continue
        out.append((int(mid), rvec, t))
    return out


def drive(pose, wl: float, wr: float, dt: float, kl: float = 0.05, kr: float = 0.05, b: float = 0.09):
    """ground-truth motion of the differential drive for one encoder interval (the model of src/aruco_slam.cpp:35-52)."""
    dsl, dsr = kl * wl * dt, kr * wr * dt
    dth, ds = (dsr - dsl) / (2 * b), 0.5 * (dsr + dsl)
    th = pose[2] + 0.5 * dth
    return np.array([pose[0] + ds * np.cos(th), pose[1] + ds * np.sin(th), pose[2] + dth])


def c5_state(n_lm: int, seed: int = 0):
    """start state of the EKF-only workload C5: Sigma0 = A A^T / N + 0.1 I, landmarks uniform in an 8 m square,
    aruco ids 100 .. 100 + n_lm - 1"""
    N = 3 + 3 * n_lm
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(N, N))
    sigma0 = A @ A.T / N + 0.1 * np.eye(N)
    mu0 = np.concatenate([[0.3, -0.2, 0.4], rng.uniform(-4, 4, 3 * n_lm)])
    return mu0, sigma0, np.arange(100, 100 + n_lm, dtype=np.int32)


# ----------------------------------------------------------------------------
# C4 (BASELINE.json config 4): camera streams of a robot driving through a room whose walls carry a map.txt-style landmark
# map; 1080p pinhole camera (fx = fy = 1400) with the distortion of /root/reference/default.yaml:16-20, DICT_ARUCO_ORIGINAL
# markers of 0.27 m (parameters.yaml:16-17), one encoder message and one frame per step.
# ----------------------------------------------------------------------------
C4_K = np.array([[1400.0, 0, 960.0], [0, 1400.0, 540.0], [0, 0, 1]])
C4_D = np.array([0.04160142651680036, -0.04771035303381654, -0.0032638387781624705, -0.003985120051161831, 0.01110263483766991])
C4_R2C = (0.12, 0.0, 0.25)
C4_DICT = 16
C4_MARKER_LENGTH = 0.27


def c4_map():
    """the reference's map/map.txt plus more markers in the same layout (id length x y z roll pitch yaw): the far wall (x = 5.1,
    facing -x) in two rows, the side walls (y = 0.6 facing -y, y = -2.2 facing +y)"""
    m = list(REFERENCE_MAP)
    nid = 7
    for y in (-0.55, -1.05, -2.0, -2.5, 0.45):
        for z in (0.3, 0.75):
            m.append((nid, 0.27, 5.10375, y, z, 0.0, -1.5708, 0.0)); nid += 1
    for x in (2.6, 3.3, 4.6):
        m.append((nid, 0.27, x, 0.6025, 0.55, 1.5708, -0.0, 0.0)); nid += 1
    for x in (2.4, 3.1, 3.8, 4.5):
        for z in (0.3, 0.8):
            m.append((nid, 0.27, x, -2.2, z, -1.5708, -0.0, 0.0)); nid += 1
    return tuple(m)


def c4_stream(stream: int, n_frames: int, dt: float = 0.1):
    """frames (n, 1080, 1920) u8, encoder readings (n, 3) = (wl, wr, dt) applied BEFORE each frame, ground-truth poses (n, 3)
    of one camera stream; streams differ in start pose and wheel speeds"""
    rng = np.random.default_rng(4000 + stream)
    cmap = c4_map()
    pose = np.array([2.45 + 0.12 * (stream % 4), -0.9 + 0.2 * (stream // 4) + rng.uniform(-0.1, 0.1), rng.uniform(-0.25, 0.1)])
    frames, enc, truth = [], [], []
    fwd = []
    for f in range(n_frames):
        turn = -1.0 if (f // 6) % 2 == 0 else 0.6
        wl, wr = 2.0 + 0.5 * turn + rng.uniform(-0.2, 0.2), 2.0 - 0.5 * turn + rng.uniform(-0.2, 0.2)
        if f % 9 == 4:
            wl = wr = 0.0                                   # standing still: the reference's stationary gate
        # a long stream drives 24 frames forward, then the same wheel speeds backwards in reverse order, and so on: the robot stays
        # in front of the same part of the map (about nine markers in view) however long the stream is
        if f % 48 < 24:
            if f < 24:
                fwd.append((wl, wr))
            else:
                wl, wr = fwd[f % 48]
        else:
            wl, wr = (-v for v in fwd[47 - f % 48])
        pose = drive(pose, wl, wr, dt)
        poses = scene_poses(cmap, pose, C4_R2C, C4_K, C4_D, 1920, 1080)
        fr = render_scene(1920, 1080, C4_DICT, C4_K, C4_D, C4_MARKER_LENGTH, poses, seed=1000 * stream + f, noise_sigma=1.0, blur_sigma=0.8)
        frames.append(fr.image)
        enc.append((wl, wr, dt))
        truth.append(pose.copy())
    return np.stack(frames), np.array(enc), np.array(truth)

"""TEST INFRASTRUCTURE ONLY (see oracle.h): numpy restatement of cv::aruco::ArucoDetector::refineDetectedMarkers (cv2 4.13
objdetect/src/aruco/aruco_detector.cpp), the board-driven recovery of rejected candidates that belongs to the detectMarkers
surface of the reference's call (src/aruco_slam.cpp:313) -- the reference itself never calls it, SURVEY 8(f) row 4.

  _projectUndetectedMarkers (no camera)   findHomography(board xy -> detected corners) + perspectiveTransform
  _projectUndetectedMarkers (camera)      solvePnP(matched board corners, detected corners) + projectPoints
  the greedy loop                         every undetected board marker takes the closest still-free rejected candidate
                                          (max over the four corners of the squared distance, the LAST corner order that beats
                                          the best so far -- cv2's loop keeps overwriting), subject to the code test
                                          getDistanceToId(bits, id, allRotations = false) < int(maxCorrectionBits * rate)

The projections are least-squares fits; cv2 reaches them with its own solvers (normalised DLT + LM), this file with numpy's,
so projected corners agree to ~1e-3 px, which only matters for candidates within that of minRepDistance.  Pinned by
tests/golden/refine_board.npz (tools/make_golden_refine.py)."""
import numpy as np

from . import oracle


def _normalise(p):
    c = p.mean(0)
    s = len(p) / np.abs(p - c).sum(0)              # cv2: count / sum |x - c| per axis
    return c, s


def find_homography(src, dst, refine_iters=10):
    """cv::findHomography(src, dst, 0): normalised DLT, then Gauss-Newton on the reprojection error"""
    src = np.asarray(src, np.float64)
    dst = np.asarray(dst, np.float64)
    cm, sm = _normalise(dst)
    cM, sM = _normalise(src)
    L = []
    for (X, Y), (x, y) in zip((src - cM) * sM, (dst - cm) * sm):
        L.append([X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x])
        L.append([0, 0, 0, X, Y, 1, -y * X, -y * Y, -y])
    L = np.array(L)
    w, v = np.linalg.eigh(L.T @ L)
    H0 = v[:, 0].reshape(3, 3)
    inv_norm = np.array([[1 / sm[0], 0, cm[0]], [0, 1 / sm[1], cm[1]], [0, 0, 1]])
    norm2 = np.array([[sM[0], 0, -cM[0] * sM[0]], [0, sM[1], -cM[1] * sM[1]], [0, 0, 1]])
    H = inv_norm @ H0 @ norm2
    H /= H[2, 2]
    if len(src) > 4:
        h = H.ravel()[:8].copy()
        for _ in range(refine_iters):
            J, r = [], []
            for (X, Y), (x, y) in zip(src, dst):
                ww = 1.0 / (h[6] * X + h[7] * Y + 1.0)
                xi, yi = (h[0] * X + h[1] * Y + h[2]) * ww, (h[3] * X + h[4] * Y + h[5]) * ww
                r += [xi - x, yi - y]
                J.append([X * ww, Y * ww, ww, 0, 0, 0, -X * ww * xi, -Y * ww * xi])
                J.append([0, 0, 0, X * ww, Y * ww, ww, -X * ww * yi, -Y * ww * yi])
            J, r = np.array(J), np.array(r)
            h = h - np.linalg.solve(J.T @ J, J.T @ r)
        H = np.append(h, 1.0).reshape(3, 3)
    return H


def perspective_transform(pts, H):
    p = np.c_[np.asarray(pts, np.float64), np.ones(len(pts))] @ H.T
    return p[:, :2] / p[:, 2:3]


def rodrigues(r):
    th = np.linalg.norm(r)
    if th < 1e-12:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def rvec_of(R):
    u, _, vt = np.linalg.svd(R)
    R = u @ vt
    th = np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))
    if th < 1e-12:
        return np.zeros(3)
    ax = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if np.linalg.norm(ax) < 1e-9:                       # theta ~ pi
        ax = np.sqrt(np.maximum((np.diag(R) + 1) / 2, 0))
        ax[1] *= np.sign(R[0, 1]) or 1.0
        ax[2] *= np.sign(R[0, 2]) or 1.0
        return th * ax / np.linalg.norm(ax)
    return th * ax / np.linalg.norm(ax)


def project_points(obj, rvec, tvec, K, D):
    """cv::projectPoints, Brown model (k1 k2 p1 p2 k3)"""
    D = np.r_[np.asarray(D, np.float64).ravel(), np.zeros(5)][:5]
    P = np.asarray(obj, np.float64) @ rodrigues(np.asarray(rvec, np.float64)).T + np.asarray(tvec, np.float64)
    x, y = P[:, 0] / P[:, 2], P[:, 1] / P[:, 2]
    r2 = x * x + y * y
    cd = 1 + r2 * (D[0] + r2 * (D[1] + r2 * D[4]))
    xd = x * cd + 2 * D[2] * x * y + D[3] * (r2 + 2 * x * x)
    yd = y * cd + D[2] * (r2 + 2 * y * y) + 2 * D[3] * x * y
    return np.c_[K[0, 0] * xd + K[0, 2], K[1, 1] * yd + K[1, 2]]


def undistort_points(img, K, D, iters=20):
    D = np.r_[np.asarray(D, np.float64).ravel(), np.zeros(5)][:5]
    x0 = (img[:, 0] - K[0, 2]) / K[0, 0]
    y0 = (img[:, 1] - K[1, 2]) / K[1, 1]
    x, y = x0.copy(), y0.copy()
    for _ in range(iters):
        r2 = x * x + y * y
        icd = 1.0 / (1 + r2 * (D[0] + r2 * (D[1] + r2 * D[4])))
        dx = 2 * D[2] * x * y + D[3] * (r2 + 2 * x * x)
        dy = D[2] * (r2 + 2 * y * y) + 2 * D[3] * x * y
        x, y = (x0 - dx) * icd, (y0 - dy) * icd
    return np.c_[x, y]


def _least_squares_pose(p, obj, img, K, D, iters=30):
    """Gauss-Newton on the reprojection error over (rvec, tvec), numeric Jacobian (cv2 runs its LM to the same minimum)"""
    f = lambda q: (project_points(obj, q[:3], q[3:], K, D) - img).ravel()
    for _ in range(iters):
        r = f(p)
        J = np.empty((len(r), 6))
        for k in range(6):
            d = np.zeros(6)
            d[k] = 1e-7
            J[:, k] = (f(p + d) - f(p - d)) / 2e-7
        step = np.linalg.solve(J.T @ J, J.T @ r)
        p = p - step
        if np.linalg.norm(step) < 1e-12:
            break
    return p[:3], p[3:]


def is_planar(obj):
    """the test of cv::solvePnP(ITERATIVE)'s start: singular values W of the scatter of the object points, W[2] / W[1] < 1e-3"""
    obj = np.asarray(obj, np.float64)
    w = np.linalg.eigvalsh((obj - obj.mean(0)).T @ (obj - obj.mean(0)))
    return w[0] / w[1] < 1e-3


def solve_pnp_dlt(obj, img, K, D, iters=30):
    """cv::solvePnP(ITERATIVE) for object points in general position (>= 6): the DLT start -- the 3 x 4 projection from the smallest
    eigenvector of L^T L over normalised image points, its left block made a rotation by SVD, the translation scaled alike --
    then least squares on the reprojection error"""
    obj = np.asarray(obj, np.float64)
    img = np.asarray(img, np.float64)
    if len(obj) < 6:
        raise ValueError("DLT algorithm needs at least 6 points")
    mn = undistort_points(img, K, D)
    L = np.zeros((2 * len(obj), 12))
    for i, (M, m) in enumerate(zip(obj, mn)):
        L[2 * i, 0:3] = M; L[2 * i, 3] = 1; L[2 * i, 8:11] = -m[0] * M; L[2 * i, 11] = -m[0]
        L[2 * i + 1, 4:7] = M; L[2 * i + 1, 7] = 1; L[2 * i + 1, 8:11] = -m[1] * M; L[2 * i + 1, 11] = -m[1]
    w, V = np.linalg.eigh(L.T @ L)
    RR = V[:, 0].reshape(3, 4)
    if np.linalg.det(RR[:, :3]) < 0:
        RR = -RR
    u, _, vt = np.linalg.svd(RR[:, :3])
    R = u @ vt
    t = RR[:, 3] * (np.linalg.norm(R) / np.linalg.norm(RR[:, :3]))
    return _least_squares_pose(np.r_[rvec_of(R), t], obj, img, K, D, iters)


def solve_pnp(obj, img, K, D):
    return solve_pnp_planar(obj, img, K, D) if is_planar(obj) else solve_pnp_dlt(obj, img, K, D)


def solve_pnp_planar(obj, img, K, D, iters=30):
    """cv::solvePnP(ITERATIVE) for (nearly) coplanar object points: pose from the plane homography, then least squares on the
    reprojection error (Gauss-Newton with a numeric Jacobian; cv2 runs its LM to the same minimum)"""
    obj = np.asarray(obj, np.float64)
    img = np.asarray(img, np.float64)
    mean = obj.mean(0)
    w, E = np.linalg.eigh(np.cov((obj - mean).T))
    if not is_planar(obj):
        raise ValueError("board is not planar")
    E = E[:, ::-1]                                        # columns: two in-plane axes, then the normal
    if np.linalg.det(E) < 0:
        E[:, 2] = -E[:, 2]
    uv = (obj - mean) @ E[:, :2]
    H = find_homography(uv, undistort_points(img, K, D), refine_iters=0)
    h1, h2, h3 = H[:, 0], H[:, 1], H[:, 2]
    lam = 2.0 / (np.linalg.norm(h1) + np.linalg.norm(h2))
    if h3[2] * lam < 0:
        lam = -lam
    r1, r2 = h1 * lam, h2 * lam
    Rh = np.c_[r1, r2, np.cross(r1, r2)]
    u, _, vt = np.linalg.svd(Rh)
    Rh = u @ vt
    R = Rh @ E.T
    t = h3 * lam - R @ mean
    return _least_squares_pose(np.r_[rvec_of(R), t], obj, img, K, D, iters)


def project_undetected(board_ids, board_obj, corners, ids, K=None, D=None):
    """-> (undetected ids in board order, their projected corners (n, 4, 2)), or ([], []) when nothing can be projected"""
    board_ids = [int(i) for i in np.asarray(board_ids).ravel()]
    board_obj = np.asarray(board_obj, np.float64).reshape(len(board_ids), 4, 3)
    ids = [int(i) for i in np.asarray(ids).ravel()]
    corners = np.asarray(corners, np.float64).reshape(len(ids), 4, 2)
    obj_pts, img_pts, und = [], [], []
    if K is None:
        for j, bid in enumerate(board_ids):               # board order, first detection with that id
            if bid in ids:
                obj_pts += list(board_obj[j]); img_pts += list(corners[ids.index(bid)])
            else:
                und.append(j)
        if not img_pts:
            return [], np.zeros((0, 4, 2))
        H = find_homography(np.array(obj_pts)[:, :2], np.array(img_pts))
        return [board_ids[j] for j in und], np.array([perspective_transform(board_obj[j][:, :2], H) for j in und]).reshape(-1, 4, 2)
    for i, did in enumerate(ids):                         # Board::matchImagePoints: detection order, first board entry with that id
        if did in board_ids:
            obj_pts += list(board_obj[board_ids.index(did)]); img_pts += list(corners[i])
    if len(obj_pts) < 4:
        return [], np.zeros((0, 4, 2))
    rvec, tvec = solve_pnp(np.array(obj_pts), np.array(img_pts), K, D)
    und = [j for j, bid in enumerate(board_ids) if bid not in ids]
    return [board_ids[j] for j in und], np.array([project_points(board_obj[j], rvec, tvec, K, D) for j in und]).reshape(-1, 4, 2)


def refine_detected_markers(gray, dic, board_ids, board_obj, corners, ids, rejected, K=None, D=None, params=None,
                            min_rep_distance=10.0, error_correction_rate=3.0, check_all_orders=True, debug=None):
    """-> (corners (n, 4, 2) f32, ids (n,) i32, rejected (m, 4, 2) f32, recovered candidate indices (k,) i32)"""
    corners = np.asarray(corners, np.float32).reshape(-1, 4, 2)
    ids = np.asarray(ids, np.int32).ravel()
    rejected = np.asarray(rejected, np.float32).reshape(-1, 4, 2)
    none = np.zeros(0, np.int32)
    if len(ids) == 0 or len(rejected) == 0:
        return corners, ids, rejected, none
    p = params or oracle.default_params()
    und_ids, und_c = project_undetected(board_ids, board_obj, corners, ids, K, D)
    und_c = und_c.astype(np.float32)                      # vector<Point2f>
    taken = np.zeros(len(rejected), bool)
    max_corr = int(float(dic.max_correction_bits) * error_correction_rate)
    bb = p.markerBorderBits
    out_c, out_i, rec = list(corners), list(ids), []
    for uid, uc in zip(und_ids, und_c):
        best_j, best_d, best_q = -1, float(min_rep_distance) * float(min_rep_distance) + 1, None
        for j in range(len(rejected)):
            if taken[j]:
                continue
            valid, rot, mind = False, 0, best_d + 1
            for c in range(4):
                dv = uc - rejected[j][(c + np.arange(4)) % 4]          # float32 differences
                cur = max(0.0, float(np.max((dv[:, 0] * dv[:, 0] + dv[:, 1] * dv[:, 1]).astype(np.float64))))
                if cur < best_d:
                    valid, rot, mind = True, c, cur
                if not check_all_orders:
                    break
            if debug is not None:
                debug.append((uid, j, mind if valid else None))
            if not valid:
                continue
            q = rejected[j][(np.arange(4) + rot) % 4] if check_all_orders else rejected[j]
            dist = 0
            if error_correction_rate >= 0:
                bits = oracle.identify_one(gray, q, dic, p)[3]
                inner = bits[bb:bits.shape[0] - bb, bb:bits.shape[0] - bb]
                dist = int(np.count_nonzero(inner != dic.bits(uid)))
            if error_correction_rate < 0 or dist < max_corr:
                best_j, best_d, best_q = j, mind, q
        if best_j >= 0:
            if p.cornerRefinementMethod == 1:
                per = float(np.sqrt(((best_q - np.roll(best_q, -1, axis=0)) ** 2).sum(1)).sum())
                nm = dic.marker_size + 2 * bb
                win = min(max(1, int(np.rint(p.relativeCornerRefinmentWinSize * per / (4.0 * nm)))), p.cornerRefinementWinSize)
                best_q = oracle.corner_subpix(gray, best_q, win, p.cornerRefinementMaxIterations, p.cornerRefinementMinAccuracy)
            taken[best_j] = True
            out_c.append(best_q); out_i.append(uid); rec.append(best_j)
    if not rec:
        return corners, ids, rejected, none
    return np.array(out_c, np.float32).reshape(-1, 4, 2), np.array(out_i, np.int32), rejected[~taken], np.array(rec, np.int32)

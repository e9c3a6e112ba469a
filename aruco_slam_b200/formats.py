"""Wire / on-disk formats either side of the path, as plain records (no ROS types).

  load_map / parse_map     <-> MapLoader::loadMap              (reference src/map_loader.cpp:7-84, map/map.txt)
  quaternion_from_rpy      <-> tf2::Quaternion::setRPY         (map_loader.cpp:90, aruco_slam.cpp:273,387)
  robot_pose(slam)         <-> ArucoSlam::toRosPose            (aruco_slam.cpp:376-407)
  detected_map(slam)       <-> detected_map_ of addImage       (aruco_slam.cpp:266-281)
  detected_markers(...)    <-> detected_markers_ of getObservations / toRosDetectedMarkers (aruco_slam.cpp:324-347)
  load_parameters(path)    <-> ArucoSlamRosNode::parseArucoSlamIniteData (aruco_slam_node.cpp:146-164) on parameters.yaml
  load_camera(path)        <-> the camera block of default.yaml (:10-20): K and (k1, k2, p1, p2, k3)

Thin ctypes wrappers over the host-side entry points of libb2aruco.so (include/b2aruco.h)."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass
class MapMarker:
    """one cube of the reference's MarkerArray: frame "world", scale (length, length, 0.01)"""
    id: int
    length: float
    x: float
    y: float
    z: float
    roll: float
    pitch: float
    yaw: float
    q: tuple          # orientation (x, y, z, w)


@dataclass
class PoseWithCovariance:
    position: np.ndarray       # (3,)
    orientation: np.ndarray    # (4,) x, y, z, w
    covariance: np.ndarray     # (6, 6)


def quaternion_from_rpy(roll: float, pitch: float, yaw: float) -> np.ndarray:
    q = (C.c_double * 4)()
    _lib.lib().b2a_quaternion_from_rpy(roll, pitch, yaw, q)
    return np.array(q[:])


def _markers(arr, n):
    return [MapMarker(m.id, m.length, m.x, m.y, m.z, m.roll, m.pitch, m.yaw, tuple(m.q[:])) for m in arr[:n]]


def parse_map(text, cap: int = 4096):
    """the markers a map text defines, under the reference loader's acceptance rules (see include/b2aruco.h)"""
    data = text.encode() if isinstance(text, str) else bytes(text)
    arr = (_lib.MapMarker * cap)()
    n = C.c_int(0)
    _lib.check(_lib.lib().b2a_map_parse(data, len(data), arr, cap, C.byref(n)))
    return _markers(arr, n.value)


def load_map(path: str, cap: int = 4096):
    arr = (_lib.MapMarker * cap)()
    n = C.c_int(0)
    _lib.check(_lib.lib().b2a_map_load(str(path).encode(), arr, cap, C.byref(n)))
    return _markers(arr, n.value)


def robot_pose(slam) -> PoseWithCovariance:
    out = _lib.PoseWithCovariance()
    _lib.check(_lib.lib().b2a_slam_robot_pose(slam._h, C.byref(out)))
    return PoseWithCovariance(np.array(out.position[:]), np.array(out.orientation[:]), np.array(out.covariance[:]).reshape(6, 6))


def robot_pose_submit(slam, slot: int = 0):
    """enqueue the read-back of the pose record as of now into pinned slot 0 .. 7 (b2a_slam_robot_pose_submit)"""
    _lib.lib().b2a_slam_robot_pose_submit.argtypes = [C.c_void_p, C.c_int]
    _lib.check(_lib.lib().b2a_slam_robot_pose_submit(slam._h, int(slot)))


def robot_pose_wait(slam, slot: int = 0) -> PoseWithCovariance:
    out = _lib.PoseWithCovariance()
    _lib.lib().b2a_slam_robot_pose_wait.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    _lib.check(_lib.lib().b2a_slam_robot_pose_wait(slam._h, int(slot), C.byref(out)))
    return PoseWithCovariance(np.array(out.position[:]), np.array(out.orientation[:]), np.array(out.covariance[:]).reshape(6, 6))


def detected_map(slam, marker_length: float = None):
    n_lm = (slam.dim - 3) // 3
    arr = (_lib.MapMarker * max(1, n_lm))()
    n = C.c_int(0)
    _lib.check(_lib.lib().b2a_slam_detected_map(slam._h, float(marker_length if marker_length is not None else slam.marker_length), arr, max(1, n_lm), C.byref(n)))
    return _markers(arr, n.value)


def pose_record(mu, sigma):
    """toRosPose's packing on host values (b2a_pack_robot_pose): position (3,), orientation (4,), covariance (36,)"""
    m = (C.c_double * 3)(*[float(v) for v in np.asarray(mu).ravel()[:3]])
    S = (C.c_double * 9)(*[float(v) for v in np.asarray(sigma)[:3, :3].ravel()])
    out = _lib.PoseWithCovariance()
    _lib.lib().b2a_pack_robot_pose(m, S, C.byref(out))
    return np.array(out.position[:]), np.array(out.orientation[:]), np.array(out.covariance[:])


def detected_map_records(mu, marker_length: float):
    """the detected-map cubes of a state vector on host values (b2a_pack_map_marker)"""
    mu = np.asarray(mu, float).ravel()
    n = (len(mu) - 3) // 3
    arr = (_lib.MapMarker * max(1, n))()
    for i in range(n):
        lm = (C.c_double * 3)(*mu[3 + 3 * i:6 + 3 * i])
        _lib.lib().b2a_pack_map_marker.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        _lib.lib().b2a_pack_map_marker(i, float(marker_length), lm, C.byref(arr[i]))
    return _markers(arr, n)


def detected_markers(ids, rvecs, tvecs, marker_length: float, useful_distance_threshold: float = 3.0, r2c_q=None, r2c_t=None):
    """toRosDetectedMarkers: the cubes of the current frame's markers inside the useful range, in detection order, in the robot
    frame (b2a_pack_detected_markers); r2c_q = (x, y, z, w) rotation and r2c_t = translation of transformStamped_r2c"""
    ids = np.ascontiguousarray(np.asarray(ids).ravel(), np.int32)
    rv = np.ascontiguousarray(np.asarray(rvecs, np.float64).reshape(-1, 3))
    tv = np.ascontiguousarray(np.asarray(tvecs, np.float64).reshape(-1, 3))
    n = len(ids)
    arr = (_lib.MapMarker * max(1, n))()
    cnt = C.c_int(0)
    q = None if r2c_q is None else (C.c_double * 4)(*[float(v) for v in r2c_q])
    t = None if r2c_t is None else (C.c_double * 3)(*[float(v) for v in r2c_t])
    _lib.check(_lib.lib().b2a_pack_detected_markers(ids.ctypes.data, rv.ctypes.data, tv.ctypes.data, n, float(marker_length), float(useful_distance_threshold),
                                                    q, t, arr, max(1, n), C.byref(cnt)))
    return _markers(arr, cnt.value)


def load_parameters(path_or_text) -> dict:
    """The rosparam tree of the reference's parameters.yaml as parseArucoSlamIniteData reads it (aruco_slam_node.cpp:146-164): keyword
    arguments for slam.ArucoSlam (`slam`: Q_k, R_x, R_y, R_theta, kl, kr, b, useful_distance_threshold), `markers_dictionary`,
    `marker_length`, and the frame / topic names.  A key the file lacks keeps ArucoSlamIniteData's default, as nh.getParam leaves it.
    The node asks for const/USEFUL_DISTANCE_THRESHOLD_ (with the trailing underscore, :160) while the shipped file spells the key
    without it, so the threshold stays at its default 3 (aruco_slam.h:58) -- reproduced here."""
    import os
    import yaml
    text = open(path_or_text).read() if os.path.exists(str(path_or_text)) else str(path_or_text)
    tree = yaml.safe_load(text) or {}
    d = _lib.SlamParams()
    _lib.lib().b2a_default_slam_params(C.byref(d))
    get = lambda sec, key, dflt: (tree.get(sec) or {}).get(key, dflt)
    slam = dict(Q_k=float(get("covariance", "Q_k", d.Q_k)), R_x=float(get("covariance", "R_x", d.R_x)), R_y=float(get("covariance", "R_y", d.R_y)),
                R_theta=float(get("covariance", "R_theta", d.R_theta)), kl=float(get("odom", "kl", d.kl)), kr=float(get("odom", "kr", d.kr)),
                b=float(get("odom", "b", d.b)), useful_distance_threshold=float(get("const", "USEFUL_DISTANCE_THRESHOLD_", 3.0)))
    return dict(slam=slam, markers_dictionary=int(get("aruco", "markers_dictionary", 16)), marker_length=float(get("aruco", "marker_length", 0.27)),
                frames=dict(tree.get("frame") or {}), topics=dict(tree.get("topic") or {}))


def load_camera(path_or_text):
    """camera: {fx, fy, cx, cy, k1, k2, p1, p2, k3} (the reference's default.yaml:10-20) -> (K 3x3, D (k1, k2, p1, p2, k3)), the shapes
    setCameraParameters takes (aruco_slam_node.cpp:121-130)"""
    import os
    import yaml
    text = open(path_or_text).read() if os.path.exists(str(path_or_text)) else str(path_or_text)
    cam = (yaml.safe_load(text) or {}).get("camera") or {}
    K = np.array([[float(cam["fx"]), 0.0, float(cam["cx"])], [0.0, float(cam["fy"]), float(cam["cy"])], [0.0, 0.0, 1.0]])
    D = np.array([float(cam.get(k, 0.0)) for k in ("k1", "k2", "p1", "p2", "k3")])
    return K, D

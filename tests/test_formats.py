"""Wire / on-disk formats: the landmark map text, tf2's setRPY, the robot pose record and the map cubes.
Host-only entry points of the C ABI, checked against independent restatements (SciPy rotations, NumPy packing); the GPU-side
records are checked in test_gpu_slam.py::test_pose_and_map_records."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from aruco_slam_b200 import formats
from aruco_slam_b200._lib import B2AError

# same layout as the reference's map/map.txt (header comment, tabs and blanks mixed, trailing blanks); values are ours
MAP_TEXT = """# id    length\tx\ty\tz\troll_x\tpitch_y\tyaw_z
0   0.27\t5.10375 0       0.3     0    -1.5708   0
1\t0.27\t5.10375 -1.5    0.3     0    -1.5708   0

7\t0.2\t4   0.6025 0.3 \t1.5708 \t-0\t0.25  
   # an indented comment
12 0.15 -2.5 1e-1
13 0.15 1 2 3
14 0.15 1 2 3 0.1
15 0.15 1 2 3 0.1 0.2
16 0.15 1 2
"""


def test_quaternion_matches_fixed_axis_rpy():
    rng = np.random.default_rng(0)
    for r, p, y in rng.uniform(-np.pi, np.pi, (50, 3)):
        q = formats.quaternion_from_rpy(r, p, y)
        want = Rotation.from_euler("xyz", [r, p, y]).as_quat()          # extrinsic x-y-z = tf2::Quaternion::setRPY
        if np.dot(q, want) < 0:
            want = -want
        assert np.allclose(q, want, atol=1e-14)
    assert np.allclose(formats.quaternion_from_rpy(0, 0, 0), [0, 0, 0, 1])


def test_parse_map_rules(tmp_path):
    ms = formats.parse_map(MAP_TEXT)
    assert [m.id for m in ms] == [0, 1, 7, 12, 13, 14, 15, 16]
    by = {m.id: m for m in ms}
    assert by[0].length == 0.27 and by[0].x == 5.10375 and by[0].y == 0 and by[0].z == 0.3
    assert (by[0].roll, by[0].pitch, by[0].yaw) == (0, -1.5708, 0)
    assert (by[7].roll, by[7].pitch, by[7].yaw) == (1.5708, 0, 0.25) and by[7].length == 0.2
    assert (by[12].x, by[12].y, by[12].z, by[12].roll, by[12].pitch, by[12].yaw) == (-2.5, 0.1, 0, 0, 0, 0)
    assert (by[13].z, by[13].roll, by[13].pitch, by[13].yaw) == (3, 0, 0, 0)
    assert (by[14].roll, by[14].pitch, by[14].yaw) == (0.1, 0, 0)
    assert (by[15].roll, by[15].pitch, by[15].yaw) == (0.1, 0.2, 0)
    assert (by[16].x, by[16].y, by[16].z, by[16].yaw) == (1, 2, 0, 0)                  # exactly the four mandatory fields
    for m in ms:
        want = Rotation.from_euler("xyz", [m.roll, m.pitch, m.yaw]).as_quat()
        assert np.allclose(m.q, want if np.dot(m.q, want) >= 0 else -want, atol=1e-14)
    # the same through a file
    f = tmp_path / "map.txt"
    f.write_text(MAP_TEXT)
    assert formats.load_map(str(f)) == ms
    with pytest.raises(B2AError):
        formats.load_map(str(tmp_path / "missing.txt"))


def test_parse_map_short_and_malformed_lines():
    assert [m.id for m in formats.parse_map("3 0.27 1\n4 0.27 1 2\n")] == [4]          # fewer than id length x y: line skipped
    assert formats.parse_map("1 0.27 1 2\n-2 0.27 1 2\n3 0.27 1 2\n") == []            # first character not a digit: the whole map is dropped
    assert formats.parse_map("") == [] and formats.parse_map("\n\n# only comments\n") == []
    assert [m.id for m in formats.parse_map("5 0.27 1 2")] == [5]                      # no trailing newline
    with pytest.raises(B2AError):
        formats.parse_map("1 1 1 1\n2 1 1 1\n", cap=1)


def test_detected_marker_cubes_vs_reference_run():
    """ArucoSlam::toRosDetectedMarkers (aruco_slam.cpp:324-347) through b2a_pack_detected_markers against the cubes the reference
    itself produced (tests/golden/slam_scene.npz, slam_synth.npz: detm_*): same markers pass the range gate, same order,
    positions and orientations to rounding"""
    import os
    for name in ("slam_scene", "slam_synth"):
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
        f = total = dropped = 0
        while "det_ids_%d" % f in g.files:
            ms = formats.detected_markers(g["det_ids_%d" % f], g["det_rvecs_%d" % f], g["det_tvecs_%d" % f], float(g["marker_length"]),
                                          float(g["useful_distance_threshold"]), None, g["r2c_t"])
            assert np.array_equal([m.id for m in ms], g["detm_id_%d" % f]), (name, f)
            if ms:
                assert np.abs(np.array([[m.x, m.y, m.z] for m in ms]) - g["detm_pos_%d" % f]).max() < 1e-12
                assert np.abs(np.array([m.q for m in ms]) - g["detm_q_%d" % f]).max() < 1e-12
                assert all(m.length == float(g["marker_length"]) for m in ms)
            total += len(ms); dropped += len(g["det_ids_%d" % f]) - len(ms); f += 1
        assert total > 50 and dropped > 0


def test_detected_marker_cubes_rotated_mount():
    """a camera mounted a quarter turn about z: positions and orientations follow tf2::doTransform (out = T * in)"""
    s = np.sqrt(0.5)
    ms = formats.detected_markers([7], [[0.0, 0.0, 0.0]], [[1.0, 2.0, 0.5]], 0.27, 3.0, (0, 0, s, s), (0.1, 0.0, 0.3))
    assert len(ms) == 1 and ms[0].id == 7
    assert np.allclose([ms[0].x, ms[0].y, ms[0].z], [-2.0 + 0.1, 1.0, 0.8], atol=1e-12)
    assert np.allclose(ms[0].q, (0, 0, s, s), atol=1e-12)
    assert formats.detected_markers([7], [[0.0, 0.0, 0.0]], [[0.0, 0.0, 3.5]], 0.27, 3.0) == []          # beyond the useful range


PARAMETERS_YAML = """
covariance:
    Q_k: 0.01           ## Error coefficient of encoder
    R_x: 100
    R_y: 100
    R_theta: 10
odom:
    kl: 0.05
    kr: 0.05
    b: 0.09
aruco:
    markers_dictionary: 16 ## cv::aruco::PREDEFINED_DICTIONARY_NAME
    marker_length: 0.27
frame:
    world_frame: "world"
    robot_frame_base: "base_link"
const:
    USEFUL_DISTANCE_THRESHOLD: 4
"""


def test_parameters_yaml_as_the_node_reads_it():
    """parseArucoSlamIniteData (aruco_slam_node.cpp:146-164): the values of the shipped parameters.yaml, and the threshold key the
    node misspells (so the file's 4 is never read and the default 3 stays)"""
    import os
    p = formats.load_parameters(PARAMETERS_YAML)
    assert p["slam"] == dict(Q_k=0.01, R_x=100.0, R_y=100.0, R_theta=10.0, kl=0.05, kr=0.05, b=0.09, useful_distance_threshold=3.0)
    assert p["markers_dictionary"] == 16 and p["marker_length"] == 0.27 and p["frames"]["robot_frame_base"] == "base_link"
    assert formats.load_parameters("const:\n    USEFUL_DISTANCE_THRESHOLD_: 2.5\n")["slam"]["useful_distance_threshold"] == 2.5
    assert formats.load_parameters("")["slam"]["R_x"] == 100.0 and formats.load_parameters("")["marker_length"] == 0.27      # defaults survive missing keys
    ref = "/root/reference/parameters.yaml"
    if os.path.exists(ref):                                    # authoring container only: the shipped file itself
        q = formats.load_parameters(ref)
        assert q["slam"] == p["slam"] and q["markers_dictionary"] == 16 and q["marker_length"] == 0.27
        assert q["topics"] == {"image": "/camera/image_raw", "encoder": "/encoder"}


def test_camera_yaml():
    import os
    K, D = formats.load_camera("camera:\n    fx: 525.5\n    fy: 524.5\n    cx: 472.0\n    cy: 264.0\n    k1: 0.04\n    k2: -0.05\n    p1: -0.003\n    p2: -0.004\n    k3: 0.01\n")
    assert np.array_equal(K, [[525.5, 0, 472.0], [0, 524.5, 264.0], [0, 0, 1]]) and np.array_equal(D, [0.04, -0.05, -0.003, -0.004, 0.01])
    ref = "/root/reference/default.yaml"
    if os.path.exists(ref):
        K, D = formats.load_camera(ref)
        assert abs(K[0, 0] - 525.2866213437447) < 1e-12 and abs(K[1, 2] - 264.77181506420266) < 1e-12 and abs(D[4] - 0.01110263483766991) < 1e-15

"""CPU tier: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/b2aruco.h declares (no compute calls here: there is no GPU in this tier)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so_path():
    from aruco_slam_b200 import _lib
    return _lib.build()


def test_header_symbols_exported(so_path):
    from aruco_slam_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "b2aruco.h")).read()
    declared = sorted(set(re.findall(r"^\s*(?:const char \*|int\s+|void \*|void\s+)(b2a_[a-z_0-9]+)\s*\(", hdr, re.M)))
    assert declared, "no declarations found"
    L = ctypes.CDLL(so_path)
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(_lib.SYMBOLS) == declared


def test_predefined_dictionary_tables_without_gpu(so_path):
    """b2a_get_predefined_dictionary is host-only data: same bytes as the package tables."""
    import numpy as np
    from aruco_slam_b200 import aruco, dictionaries as D
    for did in range(22):
        a, b = aruco.library_dictionary(did), D.getPredefinedDictionary(did)
        assert np.array_equal(a.table, b.table) and a.max_correction_bits == b.max_correction_bits


def test_create_fails_loudly_without_gpu(so_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from aruco_slam_b200 import aruco, dictionaries as D
    with pytest.raises(aruco.B2AError):
        aruco.ArucoDetector(D.getPredefinedDictionary(0), max_shape=(64, 64))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "aruco_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f


def _build_demo(so_path, tmp_path):
    import subprocess
    exe = str(tmp_path / "b2a_demo")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "b2a_demo.cpp"),
                           "-L" + os.path.dirname(so_path), "-lb2aruco", "-Wl,-rpath," + os.path.dirname(so_path), "-o", exe])
    return exe


def test_cpp_shim_compiles_and_fails_loudly_without_gpu(so_path, tmp_path):
    """include/b2aruco.hpp (the cv::aruco-shaped C++ surface) builds against the C ABI with plain g++;
    without a GPU the demo exits 1 with the library's error instead of falling back to anything."""
    import subprocess
    import torch
    exe = _build_demo(so_path, tmp_path)
    pgm = tmp_path / "f.pgm"
    with open(pgm, "wb") as f:
        f.write(b"P5\n64 48\n255\n" + bytes(64 * 48))
    r = subprocess.run([exe, str(pgm), "0"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "0 markers" in r.stdout
    else:
        assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_shim_matches_python_surface(so_path, tmp_path):
    """the C++ shim's detectMarkers / estimatePoseSingleMarkers print the same ids, corners and poses
    as the Python mirror on a rendered frame"""
    import subprocess
    import numpy as np
    from aruco_slam_b200 import aruco, synth, dictionaries as D
    exe = _build_demo(so_path, tmp_path)
    fr = synth.render_config("C1", 1).image
    H, W = fr.shape
    pgm = tmp_path / "c1.pgm"
    with open(pgm, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (W, H) + fr.tobytes())
    out = subprocess.run([exe, str(pgm), "0", "0.27"], capture_output=True, text=True, check=True).stdout.splitlines()
    det = aruco.ArucoDetector(D.getPredefinedDictionary(0), max_shape=fr.shape)
    K = np.array([[1400.0, 0, W / 2.0], [0, 1400.0, H / 2.0], [0, 0, 1]])
    r = det.detect_pose_batch(fr, 0.27, K, np.zeros(0))
    assert out[0].startswith("%d markers, %d rejected" % (len(r.ids[0]), len(r.rejected[0])))
    for line, mid, c, rv, tv in zip(out[1:], r.ids[0], r.corners[0], r.rvecs[0], r.tvecs[0]):
        nums = [float(x) for x in re.findall(r"-?\d+\.\d+", line)]
        assert line.startswith("id %d " % mid)
        assert np.allclose(nums[:8], c.reshape(-1), atol=0.051)
        assert np.allclose(nums[8:11], rv.reshape(-1), atol=2e-5) and np.allclose(nums[11:14], tv.reshape(-1), atol=2e-5)
    det.close()


def test_cpp_arucoslam_class_compiles_and_fails_loudly_without_gpu(so_path, tmp_path):
    """b2a::ArucoSlam (the reference's class interface over the C ABI, include/b2aruco.hpp) builds with plain g++; without a GPU its
    constructor throws the library's error"""
    import subprocess
    import torch
    exe = _build_demo(so_path, tmp_path)
    pgm = tmp_path / "f.pgm"
    with open(pgm, "wb") as f:
        f.write(b"P5\n64 48\n255\n" + bytes(64 * 48))
    r = subprocess.run([exe, str(pgm), "0", "0.27", "slam"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "pose " in r.stdout and "marked 64 x 48 x 1" in r.stdout
    else:
        assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_arucoslam_matches_python_mirror(so_path, tmp_path):
    """the same three addEncoder / addImage steps through b2a::ArucoSlam (C++) and through aruco_slam_b200.slam.ArucoSlam: pose record,
    mapped markers, detected-marker cubes and the marked image agree"""
    import subprocess
    import numpy as np
    from aruco_slam_b200 import slam, synth, formats
    exe = _build_demo(so_path, tmp_path)
    fr = synth.render_config("C1", 1).image
    H, W = fr.shape
    pgm = tmp_path / "c1.pgm"
    with open(pgm, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (W, H) + fr.tobytes())
    # markers of 50-90 px seen with f = 1400 px: a 0.1 m marker is 1.6-2.8 m away, inside the 3 m useful range
    out = subprocess.run([exe, str(pgm), "0", "0.1", "slam"], capture_output=True, text=True, check=True).stdout.splitlines()
    s = slam.ArucoSlam(0, 0.1, image_shape=fr.shape, r2c_tx=0.12, r2c_ty=0.0, R_x=0.1, R_y=0.1, R_theta=0.01)      # the demo's values
    K = np.array([[1400.0, 0, W / 2.0], [0, 1400.0, H / 2.0], [0, 0, 1]])
    s.setCameraParameters(K, np.zeros(0))
    s.addEncoder(0.0, 0.0, 0.0)
    for _ in range(3):
        s.addEncoder(2.0, 2.5, 0.05)
        s.addImage(fr)
    nums = lambda line: [float(x) for x in re.findall(r"-?\d+\.\d+(?:e[-+]\d+)?", line)]
    pose = formats.robot_pose(s)
    got = nums(out[0])
    assert out[0].startswith("pose ") and np.allclose(got[:3], pose.position, atol=1e-8) and np.allclose(got[3:7], pose.orientation, atol=1e-8)
    assert np.allclose(got[7:10], [pose.covariance[0, 0], pose.covariance[1, 1], pose.covariance[5, 5]], rtol=1e-6, atol=1e-12)
    mapped = [l for l in out if l.startswith("mapped ")]
    want = formats.detected_map(s)
    assert len(mapped) == len(want) == 4
    for line, m in zip(mapped, want):
        assert int(line.split()[1]) == m.id and np.allclose(nums(line)[-3:], [m.x, m.y, m.yaw], atol=1e-8)
    det = [l for l in out if l.startswith("detected ")]
    wd = s.toRosDetectedMarkers(r2c_t=(0.12, 0.0, 0.25))
    assert len(det) == len(wd) > 0
    for line, m in zip(det, wd):
        assert int(line.split()[1]) == m.id and np.allclose(nums(line)[-7:], [m.x, m.y, m.z, *m.q], atol=1e-8)
    marked = s.getMarkedImg()
    assert out[-1].startswith("marked %d x %d x 1 sum %d dim %d" % (W, H, int(marked.astype(np.int64).sum()), s.dim))
    s.close()


def test_host_staging_memory_needs_a_device(so_path):
    """b2a_host_alloc is cudaHostAlloc: without a CUDA device it reports the error instead of handing out pageable memory"""
    import torch
    from aruco_slam_b200 import _lib
    if torch.cuda.is_available():
        hb = _lib.HostBuffer((3, 5), write_combined=True)
        hb.array[...] = 7
        assert hb.array.shape == (3, 5) and int(hb.array.sum()) == 105
        hb.close()
    else:
        with pytest.raises(_lib.B2AError):
            _lib.HostBuffer((3, 5))


def test_cpp_shim_draw_wrapper_compiles(so_path, tmp_path):
    """b2a::aruco::drawDetectedMarkers (aruco_slam.cpp:318-319 through the shim) builds with plain g++ -Wall -Wextra; without a GPU the
    program stops at the detect call with the library's error"""
    import subprocess
    import torch
    src = tmp_path / "draw.cpp"
    src.write_text('''#include "b2aruco.hpp"
#include <cstdio>
int main() {
    try {
        std::vector<uint8_t> px(64 * 48, 120);
        auto dict = b2a::aruco::getPredefinedDictionary(0);
        std::vector<std::vector<b2a::Point2f>> corners; std::vector<int> ids;
        b2a::aruco::detectMarkers(b2a::Image{px.data(), 64, 48, 1, 0}, dict, corners, ids);
        b2a::aruco::drawDetectedMarkers(b2a::aruco::MutableImage{px.data(), 64, 48, 1, 0}, corners, ids);
        std::printf("%zu markers drawn\\n", ids.size());
    } catch (const b2a::Exception &e) { std::fprintf(stderr, "b2aruco error %d: %s\\n", e.code, e.what()); return 1; }
    return 0;
}
''')
    exe = str(tmp_path / "draw")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src), "-L" + os.path.dirname(so_path), "-lb2aruco",
                           "-Wl,-rpath," + os.path.dirname(so_path), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "0 markers drawn" in r.stdout
    else:
        assert r.returncode == 1 and "no CUDA device" in r.stderr

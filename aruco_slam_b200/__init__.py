"""b2aruco -- B200-native ArUco detect + pose (+ EKF landmark update) front-end.

Drop-in for the `cv::aruco::detectMarkers` / `estimatePoseSingleMarkers` path of
gitAugust/Aruco_Slam (reference src/aruco_slam.cpp:313-314) and the EKF it feeds.
All compute is hand-written CUDA for sm_100a behind the C ABI in include/b2aruco.h;
this package is the host-side mirror of the reference's call surface.
"""
from . import dictionaries, synth  # noqa: F401
from .dictionaries import Dictionary, getPredefinedDictionary  # noqa: F401

__all__ = ["dictionaries", "synth", "Dictionary", "getPredefinedDictionary", "aruco", "slam"]


def __getattr__(name):
    if name in ("aruco", "slam", "_lib"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)

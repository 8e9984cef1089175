"""invcompcamtrack_b200 — B200 (sm_100a) implementation of the inverse-compositional GN pose-tracking path of
catree/InvCompCamTrack behind a C ABI (include/ictrack.h, libictrack.so) that Python reaches through ctypes, the
way the reference's misc_src scripts reach triang.c (misc_src/func_util_geom.py:582-604).

There is no CPU fallback: importing works without a GPU (so that the symbol table can be checked), every compute
call fails loudly with IctError when the library or a CUDA device is missing.
"""
from .api import (IctError, OptParam, make_optparam, lib, lib_path, device_count, pyramid_layout, pyramid_build,
                  camera_levels, Frames, Tracker, track_pair, TRACE_FLOATS, launch_count)

__all__ = ["IctError", "OptParam", "make_optparam", "lib", "lib_path", "device_count", "pyramid_layout",
           "pyramid_build", "camera_levels", "Frames", "Tracker", "track_pair", "TRACE_FLOATS", "launch_count"]

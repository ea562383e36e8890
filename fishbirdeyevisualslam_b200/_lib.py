"""ctypes loader for libfbe_b200.so (the sm_100a CUDA library behind include/fbe_cabi.h).

There is no CPU path: if the library is missing or no CUDA device is present every entry point fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libfbe_b200.so")

FBE_OK, FBE_E_INVALID, FBE_E_CUDA, FBE_E_CAPACITY, FBE_E_UNSUPPORTED = 0, -1, -2, -3, -4

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


class FbeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fbe error {code}: {msg}")
        self.code = code


class ExtractorCfg(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32), ("max_batch", C.c_int32),
                ("device", C.c_int32)]


class FrameView(C.Structure):
    _fields_ = [("kps", C.c_void_p), ("desc", C.c_void_p), ("n", C.c_int32),
                ("min_x", C.c_float), ("min_y", C.c_float), ("inv_w", C.c_float), ("inv_h", C.c_float),
                ("gcols", C.c_int32), ("grows", C.c_int32)]


class FrustumView(C.Structure):
    """fbe_frustum_view (include/fbe_cabi.h): the Frame members Frame::isInFrustum reads."""
    _fields_ = [("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("Ow", C.c_float * 3), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float),
                ("max_y", C.c_float), ("mbf", C.c_float), ("log_scale_factor", C.c_float), ("n_levels", C.c_int32)]


class PipelineCfg(C.Structure):
    _fields_ = [("front", ExtractorCfg), ("bird", ExtractorCfg),
                ("front_rows", C.c_int32), ("front_cols", C.c_int32), ("bird_rows", C.c_int32), ("bird_cols", C.c_int32),
                ("batch", C.c_int32), ("nn_ratio", C.c_float), ("check_orientation", C.c_int32),
                ("front_window", C.c_int32), ("bird_window", C.c_int32), ("device", C.c_int32),
                ("front_fisheye", C.c_int32), ("front_K", C.c_float * 4), ("front_D", C.c_float * 4),
                ("front_row_cap", C.c_int32)]


class PipelineFeatures(C.Structure):
    """fbe_pipeline_features: host buffers for the extractor outputs of one step (any pointer may be NULL)."""
    _fields_ = [("front_kps", C.c_void_p), ("front_desc", C.c_void_p), ("bird_kps", C.c_void_p), ("bird_desc", C.c_void_p)]


PAIR_RESULT_DTYPE = np.dtype([("n_front", "<i4"), ("n_bird", "<i4"), ("front_matches", "<i4"), ("bird_matches", "<i4")])

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FbeError(FBE_E_CUDA, f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(LIB_PATH)
        lib.fbe_last_error.restype = C.c_char_p
        lib.fbe_kernel_launch_count.restype = C.c_uint64
        lib.fbe_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_size_t, C.c_void_p, C.c_void_p,
                                    C.c_int32, C.c_void_p]
        lib.fbe_extract_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_size_t, C.c_void_p,
                                          C.c_void_p, C.c_int32, C.c_void_p]
        lib.fbe_pyramid_level.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        lib.fbe_pyramid_geometry.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        lib.fbe_extract_pyramid.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_size_t, C.c_void_p, C.c_void_p,
                                            C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.fbe_grid_assign.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32,
                                        C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.fbe_matcher_create.argtypes = [C.c_float, C.c_int32, C.c_int32, C.c_void_p]
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != FBE_OK:
        raise FbeError(rc, (load().fbe_last_error() or b"").decode())


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)

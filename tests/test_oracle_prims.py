"""The oracle's OpenCV primitives (oracle/prim.hpp) against (a) committed cv2 4.13.0 known-answer vectors and
(b) cv2 itself when it is importable.  Bit-exact for the integer primitives; fastAtan2 within 1e-4 rad."""
import ctypes as C

import numpy as np
import pytest


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def o_resize(L, src, dw, dh):
    src = np.ascontiguousarray(src)
    dst = np.zeros((dh, dw), np.uint8)
    L.orc_resize_linear_u8(P(src), src.shape[1], src.shape[0], src.strides[0], P(dst), dw, dh, dw)
    return dst


def o_blur(L, src):
    src = np.ascontiguousarray(src)
    dst = np.zeros_like(src)
    L.orc_gauss7_u8(P(src), src.shape[1], src.shape[0], src.strides[0], P(dst), dst.strides[0])
    return dst


def o_fast(L, im, th):
    im = np.ascontiguousarray(im)
    out = np.zeros((im.size, 3), np.int32)
    n = L.orc_fast9_nms(P(im), im.shape[1], im.shape[0], im.strides[0], int(th), P(out), len(out))
    return out[:n].copy()


def o_border(L, src, b=19):
    src = np.ascontiguousarray(src)
    dst = np.zeros((src.shape[0] + 2 * b, src.shape[1] + 2 * b), np.uint8)
    L.orc_border_reflect101_u8(P(src), src.shape[1], src.shape[0], src.strides[0], P(dst), dst.strides[0], b)
    return dst


def test_golden_resize(oracle, prims_golden):
    g, L = prims_golden, oracle.lib()
    for i in range(4):
        dst = g[f"resize_dst{i}"]
        assert np.array_equal(o_resize(L, g[f"resize_src{i}"], dst.shape[1], dst.shape[0]), dst)


def test_golden_border_blur_fast(oracle, prims_golden):
    g, L = prims_golden, oracle.lib()
    assert np.array_equal(o_border(L, g["border_src"]), g["border_dst"])
    for i in range(3):
        assert np.array_equal(o_blur(L, g[f"blur_src{i}"]), g[f"blur_dst{i}"])
    for i in range(4):
        assert np.array_equal(o_fast(L, g[f"fast_src{i}"], int(g[f"fast_th{i}"])), g[f"fast_out{i}"])


def test_golden_atan_round(oracle, prims_golden):
    g, L = prims_golden, oracle.lib()
    got = np.array([L.orc_fast_atan2(float(y), float(x)) for y, x in g["atan_yx"]], np.float32)
    assert np.max(np.abs(got - g["atan_deg"])) * np.pi / 180 < 1e-4      # north_star tolerance (observed: 0)
    assert [L.orc_cv_round_f(float(v)) for v in g["round_in"]] == g["round_out"].tolist()


def test_blur_kernel_is_integer_exact(oracle):
    # an impulse reproduces the integer taps [18,34,48,56,48,34,18] (x56 centre) >> 16
    L = oracle.lib()
    im = np.zeros((15, 15), np.uint8)
    im[7, 7] = 255
    out = o_blur(L, im)
    k = np.array([18, 34, 48, 56, 48, 34, 18])
    exp = (np.outer(k, k) * 255 + 32768) >> 16
    assert np.array_equal(out[4:11, 4:11], exp)


def test_fast_threshold_nesting(oracle):
    # FAST(15) == {k in FAST(5): response >= 15}: the fact the single-pass GPU cell kernel relies on
    from fishbirdeyevisualslam_b200 import synth
    L = oracle.lib()
    for seed in range(6):
        im = synth.frame(37, 38, 700 + seed)
        lo, hi = o_fast(L, im, 5), o_fast(L, im, 15)
        assert np.array_equal(lo[lo[:, 2] >= 15], hi)


def test_against_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    from fishbirdeyevisualslam_b200 import synth
    L = oracle.lib()
    rng = np.random.default_rng(3)
    for (sw, sh, dw, dh) in [(640, 480, 533, 400), (384, 384, 320, 320), (179, 134, 149, 112), (950, 400, 792, 333)]:
        src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        assert np.array_equal(o_resize(L, src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR))
    src = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    assert np.array_equal(o_blur(L, src), cv2.GaussianBlur(src, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101))
    assert np.array_equal(o_border(L, src), cv2.copyMakeBorder(src, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))
    im = synth.frame(120, 160, 77)
    for th in (5, 15, 20):
        det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        ref = np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(im)], np.int32).reshape(-1, 3)
        assert np.array_equal(o_fast(L, im, th), ref)

"""Row f-3: the oracle's restatement of cv::ORB::create(2000) detect / compute (oracle/cvorb_oracle.cpp) pinned to cv2 4.13.0 --
committed known answers (tests/golden/cvorb.npz, tools/gen_golden_cvorb.py) and, where cv2 is importable, live calls:
keypoints byte for byte INCLUDING their order (std::nth_element / std::partition of KeyPointsFilter::retainBest), Harris
responses and angles as float bit patterns, descriptors bit for bit."""
import zlib

import numpy as np
import pytest

import bird_scenes as S
from fishbirdeyevisualslam_b200 import synth
from fishbirdeyevisualslam_b200._lib import KP_DTYPE


@pytest.fixture(scope="module")
def gold():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "cvorb.npz"))


def case_image(i):
    return S.bird_image(i) if i < 5 else np.ascontiguousarray(synth.road_frame(384, 384, 77))


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def test_golden_primitives(oracle, gold):
    assert np.array_equal(oracle.resize_linear_exact(gold["rs_src"], 108, 89), gold["rs_dst"])
    assert np.array_equal(oracle.cvorb_blur(gold["blur_src"]), gold["blur_dst"])


def test_golden_detect_and_compute(oracle, gold):
    for i in range(int(gold["ncases"])):
        img = case_image(i)
        assert synth.crc(img) == int(gold[f"img_crc{i}"])
        det = oracle.cvorb_detect(img, S.bird_mask(i))
        assert len(det) == int(gold[f"det_n{i}"]) and crc(det) == int(gold[f"det_crc{i}"]), f"detect case {i}"
        if i < 2:
            assert det.tobytes() == gold[f"det{i}"].tobytes()
            moved = gold[f"moved{i}"]
            k, d = oracle.cvorb_compute(img, moved)
            assert k.tobytes() == gold[f"cmp_kps{i}"].tobytes() and np.array_equal(d, gold[f"cmp_desc{i}"])


def test_golden_full_bird_block(oracle, gold):
    """detect -> cornerSubPix (oracle/bird_oracle.cpp) -> compute, all restated, equals the chain of real cv2 calls."""
    for i in range(int(gold["ncases"])):
        img = case_image(i)
        det = oracle.cvorb_detect(img, S.bird_mask(i))
        xy, _ = oracle.corner_subpix(img, np.stack([det["x"], det["y"]], 1))
        moved = det.copy()
        moved["x"], moved["y"] = xy[:, 0], xy[:, 1]
        k, d = oracle.cvorb_compute(img, moved)
        assert len(k) == int(gold[f"cmp_n{i}"]) and crc(k) == int(gold[f"cmp_kps_crc{i}"]) and crc(d) == int(gold[f"cmp_desc_crc{i}"]), f"case {i}"


def _cv2():
    return pytest.importorskip("cv2")


def _cvkps(kps):
    return np.array([(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id) for k in kps], dtype=KP_DTYPE)


def _tocv(cv2, a):
    return [cv2.KeyPoint(float(k["x"]), float(k["y"]), float(k["size"]), float(k["angle"]), float(k["response"]), int(k["octave"]),
                         int(k["class_id"])) for k in a]


@pytest.mark.parametrize("shape", [(384, 384), (300, 421), (200, 640)])
def test_live_cv2_detect(oracle, shape):
    cv2 = _cv2()
    rng = np.random.default_rng(shape[0])
    for seed in range(3):
        img = synth.frame(shape[0], shape[1], 500 + seed)
        mask = None
        if seed == 1:
            mask = (rng.integers(0, 4, shape) > 0).astype(np.uint8) * 255          # salt-and-pepper holes
            mask[: shape[0] // 3] = 3                                             # non-binary "keep"
        if seed == 2:
            mask = np.zeros(shape, np.uint8)
            mask[40:-40, 50:-50] = 200
        for nf in (2000, 500):
            ref = _cvkps(cv2.ORB_create(nf).detect(img, mask))
            got = oracle.cvorb_detect(img, mask, nf)
            assert len(ref) == len(got) and ref.tobytes() == got.tobytes(), (shape, seed, nf)


def test_live_cv2_compute(oracle):
    cv2 = _cv2()
    rng = np.random.default_rng(11)
    for seed in range(3):
        img = synth.frame(384, 384, 600 + seed) if seed < 2 else np.ascontiguousarray(synth.road_frame(384, 384, 5))
        orb = cv2.ORB_create(2000)
        det = _cvkps(orb.detect(img, None))
        for variant in ("asis", "jitter", "border", "unsorted", "angles", "few_levels"):
            k = det.copy()
            if variant == "jitter":
                k["x"] += rng.uniform(-3, 3, len(k)).astype(np.float32); k["y"] += rng.uniform(-3, 3, len(k)).astype(np.float32)
            if variant == "border":        # KeyPointsFilter::runByImageBorder rounds the (sub-pixel) position before the test
                k["x"][:200] = rng.uniform(28, 34, 200).astype(np.float32); k["y"][200:400] = rng.uniform(350, 356, 200).astype(np.float32)
                k["x"][400:420] = 30.5; k["x"][420:440] = 31.5; k["x"][440:460] = 352.5; k["x"][460:480] = 353.5
            if variant == "unsorted":      # regrouped by octave, stable inside an octave
                k = k[rng.permutation(len(k))]
            if variant == "angles":
                k["angle"] = rng.uniform(0, 360, len(k)).astype(np.float32)
            if variant == "few_levels":    # the pyramid is only built up to the largest octave present
                k = k[k["octave"] <= 2]
            rk, rd = orb.compute(img, _tocv(cv2, k))
            gk, gd = oracle.cvorb_compute(img, k)
            assert _cvkps(rk).tobytes() == gk.tobytes() and np.array_equal(rd, gd), (seed, variant)


def test_live_cv2_primitives(oracle):
    cv2 = _cv2()
    rng = np.random.default_rng(3)
    for (h, w, dh, dw) in [(384, 384, 320, 320), (320, 320, 267, 267), (480, 640, 400, 533), (107, 129, 89, 108), (50, 70, 42, 58), (64, 64, 64, 64)]:
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(oracle.resize_linear_exact(a, dw, dh), cv2.resize(a, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT))
        kern = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
        assert np.array_equal(oracle.cvorb_blur(a), cv2.sepFilter2D(a, cv2.CV_8U, kern, kern, borderType=cv2.BORDER_REFLECT_101))
    img = synth.frame(384, 384, 9)
    kern = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    assert np.array_equal(oracle.cvorb_blur(img), cv2.sepFilter2D(img, cv2.CV_8U, kern, kern, borderType=cv2.BORDER_REFLECT_101))


def test_retain_best_and_its_adversary(oracle):
    """oracle.retain_best (the real std::nth_element + std::partition) keeps exactly the k largest responses plus the ties of the
    k-th, on ordinary inputs and on the adversarial ones of oracle.antiselect (which exist to drive the GPU replay into introselect's
    heap-select exit: tests/test_gpu_bird_orb.py); the adversary's output is a permutation-valued response array."""
    rng = np.random.default_rng(3)
    for n in (50, 1000, 5000):
        for r in (rng.random(n).astype(np.float32), rng.integers(0, 8, n).astype(np.float32), oracle.antiselect(n, n // 2 - 1)):
            k = n // 2
            order, kept = oracle.retain_best(r, k)
            assert sorted(order.tolist()) == list(range(n))
            kth = np.sort(r)[::-1][k - 1]
            assert kept == int((r >= kth).sum())
            assert set(order[:kept].tolist()) == set(np.nonzero(r >= kth)[0].tolist())
    a = oracle.antiselect(2000, 999)
    assert sorted(a.tolist()) == [float(v) for v in range(1, 2001)]

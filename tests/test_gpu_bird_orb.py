"""Row f-3 on the GPU: cv::ORB::create(2000) detect / compute of the reference's bird-view block (src/Frame.cc:336-355) through the
C-ABI against the oracle restatement (pinned to cv2 4.13.0 by tests/test_cvorb_oracle.py).  Bar: everything bit-exact --
keypoint records byte for byte INCLUDING their order, float responses and angles as bit patterns, descriptors."""
import numpy as np
import pytest

import bird_scenes as S
from fishbirdeyevisualslam_b200 import synth

pytestmark = pytest.mark.gpu


def case_image(i):
    return S.bird_image(i) if i < 5 else np.ascontiguousarray(synth.road_frame(384, 384, 77))


def test_retain_best_replay_equals_std_algorithms(oracle):
    """The parallel replay of libstdc++'s introselect + partition against the real std:: algorithms (oracle side): random
    floats, heavy ties (integer FAST scores), sorted / reversed / constant inputs, tiny and large n, every n_points regime."""
    from fishbirdeyevisualslam_b200.bird_orb import retain_best
    rng = np.random.default_rng(0)
    cases = []
    for n in (1, 2, 3, 4, 5, 7, 16, 33, 100, 257, 1000, 4097, 20000):
        cases.append(rng.random(n).astype(np.float32))
        cases.append(rng.integers(20, 60, n).astype(np.float32))              # FAST-like: few distinct values
        cases.append(np.sort(rng.random(n).astype(np.float32)))
        cases.append(np.sort(rng.random(n).astype(np.float32))[::-1].copy())
        cases.append(np.full(n, 3.0, np.float32))
        cases.append(np.where(rng.random(n) < 0.5, 1.0, 2.0).astype(np.float32))
    for r in cases:
        n = len(r)
        for k in sorted({0, 1, 2, 3, n // 3, n // 2, n - 2, n - 1, n, n + 5}):
            if k < 0:
                continue
            oo, ok = oracle.retain_best(r, k)
            go, gk = retain_best(r, k)
            assert gk == ok, (n, k)
            assert np.array_equal(go[:gk], oo[:ok]), (n, k)


def test_retain_best_heap_select_fallback(oracle):
    """Adversarial responses (McIlroy's adversary played against the oracle side's real std::nth_element) drive introselect past
    its depth budget: libstdc++ then finishes with __heap_select + iter_swap, and so does the device replay -- same permutation."""
    from fishbirdeyevisualslam_b200.bird_orb import retain_best
    hit = 0
    for n in (200, 1000, 5000, 20000):
        for k in (n // 2, n // 3, (2 * n) // 3):
            r = oracle.antiselect(n, k - 1)
            assert len(np.unique(r)) == n
            oo, ok = oracle.retain_best(r, k)
            go, gk, heap = retain_best(r, k, with_flag=True)
            hit += heap
            assert gk == ok == k, (n, k)
            assert np.array_equal(go[:gk], oo[:ok]), (n, k)
            assert np.array_equal(go, oo), (n, k)               # the discarded tail is the same permutation as well
    assert hit >= 8                                            # the inputs do reach the fallback
    r = np.random.default_rng(1).random(5000).astype(np.float32)
    assert retain_best(r, 2500, with_flag=True)[2] is False    # ordinary inputs never do


@pytest.mark.parametrize("i", range(6))
def test_detect_equals_oracle(oracle, i):
    from fishbirdeyevisualslam_b200.bird_orb import BirdORB
    img, mask = case_image(i), S.bird_mask(i)
    orb = BirdORB(2000, 384, 384)
    got = orb.detect(img, mask)
    ref = oracle.cvorb_detect(img, mask)
    assert len(got) == len(ref) > 500 and got.tobytes() == ref.tobytes()
    orb.close()


def test_detect_other_sizes_and_feature_counts(oracle):
    from fishbirdeyevisualslam_b200.bird_orb import BirdORB
    rng = np.random.default_rng(4)
    for (h, w), nf in (((300, 421), 2000), ((200, 640), 500), ((384, 384), 100), ((129, 140), 2000), ((70, 90), 300)):
        img = synth.frame(h, w, 800 + h)
        mask = (rng.integers(0, 5, (h, w)) > 0).astype(np.uint8) * 200
        orb = BirdORB(nf, h, w)
        for m in (None, mask):
            got, ref = orb.detect(img, m), oracle.cvorb_detect(img, m, nf)
            assert len(got) == len(ref) and got.tobytes() == ref.tobytes(), ((h, w), nf, m is None)
        orb.close()


def test_detect_degenerate_images(oracle):
    from fishbirdeyevisualslam_b200.bird_orb import BirdORB
    orb = BirdORB(2000, 384, 384)
    flat = np.full((384, 384), 77, np.uint8)
    assert len(orb.detect(flat)) == 0 == len(oracle.cvorb_detect(flat))
    img = case_image(0)
    none = np.zeros((384, 384), np.uint8)
    assert len(orb.detect(img, none)) == 0 == len(oracle.cvorb_detect(img, none))
    sat = np.where(np.indices((384, 384)).sum(0) % 16 < 8, 255, 0).astype(np.uint8)       # stripes: massive response ties
    got, ref = orb.detect(sat), oracle.cvorb_detect(sat)
    assert got.tobytes() == ref.tobytes()
    orb.close()


def test_compute_equals_oracle(oracle):
    from fishbirdeyevisualslam_b200.bird_orb import BirdORB
    rng = np.random.default_rng(11)
    orb = BirdORB(2000, 384, 384)
    for i in (0, 1, 5):
        img = case_image(i)
        det = oracle.cvorb_detect(img, S.bird_mask(i))
        for variant in ("asis", "jitter", "border", "unsorted", "angles", "few_levels", "empty"):
            k = det.copy()
            if variant == "jitter":
                k["x"] += rng.uniform(-3, 3, len(k)).astype(np.float32); k["y"] += rng.uniform(-3, 3, len(k)).astype(np.float32)
            if variant == "border":
                k["x"][:200] = rng.uniform(28, 34, 200).astype(np.float32); k["y"][200:400] = rng.uniform(350, 356, 200).astype(np.float32)
                k["x"][400:420] = 30.5; k["x"][420:440] = 31.5; k["x"][440:460] = 352.5; k["x"][460:480] = 353.5
            if variant == "unsorted":
                k = k[rng.permutation(len(k))]
            if variant == "angles":
                k["angle"] = rng.uniform(0, 360, len(k)).astype(np.float32)
            if variant == "few_levels":
                k = k[k["octave"] <= 2]
            if variant == "empty":
                k = k[:0]
            gk, gd = orb.compute(img, k)
            rk, rd = oracle.cvorb_compute(img, k)
            assert gk.tobytes() == rk.tobytes() and np.array_equal(gd, rd), (i, variant)
    orb.close()


def test_whole_bird_block_and_ragged_batch(oracle):
    """fbe_bird_features (detect -> GuidenceKeyBirdPts -> cornerSubPix -> compute, device-resident) for a batch of different
    frames equals the oracle chain frame by frame, and equals the three separate calls."""
    from fishbirdeyevisualslam_b200.bird_orb import BirdORB
    B = 4
    imgs = np.stack([case_image(i) for i in (0, 1, 2, 5)])
    masks = np.stack([np.full((384, 384), 255, np.uint8) if S.bird_mask(i) is None else S.bird_mask(i) for i in (0, 1, 2, 5)])
    contours = np.stack([S.contour_image(i) for i in range(B)])
    orb = BirdORB(2000, 384, 384, max_batch=B)
    out = orb.features_batch(imgs, masks, contours)
    for b in range(B):
        det = oracle.cvorb_detect(imgs[b], masks[b])
        keep = oracle.bird_near_edges(contours[b], np.stack([det["x"], det["y"]], 1)).astype(bool)
        kept = det[keep]
        xy, _ = oracle.corner_subpix(imgs[b], np.stack([kept["x"], kept["y"]], 1))
        moved = kept.copy()
        moved["x"], moved["y"] = xy[:, 0], xy[:, 1]
        rk, rd = oracle.cvorb_compute(imgs[b], moved)
        gk, gd, ndet = out[b]
        assert ndet == len(det) and 0 < len(rk) < len(det)
        assert gk.tobytes() == rk.tobytes() and np.array_equal(gd, rd), b
    # no contour: every detection is refined and described
    gk, gd, ndet = orb.features(imgs[0], None, None)
    det = oracle.cvorb_detect(imgs[0], None)
    xy, _ = oracle.corner_subpix(imgs[0], np.stack([det["x"], det["y"]], 1))
    moved = det.copy()
    moved["x"], moved["y"] = xy[:, 0], xy[:, 1]
    rk, rd = oracle.cvorb_compute(imgs[0], moved)
    assert ndet == len(det) and gk.tobytes() == rk.tobytes() and np.array_equal(gd, rd)
    # batch == single calls
    single = BirdORB(2000, 384, 384)
    for b in range(B):
        sk, sd, sn = single.features(imgs[b], masks[b], contours[b])
        assert sk.tobytes() == out[b][0].tobytes() and np.array_equal(sd, out[b][1]) and sn == out[b][2]
    dets = orb.detect_batch(imgs, masks)
    for b in range(B):
        assert dets[b].tobytes() == oracle.cvorb_detect(imgs[b], masks[b]).tobytes()
    single.close()
    orb.close()

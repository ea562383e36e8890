"""Row f-3 timing probe: the reference's bird feature block (cv::ORB detect -> nearEdges -> cornerSubPix -> cv::ORB compute,
src/Frame.cc:336-355) on the GPU through the host-buffer C-ABI (uploads and downloads included, wall clock) beside cv2 4.13 on one
host core and the oracle restatement; single frame and a 64-frame batch.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import bird_scenes as S
from fishbirdeyevisualslam_b200.bird_orb import BirdORB
from oracle import oracle as O


def med(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return float(np.median(ts)) * 1e3


def oracle_block(img, mask, contour):
    det = O.cvorb_detect(img, mask)
    kept = det[O.bird_near_edges(contour, np.stack([det["x"], det["y"]], 1)).astype(bool)]
    xy, _ = O.corner_subpix(img, np.stack([kept["x"], kept["y"]], 1))
    moved = kept.copy(); moved["x"], moved["y"] = xy[:, 0], xy[:, 1]
    return O.cvorb_compute(img, moved)


def main():
    img, mask, contour = S.bird_image(1), S.bird_mask(1), S.contour_image(1)
    orb = BirdORB(2000, 384, 384)
    gk, gd, ndet = orb.features(img, mask, contour)
    rk, rd = oracle_block(img, mask, contour)
    out = {"what": "bird feature block of src/Frame.cc:336-355, 384x384, cv::ORB::create(2000)", "detected": ndet, "kept": len(gk),
           "parity": bool(gk.tobytes() == rk.tobytes() and np.array_equal(gd, rd)),
           "gpu_ms_single_frame_host_api": med(lambda: orb.features(img, mask, contour), 30),
           "gpu_ms_detect_only": med(lambda: orb.detect(img, mask), 30),
           "oracle_ms_one_core": med(lambda: oracle_block(img, mask, contour), 3, 1)}
    B = 64
    imgs = np.stack([S.bird_image(i % 5) for i in range(B)]); masks = np.stack([mask] * B); contours = np.stack([S.contour_image(i % 4) for i in range(B)])
    ob = BirdORB(2000, 384, 384, max_batch=B)
    out["gpu_ms_per_frame_batch64_host_api"] = med(lambda: ob.features_batch(imgs, masks, contours), 5) / B
    try:
        import cv2
        cv2.setNumThreads(1)
        crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 40, 0.001)

        def cvblock():
            o = cv2.ORB_create(2000)
            k = o.detect(img, mask)
            pts = np.float32([p.pt for p in k]).reshape(-1, 1, 2)
            cv2.cornerSubPix(img, pts, (5, 5), (-1, -1), crit)
            o.compute(img, k)
        out["cv2_ms_one_core_detect_subpix_compute"] = med(cvblock, 10)
        out["cv2_version"] = cv2.__version__
    except Exception as e:          # cv2 not importable on this box
        out["cv2"] = repr(e)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Seeded local-map scenes for the Frame::isInFrustum row: a camera pose, map points in front of / behind / beside it,
normals mostly facing the camera, distance ranges that accept most points and reject some on either side."""
import numpy as np

f32 = np.float32


def scene(seed, n=3000, cv2_pose=False):
    rng = np.random.default_rng(7000 + seed)
    # pose: rotation from a random axis-angle (double -> float32 like the tracker's cv::Mat), translation a few metres
    ax = rng.normal(0, 1, 3); ax /= np.linalg.norm(ax)
    ang = rng.uniform(-0.6, 0.6)
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = (np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx).astype(f32)
    t = rng.normal(0, 2, 3).astype(f32)
    if cv2_pose:
        import cv2                                   # Frame::UpdatePoseMatrices: mOw = -mRcw.t()*mtcw through cv::gemm
        Ow = cv2.gemm(R, t.reshape(3, 1), -1.0, None, 0.0, flags=cv2.GEMM_1_T).ravel().astype(f32)
    else:
        Ow = None                                    # taken from the golden file (tests) -- it is an INPUT of isInFrustum
    K = (f32(300.0 + 20 * seed), f32(310.5), f32(640.25), f32(360.75))
    bounds = (f32(-35.5), f32(1310.0), f32(-20.25), f32(745.0))
    # points: in camera coordinates first (so that most are visible), then moved to the world frame in double
    z = rng.uniform(-2, 30, n); z[rng.random(n) < 0.02] = 0.0
    x = rng.uniform(-1.5, 1.5, n) * np.abs(z) * 2.2 + rng.normal(0, 0.3, n)
    y = rng.uniform(-1.0, 1.0, n) * np.abs(z) * 1.3 + rng.normal(0, 0.3, n)
    Pc = np.stack([x, y, z], 1)
    Pw = ((Pc - t.astype(np.float64)) @ R.astype(np.float64)).astype(f32)          # R^T (Pc - t)
    cam_centre = (-R.astype(np.float64).T @ t.astype(np.float64))
    to_cam = cam_centre - Pw
    to_cam /= np.maximum(np.linalg.norm(to_cam, axis=1, keepdims=True), 1e-9)
    nrm = -to_cam + rng.normal(0, 0.6, (n, 3))      # PO = P - Ow points away from the camera; viewCos = PO.Pn/|PO|
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-9)
    d = np.linalg.norm(Pw - cam_centre, axis=1)
    max_dist = (d * rng.uniform(0.7, 6.0, n)).astype(f32)
    min_dist = (max_dist / f32(1.2 ** 7) * rng.uniform(0.5, 1.3, n)).astype(f32)
    # exact powers of the scale factor: quotients that land on (or next to) integers
    k = rng.integers(0, n, n // 10)
    max_dist[k] = (d[k] * 1.2 ** rng.integers(0, 8, len(k))).astype(f32)
    return dict(Rcw=R, tcw=t, Ow=Ow, K=K, bounds=bounds, mbf=f32(40.0), log_scale=f32(np.log(f32(1.2))), n_levels=8,
                pos=Pw, normal=nrm.astype(f32), min_dist=min_dist, max_dist=max_dist, cos_limit=0.5)

/*
 * fbe_cabi.h -- C ABI of the B200-native ORB front end (extract + grid + match).
 *
 * This is the drop-in boundary: plain C, POD arguments, no torch / OpenCV / C++ types.  Every entry point
 * replaces one interface of the reference (file:line relative to the FishBirdEyeVisualSLAM checkout) and is what
 * the drop-in host classes in fishbirdeyevisualslam_b200/host/ (ORBextractor / ORBmatcher with the reference's
 * signatures) and the Python mirror (fishbirdeyevisualslam_b200/*.py, ctypes) bind.
 *
 * Conventions
 *   - every function returns 0 (FBE_OK) or a negative FBE_E_* code; it never throws and never aborts;
 *   - there is NO CPU fallback: without a CUDA device (or without the sm_100a image) creation fails with
 *     FBE_E_CUDA and nothing else is callable;
 *   - handles are not re-entrant per instance, distinct handles may be used from distinct threads
 *     (each owns its stream and workspace), mirroring the reference (SURVEY 8b "Threading");
 *   - "device-resident" functions (*_dev) take/return device pointers and only enqueue work on the handle's
 *     stream; fbe_*_sync waits for it.
 */
#ifndef FBE_CABI_H
#define FBE_CABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FBE_API __attribute__((visibility("default")))
#else
#define FBE_API
#endif

#define FBE_OK 0
#define FBE_E_INVALID (-1)     /* bad argument */
#define FBE_E_CUDA (-2)        /* CUDA runtime error / no device; see fbe_last_error() */
#define FBE_E_CAPACITY (-3)    /* caller buffer or internal workspace too small */
#define FBE_E_UNSUPPORTED (-4) /* geometry outside what the kernels support */

#define FBE_MAX_LEVELS 16
#define FBE_HISTO_LENGTH 30 /* ORBmatcher::HISTO_LENGTH, src/ORBmatcher.cc:40 */
#define FBE_TH_HIGH 100     /* ORBmatcher::TH_HIGH,      src/ORBmatcher.cc:38 */
#define FBE_TH_LOW 50       /* ORBmatcher::TH_LOW,       src/ORBmatcher.cc:39 */

/* cv::KeyPoint layout (28 bytes): what ORBextractor::operator() fills, include/ORBextractor.h:59-61 */
typedef struct fbe_keypoint {
    float x, y;      /* pt, level-0 pixel coordinates (pt *= scale for octave > 0) */
    float size;      /* (int)(31 * scale[octave]) */
    float angle;     /* degrees [0,360), IC_Angle */
    float response;  /* FAST score */
    int32_t octave;
    int32_t class_id; /* always -1 */
} fbe_keypoint;

/* ORBextractor ctor arguments, src/ORBextractor.cc:410-412 (+ workspace sizing) */
typedef struct fbe_extractor_cfg {
    int32_t nfeatures;
    float scale_factor;
    int32_t nlevels;
    int32_t ini_th_fast;
    int32_t min_th_fast;
    int32_t max_batch; /* images processed per launch group (>=1); workspace is sized for it */
    int32_t device;    /* CUDA ordinal */
} fbe_extractor_cfg;

typedef struct fbe_extractor fbe_extractor;

FBE_API const char* fbe_last_error(void);
FBE_API int fbe_version(void);
/* number of kernels this library has launched since load (bench.py's "gpu_launches") */
FBE_API uint64_t fbe_kernel_launch_count(void);

/* ---- ORBextractor ----------------------------------------------------------------------------- */
/* ORBextractor::ORBextractor, src/ORBextractor.cc:410-470 */
FBE_API int fbe_extractor_create(const fbe_extractor_cfg* cfg, fbe_extractor** out);
FBE_API int fbe_extractor_destroy(fbe_extractor* e);
/* GetLevels / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares,
 * include/ORBextractor.h:63-83.  Pointers stay valid for the life of the handle. */
FBE_API int fbe_extractor_tables(const fbe_extractor* e, int32_t* nlevels, const float** scale, const float** inv_scale,
                         const float** sigma2, const float** inv_sigma2);
/* mnFeaturesPerLevel, src/ORBextractor.cc:432-446 */
FBE_API int fbe_extractor_features_per_level(const fbe_extractor* e, const int32_t** per_level);
/* upper bound on keypoints one image can produce (octree may overshoot nfeatures) */
FBE_API int fbe_extractor_max_keypoints(const fbe_extractor* e, int32_t rows, int32_t cols, int32_t* cap);

/* Host-only geometry query (no device needed): per level {w, h, fast_cols, fast_rows, cell_w, cell_h, nfeatures, octree_roots}
 * for an image of rows x cols -- the arithmetic of src/ORBextractor.cc:1111-1112, 768-787, 543, 432-446.
 * out: nlevels x 8 int32; scale/inv_scale/sigma2/inv_sigma2 (each nlevels floats) may be NULL. */
FBE_API int fbe_plan_query(const fbe_extractor_cfg* cfg, int32_t rows, int32_t cols, int32_t* out, float* scale,
                           float* inv_scale, float* sigma2, float* inv_sigma2);

/* ORBextractor::operator()(image, mask, keypoints, descriptors), src/ORBextractor.cc:1043-1105.
 * Host buffers.  img: rows x cols 8-bit, `step` bytes per row.  kps/desc: `capacity` entries / x32 bytes.
 * Empty image (img NULL or rows/cols <= 0) -> *n_out = 0 and outputs untouched, like the reference's early return.
 * Limits (FBE_E_UNSUPPORTED, never a silent difference): rows, cols <= 4095 (candidate coordinates are packed in 12 bits each);
 * every pyramid level must hold at least one 30-px FAST cell (the reference divides by zero there). */
FBE_API int fbe_extract(fbe_extractor* e, const uint8_t* img, int32_t rows, int32_t cols, size_t step, fbe_keypoint* kps,
                uint8_t* desc, int32_t capacity, int32_t* n_out);
/* nimg images of one size; kps is [nimg][capacity], desc [nimg][capacity][32], n_out [nimg]. */
FBE_API int fbe_extract_batch(fbe_extractor* e, const uint8_t* const* imgs, int32_t nimg, int32_t rows, int32_t cols,
                      size_t step, fbe_keypoint* kps, uint8_t* desc, int32_t capacity, int32_t* n_out);
/* public member mvImagePyramid, include/ORBextractor.h:85: level image of batch slot `slot` of the LAST call,
 * with its 19-px REFLECT_101 frame ((rows+38) x (cols+38) written to dst). dst may be NULL to query the size. */
FBE_API int fbe_pyramid_level(fbe_extractor* e, int32_t slot, int32_t level, uint8_t* dst_padded, size_t dst_step,
                      int32_t* rows, int32_t* cols);
/* All levels of one slot in one call (the drop-in ORBextractor::operator() fills mvImagePyramid with it): dst_padded[l] receives
 * the padded (rows+38) x (cols+38) level l with row stride dst_step[l]; the copies are enqueued together and waited for once.
 * Pinned destinations (fbe_host_alloc) make them true DMA transfers. */
FBE_API int fbe_pyramid_fetch(fbe_extractor* e, int32_t slot, uint8_t* const* dst_padded, const size_t* dst_step);
/* fbe_extract + fbe_pyramid_fetch of the same image in one call -- what ORBextractor::operator() does in the reference
 * (src/ORBextractor.cc:1052 ComputePyramid fills mvImagePyramid, then keypoints and descriptors follow): the level copies leave on a
 * side stream as soon as the pyramid is built, behind detection and description instead of after them.  dst_padded / dst_step as
 * in fbe_pyramid_fetch (sizes from fbe_pyramid_geometry). */
FBE_API int fbe_extract_pyramid(fbe_extractor* e, const uint8_t* img, int32_t rows, int32_t cols, size_t step, fbe_keypoint* kps,
                        uint8_t* desc, int32_t capacity, int32_t* n_out, uint8_t* const* dst_padded, const size_t* dst_step);
/* Level sizes (without the frame) an image of rows x cols gets; usable before the first extraction of that size. */
FBE_API int fbe_pyramid_geometry(fbe_extractor* e, int32_t rows, int32_t cols, int32_t* level_rows, int32_t* level_cols);

/* Stage taps for the parity tests (slot of the LAST call). Not part of the reference interface. */
/* vToDistributeKeys of one level in reference order: (x, y, score) triples in level coordinates */
FBE_API int fbe_debug_candidates(fbe_extractor* e, int32_t slot, int32_t level, int32_t* xys, int32_t cap, int32_t* n);
/* blurred level image (rows x cols, tightly packed) */
FBE_API int fbe_debug_blurred(fbe_extractor* e, int32_t slot, int32_t level, uint8_t* dst, int32_t* rows, int32_t* cols);
/* stand-alone DistributeOctTree on caller candidates (relative coords as in src/ORBextractor.cc:539): returns
 * selected candidate indices in reference output order */
FBE_API int fbe_debug_octree(const int32_t* xys, int32_t n, int32_t min_x, int32_t max_x, int32_t min_y, int32_t max_y,
                     int32_t nfeat, int32_t* sel, int32_t cap, int32_t* n_sel);

/* ---- Frame::isInFrustum (next row f-2) ------------------------------------------------------------- */
/* The pose and calibration members Frame::isInFrustum reads (src/Frame.cc:435-491): mRcw (row-major 3x3), mtcw, mOw as
 * Frame::UpdatePoseMatrices left them (:425-433 stays on the host), fx fy cx cy, the image bounds, mbf, and the scale
 * pyramid members MapPoint::PredictScale(dist, Frame*) reads (src/MapPoint.cc:402-417). */
typedef struct fbe_frustum_view {
    float Rcw[9], tcw[3], Ow[3];
    float fx, fy, cx, cy;
    float min_x, max_x, min_y, max_y; /* mnMinX, mnMaxX, mnMinY, mnMaxY */
    float mbf;
    float log_scale_factor;           /* mfLogScaleFactor */
    int32_t n_levels;                 /* mnScaleLevels */
} fbe_frustum_view;
/* Frame::isInFrustum for n map points at once (Tracking::SearchLocalPoints calls it per local map point,
 * src/Tracking.cc; its outputs are the query set of SearchByProjection(Frame&, vector<MapPoint*>&, th)).
 * pos / normal: n x 3 floats (GetWorldPos, GetNormal); min_dist / max_dist: mfMinDistance / mfMaxDistance (the 0.8 / 1.2
 * invariance factors of src/MapPoint.cc:373-383 are applied inside).  Arithmetic follows OpenCV's for these cv::Mat
 * expressions (3x3 * 3x1 + 3x1 through the small-matrix gemm path: float products and sums left to right, the addend in
 * double; cv::norm and Mat::dot accumulate in double), pinned against cv2 4.13 in tests/test_frustum.py.
 * Outputs per point: in_view (mbTrackInView), proj = (mTrackProjX, mTrackProjY), proj_xr = mTrackProjXR, level =
 * mnTrackScaleLevel, view_cos = mTrackViewCos; entries of rejected points are left as 0.  Any output pointer may be NULL.
 * Parity bar: in_view / proj / view_cos bit-equal to the oracle; level equal except where logf(ratio) / mfLogScaleFactor
 * lies within 4 ulp of an integer (device logf vs glibc logf), which the oracle flags. */
FBE_API int fbe_is_in_frustum(const fbe_frustum_view* v, const float* pos, const float* normal, const float* min_dist,
                              const float* max_dist, int32_t n, float viewing_cos_limit, int32_t device, uint8_t* in_view,
                              float* proj, float* proj_xr, int32_t* level, float* view_cos);

/* ---- Initializer RANSAC scoring (next row f-2) --------------------------------------------------- */
/* Initializer::CheckHomography (src/Initializer.cc:391-474) for K hypotheses at once: the consumer of vnMatches12 in
 * Initializer::FindHomography (:211-260), which scores one H per RANSAC iteration over all matches.  The 8-point solves
 * (cv::SVD) stay on the host; this scores them.  kps1 / kps2: mvKeys1 / mvKeys2; matches: n pairs (first, second) =
 * mvMatches12; H21 / H12: K row-major 3x3 matrices.  Outputs: scores[K]; inliers[K x n] (may be NULL).
 * Bit-exact: every float expression is evaluated in the reference's order without FMA contraction (`1.0/(...)` in double
 * as written), and the score is accumulated sequentially over the matches like the reference's `score +=`. */
FBE_API int fbe_check_homography(const fbe_keypoint* kps1, const fbe_keypoint* kps2, const int32_t* matches, int32_t n,
                                 const float* H21, const float* H12, int32_t K, float sigma, int32_t device, float* scores,
                                 uint8_t* inliers);
/* Initializer::CheckFundamental (src/Initializer.cc:476-554), same batching; F21: K row-major 3x3 matrices. */
FBE_API int fbe_check_fundamental(const fbe_keypoint* kps1, const fbe_keypoint* kps2, const int32_t* matches, int32_t n,
                                  const float* F21, int32_t K, float sigma, int32_t device, float* scores, uint8_t* inliers);

/* ---- DBoW2 vocabulary descent (next row f-4) ------------------------------------------------------ */
/* The tree of DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h) on the
 * device.  Nodes 1 .. n_nodes in creation order, node 0 is the root: exactly the rows of the text format read by
 * loadFromTextFile (:1338-1436) -- parent id, nIsLeaf flag (> 0: the node becomes the next word id), 32 descriptor bytes,
 * weight.  A node is a leaf iff nothing names it as parent (Node::isLeaf(), :328). */
typedef struct fbe_vocabulary fbe_vocabulary;
FBE_API int fbe_vocabulary_create(int32_t k, int32_t L, const int32_t* parent, const uint8_t* is_word, const uint8_t* desc,
                                  const double* weight, int32_t n_nodes, int32_t device, fbe_vocabulary** out);
FBE_API int fbe_vocabulary_destroy(fbe_vocabulary* v);
/* TemplatedVocabulary::transform(feature, word_id, weight, &nid, levelsup) (:1218-1263) for n descriptors at once: the
 * per-feature part of transform(features, BowVector&, FeatureVector&, levelsup) (:1127-1205) that Frame::ComputeBoW /
 * KeyFrame::ComputeBoW call with levelsup = 4.  At every level the child with the smallest Hamming distance wins, the
 * FIRST one on ties (`d < best_d`).  Outputs per feature: word_id, node_id = the node passed at level L - levelsup (0 = root
 * when L - levelsup <= 0), weight = the word's weight.  The caller folds them into the BowVector (addWeight in feature
 * order, then normalize) and the FeatureVector (addFeature) exactly as :1150-1204 do; features with weight 0 are skipped. */
FBE_API int fbe_bow_transform(fbe_vocabulary* v, const uint8_t* desc, int32_t n, int32_t levelsup, int32_t* word_id,
                              int32_t* node_id, double* weight);

/* ---- Frame undistortion (next row f-1) ----------------------------------------------------------- */
/* Frame::UndistortKeyPoints, src/Frame.cc:638-669: cv::fisheye::undistortPoints(pts, pts, mK, mDistCoef, Mat(), mK) on the
 * keypoint positions (OpenCV 4.13 semantics: double-precision Newton solve, <= 10 iterations, eps 1e-8; (-1e6,-1e6) when
 * it fails).  K = {fx, fy, cx, cy}, D = mDistCoef {k1, k2, p1, p2} read as the four fisheye coefficients, exactly as the
 * reference passes them.  D[0] == 0 copies the keypoints (:640-644).  out may alias kps.  Tolerance, not bit-exactness,
 * is the parity bar for this row (tan / division ulp); see tests/test_gpu_undistort.py. */
FBE_API int fbe_undistort_keypoints(const fbe_keypoint* kps, int32_t n, const float K[4], const float D[4], int32_t device,
                                    fbe_keypoint* out);
/* Frame::ComputeImageBounds, src/Frame.cc:741-795: bounds = {mnMinX, mnMaxX, mnMinY, mnMaxY} from the four undistorted
 * image corners (or the image rectangle when D[0] == 0). */
FBE_API int fbe_image_bounds(int32_t cols, int32_t rows, const float K[4], const float D[4], int32_t device, float bounds[4]);

/* ---- Bird-view keypoint guidance + sub-pixel refinement (next row f-3, the part between detect and compute) ------------ */
/* Frame::GuidenceKeyBirdPts / Frame::nearEdges, src/Frame.cc:671-684, 717-739, followed by
 * cv::cornerSubPix(mBirdviewImg, pts, Size(half_w, half_h), Size(-1,-1), TermCriteria(EPS+MAX_ITER, max_iter, eps)),
 * src/Frame.cc:345-352 (the reference passes 5, 5, 40, 0.001), in one call with no host round trip in between.
 *   contour : mBirdviewContourICP (8-bit, rows x cols, contour_step bytes per row); NULL skips the filter (all kept).
 *             A keypoint is kept when any pixel of the window [x-10, x+10) x [y-10, y+10) (clipped; float -> size_t
 *             truncation below, `i < bound` compared in float above) is >= 10.  The reference indexes the window as
 *             at<uchar>(row = x range, col = y range) -- x and y swapped -- and so does this; an address beyond the image
 *             (possible only for non-square images, undefined in the reference) reads as 0.
 *   img     : mBirdviewImg (8-bit, same size, img_step bytes per row); NULL skips the refinement.  half_w, half_h in 1..10.
 *             cornerSubPix is OpenCV arithmetic, pinned to 4.13.0: float results equal cv2's bit for bit while the sampled
 *             window stays inside the image (always the case behind cv::ORB's 31-pixel edge threshold); near the border the
 *             replicated samples may differ from cv2 by one float ulp, the stated bar there is 1e-3 px.
 * Output: keep[n] (may be NULL) = nearEdges per input keypoint; out_kps[0 .. *n_out) = the kept keypoints in input order
 * (mvKeysBird) with pt refined and every other field untouched; iters[0 .. *n_out) (may be NULL) = gradient solves per
 * kept point.  out_kps needs room for n records. */
FBE_API int fbe_bird_refine(const uint8_t* contour, size_t contour_step, const uint8_t* img, size_t img_step, int32_t rows,
                            int32_t cols, const fbe_keypoint* kps, int32_t n, int32_t half_w, int32_t half_h, int32_t max_iter,
                            double eps, int32_t device, uint8_t* keep, fbe_keypoint* out_kps, int32_t* n_out, int32_t* iters);
/* The same two steps for nframes bird-view frames of one size in one call (offline batches, config C4): contours / imgs are
 * nframes images `*_stride` bytes apart with rows of `*_step` bytes (either may be NULL as above); kps, keep, out_kps and iters
 * are nframes lists `cap` records apart, n[f] <= cap candidates in frame f, n_out[f] kept.  Every frame gets exactly the result
 * of its own fbe_bird_refine call. */
FBE_API int fbe_bird_refine_batch(const uint8_t* contours, size_t contour_step, size_t contour_stride, const uint8_t* imgs,
                                  size_t img_step, size_t img_stride, int32_t rows, int32_t cols, int32_t nframes,
                                  const fbe_keypoint* kps, const int32_t* n, int32_t cap, int32_t half_w, int32_t half_h,
                                  int32_t max_iter, double eps, int32_t device, uint8_t* keep, fbe_keypoint* out_kps,
                                  int32_t* n_out, int32_t* iters);

/* ---- The bird-view feature path the reference ships (row f-3): cv::ORB::create(2000) detect + compute ------------------ */
/* src/Frame.cc:336-338  cv::Ptr<cv::ORB> extractorBird = cv::ORB::create(2000); extractorBird->detect(mBirdviewImg, preKeysBird, mBirdviewMask);
 * src/Frame.cc:355      extractorBird->compute(mBirdviewImg, mvKeysBird, mDescriptorsBird);
 * cv::ORB with its default parameters (scale 1.2, 8 levels, edge 31, Harris score, patch 31, FAST threshold 20), in the
 * arithmetic of OpenCV 4.13 (pinned by the oracle against cv2 4.13.0): same keypoints in the same ORDER, same float
 * responses / angles, same descriptors.  One handle = one image size and up to max_batch frames per call. */
typedef struct fbe_bird_orb fbe_bird_orb;
FBE_API int fbe_bird_orb_create(int32_t nfeatures, int32_t rows, int32_t cols, int32_t max_batch, int32_t device, fbe_bird_orb** out);
FBE_API int fbe_bird_orb_destroy(fbe_bird_orb* h);
/* capacity (records per frame) of the keypoint / descriptor arrays of the calls below */
FBE_API int fbe_bird_orb_max_keypoints(const fbe_bird_orb* h, int32_t* cap);
/* detect(): imgs / masks = nframes 8-bit images `stride` bytes apart with rows of `step` bytes (masks may be NULL; any non-zero
 * mask value keeps a keypoint, on every level, as in OpenCV 4.13).  kps [nframes][cap], n [nframes].
 * FBE_E_CAPACITY when response ties make a frame exceed cap.  std::nth_element's heap-select fallback (introselect out of its
 * depth budget: adversarial inputs only) is replayed too, serially. */
/* parity-test tap: KeyPointsFilter::retainBest (std::nth_element + std::partition, larger response first) replayed on the device
 * for an arbitrary response array -> order[n] = the permutation the algorithms leave behind (indices into `response`), of
 * which the first *n_kept survive; *heap_select_used (may be NULL) = 1 when introselect hit its depth limit and libstdc++'s
 * heap select was replayed. */
FBE_API int fbe_debug_retain_best(const float* response, int32_t n, int32_t n_points, int32_t device, int32_t* order, int32_t* n_kept,
                          int32_t* heap_select_used);
FBE_API int fbe_bird_orb_detect(fbe_bird_orb* h, const uint8_t* imgs, size_t step, size_t stride, const uint8_t* masks, size_t mask_step,
                                size_t mask_stride, int32_t nframes, fbe_keypoint* kps, int32_t* n);
/* compute(): kps [nframes][cap] / n [nframes] in and out (keypoints within 31 px of the image border are removed, an unsorted
 * list is regrouped by octave, as cv::ORB does), desc [nframes][cap][32]. */
FBE_API int fbe_bird_orb_compute(fbe_bird_orb* h, const uint8_t* imgs, size_t step, size_t stride, int32_t nframes, fbe_keypoint* kps,
                                 int32_t* n, uint8_t* desc);
/* The reference's whole bird block, src/Frame.cc:336-355, device-resident: detect -> GuidenceKeyBirdPts (contours != NULL:
 * Frame::nearEdges filter on mBirdviewContourICP, kept in order) -> cv::cornerSubPix(img, pts, Size(5,5), Size(-1,-1),
 * {EPS+MAX_ITER, 40, 0.001}) -> compute.  Out: mvKeysBird [nframes][cap], their number, mDescriptorsBird [nframes][cap][32];
 * n_detected (may be NULL) = |preKeysBird| per frame. */
FBE_API int fbe_bird_features(fbe_bird_orb* h, const uint8_t* imgs, size_t step, size_t stride, const uint8_t* masks, size_t mask_step,
                              size_t mask_stride, const uint8_t* contours, size_t contour_step, size_t contour_stride, int32_t nframes,
                              fbe_keypoint* kps, int32_t* n, uint8_t* desc, int32_t* n_detected);

/* Selects the CUDA device of the CALLING THREAD for the entry points that take no handle and no device argument
 * (fbe_grid_assign).  The host shims call it once per thread with FBE_DEVICE. */
FBE_API int fbe_set_device(int32_t device);

/* ---- Frame grid -------------------------------------------------------------------------------- */
/* Frame::AssignFeaturesToGrid / PosInGrid / PosInGridBirdview, src/Frame.cc:381-411,548-570.
 * cell = (round((x-min_x)*inv_w), round((y-min_y)*inv_h)), dropped when outside gcols x grows.
 * CSR out: cell_start[gcols*grows+1] indexed [ix*grows+iy] (mGrid[ix][iy]), cell_items[n] in index order. */
FBE_API int fbe_grid_assign(const fbe_keypoint* kps, int32_t n, float min_x, float min_y, float inv_w, float inv_h,
                    int32_t gcols, int32_t grows, int32_t* cell_start, int32_t* cell_items, int32_t* n_assigned);

/* ---- ORBmatcher -------------------------------------------------------------------------------- */
/* ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:1951-1967 (host inline; never a device round trip) */
FBE_API int fbe_hamming256(const uint8_t a[32], const uint8_t b[32]);

/* A frame as the matchers see it: keypoints (mvKeysUn / mvKeysBird), descriptors and the grid geometry.
 * The CSR grid is rebuilt on the device from the keypoints with the parameters below (same as fbe_grid_assign). */
typedef struct fbe_frame_view {
    const fbe_keypoint* kps;
    const uint8_t* desc; /* n x 32 */
    int32_t n;
    float min_x, min_y, inv_w, inv_h; /* mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv */
    int32_t gcols, grows;             /* 64x48 front (Frame.h:38-39), 32x32 bird (Frame.h:40) */
} fbe_frame_view;

typedef struct fbe_matcher fbe_matcher;
/* ORBmatcher::ORBmatcher(nnratio, checkOri), src/ORBmatcher.cc:42 */
FBE_API int fbe_matcher_create(float nn_ratio, int32_t check_orientation, int32_t device, fbe_matcher** out);
FBE_API int fbe_matcher_destroy(fbe_matcher* m);
/* Frames handed in as fbe_frame_view stay on the device (keypoints, descriptors, CSR grid) in a small per-matcher cache and
 * are recognised by content, so the searches Tracking runs back to back on one Frame upload and bucket it once.
 * Diagnostic counters of that cache (either pointer may be NULL). */
FBE_API int fbe_matcher_cache_stats(const fbe_matcher* m, uint64_t* hits, uint64_t* misses);

/* ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:406-521.
 * prev_matched: n1 (x,y) pairs, in/out (vbPrevMatched). matches12: n1 ints out (-1 = none). */
FBE_API int fbe_search_for_initialization(fbe_matcher* m, const fbe_frame_view* f1, const fbe_frame_view* f2,
                                  float* prev_matched, int32_t* matches12, int32_t window_size, int32_t* nmatches);

/* ORBmatcher::BirdviewMatch with isProject == 0, src/ORBmatcher.cc:1602-1760.
 * ref side: keypoints + descriptors of the reference bird frame; cur: current frame (32x32 bird grid).
 * Output: DMatch triples (queryIdx, trainIdx, distance) appended in index order, only trainIdx > 0 (quirk Q8). */
FBE_API int fbe_birdview_match(fbe_matcher* m, const fbe_keypoint* ref_kps, const uint8_t* ref_desc, int32_t n_ref,
                       const fbe_frame_view* cur, int32_t window_size, int32_t* dmatches /* n_ref x 3 */,
                       int32_t* n_dmatches, int32_t* nmatches);

/* ORBmatcher::BirdMapPointMatch, first pass, src/ORBmatcher.cc:1763-1863 (the window search + ratio test).
 * The reference's cv::Mat arithmetic stays on the host: the drop-in shim computes, exactly as :1797-1808 do,
 * local = Tbw*world, the |z| > 0.2 rejection, Converter::BaseXY2BirdPixel and the image-bounds test, and passes the
 * resulting bird pixel per map point (mp_pix: n x 2 floats; NaN x = skipped).  It also runs the second pass
 * (:1865-1895, distance filter + mvpMapPointsBird assignment) on the returned matches12.
 * Output: matches12[n_mp] = bird keypoint index or -1. */
FBE_API int fbe_bird_map_point_match(fbe_matcher* m, const float* mp_pix, const uint8_t* mp_desc, int32_t n_mp,
                                     const fbe_frame_view* cur, int32_t window_size, int32_t* matches12,
                                     int32_t* nmatches);

/* ORBmatcher::SearchByProjection(Frame&, const Frame& LastFrame, th, bMono=true), src/ORBmatcher.cc:1329-1471.
 * The projection (:1359-1376, cv::Mat fp32) stays in the host shim: last_proj is n_last x 2 (u,v), NaN u = skipped
 * (no map point, outlier, invzc < 0, outside the image bounds).  last_kps supplies octave and angle.
 * cur_taken[k] != 0 iff CurrentFrame.mvpMapPoints[k] has Observations()>0 on entry (NULL = none);
 * last_has_obs[i] != 0 iff that map point has Observations()>0, i.e. blocks its keypoint once assigned (NULL = all).
 * Output: cur_mp[k] = index i of the last-frame map point assigned to keypoint k, -1 untouched, -2 assigned by this call
 * and then removed by the orientation check (the reference writes NULL there, whatever the keypoint held before). */
FBE_API int fbe_search_by_projection_last(fbe_matcher* m, const fbe_frame_view* cur, const fbe_keypoint* last_kps,
                                          const float* last_proj, const uint8_t* last_mp_desc, int32_t n_last,
                                          const float* scale_factors, int32_t nlevels, const uint8_t* cur_taken,
                                          const uint8_t* last_has_obs, float th, int32_t* cur_mp, int32_t* nmatches);

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), src/ORBmatcher.cc:46-130.
 * Per map point (already filtered by mbTrackInView && !isBad): proj (n x 2), predicted level, view cosine,
 * descriptor -- the fields Frame::isInFrustum fills.  scale_factors: F.mvScaleFactors.  Monocular (mvuRight < 0). */
FBE_API int fbe_search_by_projection_map(fbe_matcher* m, const fbe_frame_view* cur, const float* scale_factors,
                                         int32_t nlevels, const float* mp_proj, const int32_t* mp_level,
                                         const float* mp_viewcos, const uint8_t* mp_desc, int32_t n_mp,
                                         const uint8_t* cur_taken, const uint8_t* mp_has_obs, float th,
                                         int32_t* cur_mp, int32_t* nmatches);

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist),
 * src/ORBmatcher.cc:1473-1600 (relocalisation).  The host shim keeps the reference's cv::Mat work: per key-frame map
 * point (not bad, not already found) the projection (u,v) (:1496-1509; NaN u = rejected by the bounds or the
 * min/max-distance test :1511-1521) and nPredictedLevel = MapPoint::PredictScale (src/MapPoint.cc:402-417).
 * kf_kps[i] supplies pKF->mvKeysUn[i].angle for the rotation histogram.  cur_taken[k] != 0 iff
 * CurrentFrame.mvpMapPoints[k] is non-NULL on entry.  Output: cur_mp[k] = index i of the assigned map point, -1 none. */
FBE_API int fbe_search_by_projection_reloc(fbe_matcher* m, const fbe_frame_view* cur, const fbe_keypoint* kf_kps,
                                           const float* mp_proj, const int32_t* mp_level, const uint8_t* mp_desc, int32_t n_mp,
                                           const float* scale_factors, int32_t nlevels, const uint8_t* cur_taken, float th,
                                           int32_t orb_dist, int32_t* cur_mp, int32_t* nmatches);

/* ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched,
 * int th), src/ORBmatcher.cc:291-404 (loop closing).  Host shim: Sim3 decomposition, projection, IsInImage, depth and
 * viewing-angle rejections (:300-350; NaN u = rejected) and PredictScale.  kf: the key frame's mvKeysUn / descriptors /
 * grid (KeyFrame::GetFeaturesInArea, src/KeyFrame.cc:901-940).  kf_matched[k] != 0 iff vpMatched[k] is non-NULL on entry.
 * Accepts bestDist <= TH_LOW, levels [pl-1, pl], no orientation check.  Output: kf_mp[k] = index into vpPoints or -1. */
FBE_API int fbe_search_by_projection_loop(fbe_matcher* m, const fbe_frame_view* kf, const float* mp_proj, const int32_t* mp_level,
                                          const uint8_t* mp_desc, int32_t n_mp, const float* scale_factors, int32_t nlevels,
                                          const uint8_t* kf_matched, int32_t th, int32_t* kf_mp, int32_t* nmatches);

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, matches), src/ORBmatcher.cc:160-289.
 * Feature vectors as CSR over ascending node ids: node_ids[nn], start[nn+1], items[...].
 * kf_has_mp[i] != 0 iff the key-frame keypoint has a good map point.  Output: f_mp[k] = key-frame keypoint index
 * matched to frame keypoint k (-1 none). */
FBE_API int fbe_search_by_bow(fbe_matcher* m, const fbe_keypoint* kf_kps, const uint8_t* kf_desc, int32_t n_kf,
                      const uint8_t* kf_has_mp, const int32_t* kf_node_ids, const int32_t* kf_start,
                      const int32_t* kf_items, int32_t kf_nn, const fbe_keypoint* f_kps, const uint8_t* f_desc,
                      int32_t n_f, const int32_t* f_node_ids, const int32_t* f_start, const int32_t* f_items,
                      int32_t f_nn, int32_t* f_mp, int32_t* nmatches);

/* ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12), src/ORBmatcher.cc:523-656 (loop
 * closing).  Both sides are key frames: kfX_has_mp[i] != 0 iff feature i has a good map point.  Differences to the
 * frame overload: `bestDist1 < TH_LOW` (strict), key-frame-2 features are blocked once matched (vbMatched2), the result
 * is indexed by key-frame-1 feature.  Output: matches12[i] = key-frame-2 feature index whose map point is assigned to
 * feature i of key frame 1, -1 none. */
FBE_API int fbe_search_by_bow_kf(fbe_matcher* m, const fbe_keypoint* kf1_kps, const uint8_t* kf1_desc, int32_t n1,
                                 const uint8_t* kf1_has_mp, const int32_t* kf1_node_ids, const int32_t* kf1_start,
                                 const int32_t* kf1_items, int32_t kf1_nn, const fbe_keypoint* kf2_kps, const uint8_t* kf2_desc,
                                 int32_t n2, const uint8_t* kf2_has_mp, const int32_t* kf2_node_ids, const int32_t* kf2_start,
                                 const int32_t* kf2_items, int32_t kf2_nn, int32_t* matches12, int32_t* nmatches);

/* ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:658-824 (LocalMapping::CreateNewMapPoints; next row f-4).
 * kfN_skip[i] != 0 iff feature i cannot take part: it already has a map point (GetMapPoint(i) != NULL), or bOnlyStereo is
 * set and mvuRight[i] < 0.  kfN_stereo[i] != 0 iff mvuRight[i] >= 0.  F12: row-major 3x3 fundamental matrix;
 * (ex, ey): the epipole in key frame 2, computed by the host shim exactly as :666-672 do (cv::Mat arithmetic);
 * kf2_scale_factors / kf2_level_sigma2: mvScaleFactors / mvLevelSigma2 of key frame 2.
 * Output: matches12[n1] = key-frame-2 feature or -1 (vMatchedPairs is the list of (i, matches12[i]) with matches12[i] >= 0
 * in index order); nmatches as returned by the reference.  Note that the reference never sets vbMatched2 in this method:
 * a key-frame-2 feature may be matched by several key-frame-1 features, and so it is here. */
FBE_API int fbe_search_for_triangulation(fbe_matcher* m, const fbe_keypoint* kf1_kps, const uint8_t* kf1_desc, int32_t n1,
                                         const uint8_t* kf1_skip, const uint8_t* kf1_stereo, const int32_t* kf1_node_ids,
                                         const int32_t* kf1_start, const int32_t* kf1_items, int32_t kf1_nn,
                                         const fbe_keypoint* kf2_kps, const uint8_t* kf2_desc, int32_t n2,
                                         const uint8_t* kf2_skip, const uint8_t* kf2_stereo, const int32_t* kf2_node_ids,
                                         const int32_t* kf2_start, const int32_t* kf2_items, int32_t kf2_nn, const float F12[9],
                                         float ex, float ey, const float* kf2_scale_factors, const float* kf2_level_sigma2,
                                         int32_t nlevels, int32_t* matches12, int32_t* nmatches);

/* The candidate search of both ORBmatcher::Fuse overloads, src/ORBmatcher.cc:826-976 (local mapping) and :978-1101 (loop
 * closing; next row f-4).  The projection, depth / viewing-angle rejections, MapPoint::PredictScale and the radius
 * (th * mvScaleFactors[level]) stay in the host shim like for the other projection searches: proj is n x 2 (u, v), NaN u =
 * skipped; proj_ur = u - bf*invz (only read when check_chi2 != 0 and the key frame has stereo features; may be NULL);
 * level = nPredictedLevel; radius per point.  kf: the key frame (mvKeysUn, mDescriptors, its 64x48 grid geometry);
 * kf_uright = mvuRight (NULL = monocular: all -1); inv_level_sigma2 = mvInvLevelSigma2.
 * check_chi2 != 0 applies the reprojection gate of the first overload (:911-937: 5.99 mono / 7.8 stereo), 0 = second
 * overload (no gate).  Output per point: best_idx = the most similar key-frame feature among the candidates of levels
 * [level-1, level] (first on ties, -1 = none), best_dist = its distance (INT_MAX = none).  The caller applies
 * `bestDist <= TH_LOW` and replays Replace / AddObservation / AddMapPoint / vpReplacePoint in map-point order, re-checking
 * isBad() / IsInKeyFrame at that moment (the search itself does not depend on map state). */
FBE_API int fbe_fuse_search(fbe_matcher* m, const fbe_frame_view* kf, const float* kf_uright, const float* inv_level_sigma2,
                            int32_t nlevels, const float* proj, const float* proj_ur, const int32_t* level, const float* radius,
                            const uint8_t* mp_desc, int32_t n, int32_t check_chi2, int32_t* best_idx, int32_t* best_dist);

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:242-307; MapPointBird.cc:90-155 is the same computation) for
 * npts map points at once (next row f-4).  The descriptors observed for point p are rows start[p] .. start[p+1]-1 of
 * `desc` (x 32 bytes), in the order the reference walks them (std::map<KeyFrame*, size_t> iteration order; the bird
 * variant may put its current mDescriptor first).  For each point: all pairwise Hamming distances, per row the median
 * `sorted_row[(int)(0.5*(N-1))]` (the row contains its own 0), and the FIRST row with the least median.
 * Output: best[p] = index inside the point's own list (the reference then clones that descriptor), -1 for an empty list;
 * best_median[p] (may be NULL) = that row's median. */
FBE_API int fbe_distinctive_descriptors(fbe_matcher* m, const uint8_t* desc, const int32_t* start, int32_t npts, int32_t* best,
                                        int32_t* best_median);

/* Brute-force Hamming top-2 (stress config C5): for each of nq queries best / second-best over nt targets,
 * ties -> lowest target index (traversal order). */
FBE_API int fbe_bruteforce_top2(fbe_matcher* m, const uint8_t* q_desc, int32_t nq, const uint8_t* t_desc, int32_t nt,
                        int32_t* best_idx, int32_t* best_dist, int32_t* second_dist);

/* ---- Device-resident batch pipeline (bench / offline batches; config C4) ------------------------- */
typedef struct fbe_pipeline fbe_pipeline;
typedef struct fbe_pipeline_cfg {
    fbe_extractor_cfg front; /* e.g. 2000 features */
    fbe_extractor_cfg bird;  /* e.g. 1000 features */
    int32_t front_rows, front_cols, bird_rows, bird_cols;
    int32_t batch; /* frame pairs per step */
    float nn_ratio;
    int32_t check_orientation;
    int32_t front_window; /* SearchForInitialization window (100) */
    int32_t bird_window;  /* BirdviewMatch window (10) */
    int32_t device;
    /* Front camera model.  front_fisheye != 0: Frame::UndistortKeyPoints (src/Frame.cc:638-669) runs on the device between
     * the descriptors and the grid, the grid spans Frame::ComputeImageBounds (:741-795) and the matching works on mvKeysUn,
     * exactly like the reference's Frame constructor orders it -- no host round trip.  K = {fx, fy, cx, cy}, D = mDistCoef
     * {k1, k2, p1, p2}.  0 = k1 == 0 behaviour: mvKeysUn = mvKeys, bounds 0 .. cols / rows. */
    int32_t front_fisheye;
    float front_K[4], front_D[4];
    /* Candidates kept per front query in the window search (0 = 256).  The reference has no such limit: a search window that
     * holds more keypoints than this makes THAT step fail with FBE_E_CAPACITY (reported by the fetch / wait of the step, with the
     * offending pair in fbe_last_error(); later steps are unaffected) -- re-create the pipeline with a larger value. */
    int32_t front_row_cap;
} fbe_pipeline_cfg;

typedef struct fbe_pair_result { /* per frame pair, fixed stride */
    int32_t n_front, n_bird, front_matches, bird_matches;
} fbe_pair_result;

FBE_API int fbe_pipeline_create(const fbe_pipeline_cfg* cfg, fbe_pipeline** out);
FBE_API int fbe_pipeline_destroy(fbe_pipeline* p);
/* One step = `batch` frame pairs: extract front + bird, build both grids, match pair i against pair i-1
 * (pair 0 against the last pair of the previous step; the very first pair of a run has no match).
 * d_front / d_bird: device pointers, [batch][rows][cols] u8 tightly packed.  Asynchronous on the pipeline
 * streams; results stay on the device until fetched. */
FBE_API int fbe_pipeline_step_dev(fbe_pipeline* p, const uint8_t* d_front, const uint8_t* d_bird);
/* Same through HOST (ideally pinned) buffers: H2D copies, step, D2H of the per-pair results + match lists. */
FBE_API int fbe_pipeline_step_host(fbe_pipeline* p, const uint8_t* h_front, const uint8_t* h_bird, fbe_pair_result* res,
                           int32_t* front_matches12 /* [batch][front_cap] or NULL */,
                           int32_t* bird_matches12 /* [batch][bird_cap] or NULL */);
/* Asynchronous variant: enqueue H2D of the inputs (own copy stream), the step and the D2H of the results into the
 * caller's buffers; returns a ticket.  fbe_pipeline_wait(ticket) blocks until that step's results are in the host
 * buffers.  At most three steps may be in flight (submit N+1, N+2 before waiting for N): the inputs of the next steps
 * are copied while step N computes.  Host buffers must stay valid (and should be pinned) until the wait returns. */
FBE_API int fbe_pipeline_submit_host(fbe_pipeline* p, const uint8_t* h_front, const uint8_t* h_bird, fbe_pair_result* res,
                                     int32_t* front_matches12, int32_t* bird_matches12, int32_t* ticket);
FBE_API int fbe_pipeline_wait(fbe_pipeline* p, int32_t ticket);
/* The same step returning what ORBextractor::operator() returns for every frame of the batch (src/ORBextractor.cc:1041-1104:
 * keypoints + descriptors) next to the match results: fixed-stride host arrays, entries past n_front / n_bird of a pair
 * (fbe_pair_result) are unspecified.  Any pointer may be NULL.  The copies run on their own stream beside the matching. */
typedef struct fbe_pipeline_features {
    fbe_keypoint* front_kps; /* [batch][front_cap] */
    uint8_t* front_desc;     /* [batch][front_cap][32] */
    fbe_keypoint* bird_kps;  /* [batch][bird_cap] */
    uint8_t* bird_desc;      /* [batch][bird_cap][32] */
} fbe_pipeline_features;
FBE_API int fbe_pipeline_submit_host_features(fbe_pipeline* p, const uint8_t* h_front, const uint8_t* h_bird, fbe_pair_result* res,
                                              int32_t* front_matches12, int32_t* bird_matches12, const fbe_pipeline_features* feat,
                                              int32_t* ticket);
/* Device addresses of the last step's match records, for a caller that moves them between GPUs itself (NCCL all-gather of
 * the shards' results, north_star): res = fbe_pair_result[batch], front_matches12 = int32[batch][front_cap],
 * bird_matches12 = int32[batch][bird_cap].  Valid until the pipeline is destroyed; contents change with every step
 * (order reads after fbe_pipeline_join() on the stream of fbe_pipeline_stream()). */
FBE_API int fbe_pipeline_device_results(fbe_pipeline* p, void** res, void** front_matches12, void** bird_matches12);
/* diagnostic: device milliseconds the H2D input copy of a (still current) ticket took */
FBE_API int fbe_pipeline_copy_ms(fbe_pipeline* p, int32_t ticket, float* ms);
FBE_API int fbe_pipeline_sync(fbe_pipeline* p);
FBE_API int fbe_pipeline_caps(const fbe_pipeline* p, int32_t* front_cap, int32_t* bird_cap);
/* copy back the results of the last step */
FBE_API int fbe_pipeline_fetch(fbe_pipeline* p, fbe_pair_result* res, int32_t* front_matches12, int32_t* bird_matches12);
/* full outputs of one pair of the last step (for parity tests at batch sizes) */
/* mvKeysUn of one front frame of the last step (equals the keypoints of fbe_pipeline_fetch_pair when front_fisheye == 0)
 * and the grid geometry in use: bounds = {mnMinX, mnMaxX, mnMinY, mnMaxY}.  Either pointer may be NULL. */
FBE_API int fbe_pipeline_fetch_front_undistorted(fbe_pipeline* p, int32_t pair, fbe_keypoint* front_kps_un, float bounds[4]);
FBE_API int fbe_pipeline_fetch_pair(fbe_pipeline* p, int32_t pair, fbe_keypoint* front_kps, uint8_t* front_desc,
                            fbe_keypoint* bird_kps, uint8_t* bird_desc);
/* elapsed GPU milliseconds between the first kernel and the last kernel of the last step (CUDA events) */
FBE_API int fbe_pipeline_last_step_ms(fbe_pipeline* p, float* ms);
/* raw stream handle (cudaStream_t) so that a caller can record its own events around steps; a step starts on this
 * stream but ends on the pipeline's matching stream: call fbe_pipeline_join() (stream-level wait, no host sync) before
 * recording the closing event */
FBE_API int fbe_pipeline_stream(fbe_pipeline* p, void** stream);
FBE_API int fbe_pipeline_join(fbe_pipeline* p);

/* live per-stage device timing (CUDA events on the launching streams) for bench.py's roofline block.
 * ms: 12 doubles summed over `steps` steps = front[pyramid, fast, octree, describe, grid, blur], bird[same]. */
FBE_API int fbe_pipeline_stage_timing(fbe_pipeline* p, int32_t enable);
FBE_API int fbe_pipeline_stage_ms(fbe_pipeline* p, double* ms, int32_t* steps);

/* pinned (page-locked) host memory for the host-buffer entry points */
FBE_API int fbe_host_alloc(void** ptr, size_t bytes);
FBE_API int fbe_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* FBE_CABI_H */

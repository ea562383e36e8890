// Device-side building blocks of the ORBmatcher replacements (declarations + launch wrappers in match.cu).
#pragma once
#include "fbe_internal.cuh"

namespace fbe {

// A batch of frames laid out with uniform strides (batch of 1 for the host C-ABI).
struct FrameDev {
    const fbe_keypoint* kps;   // [nb][kp_stride]
    const uint8_t* desc;       // [nb][kp_stride][32]
    const int* n;              // [nb]
    const int* start;          // [nb][gcols*grows+1]
    const int* items;          // [nb][kp_stride]
    int kp_stride;
    float min_x, min_y, inv_w, inv_h;
    int gcols, grows;
};

// Queries of a batch of problems.  q = (x, y, r, -): r < 0 marks "no search for this query".
struct QueryDev {
    const float4* q;           // [nb][stride]
    const int2* lv;            // [nb][stride]  (minLevel, maxLevel)
    const uint8_t* desc;       // [nb][stride][32]
    const int* nq;             // [nb]
    int stride;
};

// Reprojection gate of ORBmatcher::Fuse: target-side mvuRight (NULL = monocular) and mvInvLevelSigma2
struct ReprojGate { const float* t_uright = nullptr; float inv_sigma2[FBE_MAX_LEVELS] = {}; };

constexpr unsigned kNoKey = 0xFFFFFFFFu;     // "no candidate" in packed (dist << 20 | rank) keys
constexpr int kRowDistBits = 9;              // rows: (idx << 9) | dist, dist in [0,256]

enum ResolveMode { kResolveInit = 0, kResolveLast = 1, kResolveMap = 2, kResolveBow = 3 };

struct ResolveArgs {
    int mode;
    int C;                       // row capacity
    const unsigned* rows;        // [nb][q_stride][C]
    const int* cnt;              // [nb][q_stride]
    const int* nq;               // [nb]
    int q_stride;
    int t_stride;                // target-side stride (matched_dist, match21, taken, cur_mp)
    const int* nt;               // [nb] number of targets
    const fbe_keypoint* q_kps;   // source keypoints of the queries (angle / octave), [nb][q_stride]
    const int* q_src;            // BoW: query -> key-frame keypoint index ([nb][q_stride]); NULL = identity
    const fbe_keypoint* t_kps;   // target keypoints [nb][t_stride]
    const uint8_t* q_has_obs;    // LAST/MAP: assigned map point blocks its keypoint (NULL = all)
    float nn_ratio;
    int check_ori;
    int th_dist;                 // LAST: acceptance bound (0 = TH_HIGH; reloc passes ORBdist, loop closing TH_LOW); BOW: 0 = TH_LOW,
                                 //       TH_LOW - 1 for the key-frame/key-frame overload (`< TH_LOW`)
    // state / outputs
    int* matched_dist;           // INIT: [nb][t_stride] scratch (vMatchedDistance)
    int* match21;                // INIT: [nb][t_stride] scratch (vnMatches21)
    uint8_t* taken;              // LAST/MAP/BOW: [nb][t_stride] in/out scratch
    int* matches12;              // INIT: [nb][q_stride] out
    int* cur_mp;                 // LAST/MAP/BOW: [nb][t_stride] out
    float2* prev_matched;        // INIT: [nb][q_stride] in/out (NULL = do not update)
    int* q_bin;                  // [nb][q_stride] scratch
    int* q_hit;                  // [nb][q_stride] scratch (target index recorded with the bin)
    int* nmatches;               // [nb] out
};

struct BirdFinishArgs {
    const int* best_idx; const int* best_dist; const int* second_dist;   // [nb][q_stride]
    const int* nq; int q_stride;
    const fbe_keypoint* q_kps; const fbe_keypoint* t_kps; int t_stride;
    float nn_ratio; int check_ori;
    int* matches12;      // [nb][q_stride]
    int* dmatches;       // [nb][q_stride][3] or NULL
    int* n_dmatches;     // [nb] or NULL
    int* nmatches;       // [nb]
    int* q_bin;          // scratch
};

int launch_queries_from_kps(const fbe_keypoint* kps, const float2* pos, const uint8_t* desc_unused, const int* n, int stride,
                            int nb, float window, float4* q, int2* lv, cudaStream_t st);
int launch_window_rows(const FrameDev& f, const QueryDev& qs, int nb, int max_nq, bool upper_inclusive, int C, unsigned* rows,
                       int* cnt, int* overflow, cudaStream_t st);
int launch_window_top2(const FrameDev& f, const QueryDev& qs, int nb, int max_nq, bool upper_inclusive, int* best_idx,
                       int* best_dist, int* second_dist, cudaStream_t st);
int launch_window_top2_reproj(const FrameDev& f, const QueryDev& qs, int max_nq, const ReprojGate& rg, int* best_idx, int* best_dist,
                              int* second_dist, cudaStream_t st);
int launch_resolve(const ResolveArgs& a, int nb, cudaStream_t st);
int launch_bird_finish(const BirdFinishArgs& a, int nb, cudaStream_t st);
int launch_map_finish(const int* best_idx, const int* best_dist, const int* second_dist, int n, float nn_ratio, int th_dist,
                      int* matches12, int* nmatches, cudaStream_t st);
int launch_bow_rows(const uint8_t* kf_desc, const uint8_t* f_desc, const int* q_src, const int* q_beg, const int* q_end,
                    const int* f_items, int nq, int C, unsigned* rows, int* cnt, cudaStream_t st);
// SearchForTriangulation (src/ORBmatcher.cc:658-824): per-query arguments of the candidate kernel and its epilogue
struct TriArgs {
    const fbe_keypoint* kps1; const uint8_t* desc1; const uint8_t* stereo1;     // key frame 1
    const fbe_keypoint* kps2; const uint8_t* desc2; const uint8_t* stereo2; const uint8_t* skip2; const int* items2;   // key frame 2
    const int* q_src; const int* q_beg; const int* q_end; int nq;               // query = kf1 feature, candidates = items2[beg, end)
    float F[9]; float ex, ey; float scale[FBE_MAX_LEVELS]; float sigma2[FBE_MAX_LEVELS];
    int check_ori;
    int* q_best; int* q_bin;                                                     // per query: chosen kf2 feature / rotation bin
    int* matches12; int* nmatches;                                               // per kf1 feature (pre-set to -1) / total
};
int launch_triangulation(const TriArgs& a, cudaStream_t st);
int launch_distinctive(const uint8_t* desc, const int* start, int npts, int* best, int* best_median, cudaStream_t st);
int launch_bruteforce(const uint8_t* q, int nq, const uint8_t* t, int nt, unsigned* partial, int nchunks, int* best_idx,
                      int* best_dist, int* second_dist, cudaStream_t st);

}  // namespace fbe

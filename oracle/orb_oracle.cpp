// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of the reference's ORB extraction path, written from the algorithm (not a copy):
//   ORBextractor::ORBextractor          /root/reference/src/ORBextractor.cc:410-470
//   ORBextractor::ComputePyramid        :1107-1132
//   ComputeKeyPointsOctTree cell loop   :765-829   (+ fix-up :837-847)
//   DistributeOctTree / DivideNode      :539-763 / :481-537   (restated as a creation-sequence replay,
//                                        SURVEY Appendix A; ties -> most recently created node first)
//   IC_Angle / computeOrientation       :77-104 / :472-479
//   GaussianBlur + computeOrbDescriptor :1085-1090 / :108-147
//   operator() epilogue                 :1043-1105
// The OpenCV primitives are the cv2-pinned scalars in prim.hpp.  This file is pinned against the
// reference's own translation unit compiled verbatim (oracle/_ref/libfbe_ref.so, tests/test_oracle_vs_ref.py)
// and against committed fixtures (tests/golden/).  Build: -O2 -ffp-contract=off (no FMA contraction).
#include <cstdio>
#include "prim.hpp"

using namespace fbe_oracle;

namespace {

const int kEdge = 19;        // EDGE_THRESHOLD
const int kHalfPatch = 15;   // HALF_PATCH_SIZE
const int kPatch = 31;       // PATCH_SIZE

static const int8_t kPattern[256 * 4] = {
#include "orb_pattern.inc"
};

struct Key { float x, y, size, angle, response; int octave, class_id; };
struct Cand { int x, y, score; };   // x,y relative to (16,16) in the level image, as in the reference

struct Level {
    int w, h;                       // un-padded size
    int pw, ph;                     // padded size (w+38, h+38)
    std::vector<uint8_t> pad;       // padded image, row stride pw
    std::vector<uint8_t> blur;      // blurred un-padded image, row stride w (empty if level has no keypoint)
    std::vector<Cand> cand;         // vToDistributeKeys
    std::vector<Key> keys;          // after octree + fix-up + orientation (level coordinates)
    int n_cells_fallback, n_cells_empty;
    const uint8_t* roi() const { return pad.data() + (size_t)kEdge * pw + kEdge; }
};

struct Extractor {
    int nfeatures, nlevels, ini_th, min_th;
    double scale_factor;            // the member is a double initialised from a float (ORBextractor.h:98)
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> per_level;
    int umax[kHalfPatch + 1];
    std::vector<Level> lv;
    std::vector<Key> out_keys;
    std::vector<uint8_t> out_desc;
    std::vector<uint8_t> out_boundary;   // 1 if any BRIEF sample coordinate lies within 1e-3 of a .5 tie
};

void init_tables(Extractor& e, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
    e.nfeatures = nfeatures; e.nlevels = nlevels; e.ini_th = ini_th; e.min_th = min_th;
    e.scale_factor = scale_factor;
    e.scale.assign(nlevels, 1.f); e.sigma2.assign(nlevels, 1.f);
    for (int i = 1; i < nlevels; ++i) {
        // float * double -> rounded back to float, as `mvScaleFactor[i-1]*scaleFactor` does (:421)
        e.scale[i] = (float)(e.scale[i - 1] * e.scale_factor);
        e.sigma2[i] = e.scale[i] * e.scale[i];
    }
    e.inv_scale.resize(nlevels); e.inv_sigma2.resize(nlevels);
    for (int i = 0; i < nlevels; ++i) { e.inv_scale[i] = 1.0f / e.scale[i]; e.inv_sigma2[i] = 1.0f / e.sigma2[i]; }

    e.per_level.assign(nlevels, 0);
    float factor = (float)(1.0f / e.scale_factor);
    float want = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        e.per_level[l] = cv_round(want);
        sum += e.per_level[l];
        want *= factor;
    }
    e.per_level[nlevels - 1] = std::max(nfeatures - sum, 0);

    // half-widths of the radius-15 disc rows (:452-469)
    int vmax = cv_floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    int vmin = cv_ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= kHalfPatch; ++v) e.umax[v] = 0;
    for (int v = 0; v <= vmax; ++v) e.umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (e.umax[v0] == e.umax[v0 + 1]) ++v0;
        e.umax[v] = v0;
        ++v0;
    }
}

void build_pyramid(Extractor& e, const uint8_t* img, int rows, int cols, size_t step) {
    e.lv.assign(e.nlevels, Level());
    for (int l = 0; l < e.nlevels; ++l) {
        Level& L = e.lv[l];
        float s = e.inv_scale[l];
        L.w = cv_round((float)cols * s);
        L.h = cv_round((float)rows * s);
        L.pw = L.w + 2 * kEdge; L.ph = L.h + 2 * kEdge;
        L.pad.assign((size_t)L.pw * L.ph, 0);
        uint8_t* roi = L.pad.data() + (size_t)kEdge * L.pw + kEdge;
        if (l == 0) {
            border_reflect101_u8(img, cols, rows, step, L.pad.data(), L.pw, kEdge);
        } else {
            const Level& P = e.lv[l - 1];
            resize_linear_u8(P.roi(), P.w, P.h, P.pw, roi, L.w, L.h, L.pw);
            border_reflect101_u8(roi, L.w, L.h, L.pw, L.pad.data(), L.pw, kEdge);
        }
    }
}

// Per-cell FAST with threshold fallback (:765-829).  Coordinates in `cand` are relative to (16,16).
void detect_level(const Extractor& e, Level& L) {
    const int minBX = kEdge - 3, minBY = minBX;
    const int maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
    L.cand.clear(); L.n_cells_fallback = 0; L.n_cells_empty = 0;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const float W = 30;
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols <= 0 || nRows <= 0) return;   // reference would divide by zero; callers never get here
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<FastKp> cell;
    for (int i = 0; i < nRows; ++i) {
        const float iniY = (float)(minBY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = (float)maxBY;
        for (int j = 0; j < nCols; ++j) {
            const float iniX = (float)(minBX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBX - 6) continue;
            if (maxX > maxBX) maxX = (float)maxBX;
            const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
            const uint8_t* p = L.roi() + (ptrdiff_t)y0 * L.pw + x0;
            fast9_nms(p, cw, ch, L.pw, e.ini_th, cell);
            if (cell.empty()) {
                fast9_nms(p, cw, ch, L.pw, e.min_th, cell);
                if (cell.empty()) ++L.n_cells_empty; else ++L.n_cells_fallback;
            }
            for (size_t k = 0; k < cell.size(); ++k)
                L.cand.push_back({cell[k].x + j * wCell, cell[k].y + i * hCell, cell[k].score});
        }
    }
}

// DistributeOctTree as a creation-sequence replay.
// Invariant used: after the roots, the only list insertion is push_front, so the std::list is always
// "live nodes ordered by creation sequence, newest first" (roots: root 0 is front).  A sweep splits every
// splittable live node in list order; a refinement round splits them in (size, creation seq) descending
// order and stops once the live count reaches N.
struct Node {
    int ulx, uly, urx, bry;      // UL.x, UL.y, UR.x, BR.y   (BL.x=UL.x, BL.y=BR.y, ...)
    std::vector<int> keys;       // candidate indices in candidate order
    long seq;
    bool no_more, alive;
};

void split_node(const std::vector<Cand>& c, const Node& p, Node ch[4]) {
    const int halfX = (int)std::ceil((float)(p.urx - p.ulx) / 2);
    const int halfY = (int)std::ceil((float)(p.bry - p.uly) / 2);
    const int mx = p.ulx + halfX, my = p.uly + halfY;
    ch[0] = Node{p.ulx, p.uly, mx, my, {}, 0, false, true};
    ch[1] = Node{mx, p.uly, p.urx, my, {}, 0, false, true};
    ch[2] = Node{p.ulx, my, mx, p.bry, {}, 0, false, true};
    ch[3] = Node{mx, my, p.urx, p.bry, {}, 0, false, true};
    for (size_t i = 0; i < p.keys.size(); ++i) {
        const Cand& k = c[p.keys[i]];
        int q = ((float)k.x < (float)mx ? 0 : 1) + ((float)k.y < (float)my ? 0 : 2);
        ch[q].keys.push_back(p.keys[i]);
    }
    for (int q = 0; q < 4; ++q) if (ch[q].keys.size() == 1) ch[q].no_more = true;
}

std::vector<int> distribute_octree(const std::vector<Cand>& c, int minX, int maxX, int minY, int maxY, int N) {
    std::vector<int> result;
    const int W = maxX - minX, H = maxY - minY;
    const int nIni = (int)std::round((float)W / H);
    if (nIni <= 0) return result;   // reference indexes an empty vector here (UB); unsupported aspect ratio
    const float hX = (float)W / nIni;

    std::vector<Node> nodes;        // all nodes ever created; `alive` marks list membership
    nodes.reserve(4096);
    for (int i = 0; i < nIni; ++i) {
        Node n{(int)(hX * (float)i), 0, (int)(hX * (float)(i + 1)), H, {}, -(long)i, false, true};
        nodes.push_back(n);
    }
    for (size_t i = 0; i < c.size(); ++i) nodes[(size_t)((float)c[i].x / hX)].keys.push_back((int)i);
    for (int i = 0; i < nIni; ++i) {
        if (nodes[i].keys.size() == 1) nodes[i].no_more = true;
        else if (nodes[i].keys.empty()) nodes[i].alive = false;
    }
    long next_seq = 1;
    auto live_count = [&]() { int n = 0; for (auto& x : nodes) n += x.alive; return n; };
    // indices of live nodes in list order (front first) = creation sequence descending
    auto list_order = [&]() {
        std::vector<int> v;
        for (size_t i = 0; i < nodes.size(); ++i) if (nodes[i].alive) v.push_back((int)i);
        std::sort(v.begin(), v.end(), [&](int a, int b) { return nodes[a].seq > nodes[b].seq; });
        return v;
    };
    // splits node `idx`; children appended to `nodes`; multi-key children recorded in `expand`
    auto do_split = [&](int idx, std::vector<int>& expand) {
        Node ch[4];
        split_node(c, nodes[idx], ch);
        for (int q = 0; q < 4; ++q) {
            if (ch[q].keys.empty()) continue;
            ch[q].seq = next_seq++;
            nodes.push_back(ch[q]);
            if (ch[q].keys.size() > 1) expand.push_back((int)nodes.size() - 1);
        }
        nodes[idx].alive = false;
    };

    bool finish = false;
    std::vector<int> expand;
    while (!finish) {
        const int prev = live_count();
        std::vector<int> order = list_order();
        expand.clear();
        for (size_t k = 0; k < order.size(); ++k)
            if (!nodes[order[k]].no_more) do_split(order[k], expand);
        int sz = live_count();
        if (sz >= N || sz == prev) {
            finish = true;
        } else if (sz + (int)expand.size() * 3 > N) {
            while (!finish) {
                const int prev2 = live_count();
                std::vector<int> todo = expand;
                expand.clear();
                // ascending (size, creation seq); processed from the back
                std::sort(todo.begin(), todo.end(), [&](int a, int b) {
                    if (nodes[a].keys.size() != nodes[b].keys.size()) return nodes[a].keys.size() < nodes[b].keys.size();
                    return nodes[a].seq < nodes[b].seq;
                });
                int live = prev2;
                for (int j = (int)todo.size() - 1; j >= 0; --j) {
                    size_t before = nodes.size();
                    do_split(todo[j], expand);
                    live += (int)(nodes.size() - before) - 1;
                    if (live >= N) break;
                }
                if (live >= N || live == prev2) finish = true;
            }
        }
    }
    std::vector<int> order = list_order();
    for (size_t k = 0; k < order.size(); ++k) {
        const Node& n = nodes[order[k]];
        int best = n.keys[0];
        for (size_t i = 1; i < n.keys.size(); ++i)
            if ((float)c[n.keys[i]].score > (float)c[best].score) best = n.keys[i];
        result.push_back(best);
    }
    return result;
}

float ic_angle(const Extractor& e, const Level& L, int x, int y) {
    const uint8_t* c = L.pad.data() + (size_t)(y + kEdge) * L.pw + (x + kEdge);
    int m01 = 0, m10 = 0;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vs = 0, d = e.umax[v];
        for (int u = -d; u <= d; ++u) {
            int p = c[u + v * L.pw], m = c[u - v * L.pw];
            vs += p - m;
            m10 += u * (p + m);
        }
        m01 += v * vs;
    }
    return fast_atan2_deg((float)m01, (float)m10);
}

// Rotated BRIEF-256 (:108-147).  Returns 1 if some sample coordinate is so close to a rounding tie that a
// 1-ulp change of sinf/cosf (libm vs CUDA) could flip it; the parity tests tolerate a descriptor difference
// only on keypoints flagged here (north_star: "traced to an angle-rounding boundary").
int describe(const uint8_t* img, size_t step, const Key& k, uint8_t* desc) {
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float angle = k.angle * factorPI;
    const float a = cosf(angle), b = sinf(angle);
    const uint8_t* c = img + (size_t)cv_round(k.y) * step + cv_round(k.x);
    int boundary = 0;
    auto tie = [](float v) { float f = v - std::floor(v); return std::fabs(f - 0.5f) < 2e-5f; };
    auto sample = [&](int px, int py) -> int {
        float fy = px * b + py * a, fx = px * a - py * b;
        if (tie(fy) || tie(fx)) boundary = 1;
        return c[(ptrdiff_t)cv_round(fy) * (ptrdiff_t)step + cv_round(fx)];
    };
    for (int i = 0; i < 32; ++i) {
        int val = 0;
        for (int j = 0; j < 8; ++j) {
            const int8_t* p = kPattern + (size_t)(i * 8 + j) * 4;
            int t0 = sample(p[0], p[1]), t1 = sample(p[2], p[3]);
            val |= (t0 < t1) << j;
        }
        desc[i] = (uint8_t)val;
    }
    return boundary;
}

void run(Extractor& e, const uint8_t* img, int rows, int cols, size_t step) {
    e.out_keys.clear(); e.out_desc.clear(); e.out_boundary.clear();
    e.lv.clear();
    if (!img || rows <= 0 || cols <= 0) return;
    build_pyramid(e, img, rows, cols, step);
    for (int l = 0; l < e.nlevels; ++l) {
        Level& L = e.lv[l];
        detect_level(e, L);
        const int minBX = kEdge - 3, minBY = minBX, maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
        std::vector<int> sel;
        if (maxBX > minBX && maxBY > minBY) sel = distribute_octree(L.cand, minBX, maxBX, minBY, maxBY, e.per_level[l]);
        const int patch = (int)(kPatch * e.scale[l]);
        L.keys.clear();
        for (size_t i = 0; i < sel.size(); ++i) {
            const Cand& c = L.cand[sel[i]];
            Key k;
            k.x = (float)c.x + minBX; k.y = (float)c.y + minBY;
            k.size = (float)patch; k.angle = -1.f; k.response = (float)c.score; k.octave = l; k.class_id = -1;
            L.keys.push_back(k);
        }
    }
    for (int l = 0; l < e.nlevels; ++l) {
        Level& L = e.lv[l];
        for (size_t i = 0; i < L.keys.size(); ++i) L.keys[i].angle = ic_angle(e, L, (int)L.keys[i].x, (int)L.keys[i].y);
    }
    for (int l = 0; l < e.nlevels; ++l) {
        Level& L = e.lv[l];
        if (L.keys.empty()) continue;
        L.blur.resize((size_t)L.w * L.h);
        gauss7_u8(L.roi(), L.w, L.h, L.pw, L.blur.data(), L.w);
        for (size_t i = 0; i < L.keys.size(); ++i) {
            uint8_t d[32];
            int bd = describe(L.blur.data(), L.w, L.keys[i], d);
            e.out_desc.insert(e.out_desc.end(), d, d + 32);
            e.out_boundary.push_back((uint8_t)bd);
            Key k = L.keys[i];
            if (l != 0) { k.x *= e.scale[l]; k.y *= e.scale[l]; }
            e.out_keys.push_back(k);
        }
    }
}

}  // namespace

extern "C" {

void* orc_ext_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
    Extractor* e = new Extractor();
    init_tables(*e, nfeatures, scale_factor, nlevels, ini_th, min_th);
    return e;
}
void orc_ext_destroy(void* h) { delete (Extractor*)h; }

void orc_ext_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int32_t* per_level, int32_t* umax) {
    Extractor* e = (Extractor*)h;
    for (int i = 0; i < e->nlevels; ++i) {
        scale[i] = e->scale[i]; inv_scale[i] = e->inv_scale[i]; sigma2[i] = e->sigma2[i]; inv_sigma2[i] = e->inv_sigma2[i];
        per_level[i] = e->per_level[i];
    }
    for (int v = 0; v <= kHalfPatch; ++v) umax[v] = e->umax[v];
}

// Full operator(): returns keypoint count.
int orc_ext_run(void* h, const uint8_t* img, int rows, int cols, int step) {
    Extractor* e = (Extractor*)h;
    run(*e, img, rows, cols, (size_t)step);
    return (int)e->out_keys.size();
}
// kps: n x 28 B (cv::KeyPoint layout); desc: n x 32; boundary: n flags
void orc_ext_result(void* h, void* kps, uint8_t* desc, uint8_t* boundary) {
    Extractor* e = (Extractor*)h;
    size_t n = e->out_keys.size();
    if (n == 0) return;
    if (kps) std::memcpy(kps, e->out_keys.data(), n * sizeof(Key));
    if (desc) std::memcpy(desc, e->out_desc.data(), n * 32);
    if (boundary) std::memcpy(boundary, e->out_boundary.data(), n);
}
void orc_ext_level_size(void* h, int level, int32_t* w, int32_t* hh) {
    Extractor* e = (Extractor*)h;
    *w = e->lv[level].w; *hh = e->lv[level].h;
}
// padded level image, (h+38) x (w+38), tightly packed
void orc_ext_level_padded(void* h, int level, uint8_t* dst) {
    Extractor* e = (Extractor*)h;
    std::memcpy(dst, e->lv[level].pad.data(), e->lv[level].pad.size());
}
// blurred level (h x w) or zeros if the level had no keypoints; returns 1 if present
int orc_ext_level_blurred(void* h, int level, uint8_t* dst) {
    Extractor* e = (Extractor*)h;
    const Level& L = e->lv[level];
    if (L.blur.empty()) return 0;
    std::memcpy(dst, L.blur.data(), L.blur.size());
    return 1;
}
// candidates of a level in reference order; x,y are LEVEL coordinates (reference-relative + 16)
int orc_ext_candidates(void* h, int level, int32_t* xys, int cap) {
    Extractor* e = (Extractor*)h;
    const Level& L = e->lv[level];
    int n = (int)L.cand.size();
    for (int i = 0; i < n && i < cap; ++i) { xys[3 * i] = L.cand[i].x + 16; xys[3 * i + 1] = L.cand[i].y + 16; xys[3 * i + 2] = L.cand[i].score; }
    return n;
}
int orc_ext_level_nkeys(void* h, int level) { return (int)((Extractor*)h)->lv[level].keys.size(); }
void orc_ext_cell_stats(void* h, int level, int32_t* n_fallback, int32_t* n_empty) {
    Extractor* e = (Extractor*)h;
    *n_fallback = e->lv[level].n_cells_fallback; *n_empty = e->lv[level].n_cells_empty;
}

// Stand-alone octree entry for adversarial tests: cand (x,y,score) relative coords as in the reference.
int orc_octree(const int32_t* xys, int n, int minX, int maxX, int minY, int maxY, int N, int32_t* sel, int cap) {
    std::vector<Cand> c(n);
    for (int i = 0; i < n; ++i) c[i] = Cand{xys[3 * i], xys[3 * i + 1], xys[3 * i + 2]};
    std::vector<int> r = distribute_octree(c, minX, maxX, minY, maxY, N);
    for (size_t i = 0; i < r.size() && (int)i < cap; ++i) sel[i] = r[i];
    return (int)r.size();
}

}  // extern "C"

// Host-side geometry: scale tables, per-level sizes, FAST cell grid, octree capacities and the fixed-point
// bilinear tables.  Written from the reference's arithmetic (file:line in comments), evaluated in the same
// precision (fp32 unless stated) so that every derived integer matches.
#include <algorithm>
#include <cmath>
#include "fbe_internal.cuh"

namespace fbe {

static inline int round_half_even_f(float v) { return (int)nearbyintf(v); }   // cvRound(float)
static inline int round_half_even_d(double v) { return (int)nearbyint(v); }   // cvRound(double)

static inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// ORBextractor::ORBextractor, src/ORBextractor.cc:410-470
void compute_extractor_tables(int nfeatures, float scale_factor, int nlevels, std::vector<float>& scale,
                              std::vector<float>& inv_scale, std::vector<float>& sigma2, std::vector<float>& inv_sigma2,
                              std::vector<int>& per_level, int umax[16]) {
    const double sf = (double)scale_factor;   // member `double scaleFactor` initialised from the float argument
    scale.assign(nlevels, 1.0f);
    sigma2.assign(nlevels, 1.0f);
    for (int i = 1; i < nlevels; ++i) {
        scale[i] = (float)((double)scale[i - 1] * sf);
        sigma2[i] = scale[i] * scale[i];
    }
    inv_scale.resize(nlevels);
    inv_sigma2.resize(nlevels);
    for (int i = 0; i < nlevels; ++i) {
        inv_scale[i] = 1.0f / scale[i];
        inv_sigma2[i] = 1.0f / sigma2[i];
    }
    per_level.assign(nlevels, 0);
    const float factor = (float)(1.0 / sf);   // 1.0f / scaleFactor : float/double -> double -> float
    float desired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        per_level[l] = round_half_even_f(desired);
        sum += per_level[l];
        desired *= factor;
    }
    per_level[nlevels - 1] = std::max(nfeatures - sum, 0);

    // row half-widths of the radius-15 disc (:452-469)
    const int vmax = (int)floor(kHalfPatch * sqrtf(2.f) / 2 + 1);
    const int vmin = (int)ceil(kHalfPatch * sqrtf(2.f) / 2);
    for (int v = 0; v < 16; ++v) umax[v] = 0;
    for (int v = 0; v <= vmax; ++v) umax[v] = round_half_even_d(sqrt((double)(kHalfPatch * kHalfPatch) - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

// cv::resize INTER_LINEAR 8U coefficient for one destination coordinate (OpenCV >= 3 fixed-point path):
// 11-bit weights, source offset clamped into [0, src_len-1].
static ResizeTab linear_coeff(int d, int src_len, int dst_len) {
    const double sc = (double)src_len / dst_len;
    float f = (float)((d + 0.5) * sc - 0.5);
    int s = (int)floorf(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= src_len - 1) { s = src_len - 1; f = 0.f; }
    ResizeTab t;
    t.ofs = s;
    t.a0 = (short)std::min(std::max(round_half_even_f((1.f - f) * 2048.f), -32768), 32767);
    t.a1 = (short)std::min(std::max(round_half_even_f(f * 2048.f), -32768), 32767);
    return t;
}

int build_plan(const fbe_extractor_cfg& cfg, const std::vector<float>& scale, const std::vector<float>& inv_scale,
               const std::vector<int>& per_level, const int umax[16], int rows, int cols, Plan& p,
               std::vector<ResizeTab>& tabs) {
    p = Plan();
    p.nlevels = cfg.nlevels;
    p.rows = rows; p.cols = cols;
    p.ini_th = cfg.ini_th_fast; p.min_th = cfg.min_th_fast;
    for (int v = 0; v < 16; ++v) p.umax[v] = umax[v];
    if (rows > kMaxDim || cols > kMaxDim) { set_error("image larger than 4095 px per side"); return FBE_E_UNSUPPORTED; }
    tabs.clear();
    int img_off = 0, cell_base = 0, slot_base = 0, node_base = 0, kp_base = 0, grp_base = 0, blur_base = 0;
    p.max_cell_w = p.max_cell_h = 0;
    for (int l = 0; l < cfg.nlevels; ++l) {
        LevelGeom& g = p.lv[l];
        // ComputePyramid (:1111-1112): size always derives from the level-0 size
        g.w = round_half_even_f((float)cols * inv_scale[l]);
        g.h = round_half_even_f((float)rows * inv_scale[l]);
        if (g.w < 1 || g.h < 1) { set_error("pyramid level collapses to zero size"); return FBE_E_UNSUPPORTED; }
        g.pitch = (g.w + 2 * kEdge + 15) & ~15;
        g.ph = g.h + 2 * kEdge;
        g.img_off = img_off;
        img_off += (g.pitch * g.ph + 255) & ~255;
        g.scale = scale[l];
        g.patch_size = (float)(int)(kPatch * scale[l]);   // (:837) int truncation, stored as float
        g.nfeat = per_level[l];

        // FAST cell grid (:768-787), all in fp32 like the reference
        const int minB = kEdge - 3;
        const int maxBX = g.w - kEdge + 3, maxBY = g.h - kEdge + 3;
        const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
        const int ncols = (int)(width / 30.f), nrows = (int)(height / 30.f);
        if (ncols <= 0 || nrows <= 0) {
            // the reference divides by zero here (ceil(width/0)); such levels are not extractable
            set_error("pyramid level smaller than one 30-px FAST cell (reference divides by zero)");
            return FBE_E_UNSUPPORTED;
        }
        g.ncols = ncols; g.nrows = nrows;
        g.wcell = (int)ceilf(width / ncols);
        g.hcell = (int)ceilf(height / nrows);
        g.wcell_recip = 65536 / g.wcell + 1;
        g.cell_base = cell_base;
        g.cell_cap = ((g.wcell + 1) / 2) * ((g.hcell + 1) / 2);
        g.slot_base = slot_base;
        g.key_cap = ncols * nrows * g.cell_cap;
        cell_base += ncols * nrows;
        if (g.wcell > 64 || g.hcell > 64) { set_error("FAST cell larger than 64 px (cannot happen for 30-px cells)"); return FBE_E_UNSUPPORTED; }
        g.gcells = std::max(1, std::min(ncols, kFastGroupW / g.wcell));
        g.ngrp = (ncols + g.gcells - 1) / g.gcells;
        g.grp_base = grp_base;
        grp_base += g.ngrp * nrows;
        g.blur_ntx = (kEdge + g.w + kBlurTW - 1) / kBlurTW;
        g.blur_nty = (g.h + kBlurTH - 1) / kBlurTH;
        g.blur_base = blur_base;
        blur_base += g.blur_ntx * g.blur_nty;
        slot_base += g.key_cap;
        p.max_cell_w = std::max(p.max_cell_w, g.wcell);
        p.max_cell_h = std::max(p.max_cell_h, g.hcell);

        // DistributeOctTree roots (:543-545): W,H of the detection band [16, cols-16) x [16, rows-16)
        const int W = maxBX - minB, H = maxBY - minB;
        g.nini = (int)roundf((float)W / (float)H);
        if (g.nini <= 0) { set_error("aspect ratio gives zero octree roots (reference is undefined here)"); return FBE_E_UNSUPPORTED; }
        g.hx = (float)W / (float)g.nini;
        g.node_cap = std::max(g.nfeat, 4 * g.nini) + 8;
        g.node_base = node_base;
        node_base += g.node_cap;
        g.kp_cap = g.node_cap;
        g.kp_base = kp_base;
        kp_base += g.kp_cap;

        // bilinear tables over PADDED destination coordinates (frame pixels map through REFLECT_101)
        if (l > 0) {
            const LevelGeom& s = p.lv[l - 1];
            while (tabs.size() & 3) tabs.push_back(ResizeTab{0, 0, 0});     // the resize kernel reads 4 column entries as 2 x 16 bytes
            g.tabx_off = (int)tabs.size();
            for (int x = 0; x < g.pitch; ++x) {
                int xi = reflect101(std::min(x, g.w + 2 * kEdge - 1) - kEdge, g.w);
                ResizeTab t = linear_coeff(xi, s.w, g.w);
                t.ofs += kEdge;
                tabs.push_back(t);
            }
            g.taby_off = (int)tabs.size();
            for (int y = 0; y < g.ph; ++y) {
                int yi = reflect101(y - kEdge, g.h);
                ResizeTab t = linear_coeff(yi, s.h, g.h);
                t.ofs += kEdge;
                tabs.push_back(t);
            }
            // source tile of every 128x64 output tile (resize kernel stages it with TMA)
            int bw = 0, bh = 0;
            const int ntx = (g.pitch + kRsTW - 1) / kRsTW, nty = (g.ph + kRsTH - 1) / kRsTH;
            for (int tx = 0; tx < ntx; ++tx) {
                int lo = 1 << 30, hi = -1;
                for (int x = tx * kRsTW; x < std::min((tx + 1) * kRsTW, g.pitch); ++x) {
                    lo = std::min(lo, tabs[g.tabx_off + x].ofs); hi = std::max(hi, tabs[g.tabx_off + x].ofs);
                }
                lo &= ~15;
                p.rs_x0[l][tx] = (short)lo;
                bw = std::max(bw, hi + 2 - lo);
            }
            for (int ty = 0; ty < nty; ++ty) {
                int lo = 1 << 30, hi = -1;
                for (int y = ty * kRsTH; y < std::min((ty + 1) * kRsTH, g.ph); ++y) {
                    lo = std::min(lo, tabs[g.taby_off + y].ofs); hi = std::max(hi, tabs[g.taby_off + y].ofs);
                }
                p.rs_y0[l][ty] = (short)lo;
                bh = std::max(bh, hi + 2 - lo);
            }
            bw = (bw + 15) & ~15;
            // the TMA kernel reads the sources of four adjacent columns out of one realigned 8-byte span (pyramid.cu)
            bool narrow = true;
            for (int x4 = 0; x4 + 3 < (int)(((size_t)g.pitch + 3) & ~(size_t)3) && narrow; x4 += 4) {
                int lo = 1 << 30, hi = -1;
                for (int i = 0; i < 4; ++i) {
                    const size_t e = (size_t)g.tabx_off + x4 + i;
                    if (e >= tabs.size()) break;
                    lo = std::min(lo, tabs[e].ofs); hi = std::max(hi, tabs[e].ofs);
                }
                if (hi - lo > 6) narrow = false;
            }
            if (narrow && bw <= 256 && bh <= 256 && (size_t)bw * bh <= 64 * 1024) { g.rs_bw = bw; g.rs_bh = bh; }
            else g.rs_bw = g.rs_bh = 0;       // scale factor too large for one TMA box: direct-global kernel
        } else {
            g.tabx_off = g.taby_off = 0;
            g.rs_bw = g.rs_bh = 0;
        }
    }
    p.pyr_bytes = img_off;
    p.ncells_total = cell_base;
    p.ngroups_total = grp_base;
    p.blur_tiles_total = blur_base;
    p.slots_total = slot_base;
    p.nodes_total = node_base;
    p.kp_cap_total = kp_base;
    return FBE_OK;
}

}  // namespace fbe

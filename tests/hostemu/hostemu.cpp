// hostemu.cpp -- TEST INFRASTRUCTURE ONLY.  Single-lane host emulation of the product's kernel
// logic (aruco_slam_b200/csrc/core.h, frame_logic.h, pose_core.h compiled for the host) so that
// the CPU-only test tier (`pytest -m "not gpu"`) can check that logic against the oracle without
// a GPU.  It is not built into, shipped with or reachable from the product library; the CUDA
// kernels proper (thread mapping, shared memory, ballots) are checked by the `-m gpu` tests.
#include "../../aruco_slam_b200/csrc/core.h"
#include "../../aruco_slam_b200/csrc/frame_logic.h"
#include "../../aruco_slam_b200/csrc/pose_core.h"
#include "../../aruco_slam_b200/csrc/draw_core.h"
#include "../../aruco_slam_b200/csrc/refine_core.h"
#include "../../aruco_slam_b200/csrc/board_core.h"
#include "../../aruco_slam_b200/csrc/pyr_core.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace b2a;

struct HostCtx {
    int tid() const { return 0; }
    int nthreads() const { return 1; }
    void sync() const {}
    void mark() const {}
    int lane() const { return 0; }
    int lanes() const { return 1; }
    int warp() const { return 0; }
    int warps() const { return 1; }
    uint32_t ballot(bool p) const { return p ? 1u : 0u; }
    uint32_t warp_or(uint32_t v) const { return v; }
    void atomic_or(uint32_t *p, uint32_t v) const { *p |= v; }
    int exclusive_scan(int flag, int &total) const { total = flag ? 1 : 0; return 0; }
    int block_sum(int v) const { return v; }
};

// MaskView with a bounds check (a walk must never leave the padded plane)
struct CheckedView {
    const uint32_t *plane; int PWW, W, H; mutable bool bad = false; mutable long calls = 0;
    unsigned win9(int x, int y) const
    {
        ++calls;
        if (x < 0 || x >= W || y < 0 || y >= H) { bad = true; return 0; }
        return MaskView{plane, PWW}.win9(x, y);
    }
};

struct EmuParams {
    int nScales; int radius[8]; int Cfloor;
    double minPerimRate, maxPerimRate, approxRate, minCornerDistRate;
    int minDistanceToBorder; float minMarkerDistanceRate, minGroupDistance;
    int markerSize, borderBits, cellSize, cellMargin, nMarkers, maxCorr, maxBorderErr;
    double minOtsuStdDev;
    int max_cand, max_markers, surv_cap;
    int detectInverted;
};

// packed masks exactly as the device lays them out (the bits themselves come from a plain loop:
// k_threshold's tiling / ballots are device-only and are checked on the GPU)
static void emu_threshold(const uint8_t *g, int W, int H, int r, int C, uint32_t *plane, int PWW)
{
    const int k = 2 * r + 1, kk = k * k;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            long long S = 0;
            for (int dy = -r; dy <= r; ++dy) {
                int yy = std::min(std::max(y + dy, 0), H - 1);
                for (int dx = -r; dx <= r; ++dx) { int xx = std::min(std::max(x + dx, 0), W - 1); S += g[(size_t)yy * W + xx]; }
            }
            if (2 * S >= (long long)(2 * ((int)g[(size_t)y * W + x] + C) - 1) * kk)
                plane[(size_t)(y + 1) * PWW + (x >> 5) + 1] |= 1u << (x & 31);
        }
}

// masks_in: optional [nScales][H][W] u8 masks to use instead of thresholding (0 = compute)
// outputs: n_acc/n_rej, corners[max_markers*8], ids, rejected; debug: n_contours[nScales] (incl. one-point borders),
// n_cand, cand[max_cand*8]; kept contour lengths / points of scale dbg_scale
// pyr (ArUco3): gray is the segmentation image, candidates are identified in the pyramid level pyr_opt_level picks
static int emu_detect_impl(const uint8_t *gray, int W, int H, const uint8_t *masks_in, const unsigned long long *dict, const EmuParams *ep,
               int *n_acc, int *n_rej, float *corners, int32_t *ids, float *rejected,
               int *n_contours, int *n_cand, float *cand,
               int dbg_scale, int *dbg_nkept, int *dbg_len, int dbg_cap, int16_t *dbg_pts, int dbg_pts_cap, int anchor_R, const PyrLevels *pyr)
{
    const int nS = ep->nScales, WW = (W + 31) / 32, PWW = WW + 2, KS = W + 1;
    const size_t plane_words = (size_t)PWW * (H + 2);
    const int maxWH = std::max(W, H);
    const int minPerim = pyr ? pyr->minPerimeter : (int)(unsigned)(ep->minPerimRate * maxWH), maxPerim = (int)(unsigned)(ep->maxPerimRate * maxWH);
    std::vector<uint32_t> masks(plane_words * nS, 0);
    for (int s = 0; s < nS; ++s) {
        uint32_t *pl = masks.data() + plane_words * s;
        if (masks_in) {
            for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x)
                if (masks_in[((size_t)s * H + y) * W + x]) pl[(size_t)(y + 1) * PWW + (x >> 5) + 1] |= 1u << (x & 31);
        } else emu_threshold(gray, W, H, ep->radius[s], ep->Cfloor, pl, PWW);
    }
    static WalkTables wt;
    for (int i = 0; i < 4096; ++i) build_walk_table_entry(wt, i);
    const int surv_cap = ep->surv_cap;
    std::vector<int32_t> s_count(nS, 0), q_len((size_t)nS * surv_cap), q_xy((size_t)nS * surv_cap * 8);
    std::vector<uint8_t> q_ok((size_t)nS * surv_cap, 0);
    int status = 0;
    for (int s = 0; s < nS; ++s) {
        const uint32_t *pl = masks.data() + plane_words * s;
        MaskView mv{pl, PWW};
        struct Surv { uint32_t key; int len; uint32_t who; unsigned kind; };
        std::vector<Surv> surv;
        int ncont = 0;
        // --- anchors and start candidates (k_anchors, one word at a time) ---
        const bool device_like = anchor_R < 0;
        const int R = device_like ? -anchor_R : anchor_R;
        const int Rm = R - 1, Rm2 = 8 * R - 1;
        std::vector<uint32_t> ax, ay, as_, asup, cx, cy, cs;
        std::vector<Seg> seg;
        std::vector<uint32_t> amap((size_t)WW * H, A_NONE), minoff;
        for (int y = 0; y < H; ++y)
            for (int wx = 0; wx < WW; ++wx) {
                const uint32_t *row = pl + (size_t)(y + 1) * PWW + wx + 1;
                const uint32_t m = row[0];
                if (!m) continue;
                // word-parallel enumeration (k_anchors) ...
                uint32_t A[4], SU[4], U[4], hi[4], iso;
                anchor_words(m, row[-1], row[1], row[-PWW], row[-PWW - 1], row[-PWW + 1], row[PWW], row[PWW - 1], row[PWW + 1],
                             !(y & Rm), grid_cols(wx, Rm), !(y & Rm2), grid_cols(wx, Rm2), A, SU, U, hi, iso);
                ncont += __builtin_popcount(iso);
                if (A[0] | A[1] | A[2] | A[3]) amap[(size_t)y * WW + wx] = (uint32_t)ax.size();
                // ... cross-checked against the per-pixel window tables the walkers use
                for (int b = 0; b < 32; ++b) {
                    const int x = wx * 32 + b;
                    if (x >= W) break;
                    const unsigned w9 = mv.win9(x, y);
                    const uint32_t p = ((w9 >> 4) & 1u) ? wt.pix[w9] : 0u;
                    if ((bool)((iso >> b) & 1u) != (((w9 >> 4) & 1u) && MaskView::code_of_win9(w9) == 0)) return -108;
                    for (int k = 0; k < 4; ++k) {
                        const unsigned e = (p >> (8 * k)) & 0xFFu;
                        const unsigned fl = (e << 2) & (ST_ROW | ST_COL | ST_UNC);
                        const bool anchor = (e & 0x80u) && is_anchor(fl, x, y, Rm), sup = (e & 0x80u) && is_anchor(fl, x, y, Rm2);
                        const bool cand = (e & 0x80u) && (fl & ST_UNC);
                        if (anchor != (bool)((A[k] >> b) & 1u) || sup != (bool)((SU[k] >> b) & 1u) || cand != (bool)((U[k] >> b) & 1u)) return -109;
                        if ((anchor || cand) && (int)(e & 7u) != state_dir(k, hi[k], b)) return -110;
                        if (sup && !anchor) return -111;
                        if (anchor) { ax.push_back(x); ay.push_back(y); as_.push_back(e & 7u); asup.push_back(sup); }
                        if (cand) { cx.push_back(x); cy.push_back(y); cs.push_back(e & 7u); }
                    }
                }
            }
        const uint32_t nA = (uint32_t)ax.size();
        seg.assign(nA, Seg{A_NONE, A_NONE, 0u, A_NONE});
        minoff.assign(nA, 0);
        std::vector<uint32_t> codes((size_t)nA * SEG_CODE_WORDS + 1, 0);
        // anchor_R < 0: behave like the product call (walks give up after maxPerimeter steps; overflowed
        // segments and their borders are dropped) instead of the exact-count debug mode
        const int max_len = device_like ? maxPerim : 2 * W * H + 16;
        // --- segments (k_segments) ---
        for (uint32_t i = 0; i < nA; ++i) {
            int x = (int)ax[i], y = (int)ay[i], st = (int)as_[i];
            uint32_t len, mk, mo;
            CheckedView cv{pl, PWW, W, H};
            seg_walk(cv, wt.succ, KS, Rm, max_len, x, y, st, len, mk, mo, codes.data() + (size_t)i * SEG_CODE_WORDS);
            if (cv.bad) return -106;
            {   // the kernel walks with the cached window (CachedMaskView): same segment, same codes
                int x2 = (int)ax[i], y2 = (int)ay[i], st2 = (int)as_[i];
                uint32_t len2, mk2, mo2, codes2[SEG_CODE_WORDS] = {0};
                CachedMaskView cached(pl, PWW);
                seg_walk(cached, wt.succ, KS, Rm, max_len, x2, y2, st2, len2, mk2, mo2, codes2);
                if (len2 != len || mk2 != mk || mo2 != mo || x2 != x || y2 != y || st2 != st) return -116;
                const int nw = len < (uint32_t)(SEG_CODE_WORDS * 10) ? (int)((len + 9) / 10) : SEG_CODE_WORDS;
                for (int k = 0; k < nw; ++k) if (codes2[k] != codes[(size_t)i * SEG_CODE_WORDS + k]) return -117;
            }
            seg[i].len = len | (asup[i] ? SEG_SUPER : 0u); seg[i].minkey = mk; minoff[i] = mo;
            if (len == SEG_OVERFLOW) { if (device_like) continue; return -101; }
            const uint32_t base = amap[(size_t)y * WW + (x >> 5)];
            if (base == A_NONE) return -102;
            const int r = anchor_rank_in_word(mv, x, y, st, Rm);
            if (r < 0) return -112;
            const uint32_t j = base + (uint32_t)r;
            if (j >= nA || (int)ax[j] != x || (int)ay[j] != y || (int)as_[j] != st) return -103;
            seg[i].next = j;
            if (seg[j].prev != A_NONE) return -104;
            seg[j].prev = i;
        }
        // --- super anchors (k_skip) ---
        auto seg_at = [&](uint32_t i) { return seg[i]; };
        auto minoff_at = [&](uint32_t i) { return minoff[i]; };
        std::vector<Seg> sseg(nA, Seg{A_NONE, A_NONE, 0u, A_NONE});
        std::vector<uint32_t> ssoff(nA, 0);
        for (uint32_t i = 0; i < nA; ++i) {
            if (!asup[i]) continue;
            uint32_t snext, slen, smin, soff;
            super_skip(seg_at, minoff_at, i, max_len, snext, slen, smin, soff);
            sseg[i].next = snext; sseg[i].len = slen; sseg[i].minkey = smin; ssoff[i] = soff;
            if (snext != A_NONE) { if (sseg[snext].prev != A_NONE) return -107; sseg[snext].prev = i; }
        }
        auto sseg_at = [&](uint32_t i) { return sseg[i]; };
        // --- cycles (k_cycles): borders with anchors from their leaders, the others from their start candidates ---
        for (uint32_t i = 0; i < nA; ++i) {
            const bool sup = asup[i];
            const uint32_t len = sup ? cycle_leader(sseg_at, NeverStop{}, i, max_len) : cycle_leader(seg_at, StopAtSuper{}, i, max_len);
            if (!len) continue;
            ++ncont;
            if ((int)len >= minPerim && (int)len <= maxPerim) surv.push_back({sup ? sseg[i].minkey : seg[i].minkey, (int)len, i, sup ? 1u : 0u});
        }
        for (size_t c = 0; c < cx.size(); ++c) {
            const int x = (int)cx[c], y = (int)cy[c], s0 = (int)cs[c];
            const unsigned e0 = wt.succ[mv.win9(x, y) | ((unsigned)s0 << 9)];
            if (!(e0 & WT_ELIG)) return -113;
            const uint32_t key0 = key_of(x, y, e0, KS);
            CheckedView cv{pl, PWW, W, H};
            const int len = direct_walk(cv, cv, wt.succ, wt.pred, KS, Rm, x, y, s0, key0, max_len);
            if (cv.bad) return -114;
            {   // the kernel's dense walks use one cached window per walker
                const CachedMaskView wf(pl, PWW), wb(pl, PWW);
                if (direct_walk(wf, wb, wt.succ, wt.pred, KS, Rm, x, y, s0, key0, max_len) != len) return -118;
            }
            if (std::getenv("EMU_WALK_STATS")) { static long tot = 0, mx = 0, cnt = 0, big = 0; tot += cv.calls; cnt++; if (cv.calls > mx) mx = cv.calls; if (cv.calls > 40) big++;
                if (c + 1 == cx.size()) std::fprintf(stderr, "scale %d: %ld candidates, window reads total %ld max %ld, >40 reads: %ld\n", s, cnt, tot, mx, big); }
            if (len <= 0) continue;
            ++ncont;
            if (len >= minPerim && len <= maxPerim) surv.push_back({key0, len, (uint32_t)x | ((uint32_t)y << 16), 2u | ((unsigned)s0 << 8)});
        }
        if (n_contours) n_contours[s] = ncont;
        std::sort(surv.begin(), surv.end(), [](const Surv &a, const Surv &b) { return a.key > b.key; });
        if ((int)surv.size() > surv_cap) { surv.resize(surv_cap); status = 3; }
        s_count[s] = (int)surv.size();
        int w = 0;
        if (s == dbg_scale && dbg_nkept) *dbg_nkept = (int)surv.size();
        for (size_t i = 0; i < surv.size(); ++i) {
            const Surv &e = surv[i];
            // --- assign + emit (k_assign, k_assign_sub, k_emit) ---
            std::vector<uint32_t> pts(e.len, 0xFFFFFFFFu);
            std::vector<std::pair<uint32_t, int>> placed;
            auto place = [&](uint32_t a, int pos) { placed.push_back({a, pos}); };
            if ((e.kind & 3u) == 1u) {
                cycle_assign(sseg_at, place, e.who, e.len, (int)ssoff[e.who]);
                const size_t nsup = placed.size();
                for (size_t k = 0; k < nsup; ++k) super_assign(seg_at, place, placed[k].first, placed[k].second);
            } else if ((e.kind & 3u) == 0u) cycle_assign(seg_at, place, e.who, e.len, (int)minoff[e.who]);
            else seg_emit(mv, wt.succ, (int)(e.who & 0xFFFFu), (int)(e.who >> 16), (int)(e.kind >> 8), e.len, 0, e.len, pts.data());
            for (auto &pr : placed) {
                const int slen = (int)(seg[pr.first].len & SEG_LEN);
                if (slen <= SEG_CODE_WORDS * 10) seg_emit_codes(codes.data() + (size_t)pr.first * SEG_CODE_WORDS, (int)ax[pr.first], (int)ay[pr.first], slen, pr.second, e.len, pts.data());
                else seg_emit(mv, wt.succ, (int)ax[pr.first], (int)ay[pr.first], (int)as_[pr.first], slen, pr.second, e.len, pts.data());
            }
            for (int k = 0; k < e.len; ++k) if (pts[k] == 0xFFFFFFFFu) return -105;
            if (s == dbg_scale) {
                if (dbg_len && (int)i < dbg_cap) dbg_len[i] = e.len;
                if (dbg_pts) for (int k = 0; k < e.len && w < dbg_pts_cap; ++k, ++w) { dbg_pts[2 * w] = (int16_t)px_of(pts[k]); dbg_pts[2 * w + 1] = (int16_t)py_of(pts[k]); }
            }
            int ox[8], oy[8];
            SingleLane lg;
            const int m = approx_closed(lg, pts.data(), e.len, (double)e.len * ep->approxRate, ox, oy);
            const bool ok = (m == 4) && quad_passes(ox, oy, m, e.len, maxWH, ep->minCornerDistRate);
            const size_t slot = (size_t)s * surv_cap + i;
            q_ok[slot] = ok; q_len[slot] = e.len;
            if (ok) for (int k = 0; k < 4; ++k) { q_xy[slot * 8 + 2 * k] = ox[k]; q_xy[slot * 8 + 2 * k + 1] = oy[k]; }
        }
    }
    // per-frame logic
    const int MC = ep->max_cand;
    std::vector<float> cq((size_t)MC * 8), tq((size_t)MC * 8), tper(MC), cent((size_t)MC * 3), wq((size_t)MC * 8);
    std::vector<int32_t> clen(MC), gid(MC), sel(MC), gstart(MC + 1), gfill(MC), members(MC), closeIdx(MC), closeCnt(MC), S(MC), parent(MC), depth(MC),
        selGroup(MC), wres(MC), closeStart(MC), closeNum(MC), counters(8, 0), tlen(MC), wlen(MC);
    std::vector<uint32_t> closeM((size_t)MC * 2 * ((MC + 31) / 32));
    FrameScratch fs{cq.data(), clen.data(), tq.data(), tper.data(), cent.data(), gid.data(), sel.data(), gstart.data(), gfill.data(), members.data(),
                    closeIdx.data(), closeCnt.data(), S.data(), parent.data(), depth.data(), selGroup.data(), closeM.data(),
                    wq.data(), wres.data(), closeStart.data(), closeNum.data(), counters.data(), pyr ? tlen.data() : nullptr, pyr ? wlen.data() : nullptr};
    FrameParams fp;
    fp.W = W; fp.H = H; fp.nScales = nS; fp.surv_cap = surv_cap; fp.max_cand = MC; fp.max_markers = ep->max_markers;
    fp.markerSize = ep->markerSize; fp.borderBits = ep->borderBits; fp.minDistanceToBorder = ep->minDistanceToBorder;
    fp.minMarkerDistanceRate = ep->minMarkerDistanceRate; fp.minGroupDistance = ep->minGroupDistance;
    fp.detectInverted = ep->detectInverted;
    ScaleQuads sq{s_count.data(), q_ok.data(), q_xy.data(), q_len.data()};
    HostCtx ctx;
    frame_group(ctx, fp, sq, fs, nullptr, 0, 0);
    if (n_cand) *n_cand = counters[FC_NCAND];
    if (cand) std::memcpy(cand, cq.data(), (size_t)counters[FC_NCAND] * 8 * sizeof(float));
    // identification of every work item (k_identify's body, one lane)
    const int nb = ep->markerSize + 2 * ep->borderBits, Sz = nb * ep->cellSize, m0 = ep->cellSize / 2;
    for (int w = 0; w < counters[FC_NWORK]; ++w) {
        double M[9];
        const float *q = wq.data() + (size_t)w * 8;
        float sq8[8];
        const uint8_t *im = gray; int iW = W, iH = H; size_t ipitch = (size_t)W;
        if (pyr) {               // k_homography / k_identify<true>
            const int lvl = pyr_opt_level(pyr->W, pyr->n, pyr->segW, wlen[w], pyr->minPerimeter);
            pyr_scale_quad(q, f_div((float)pyr->W[lvl], (float)pyr->segW), sq8);
            q = sq8;
            im = pyr->base[lvl]; iW = pyr->W[lvl]; iH = pyr->H[lvl]; ipitch = pyr->pitch[lvl];
        }
        perspective_inverse(q, Sz, M);
        std::vector<uint8_t> patch((size_t)Sz * Sz);
        int hist[256] = {0};
        long long sum = 0, sqs = 0;
        for (int p = 0; p < Sz * Sz; ++p) {
            const int y = p / Sz, x = p - y * Sz;
            const unsigned v = warp_sample(im, iW, iH, ipitch, M, x, y);
            patch[p] = (uint8_t)v; hist[v]++;
            if (x >= m0 && x < Sz - m0 && y >= m0 && y < Sz - m0) { sum += v; sqs += v * v; }
        }
        int mode, thr;
        ident_decide(sum, sqs, Sz, m0, ep->minOtsuStdDev, hist, mode, thr);
        std::vector<uint8_t> bits((size_t)nb * nb);
        for (int c = 0; c < nb * nb; ++c) bits[c] = (uint8_t)((mode < 2) ? mode : ident_cell_bit(patch.data(), Sz, ep->cellSize, ep->cellMargin, c / nb, c % nb, thr));
        unsigned long long code;
        int res = 0;
        if (ident_border_code(bits.data(), ep->markerSize, ep->borderBits, ep->maxBorderErr, code, ep->detectInverted != 0))
            for (int m = 0; m < ep->nMarkers; ++m) {
                int rot;
                if (ident_marker_distance(dict + (size_t)m * 4, code, ep->markerSize, rot) <= ep->maxCorr) { res = (int)(0x80000000u | ((unsigned)m << 8) | (unsigned)rot); break; }
            }
        wres[w] = res;
    }
    int st = status;
    FrameOutputs fo{n_acc, n_rej, corners, ids, rejected, &st};
    frame_finalize(ctx, fp, fs, fo);
    return st;
}

extern "C" {

int emu_detect(const uint8_t *gray, int W, int H, const uint8_t *masks_in, const unsigned long long *dict, const EmuParams *ep,
               int *n_acc, int *n_rej, float *corners, int32_t *ids, float *rejected,
               int *n_contours, int *n_cand, float *cand,
               int dbg_scale, int *dbg_nkept, int *dbg_len, int dbg_cap, int16_t *dbg_pts, int dbg_pts_cap, int anchor_R)
{
    return emu_detect_impl(gray, W, H, masks_in, dict, ep, n_acc, n_rej, corners, ids, rejected, n_contours, n_cand, cand, dbg_scale, dbg_nkept, dbg_len, dbg_cap,
                           dbg_pts, dbg_pts_cap, anchor_R, nullptr);
}

// the interior ranges k_pyr_down iterates over against the per-group test
int emu_pyr_ranges_ok(int W, int H)
{
    const int dW = (W + 1) / 2, dH = (H + 1) / 2, gW = (dW + 3) / 4, gR = pyr_interior_gx_end(W, dW), yB = pyr_interior_y_end(H);
    if (gR < 1 || gR > gW || yB < 1) return 0;
    for (int y = 0; y < dH; ++y)
        for (int gx = 0; gx < gW; ++gx) {
            const bool by_range = gx >= 1 && gx < gR && y >= 1 && y < yB;
            if (by_range != pyr_down_is_interior4(W, H, dW, 4 * gx, y)) return 0;
            if (by_range && !(2 * (4 * gx) - 4 >= 0)) return 0;          // the word path's left bound holds for every interior group
        }
    return 1;
}

// ArUco3 (pyr_core.h): one pyrDown / one resize, a pixel at a time as the kernels compute them
void emu_pyr_down(const uint8_t *src, int W, int H, uint8_t *dst)
{
    const int dW = (W + 1) / 2, dH = (H + 1) / 2;
    for (int y = 0; y < dH; ++y)
        for (int x0 = 0; x0 < dW; x0 += 4) {                     // k_pyr_down: four pixels at a time away from the border
            if (pyr_down_is_interior4(W, H, dW, x0, y)) {
                // the word form needs word-aligned rows: the emulation takes it when the image width allows, as the kernel does for its pitched levels
                const bool words = (W & 3) == 0 && (((size_t)src) & 3) == 0 && 2 * x0 - 4 >= 0 && 2 * x0 + 11 < W;
                const uint32_t w = words ? pyr_down_interior4<true>(src, (size_t)W, x0, y) : pyr_down_interior4<false>(src, (size_t)W, x0, y);
                for (int q = 0; q < 4; ++q) dst[(size_t)y * dW + x0 + q] = (uint8_t)(w >> (8 * q));
            } else if (W >= 3 && H >= 3) {
                const uint32_t w = pyr_down_border4(src, W, H, (size_t)W, dW, x0, y);
                for (int q = 0; q < 4 && x0 + q < dW; ++q) dst[(size_t)y * dW + x0 + q] = (uint8_t)(w >> (8 * q));
            } else for (int x = x0; x < x0 + 4 && x < dW; ++x) dst[(size_t)y * dW + x] = pyr_down_pixel(src, W, H, (size_t)W, x, y);
        }
}
void emu_resize_linear(const uint8_t *src, int W, int H, uint8_t *dst, int dW, int dH)
{
    const bool area = (W == 2 * dW && H == 2 * dH);
    std::vector<int> tab((size_t)3 * (dW + dH));                // k_resize_tabs, then k_resize_linear
    for (int i = 0; i < dW + dH; ++i) { if (i < dW) resize_tab(i, dW, W, tab[3 * i], tab[3 * i + 1], tab[3 * i + 2]); else resize_tab(i - dW, dH, H, tab[3 * i], tab[3 * i + 1], tab[3 * i + 2]); }
    for (int y = 0; y < dH; ++y)
        for (int x = 0; x < dW; ++x) {
            const int *tx = &tab[3 * x], *ty = &tab[3 * (dW + y)];
            dst[(size_t)y * dW + x] = area ? resize_pixel(src, W, H, (size_t)W, dW, dH, x, y) : resize_pixel_tab(src, W, H, (size_t)W, tx[0], tx[1], tx[2], ty[0], ty[1], ty[2]);
        }
}

// the ArUco3 front half of a call as run_front / run_back arrange it: plan, pyramid, segmentation image, detection in it with the
// identification in the pyramid; the accepted corners come back UNREFINED in segmentation-image coordinates (the refinement chain
// is k_subpix, device only) together with the plan (seg size, closest level, level-0-to-seg scale) for the caller
int emu_detect_aruco3(const uint8_t *gray, int W, int H, const unsigned long long *dict, const EmuParams *ep, int minSide, float ratio,
                      int *n_acc, int *n_rej, float *corners, int32_t *ids, float *rejected, int *plan_out /* segW, segH, numLevels, closestIdx */)
{
    Aruco3Plan plan;
    if (!aruco3_plan(W, H, minSide, ratio, plan)) return -1;
    std::vector<std::vector<uint8_t>> lv((size_t)plan.numLevels + 1);
    PyrLevels pl;
    pl.n = plan.numLevels + 1; pl.segW = plan.segW; pl.minPerimeter = 4 * minSide;
    pl.W[0] = W; pl.H[0] = H; pl.base[0] = gray; pl.pitch[0] = (size_t)W; pl.frame_stride[0] = 0;
    for (int l = 1; l <= plan.numLevels; ++l) {
        lv[l].resize((size_t)plan.W[l] * plan.H[l]);
        emu_pyr_down(pl.base[l - 1], pl.W[l - 1], pl.H[l - 1], lv[l].data());
        pl.W[l] = plan.W[l]; pl.H[l] = plan.H[l]; pl.base[l] = lv[l].data(); pl.pitch[l] = (size_t)plan.W[l]; pl.frame_stride[l] = 0;
    }
    std::vector<uint8_t> seg;
    const uint8_t *sg = gray;
    if (plan.fxfy != 1.f) { seg.resize((size_t)plan.segW * plan.segH); emu_resize_linear(gray, W, H, seg.data(), plan.segW, plan.segH); sg = seg.data(); }
    plan_out[0] = plan.segW; plan_out[1] = plan.segH; plan_out[2] = plan.numLevels; plan_out[3] = plan.closestIdx;
    std::vector<int> ncont(ep->nScales);
    std::vector<float> cand((size_t)ep->max_cand * 8);
    int n_cand = 0, nk = 0;
    return emu_detect_impl(sg, plan.segW, plan.segH, nullptr, dict, ep, n_acc, n_rej, corners, ids, rejected, ncont.data(), &n_cand, cand.data(), -1, &nk, nullptr, 0,
                           nullptr, 0, 32, &pl);
}

// approxPolyDP(closed) of the product (one lane) on an integer contour; returns the vertex count (-1: gave up, > 8 vertices)
int emu_approx(const int32_t *xy, int n, double eps, int32_t *out_xy)
{
    std::vector<uint32_t> P(n);
    for (int i = 0; i < n; ++i) P[i] = (uint32_t)xy[2 * i] | ((uint32_t)xy[2 * i + 1] << 16);
    int ox[8], oy[8];
    SingleLane lg;
    const int m = approx_closed(lg, P.data(), n, eps, ox, oy);
    for (int k = 0; k < m && k < 8; ++k) { out_xy[2 * k] = ox[k]; out_xy[2 * k + 1] = oy[k]; }
    return m;
}

// restated Otsu vs the textbook loop on one histogram
void emu_otsu(const int *h, int n, int *thr_new, int *thr_seq)
{
    *thr_new = otsu_threshold(h, n);
    *thr_seq = otsu_threshold_sequential(h, n);
}

void emu_pose(const float *corners, int n, const double *K9, const double *D5, float marker_length, double *rvecs, double *tvecs)
{
    Camera cam{K9[0], K9[4], K9[2], K9[5], D5[0], D5[1], D5[2], D5[3], D5[4]};
    double sh[POSE_SH];
    for (int i = 0; i < n; ++i) solve_marker_pose(OneLane{}, cam, marker_length, corners + 8 * i, sh, rvecs + 3 * i, tvecs + 3 * i);
}

// returns kept flag; obs = (x, y, theta, cov[9])
int emu_observation(const float *corners, int id, const double *rvec, const double *tvec, const double *K9, const double *D5,
                    double R_x, double R_y, double R_theta, double marker_length, double tx, double ty, float thr, double *obs12)
{
    Camera cam{K9[0], K9[4], K9[2], K9[5], D5[0], D5[1], D5[2], D5[3], D5[4]};
    ObsParams op{R_x, R_y, R_theta, marker_length, tx, ty, thr};
    Observation o;
    if (!make_observation(cam, op, corners, id, rvec, tvec, o)) return 0;
    obs12[0] = o.x; obs12[1] = o.y; obs12[2] = o.theta;
    for (int i = 0; i < 9; ++i) obs12[3 + i] = o.cov[i];
    return 1;
}

// drawDetectedMarkers through the product's draw_core.h, one thread
void emu_draw(uint8_t *img, int W, int H, int channels, const float *corners, const int32_t *ids, int n, const uint8_t *border)
{
    static const OverlayTables t = make_overlay_tables();
    OverlayImage im{img, W, H, channels, (size_t)W * channels};
    overlay_draw_sequential(im, t, corners, ids, n, border);
}

// CORNER_REFINE_CONTOUR of one marker through the product's refine_core.h, one lane.  contour: n x (x, y)
struct OneRefineLane {
    int lane() const { return 0; }
    int nlanes() const { return 1; }
    uint32_t ballot(bool p) const { return p ? 1u : 0u; }
    int bcast(int v, int) const { return v; }
    long long sum(long long v) const { return v; }
    int imin(int v) const { return v; }
    int imax(int v) const { return v; }
};
int emu_refine_lines(const int32_t *contour, int n, const float *cin, float *cout)
{
    std::vector<uint32_t> P((size_t)n);
    for (int i = 0; i < n; ++i) P[i] = (uint32_t)contour[2 * i] | ((uint32_t)contour[2 * i + 1] << 16);
    return refine_marker_lines(OneRefineLane{}, P.data(), n, cin, cout) ? 1 : 0;
}

// board_core.h: the least-squares homography and the planar board pose
int emu_homography(const double *src, const double *dst, int n, double *H, int refine_iters) { return homography_ls(src, dst, n, H, refine_iters) ? 1 : 0; }
int emu_board_pose(const double *K9, const double *D5, const double *obj, const double *img, int n, double *rvec, double *tvec)
{
    Camera cam{K9[0], K9[4], K9[2], K9[5], D5[0], D5[1], D5[2], D5[3], D5[4]};
    return board_pose(cam, obj, img, n, rvec, tvec);
}

}  // extern "C"

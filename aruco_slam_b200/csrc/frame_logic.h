// frame_logic.h -- per-frame (one CTA per frame) stages of the detect path:
//   frame_group    : candidate gathering, A4 reorder, A5 filterTooCloseCandidates + hierarchy,
//                    identification work list                       
//   frame_finalize : A6 depth-ordered acceptance, corner rotation, accepted / rejected output
// Replaces the corresponding parts of cv::aruco::detectMarkers (reference
// src/aruco_slam.cpp:313).  Written against a small "Ctx" (tid / nthreads / sync / scan) so
// that the identical source runs as a CUDA block on the device and as one sequential lane in
// tests/hostemu (CPU-only logic check against the CPU restatement; test infrastructure only).
#pragma once
#include "core.h"

namespace b2a {

struct FrameParams {
    int W, H;
    int nScales;
    int surv_cap;            // capacity of the per-(frame,scale) survivor arrays
    int max_cand;            // capacity of candidates per frame
    int max_markers;         // capacity of accepted / rejected per frame
    int markerSize, borderBits;
    int minDistanceToBorder;
    float minMarkerDistanceRate;   // (float)params.minMarkerDistanceRate
    float minGroupDistance;
    int detectInverted;      // detectInvertedMarker: a group is walked from its smallest contour (cv2 sorts the group descending)
};

// per-(frame,scale) outputs of the contour / polygon stages, in cv2.findContours list order
struct ScaleQuads {
    const int32_t *count;        // [nScales]      survivors of this frame
    const uint8_t *quad_ok;      // [nScales][surv_cap]
    const int32_t *quad_xy;      // [nScales][surv_cap][8]  x0,y0,...,x3,y3
    const int32_t *len;          // [nScales][surv_cap]      contour length
};

// per-frame scratch in global memory (all indexed within the frame)
struct FrameScratch {
    float   *cq;        // [max_cand][8] candidate quads in candidate order (after A4)
    int32_t *clen;      // [max_cand]
    float   *tq;        // [max_cand][8] quads in T order (sorted by perimeter, descending, stable)
    float   *tper;      // [max_cand]
    float   *cent;      // [max_cand][3]  T order: 4 * centroid x, 4 * centroid y, perimeter * minMarkerDistanceRate
    int32_t *gid;       // [max_cand]
    int32_t *sel;       // [max_cand]
    int32_t *gstart;    // [max_cand + 1]
    int32_t *gfill;     // [max_cand]
    int32_t *members;   // [max_cand]
    int32_t *closeIdx;  // [max_cand]  per group segment: T indices of "close" candidates
    int32_t *closeCnt;  // [max_cand]  per group
    int32_t *S;         // [max_cand]  selected -> T index
    int32_t *parent;    // [max_cand]
    int32_t *depth;     // [max_cand]
    int32_t *selGroup;  // [max_cand]  selected -> group id or -1
    uint32_t *closeM;   // 2 x [max_cand][max_cand/32] closeness bit matrix (row i: bits j > i) and its transpose, global fallback
    // identification work list: item w < n_sel is selected candidate w; then the close candidates
    float   *wq;        // [max_cand][8]
    int32_t *wres;      // [max_cand]  result: bit 31 valid, bits 8..: id, bits 0..1: rot   (written by identify)
    int32_t *closeStart;// [max_cand]  selected -> first work item of its close list
    int32_t *closeNum;  // [max_cand]
    int32_t *counters;  // [8]: 0 n_cand, 1 n_sel, 2 n_work, 3 status
    // ArUco3 (null when the mode is off): contour lengths in T order, and per work item the contour length of the selected
    // candidate whose list it belongs to (a group's close candidates are read in their representative's pyramid level)
    int32_t *tlen;      // [max_cand]
    int32_t *wlen;      // [max_cand]
};

enum { FC_NCAND = 0, FC_NSEL = 1, FC_NWORK = 2, FC_STATUS = 3 };

// --------------------------------------------------------------------------------------------
// One row-word of the closeness matrix: bits j (in word w) of row i, j > i, and the transposed bits.
// lane = candidate j; tq / cent are the T-order quads and (4 * centroid, threshold) of part A.
template <class Ctx>
B2A_HD void close_task(Ctx &ctx, const FrameParams &fp, const float *tq, const float *cent, int n, int wpr, int t, uint32_t *M, uint32_t *MT)
{
    const int lane = ctx.lane(), nl = ctx.lanes();
    const int i = t / wpr, w = t - i * wpr;
    uint32_t bits = 0;
    if (w * 32 + 31 > i) {
        const float *a = tq + (size_t)i * 8;
        const float acx = cent[3 * i], acy = cent[3 * i + 1];
        for (int b = lane; b < 32; b += nl) {
            const int j = w * 32 + b;
            if (j <= i || j >= n) continue;
            const float thr = cent[3 * j + 2];
            // |centroid_a - centroid_b| <= avgDist: cheap exact-safe rejection (integer coordinates)
            const float dcx = (acx - cent[3 * j]) * 0.25f, dcy = (acy - cent[3 * j + 1]) * 0.25f;
            if (dcx * dcx + dcy * dcy > thr * thr * 1.01f + 1.0f) continue;
            if (quad_avg_distance(a, tq + (size_t)j * 8) < thr) { bits |= 1u << b; ctx.atomic_or(&MT[j * wpr + (i >> 5)], 1u << (i & 31)); }
        }
    }
    bits = ctx.warp_or(bits);
    if (lane == 0) M[t] = bits;
}

// phase 0: the whole stage in one go (host emulation); on the device the closeness matrix is filled by many
// CTAs per frame in between: phase 1 = up to the T-order staging, [k_close], phase 2 = the rest.
template <class Ctx>
B2A_HD void frame_group(Ctx &ctx, const FrameParams &fp, const ScaleQuads &sq, const FrameScratch &fs,
                        uint32_t *smemM, int smemM_words, int phase)
{
    const int tid = ctx.tid(), nt = ctx.nthreads();
    int n = 0;
    if (phase != 2) {
    if (tid == 0) fs.counters[FC_STATUS] = 0;
    ctx.sync();
    // ---- 1. gather candidates in reference order (scale by scale, contour list order) + A4 ----
    for (int s = 0; s < fp.nScales; ++s) {
        const int cnt = sq.count[s] < fp.surv_cap ? sq.count[s] : fp.surv_cap;
        for (int base = 0; base < cnt; base += nt) {
            const int i = base + tid;
            const int flag = (i < cnt) ? (int)sq.quad_ok[s * fp.surv_cap + i] : 0;
            int total;
            const int pos = ctx.exclusive_scan(flag, total);
            if (flag && n + pos < fp.max_cand) {
                float *c = fs.cq + (size_t)(n + pos) * 8;
                const int32_t *q = sq.quad_xy + ((size_t)s * fp.surv_cap + i) * 8;
                for (int k = 0; k < 8; ++k) c[k] = (float)q[k];
                quad_make_clockwise(c);
                fs.clen[n + pos] = sq.len[s * fp.surv_cap + i];
            }
            n += total;
        }
    }
    if (n > fp.max_cand) { if (tid == 0) fs.counters[FC_STATUS] = 3; n = fp.max_cand; }
    ctx.sync();
    // ---- 2. stable sort by perimeter, descending: rank = #greater + #equal-before ----
    for (int i = tid; i < n; i += nt) fs.tper[i] = quad_perimeter(fs.cq + (size_t)i * 8);   // tper used as temp (candidate order)
    ctx.sync();
    for (int i = tid; i < n; i += nt) {
        const float p = fs.tper[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) { const float q = fs.tper[j]; rank += (q > p) || (q == p && j < i); }
        fs.members[i] = rank;       // temp: candidate i -> T position
    }
    ctx.sync();
    // T order: quads, perimeters, (4 * centroid, threshold) in global memory
    for (int i = tid; i < n; i += nt) {
        const int r = fs.members[i];
        for (int k = 0; k < 8; ++k) fs.tq[(size_t)r * 8 + k] = fs.cq[(size_t)i * 8 + k];
        if (fs.tlen) fs.tlen[r] = fs.clen[i];
    }
    ctx.sync();
    for (int i = tid; i < n; i += nt) {
        const float *a = fs.tq + (size_t)i * 8;
        const float per = quad_perimeter(a);
        fs.tper[i] = per;
        fs.cent[3 * i] = (a[0] + a[2]) + (a[4] + a[6]); fs.cent[3 * i + 1] = (a[1] + a[3]) + (a[5] + a[7]); fs.cent[3 * i + 2] = f_mul(per, fp.minMarkerDistanceRate);
    }
    {   // the transposed matrix is filled with atomic ORs: clear it
        const int wpr0 = (n + 31) >> 5;
        uint32_t *MT0 = fs.closeM + (size_t)n * wpr0;
        for (int t = tid; t < n * wpr0; t += nt) MT0[t] = 0;
    }
    if (tid == 0) fs.counters[FC_NCAND] = n;
    ctx.sync();
    if (phase == 1) return;
    }   // phase != 2
    else n = fs.counters[FC_NCAND];
    const int wpr = (n + 31) >> 5;
    // ---- 3. closeness bit matrix (global): M row i, bits j > i iff avgDist(T[i],T[j]) < perimeter_j * rate; MT the transposed
    // bits (row j, bits i < j), so that both "smallest close neighbour below" and "above" are first-set-bit scans.
    if (phase == 0) {
        for (int t = ctx.warp(); t < n * wpr; t += ctx.warps()) close_task(ctx, fp, fs.tq, fs.cent, n, wpr, t, fs.closeM, fs.closeM + (size_t)n * wpr);
        ctx.sync();
    }
    // shared-memory staging (device): T-order quads, the two matrices, later the selected-candidate arrays
    const int lane = ctx.lane(), nl = ctx.lanes();
    float *tqs = nullptr;
    uint32_t *Msm = smemM;
    int Mwords = smemM_words;
    if (smemM && 8 * n + 6 * fp.max_cand <= smemM_words) {
        tqs = reinterpret_cast<float *>(smemM);
        Msm = smemM + 8 * n; Mwords = smemM_words - 8 * n;
        for (int t = tid; t < 8 * n; t += nt) tqs[t] = fs.tq[t];
    }
    const float *tq = tqs ? tqs : fs.tq;
    uint32_t *M = fs.closeM, *MT = fs.closeM + (size_t)n * wpr;
    if (Msm && (long long)2 * n * wpr + 6 * fp.max_cand <= (long long)Mwords) {
        for (int t = tid; t < 2 * n * wpr; t += nt) Msm[t] = fs.closeM[t];
        M = Msm; MT = Msm + (size_t)n * wpr;
        Msm += 2 * n * wpr; Mwords -= 2 * n * wpr;
    }
    for (int i = tid; i < n; i += nt) { fs.gid[i] = -1; fs.sel[i] = 1; }
    ctx.sync();
    // ---- 4. group assignment (A5).  OpenCV walks the close pairs (i, j), i < j, in lexicographic order:
    //   both ungrouped -> new group; one grouped -> the other joins it; both grouped -> nothing (no merge).
    // A candidate is therefore decided at its first pair in that order, which gives the closed form
    //   pm(x) = smallest close neighbour < x,  jm(x) = smallest close neighbour > x
    //   par(x) = pm(x)                      if it exists
    //          = ROOT (opens a group)       if pm(jm(x)) == x
    //          = pm(jm(x))                  if x has only larger neighbours and jm(x) was taken by a smaller one
    //          = ISOLATED                   without neighbours
    //   gid(x) = gid(par(x)), par(x) < x;  groups are numbered in the order of their roots
    // (checked against the sequential loop on random graphs and by the golden frames).
    const int G_ROOT = -2, G_ISO = -3;
    int32_t *pm = fs.closeIdx, *jm = fs.closeCnt, *par = fs.gfill, *rootrank = fs.members;      // temporaries
    for (int x = tid; x < n; x += nt) {
        int p = -1, j = -1;
        const int xw = x >> 5;
        for (int w = 0; w <= xw; ++w) { const uint32_t bits = MT[x * wpr + w]; if (bits) { p = w * 32 + ffs32(bits) - 1; break; } }
        for (int w = xw; w < wpr; ++w) { const uint32_t bits = M[x * wpr + w]; if (bits) { j = w * 32 + ffs32(bits) - 1; break; } }
        pm[x] = p; jm[x] = j;
    }
    ctx.sync();
    int ng = 0;
    for (int base = 0; base < n; base += nt) {
        const int x = base + tid;
        int pr = G_ISO;
        if (x < n) {
            if (pm[x] >= 0) pr = pm[x];
            else if (jm[x] >= 0) pr = (pm[jm[x]] == x) ? G_ROOT : pm[jm[x]];
            par[x] = pr;
        }
        int total;
        const int pos = ctx.exclusive_scan(x < n && pr == G_ROOT, total);
        if (x < n) rootrank[x] = ng + pos;
        ng += total;
    }
    ctx.sync();
    for (int x = tid; x < n; x += nt) {
        int r = x;
        while (par[r] >= 0) r = par[r];
        fs.gid[x] = (par[r] == G_ROOT) ? rootrank[r] : -1;
        fs.sel[x] = (par[x] == G_ISO) ? 1 : 0;
    }
    ctx.sync();
    // group segments: members grouped by gid, ascending T index inside a group
    for (int g = tid; g <= ng; g += nt) {
        int c = 0;
        for (int i = 0; i < n; ++i) c += (fs.gid[i] >= 0 && fs.gid[i] < g);
        fs.gstart[g] = c;
    }
    ctx.sync();
    for (int x = tid; x < n; x += nt) {
        const int g = fs.gid[x];
        if (g < 0) continue;
        int rank = 0;
        for (int i = 0; i < x; ++i) rank += (fs.gid[i] == g);
        par[x] = fs.gstart[g] + rank;                    // par is dead; members still holds rootrank
    }
    ctx.sync();
    for (int x = tid; x < n; x += nt) if (fs.gid[x] >= 0) fs.members[par[x]] = x;
    ctx.sync();
    // ---- 5. per group: representative + close contours ----
    for (int g = tid; g < ng; g += nt) {
        const int b = fs.gstart[g], e = fs.gstart[g + 1];
        const bool rev = fp.detectInverted != 0;
        int cur = fs.members[rev ? e - 1 : b];
        fs.sel[cur] = 1;
        int nc = 0;
        for (int kk = b + 1; kk < e; ++kk) {
            const int k = rev ? (b + e - 1 - kk) : kk;
            const int id = fs.members[k];
            const float dist = quad_avg_distance(tq + (size_t)id * 8, tq + (size_t)cur * 8);
            const float ms = quad_module_size(tq + (size_t)id * 8, fp.markerSize, fp.borderBits);
            if (dist > f_mul(fp.minGroupDistance, ms)) { cur = id; fs.closeIdx[b + nc] = id; ++nc; }
        }
        fs.closeCnt[g] = nc;
    }
    ctx.sync();
    // the closeness matrix is dead: its shared-memory region now holds the selected-candidate arrays
    const int mc = fp.max_cand;
    const bool sm6 = Msm && Mwords >= 6 * mc;
    int32_t *S = sm6 ? (int32_t *)Msm : fs.S, *selGroup = sm6 ? (int32_t *)Msm + mc : fs.selGroup;
    int32_t *parent = sm6 ? (int32_t *)Msm + 2 * mc : fs.parent, *depth = sm6 ? (int32_t *)Msm + 3 * mc : fs.depth;
    int32_t *closeStart = sm6 ? (int32_t *)Msm + 4 * mc : fs.closeStart, *closeNum = sm6 ? (int32_t *)Msm + 5 * mc : fs.closeNum;
    // ---- 6. selected candidates (minus the ones near the image border), in T order ----
    int nS = 0;
    for (int base = 0; base < n; base += nt) {
        const int i = base + tid;
        const int flag = (i < n) && fs.sel[i] && !quad_near_border(tq + (size_t)i * 8, fp.W, fp.H, fp.minDistanceToBorder);
        int total;
        const int pos = ctx.exclusive_scan(flag, total);
        if (flag) { S[nS + pos] = i; selGroup[nS + pos] = fs.gid[i]; }
        nS += total;
    }
    ctx.sync();
    // ---- 7. containment hierarchy ----
    for (int i = ctx.warp(); i < nS; i += ctx.warps()) {           // one warp per selected candidate, lanes scan the earlier ones
        int pr = -1;
        const float *a = tq + (size_t)S[i] * 8;
        for (int jb = i - 1; jb >= 0 && pr < 0; jb -= nl) {
            const int j = jb - lane;
            const bool in = j >= 0 && quad_inside_quad(a, tq + (size_t)S[j] * 8);
            const uint32_t hit = ctx.ballot(in);
            if (hit) pr = jb - (ffs32(hit) - 1);                   // the closest earlier candidate that contains it
        }
        if (lane == 0) parent[i] = pr;
    }
    ctx.sync();
    // depth = height of the containment subtree (children always come after their parent)
    for (int i = tid; i < nS; i += nt) {
        int dmax = 0;
        for (int j = i + 1; j < nS; ++j) {
            int a = j, k = 0;
            while (a > i) { a = parent[a]; ++k; }
            if (a == i && k > dmax) dmax = k;
        }
        depth[i] = dmax;
    }
    ctx.sync();
    if (tid == 0) {
        // ---- 8. identification work list ----
        int nw = nS;
        for (int v = 0; v < nS; ++v) {
            const int g = selGroup[v];
            int cnt = 0;
            // only a group's representative (its first member) owns the close list
            if (g >= 0 && fs.members[fs.gstart[g]] == S[v]) cnt = fs.closeCnt[g];
            closeStart[v] = nw;
            if (nw + cnt > fp.max_cand) { cnt = fp.max_cand - nw; fs.counters[FC_STATUS] = 3; }
            closeNum[v] = cnt;
            nw += cnt;
        }
        fs.counters[FC_NCAND] = n;
        fs.counters[FC_NSEL] = nS;
        fs.counters[FC_NWORK] = nw;
        fs.counters[4] = ng;
    }
    ctx.sync();
    if (sm6)
        for (int v = tid; v < nS; v += nt) {
            fs.S[v] = S[v]; fs.selGroup[v] = selGroup[v]; fs.parent[v] = parent[v]; fs.depth[v] = depth[v];
            fs.closeStart[v] = closeStart[v]; fs.closeNum[v] = closeNum[v];
        }
    const int nw = fs.counters[FC_NWORK];
    for (int v = tid; v < nS; v += nt) {
        for (int k = 0; k < 8; ++k) fs.wq[(size_t)v * 8 + k] = tq[(size_t)S[v] * 8 + k];
        if (fs.wlen) fs.wlen[v] = fs.tlen[S[v]];
        const int g = selGroup[v];
        for (int c = 0; c < closeNum[v]; ++c) {
            const int id = fs.closeIdx[fs.gstart[g] + c];
            for (int k = 0; k < 8; ++k) fs.wq[(size_t)(closeStart[v] + c) * 8 + k] = tq[(size_t)id * 8 + k];
            if (fs.wlen) fs.wlen[closeStart[v] + c] = fs.tlen[S[v]];
        }
    }
    for (int w = tid; w < nw; w += nt) fs.wres[w] = 0;
    ctx.sync();
}

// --------------------------------------------------------------------------------------------
struct FrameOutputs {
    int32_t *n_accepted;   // [1]
    int32_t *n_rejected;   // [1]
    float   *corners;      // [max_markers][8]
    int32_t *ids;          // [max_markers]
    float   *rejected;     // [max_markers][8]
    int32_t *status;       // [1]
};

// A6 + output: the depth-ordered acceptance and the output slots by all threads (parallel steps per depth level, ranks by
// block-wide scans), then all threads copy the accepted / rejected quads to their output slots
template <class Ctx>
B2A_HD void frame_finalize(Ctx &ctx, const FrameParams &fp, const FrameScratch &fs, const FrameOutputs &fo)
{
    const int nS = fs.counters[FC_NSEL];
    // valid / was flags live in gid / sel (free after grouping); chosen work item in gfill; after the
    // decision `was` is reused as the output slot: >= 0 accepted slot, <= -2 rejected slot -(slot + 2), -1 dropped
    int32_t *valid = fs.gid, *was = fs.sel, *chosen = fs.gfill;
    // whether a candidate (or one of its close contours, first match) identifies is independent of the order:
    // all threads; valid = 2 marks "would be valid", the single-lane loop below decides which of them are reached
    for (int v = ctx.tid(); v < nS; v += ctx.nthreads()) {
        int ok = 0, ch = v;
        if (fs.wres[v] < 0) ok = 2;
        else
            for (int c = 0; c < fs.closeNum[v]; ++c)
                if (fs.wres[fs.closeStart[v] + c] < 0) { ok = 2; ch = fs.closeStart[v] + c; break; }
        valid[v] = ok; chosen[v] = ch; was[v] = 0;
    }
    ctx.sync();
    // The depth-ordered loop of identifyCandidates, all threads.  What the order decides is only `counter` (a candidate counts
    // once at its own depth level and once more when a valid descendant marks it first) and with it how many depth levels are
    // reached; within a level the marks are a set union, so the level is done in three parallel steps.
    int maxDepth = 0;
    for (int v = 0; v < nS; ++v) if (fs.depth[v] > maxDepth) maxDepth = fs.depth[v];
    int counter = 0;
    for (int depth = 0; counter < nS && depth <= maxDepth; ++depth) {
        int mine = 0;
        for (int v = ctx.tid(); v < nS; v += ctx.nthreads()) {
            if (fs.depth[v] != depth) continue;
            was[v] = 1;
            if (valid[v] == 2) valid[v] = 1;
            ++mine;
        }
        ctx.sync();
        for (int v = ctx.tid(); v < nS; v += ctx.nthreads()) {
            if (fs.depth[v] != depth || valid[v] != 1) continue;
            for (int par = fs.parent[v]; par != -1; par = fs.parent[par]) if (was[par] == 0) was[par] = 2;     // 2 = first marked in this level (racing writers agree)
        }
        ctx.sync();
        for (int v = ctx.tid(); v < nS; v += ctx.nthreads()) if (was[v] == 2) { was[v] = 1; ++mine; }
        counter += ctx.block_sum(mine);
    }
    for (int v = ctx.tid(); v < nS; v += ctx.nthreads()) if (valid[v] != 1) valid[v] = 0;        // never reached (the counter ran out first)
    ctx.sync();
    // output slots in candidate order: accepted and rejected candidates each get their rank
    int na = 0, nr = 0;
    for (int base = 0; base < nS; base += ctx.nthreads()) {
        const int v = base + ctx.tid();
        const bool isv = v < nS && valid[v] != 0, isr = v < nS && valid[v] == 0;
        int totv, totr;
        const int rv = ctx.exclusive_scan(isv ? 1 : 0, totv), rr = ctx.exclusive_scan(isr ? 1 : 0, totr);
        if (isv) was[v] = (na + rv < fp.max_markers) ? na + rv : -1;
        if (isr) was[v] = (nr + rr < fp.max_markers) ? -(nr + rr) - 2 : -1;
        na += totv; nr += totr;
    }
    if (ctx.tid() == 0) {
        int status = fs.counters[FC_STATUS];
        if (*fo.status != 0) status = *fo.status;               // overflow flagged by an earlier stage
        if (na > fp.max_markers || nr > fp.max_markers) status = 3;
        *fo.n_accepted = na < fp.max_markers ? na : fp.max_markers;
        *fo.n_rejected = nr < fp.max_markers ? nr : fp.max_markers;
        *fo.status = status;
    }
    ctx.sync();
    for (int v = ctx.tid(); v < nS; v += ctx.nthreads()) {
        const int slot = was[v];
        if (slot >= 0) {
            const int w = chosen[v];
            const uint32_t res = (uint32_t)fs.wres[w];
            const int rot = (int)(res & 3u), id = (int)((res >> 8) & 0x7FFFFFu);
            const float *c = fs.wq + (size_t)w * 8;
            float *o = fo.corners + (size_t)slot * 8;
            for (int j = 0; j < 4; ++j) {       // std::rotate(begin, begin + 4 - rot, end)
                const int src = (j + 4 - rot) & 3;
                o[2 * j] = c[2 * src]; o[2 * j + 1] = c[2 * src + 1];
            }
            fo.ids[slot] = id;
        } else if (slot <= -2) {
            const float *c = fs.wq + (size_t)v * 8;
            float *o = fo.rejected + (size_t)(-slot - 2) * 8;
            for (int k = 0; k < 8; ++k) o[k] = c[k];
        }
    }
}

}  // namespace b2a

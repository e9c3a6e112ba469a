// core.h -- per-thread / per-warp building blocks of the detect path, written so that the
// same source compiles for the device (inside the kernels of detect_kernels.cuh) and for the
// host (tests/hostemu: a single-lane emulation used by the CPU-only tests to check the
// kernel logic against the CPU restatement; it is test infrastructure, never a product path).
//
// What each block replaces inside cv::aruco::detectMarkers (reference src/aruco_slam.cpp:313;
// algorithm per SURVEY.md Appendix A, OpenCV 4.13.0):
//   border states / walks      -> findContours(RETR_LIST, CHAIN_APPROX_NONE)   (A3a, marking-free form)
//   approx_closed              -> approxPolyDP(closed)                          (A3b)
//   quad tests                 -> isContourConvex + corner-distance gate        (A3)
//   perimeter / avg distance   -> filterTooCloseCandidates arithmetic           (A5)
//   point_in_quad              -> pointPolygonTest(measureDist=false)           (A5 hierarchy)
//   perspective / otsu / bits  -> _extractBits / Dictionary::identify           (A7)
#pragma once
#include <stdint.h>
#include <float.h>
#include <math.h>

#if defined(__CUDACC__)
#define B2A_HD __host__ __device__ __forceinline__
#else
#define B2A_HD inline
#endif
// big helpers that are called from several places: one copy keeps the instruction footprint (and the cold
// instruction-cache misses of latency-bound kernels) small
#if defined(__CUDACC__)
#define B2A_HD_NOINLINE __host__ __device__ __noinline__
#else
#define B2A_HD_NOINLINE inline
#endif
// full unrolling where compile-time indices keep small arrays in registers on the device
#if defined(__CUDA_ARCH__)
#define B2A_UNROLL _Pragma("unroll")
#else
#define B2A_UNROLL
#endif

namespace b2a {

// ---------------------------------------------------------------------------------------------
// bit helpers
// ---------------------------------------------------------------------------------------------
B2A_HD int ffs32(unsigned v)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v);
#else
    return __builtin_ffs((int)v);
#endif
}
B2A_HD int clz32(unsigned v)
{
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}
B2A_HD int popc64(unsigned long long v)
{
#if defined(__CUDA_ARCH__)
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}

// ---------------------------------------------------------------------------------------------
// Border states.  A mask pixel's "code" byte has bit d set iff its neighbour in direction d is
// set (d = 0..7: E, NE, N, NW, W, SW, S, SE; y grows downward); code == 0 means the pixel is
// not set (an isolated set pixel has no border longer than one point and is stored as 0 too).
// A border state is (pixel, s_in) with s_in the direction towards the previous border pixel.
// Suzuki-Abe border following is the permutation succ() on these states; every border is one
// cycle, its first point is the state with the smallest "start key" on the cycle.
// Start keys are strictly monotone in the raster position at which the sequential scan would
// detect the border (outer borders at the pixel itself, hole borders at the 0-pixel to its
// right), outer before hole (key_of below).
// ---------------------------------------------------------------------------------------------
B2A_HD int dir_dx(int d) { return (int)((0x901Au >> (2 * d)) & 3u) - 1; }
B2A_HD int dir_dy(int d) { return (int)((0xA901u >> (2 * d)) & 3u) - 1; }
B2A_HD unsigned ror8(unsigned v, int r)
{
    r &= 7;
    return ((v >> r) | (v << (8 - r))) & 0xFFu;
}
// direction of the next border pixel: first set neighbour scanning s_in+1, s_in+2, ... (ccw)
B2A_HD int succ_dir(unsigned code, int s_in)
{
    unsigned r = ror8(code, s_in + 1);
    return (s_in + ffs32(r)) & 7;            // s_in + 1 + (ffs - 1)
}
// first set neighbour scanning from-1, from-2, ..., from-7 (cw); -1 if none
B2A_HD int first_cw(unsigned code, int from)
{
    unsigned r = ror8(code, from) & 0xFEu;
    if (!r) return -1;
    return (from + (31 - clz32(r))) & 7;
}
// local necessary conditions for a pixel to carry the first state of a border
B2A_HD bool outer_start_candidate(unsigned code) { return code != 0 && (code & 0x1Eu) == 0; }   // W,NW,N,NE clear
B2A_HD bool hole_start_candidate(unsigned code) { return (code & 3u) == 2u; }                   // E clear, NE set

// Packed mask access: `plane` points at row -1; pixel (x,y) is bit (x & 31) of word
// (y + 1) * PWW + (x >> 5) + 1.  One zero word left and >= 1 right of every row and zero rows
// above and below, so 3x3 neighbourhoods never need bounds checks.
B2A_HD uint32_t ld_ro(const uint32_t *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
struct MaskView {
    const uint32_t *plane;
    int PWW;
    B2A_HD unsigned win3(int x, int y) const      // bits x-1, x, x+1 of row y
    {
        // two unconditional loads + funnel shift: no data-dependent branch between the loads, so
        // the six loads of a 3x3 neighbourhood are all in flight together
        const int bit = x + 31;                                   // padded bit position of pixel x-1
        const uint32_t *row = plane + (size_t)(y + 1) * PWW + (bit >> 5);
        const uint32_t lo = ld_ro(row), hi = ld_ro(row + 1);
#if defined(__CUDA_ARCH__)
        return __funnelshift_r(lo, hi, bit & 31) & 7u;
#else
        return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> (bit & 31)) & 7u;
#endif
    }
    B2A_HD uint32_t word(int wx, int y) const { return ld_ro(plane + (size_t)(y + 1) * PWW + wx + 1); }   // wx = -1 / WW are the zero pads
    // 3x3 window of pixel (x,y) as 9 bits: rows y-1, y, y+1 at bits 0-2, 3-5, 6-8 (bit j = column x-1+j)
    B2A_HD unsigned win9(int x, int y) const { return win3(x, y - 1) | (win3(x, y) << 3) | (win3(x, y + 1) << 6); }
    // neighbour code of pixel (x,y): bit d = neighbour in direction d
    B2A_HD unsigned operator()(int x, int y) const { return code_of_win9(win9(x, y)); }
    B2A_HD static unsigned code_of_win9(unsigned w)
    {
        const unsigned u = w & 7u, m = (w >> 3) & 7u, d = (w >> 6) & 7u;
        return ((m >> 2) & 1u) | (((u >> 2) & 1u) << 1) | (((u >> 1) & 1u) << 2) | ((u & 1u) << 3) |
               ((m & 1u) << 4) | ((d & 1u) << 5) | (((d >> 1) & 1u) << 6) | (((d >> 2) & 1u) << 7);
    }
};

// The same 3x3 window for a border walk: consecutive pixels of a walk are neighbours, so the three 64-bit row pieces of the
// previous step are kept and a step loads at most one new row (two words) instead of six words with three row addresses.
// A piece covers the pixels 32 w - 31 .. 32 w + 32; it is kept while the window's first bit stays inside its first 62 bits
// (hysteresis: a border that wiggles across a word boundary does not reload every step).
struct CachedMaskView {
    const uint32_t *plane;
    int PWW;
    mutable int cy, cw;                           // rows cy - 1, cy, cy + 1 from word index cw (cw < 0: nothing cached)
    mutable unsigned long long r0, r1, r2;
    B2A_HD CachedMaskView(const uint32_t *plane_, int PWW_) : plane(plane_), PWW(PWW_), cy(0), cw(-1), r0(0), r1(0), r2(0) {}
    B2A_HD unsigned long long piece(int y, int w) const
    {
        const uint32_t *row = plane + (size_t)(y + 1) * PWW + w;
        return (unsigned long long)ld_ro(row) | ((unsigned long long)ld_ro(row + 1) << 32);
    }
    B2A_HD unsigned win9(int x, int y) const
    {
        const int bit = x + 31;                                   // padded bit position of pixel x - 1
        int rel = bit - (cw << 5);
        const int dy = y - cy;
        if (cw < 0 || (unsigned)rel > 61u || (unsigned)(dy + 1) > 2u) {
            cw = bit >> 5; rel = bit & 31;
            r0 = piece(y - 1, cw); r1 = piece(y, cw); r2 = piece(y + 1, cw);
        } else if (dy == 1) { r0 = r1; r1 = r2; r2 = piece(y + 1, cw); }
        else if (dy == -1) { r2 = r1; r1 = r0; r0 = piece(y - 1, cw); }
        cy = y;
        return ((unsigned)(r0 >> rel) & 7u) | (((unsigned)(r1 >> rel) & 7u) << 3) | (((unsigned)(r2 >> rel) & 7u) << 6);
    }
};

// ---------------------------------------------------------------------------------------------
// Border graph.  The states that lie on borders are locally enumerable (checked against
// cv2.findContours on every golden mask and on random masks: the state set below has exactly
// one element per traced border point and its succ-cycles are exactly the borders longer than
// one point):
//     a set pixel c with code != 0 carries one state per maximal run of clear neighbours around
//     its 8-ring that contains a 4-neighbour D in {E, N, W, S};  s_in = first_cw(code, D).
// The run's clockwise-most 4-neighbour is the state's "canonical" D: D clear and (D-1 set or D-2 set).
// Two nested, locally decidable grids of states structure every border that is long enough to cross them:
//     ANCHORS        the state hugs a clear W or E neighbour and y % R == 0, or a clear N or S neighbour and x % R == 0
//     SUPER ANCHORS  the same with R2 = 8 R (a subset of the anchors)
// and the borders that carry no anchor at all (confined to a grid cell) are found from the
//     START CANDIDATES  states with the local necessary conditions of a border's first state.
// The stage runs as
//     anchors    : enumerate anchors (+ super flag) and start candidates, word-parallel
//     segments   : every anchor walks to the next anchor (<= ~R steps, all in parallel)
//     skip       : every super anchor hops over the plain anchors to the next super anchor
//     cycles     : leader (= the segment holding the border's smallest start key = cv2's first point) and length,
//                  by hops over the super anchors (or the plain anchors of a border without super anchors)
//     direct     : every start candidate walks its border both ways and gives up at the first anchor or smaller
//                  key; what closes is a border without anchors, reported with its length
//     order / offsets, assign (position of every segment inside its border), emit (points)
// Tables, indexed by the 3x3 window w9 and a direction:
//   succ[w9 | s_in << 9] : successor direction of state (pixel, s_in) | its flags
//   pred[w9 | t    << 9] : w9 = window of the previous border pixel p, t = direction p -> c:
//                          s_in of the predecessor state (p, .) | that state's flags
//   pix[w9]              : byte k (canonical D = 2k): 0x80 present | flags >> 2 | s_in   (host cross-check only)
// flags: WT_OUTER / WT_HOLE start eligibility, ST_ROW / ST_COL grid classes, ST_UNC start candidate, ST_VALID
// ---------------------------------------------------------------------------------------------
enum { WT_OUTER = 8, WT_HOLE = 16, WT_ELIG = 24, ST_ROW = 32, ST_COL = 64, ST_UNC = 128, ST_VALID = 256 };
enum : uint32_t { A_NONE = 0xFFFFFFFFu, SEG_OVERFLOW = 0x7FFFFFFFu, SEG_LEN = 0x7FFFFFFFu, SEG_SUPER = 0x80000000u };
struct WalkTables { uint16_t succ[4096]; uint16_t pred[4096]; uint32_t pix[512]; };

B2A_HD unsigned state_eligibility(unsigned code, int s)
{
    if (!(code & 16u) && first_cw(code, 4) == s) return WT_OUTER;
    if (!(code & 1u) && first_cw(code, 0) == s) return WT_HOLE;
    return 0;
}
// flags of state (code, s); 0 if (code, s) is not a border state
B2A_HD unsigned state_flags(unsigned code, int s)
{
    if (!code) return 0;
    unsigned types = 0;
    for (int D = 0; D < 8; D += 2)
        if (!((code >> D) & 1u) && first_cw(code, D) == s) types |= 1u << (D >> 1);
    if (!types) return 0;
    unsigned f = ST_VALID | state_eligibility(code, s);
    if (types & 5u) f |= ST_ROW;                 // E or W
    if (types & 10u) f |= ST_COL;                // N or S
    if (((f & WT_OUTER) && outer_start_candidate(code)) || ((f & WT_HOLE) && hole_start_candidate(code))) f |= ST_UNC;
    return f;
}
B2A_HD void build_walk_table_entry(WalkTables &t, int idx)
{
    const unsigned w9 = (unsigned)idx & 511u;
    const unsigned code = ((w9 >> 4) & 1u) ? MaskView::code_of_win9(w9) : 0u;
    const int s = idx >> 9;
    const unsigned f = state_flags(code, s);
    t.succ[idx] = (uint16_t)(f ? ((unsigned)succ_dir(code, s) | f) : 0u);
    {   // here s plays the role of t (direction p -> c); the predecessor state is (p, sp)
        const int cw = first_cw(code, s);
        const int sp = cw < 0 ? s : cw;
        t.pred[idx] = (uint16_t)(code ? ((unsigned)sp | state_flags(code, sp)) : 0u);
    }
    if (s == 0) {
        uint32_t p = 0;
        for (int k = 0; k < 4; ++k) {
            const int D = 2 * k, d1 = (D + 7) & 7, d2 = (D + 6) & 7;
            if (!code || ((code >> D) & 1u)) continue;
            int sk;
            if ((code >> d1) & 1u) sk = d1; else if ((code >> d2) & 1u) sk = d2; else continue;
            const unsigned fk = state_flags(code, sk);
            p |= (0x80u | ((fk & (ST_ROW | ST_COL | ST_UNC)) >> 2) | (unsigned)sk) << (8 * k);
        }
        t.pix[w9] = p;
    }
}
B2A_HD uint32_t key_of(int x, int y, unsigned elig, int KS)
{
    return (elig & WT_OUTER) ? (uint32_t)((y * KS + x) * 2) : (uint32_t)((y * KS + x + 1) * 2 + 1);
}
// Rm = R - 1 (R a power of two); flags carry ST_ROW / ST_COL
B2A_HD bool is_anchor(unsigned flags, int x, int y, int Rm)
{
    return ((flags & ST_ROW) && !(y & Rm)) || ((flags & ST_COL) && !(x & Rm));
}
// bits of a mask word whose pixel x = 32 wx + b has x % R == 0 (R = Rm + 1, a power of two)
B2A_HD uint32_t grid_cols(int wx, int Rm)
{
    if (Rm >= 32) return ((wx << 5) & Rm) ? 0u : 1u;
    return 0xFFFFFFFFu / ((Rm >= 31) ? 0xFFFFFFFFu : ((2u << Rm) - 1u));
}

// Word-parallel enumeration for the 32 pixels of mask word m (row y; ml / mr = the words left / right of
// it, u* the row above, d* the row below; bit i = pixel 32 w + i), per canonical direction D = 2k:
//   A[k]  pixels whose state k is an anchor (rowflag: y % R == 0, cols: bits with x % R == 0)
//   SU[k] ... is a super anchor (rowflag2 / cols2 for R2); a subset of A[k]
//   U[k]  ... is a start candidate
//   hi[k] s_in = D - 1 where set, else D - 2
// A pixel's states are ordered by k; the states of a word by pixel, then k.
B2A_HD void anchor_words(uint32_t m, uint32_t ml, uint32_t mr, uint32_t u, uint32_t ul, uint32_t ur, uint32_t d, uint32_t dl, uint32_t dr,
                         bool rowflag, uint32_t cols, bool rowflag2, uint32_t cols2,
                         uint32_t A[4], uint32_t SU[4], uint32_t U[4], uint32_t hi[4], uint32_t &iso)
{
    uint32_t n[8];
    n[0] = (m >> 1) | (mr << 31); n[1] = (u >> 1) | (ur << 31); n[2] = u; n[3] = (u << 1) | (ul >> 31);
    n[4] = (m << 1) | (ml >> 31); n[5] = (d << 1) | (dl >> 31); n[6] = d; n[7] = (d >> 1) | (dr << 31);
    iso = m & ~(n[0] | n[1] | n[2] | n[3] | n[4] | n[5] | n[6] | n[7]);
    const uint32_t outerCand = ~(n[1] | n[2] | n[3] | n[4]), holeCand = ~n[0] & n[1];
    const uint32_t rows = rowflag ? 0xFFFFFFFFu : 0u, rows2 = rowflag2 ? 0xFFFFFFFFu : 0u;
    B2A_UNROLL
    for (int k = 0; k < 4; ++k) {
        const int D = 2 * k;
        const uint32_t canon = m & ~n[D] & (n[(D + 7) & 7] | n[(D + 6) & 7]);
        // 4-neighbours inside the state's clear run, walking counter-clockwise from D
        const uint32_t c2 = canon & ~n[(D + 1) & 7] & ~n[(D + 2) & 7];
        const uint32_t c4 = c2 & ~n[(D + 3) & 7] & ~n[(D + 4) & 7];
        const uint32_t c6 = c4 & ~n[(D + 5) & 7] & ~n[(D + 6) & 7];
        uint32_t has[4];                         // has[j]: run contains direction 2 j
        has[k] = canon; has[(k + 1) & 3] = c2; has[(k + 2) & 3] = c4; has[(k + 3) & 3] = c6;
        const uint32_t row_t = has[0] | has[2], col_t = has[1] | has[3];
        A[k] = canon & ((row_t & rows) | (col_t & cols));
        SU[k] = canon & ((row_t & rows2) | (col_t & cols2));
        U[k] = (has[2] & outerCand) | (has[0] & ~has[2] & holeCand);
        hi[k] = n[(D + 7) & 7];
    }
}
B2A_HD int state_dir(int k, uint32_t hi_k, int b) { return (2 * k + (((hi_k >> b) & 1u) ? 7 : 6)) & 7; }
B2A_HD int popc32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
// Rank of anchor state (x, y, s) among the anchors of its mask word (the word map holds the index of the
// word's first anchor); -1 if the state is not an anchor of the word.  word(wx, y) reads a mask word.
template <class Mask>
B2A_HD int anchor_rank_in_word(const Mask &mk, int x, int y, int s, int Rm)
{
    const int wx = x >> 5, b = x & 31;
    uint32_t A[4], SU[4], U[4], hi[4], iso;
    anchor_words(mk.word(wx, y), mk.word(wx - 1, y), mk.word(wx + 1, y), mk.word(wx, y - 1), mk.word(wx - 1, y - 1), mk.word(wx + 1, y - 1),
                 mk.word(wx, y + 1), mk.word(wx - 1, y + 1), mk.word(wx + 1, y + 1), !(y & Rm), grid_cols(wx, Rm), false, 0u, A, SU, U, hi, iso);
    const uint32_t below = (1u << b) - 1u;
    int r = popc32(A[0] & below) + popc32(A[1] & below) + popc32(A[2] & below) + popc32(A[3] & below);
    B2A_UNROLL
    for (int k = 0; k < 4; ++k) {
        if (!((A[k] >> b) & 1u)) continue;
        if (state_dir(k, hi[k], b) == s) return r;
        ++r;
    }
    return -1;
}

// Walk from anchor state (x,y,s) to the next anchor.  len = number of states of the segment
// (SEG_OVERFLOW once more than max_len steps were taken), minkey / minoff = smallest start key
// among the segment's start-eligible states and its offset (A_NONE if none); (x,y,s) is left at the next anchor.
// The first SEG_CODE_WORDS * 10 successor directions of the segment are also left in `codes` (3 bits each,
// 10 per word) when it is not null: the emit pass then rebuilds the points without touching the mask.
constexpr int SEG_CODE_WORDS = 8;
template <class Win>
B2A_HD void seg_walk(const Win &win, const uint16_t *__restrict__ succ, int KS, int Rm, int max_len,
                     int &x, int &y, int &s, uint32_t &len, uint32_t &minkey, uint32_t &minoff, uint32_t *__restrict__ codes)
{
    unsigned e = succ[win.win9(x, y) | ((unsigned)s << 9)];
    uint32_t n = 0, cw = 0;
    int ci = 0, cshift = 0;
    minkey = A_NONE; minoff = 0;
    for (;;) {
        if (e & WT_ELIG) { const uint32_t k = key_of(x, y, e, KS); if (k < minkey) { minkey = k; minoff = n; } }
        const int so = (int)(e & 7u);
        cw |= (uint32_t)so << cshift; cshift += 3;
        if (cshift == 30) { if (codes && ci < SEG_CODE_WORDS) codes[ci] = cw; ++ci; cw = 0; cshift = 0; }
        x += dir_dx(so); y += dir_dy(so); s = so ^ 4;
        ++n;
        e = succ[win.win9(x, y) | ((unsigned)s << 9)];
        if (is_anchor(e, x, y, Rm)) { len = n; break; }
        if (n > (uint32_t)max_len) { len = SEG_OVERFLOW; break; }
    }
    if (codes && cshift && ci < SEG_CODE_WORDS) codes[ci] = cw;
}
// points of a segment from its stored direction codes (len <= SEG_CODE_WORDS * 10)
B2A_HD void seg_emit_codes(const uint32_t *__restrict__ codes, int x, int y, int len, int pos, int n, uint32_t *__restrict__ out)
{
    uint32_t cw = 0;
    for (int k = 0; k < len; ++k) {
        const int q = pos + k;
        out[q < 0 ? q + n : q] = (uint32_t)x | ((uint32_t)y << 16);
        const int r = k % 10;
        if (r == 0) cw = codes[k / 10];
        const int so = (int)((cw >> (3 * r)) & 7u);
        x += dir_dx(so); y += dir_dy(so);
    }
}
// Emit the len points of the segment that starts at state (x,y,s): point k goes to
// out[pos + k] (+ n when negative: only the leader's segment wraps)
template <class Win>
B2A_HD void seg_emit(const Win &win, const uint16_t *__restrict__ succ, int x, int y, int s, int len, int pos, int n, uint32_t *__restrict__ out)
{
    for (int k = 0;; ) {
        const int q = pos + k;
        out[q < 0 ? q + n : q] = (uint32_t)x | ((uint32_t)y << 16);
        if (++k >= len) break;
        const int so = (int)(succ[win.win9(x, y) | ((unsigned)s << 9)] & 7u);
        x += dir_dx(so); y += dir_dy(so); s = so ^ 4;
    }
}
// A start candidate (x0,y0,s0) with start key key0 walks its border in both directions at once.
// Returns the border length if the border carries no anchor and (x0,y0,s0) is its first state;
// 0 as soon as an anchor or a start-eligible state with a smaller key is met (the border is then
// reported through its anchors, or by that other state); -1 when undecided after max_len steps.
// (winf / winb: the window views of the forward and of the backward walker -- two objects so that a caching view keeps one
// neighbourhood per walker; a stateless view can be passed twice)
template <class Win>
B2A_HD int direct_walk(const Win &winf, const Win &winb, const uint16_t *__restrict__ succ, const uint16_t *__restrict__ pred, int KS, int Rm,
                       int x0, int y0, int s0, uint32_t key0, int max_len)
{
    int xf = x0, yf = y0, sf = s0, xb = x0, yb = y0, sb = s0;
    const unsigned e0 = succ[winf.win9(x0, y0) | ((unsigned)s0 << 9)];
    if (is_anchor(e0, x0, y0, Rm)) return 0;
    int so = (int)(e0 & 7u);
    int n = 0;
    for (;;) {
        xf += dir_dx(so); yf += dir_dy(so); sf = so ^ 4;
        ++n;
        if (xf == xb && yf == yb && sf == sb) return n;
        const int xp = xb + dir_dx(sb), yp = yb + dir_dy(sb);
        const unsigned wf = winf.win9(xf, yf), wp = winb.win9(xp, yp);
        const unsigned ef = succ[wf | ((unsigned)sf << 9)], ep = pred[wp | ((unsigned)(sb ^ 4) << 9)];
        if (is_anchor(ef, xf, yf, Rm) || ((ef & WT_ELIG) && key_of(xf, yf, ef, KS) < key0)) return 0;
        if (n > max_len) return -1;
        sb = (int)(ep & 7u); xb = xp; yb = yp;
        ++n;
        if (xf == xb && yf == yb && sf == sb) return n;
        if (is_anchor(ep, xb, yb, Rm) || ((ep & WT_ELIG) && key_of(xb, yb, ep, KS) < key0)) return 0;
        if (n > max_len) return -1;
        so = (int)(ef & 7u);
    }
}

// seg[i] = (next anchor, previous anchor, segment length | SEG_SUPER, segment min key).
// Hop over the anchors of anchor i's border in both directions.  Returns the border length if
// i's segment holds the border's smallest start key (i is the border's leader), 0 otherwise
// (also when the border is longer than max_len, a segment overflowed, or stop(segment) is true
// for an anchor of the border other than i).
struct Seg { uint32_t next, prev, len, minkey; };
struct NeverStop { B2A_HD bool operator()(const Seg &) const { return false; } };
struct StopAtSuper { B2A_HD bool operator()(const Seg &s) const { return (s.len & SEG_SUPER) != 0; } };
template <class SegAt, class Stop>
B2A_HD uint32_t cycle_leader(const SegAt &seg_at, const Stop &stop, uint32_t i, int max_len)
{
    const Seg me = seg_at(i);
    if (me.minkey == A_NONE || (me.len & SEG_LEN) == SEG_OVERFLOW) return 0;
    uint32_t total = me.len & SEG_LEN, f = i, b = i, fn = me.next, bp = me.prev;
    for (;;) {
        if (fn == A_NONE || bp == A_NONE) return 0;
        if (fn == b) break;
        const Seg sf = seg_at(fn), sb = seg_at(bp);
        if (stop(sf) || sf.minkey < me.minkey || (sf.len & SEG_LEN) == SEG_OVERFLOW) return 0;
        total += sf.len & SEG_LEN; f = fn; fn = sf.next;
        if (total > (uint32_t)max_len) return 0;
        if (bp == f) break;
        if (stop(sb) || sb.minkey < me.minkey || (sb.len & SEG_LEN) == SEG_OVERFLOW) return 0;
        total += sb.len & SEG_LEN; b = bp; bp = sb.prev;
        if (total > (uint32_t)max_len) return 0;
    }
    return total;
}
// Leader i of a kept border of n points whose first point sits minoff states into i's segment:
// give every anchor of the border the position of its segment's first state (set(a, pos)).
template <class SegAt, class Set>
B2A_HD void cycle_assign(const SegAt &seg_at, const Set &set, uint32_t i, int n, int minoff)
{
    const Seg me = seg_at(i);
    set(i, -minoff);
    int pf = (int)(me.len & SEG_LEN) - minoff, pb = n - minoff;
    uint32_t f = i, b = i, fn = me.next, bp = me.prev;
    for (;;) {
        if (fn == b) break;
        const Seg sf = seg_at(fn), sb = seg_at(bp);
        set(fn, pf); pf += (int)(sf.len & SEG_LEN); f = fn; fn = sf.next;
        if (bp == f) break;
        pb -= (int)(sb.len & SEG_LEN); set(bp, pb); b = bp; bp = sb.prev;
    }
}
// Second level: a super anchor's segment runs to the next super anchor of its border (super_skip hops over
// the plain anchors in between; the R2 grid bounds their number), so the leader / assign hops over a long
// border touch ~8x fewer nodes; borders without a super anchor are resolved on the plain anchors.
// out: next super anchor (A_NONE on overflow), super segment length, min key, offset of the min-key state
template <class SegAt, class MinoffAt>
B2A_HD void super_skip(const SegAt &seg_at, const MinoffAt &minoff_at, uint32_t S, int max_len,
                       uint32_t &snext, uint32_t &slen, uint32_t &smin, uint32_t &soff)
{
    uint32_t cur = S;
    Seg sg = seg_at(S);
    slen = 0; smin = A_NONE; soff = 0;
    for (;;) {
        if ((sg.len & SEG_LEN) == SEG_OVERFLOW || sg.next == A_NONE) { snext = A_NONE; slen = SEG_OVERFLOW; return; }
        if (sg.minkey < smin) { smin = sg.minkey; soff = slen + minoff_at(cur); }
        slen += sg.len & SEG_LEN;
        if (slen > (uint32_t)max_len) { snext = A_NONE; slen = SEG_OVERFLOW; return; }
        cur = sg.next;
        sg = seg_at(cur);
        if (sg.len & SEG_SUPER) { snext = cur; return; }
    }
}
// positions of the plain anchors behind super anchor S (whose own position is pos)
template <class SegAt, class Set>
B2A_HD void super_assign(const SegAt &seg_at, const Set &set, uint32_t S, int pos)
{
    Seg sg = seg_at(S);
    for (;;) {
        pos += (int)(sg.len & SEG_LEN);
        const uint32_t cur = sg.next;
        sg = seg_at(cur);
        if (sg.len & SEG_SUPER) return;
        set(cur, pos);
    }
}

// ---------------------------------------------------------------------------------------------
// approxPolyDP(closed = true) on integer points, cooperative over a group of lanes.
// LG supplies lane(), nlanes() and argmax_first(d, pos): the maximum d over the lanes and,
// among equal maxima, the smallest pos (= OpenCV's strict '>' first-maximum rule).
// Only results with <= 8 vertices before the clean-up pass can end as quadrilaterals, so the
// recursion stops (returns -1) as soon as more than 8 vertices or 16 pending slices exist.
// ---------------------------------------------------------------------------------------------
struct SingleLane {
    B2A_HD int lane() const { return 0; }
    B2A_HD int nlanes() const { return 1; }
    B2A_HD void argmax_first(long long &, int &) const {}
};

B2A_HD int px_of(uint32_t p) { return (int)(p & 0xFFFFu); }
B2A_HD int py_of(uint32_t p) { return (int)(p >> 16); }

template <class LG>
B2A_HD int approx_closed(const LG &lg, const uint32_t *__restrict__ P, int count, double eps, int *ox, int *oy)
{
    const double eps2 = eps * eps;
    const int lane = lg.lane(), nl = lg.nlanes();
    int right = 0, pos = 0, startp = 0;
    long long maxd = 0;
    for (int it = 0; it < 3; ++it) {
        pos = (pos + right) % count;
        startp = pos;
        const int sx = px_of(P[pos]), sy = py_of(P[pos]);
        pos = (pos + 1) % count;
        // coordinates are below 2^15: squared distances fit 32 bits
        uint32_t bd32 = 0; int bj = 0x7FFFFFFF;
        for (int j0 = 1 + lane; j0 < count; j0 += 4 * nl) {          // four points per lane in flight
            uint32_t q[4];
            const int jb = j0 - lane;                                // same on every lane: slots past the end are skipped by the whole group
            B2A_UNROLL
            for (int k = 0; k < 4; ++k) {
                if (jb + k * nl >= count) break;
                const int j = j0 + k * nl;
                int idx = pos + j - 1; if (idx >= count) idx -= count;
                q[k] = j < count ? P[idx] : 0u;
            }
            B2A_UNROLL
            for (int k = 0; k < 4; ++k) {
                if (jb + k * nl >= count) break;
                const int j = j0 + k * nl;
                const int dx = px_of(q[k]) - sx, dy = py_of(q[k]) - sy;
                const uint32_t d = (uint32_t)(dx * dx) + (uint32_t)(dy * dy);
                if (j < count && d > bd32) { bd32 = d; bj = j; }
            }
        }
        long long bd = (long long)bd32;
        lg.argmax_first(bd, bj);
        maxd = bd;
        if (bd > 0) right = bj;
        pos = (pos + count - 1) % count;
    }
    int m = 0;
    if ((double)maxd <= eps2) { ox[0] = px_of(P[startp]); oy[0] = py_of(P[startp]); return 1; }
    int st_s[16], st_e[16], top = 0;
    {
        int a = pos % count, b = (right + a) % count;
        st_s[0] = b; st_e[0] = a; st_s[1] = a; st_e[1] = b; top = 2;
    }
    while (top > 0) {
        --top;
        const int s = st_s[top], e = st_e[top];
        const int ex = px_of(P[e]), ey = py_of(P[e]);
        const int sx = px_of(P[s]), sy = py_of(P[s]);
        int inner = e - s - 1; if (inner < 0) inner += count;     // points strictly between s and e (count-1 when s == e)
        bool le = true;
        int split = 0;
        if (inner > 0) {
            // |coordinate differences| < 2^15: dot and cross products fit int32 (|a b + c d| < 2^31), squared
            // lengths fit uint32; only the final products need 64 bits (one widening multiply each)
            const int dx = ex - sx, dy = ey - sy;
            const uint32_t L2u = (uint32_t)(dx * dx) + (uint32_t)(dy * dy);
            const long long L2 = (long long)L2u;
            unsigned long long bdu = 0; int bt = 0x7FFFFFFF;
            for (int t0 = lane; t0 < inner; t0 += 4 * nl) {            // four points per lane in flight
                uint32_t qq[4];
                const int tb = t0 - lane;                            // same on every lane
                B2A_UNROLL
                for (int k = 0; k < 4; ++k) {
                    if (tb + k * nl >= inner) break;
                    const int t = t0 + k * nl;
                    int idx = s + 1 + t; if (idx >= count) idx -= count;
                    qq[k] = t < inner ? P[idx] : 0u;
                }
                B2A_UNROLL
                for (int k = 0; k < 4; ++k) {
                    if (tb + k * nl >= inner) break;
                    const int t = t0 + k * nl;
                    const int px = px_of(qq[k]) - sx, py = py_of(qq[k]) - sy;
                    const int dot = px * dx + py * dy;
                    // one widening multiply a * b with selected factors, no divergent branch: before the segment |p - s|^2 * L2,
                    // past it |p - e|^2 * L2, beside it cross^2
                    const int qx = px - dx, qy = py - dy, cr = py * dx - px * dy;
                    const uint32_t acr = (uint32_t)(cr < 0 ? -cr : cr);
                    const bool before = dot < 0, past = !before && (uint32_t)dot > L2u;
                    const uint32_t fa = before ? (uint32_t)(px * px) + (uint32_t)(py * py) : past ? (uint32_t)(qx * qx) + (uint32_t)(qy * qy) : acr;
                    const uint32_t fb = (before || past) ? L2u : acr;
                    const unsigned long long d = (unsigned long long)fa * fb;
                    if (t < inner && d > bdu) { bdu = d; bt = t; }
                }
            }
            long long bd = (long long)bdu;
            lg.argmax_first(bd, bt);
            if (bd > 0) { split = s + 1 + bt; if (split >= count) split -= count; }
            le = (double)bd <= eps2 * (double)L2;
        }
        if (le) {
            if (m >= 8) return -1;
            ox[m] = sx; oy[m] = sy; ++m;
        } else {
            if (top + 2 > 16) return -1;
            st_s[top] = split; st_e[top] = e; ++top;
            st_s[top] = s; st_e[top] = split; ++top;
        }
    }
    // clean-up pass (sequential, <= 8 points; every lane computes the same thing)
    int new_count = m;
    {
        int r = 0, w = 0;
        double stx = ox[m - 1], sty = oy[m - 1];
        double ptx = ox[0], pty = oy[0];
        r = 1 % m;
        for (int i = 0; i < m && new_count > 2; ++i) {
            const double ex = ox[r], ey = oy[r];
            r = (r + 1) % m;
            const double dx = ex - stx, dy = ey - sty;
            const double dist = fabs((ptx - stx) * dy - (pty - sty) * dx);
            const double sip = (ptx - stx) * (ex - ptx) + (pty - sty) * (ey - pty);
            if (dist * dist <= 0.5 * eps2 * (dx * dx + dy * dy) && dx != 0 && dy != 0 && sip >= 0) {
                --new_count;
                ox[w] = (int)ex; oy[w] = (int)ey; w = (w + 1) % m;
                stx = ex; sty = ey;
                ptx = ox[r]; pty = oy[r];
                r = (r + 1) % m;
                ++i;
                continue;
            }
            ox[w] = (int)ptx; oy[w] = (int)pty; w = (w + 1) % m;
            stx = ptx; sty = pty;
            ptx = ex; pty = ey;
        }
    }
    return new_count;
}

// isContourConvex on 4 integer points (any collinear consecutive edge pair -> false)
B2A_HD bool quad_is_convex(const int *qx, const int *qy)
{
    long long px = qx[2], py = qy[2], cx = qx[3], cy = qy[3];
    long long dx0 = cx - px, dy0 = cy - py;
    int o = 0;
    for (int i = 0; i < 4; ++i) {
        px = cx; py = cy; cx = qx[i]; cy = qy[i];
        long long dx = cx - px, dy = cy - py;
        long long a = dx * dy0, b = dy * dx0;
        o |= (b > a) ? 1 : ((b < a) ? 2 : 3);
        if (o == 3) return false;
        dx0 = dx; dy0 = dy;
    }
    return true;
}

// A3 gate after approxPolyDP: 4 vertices, convex, shortest side >= n * minCornerDistanceRate
B2A_HD bool quad_passes(const int *qx, const int *qy, int n_approx, int contour_len, int maxWH, double minCornerDistanceRate)
{
    if (n_approx != 4 || !quad_is_convex(qx, qy)) return false;
    double minDistSq = (double)maxWH * (double)maxWH;
    for (int j = 0; j < 4; ++j) {
        double dx = qx[j] - qx[(j + 1) & 3], dy = qy[j] - qy[(j + 1) & 3];
        double d = dx * dx + dy * dy;
        if (d < minDistSq) minDistSq = d;
    }
    double mcd = (double)contour_len * minCornerDistanceRate;
    return !(minDistSq < mcd * mcd);
}

// A4: make the corner order clockwise (swap corners 1 and 3 when the cross product is negative)
B2A_HD void quad_make_clockwise(float *c)
{
    double dx1 = c[2] - c[0], dy1 = c[3] - c[1], dx2 = c[4] - c[0], dy2 = c[5] - c[1];
    if (dx1 * dy2 - dy1 * dx2 < 0.0) {
        float tx = c[2], ty = c[3];
        c[2] = c[6]; c[3] = c[7]; c[6] = tx; c[7] = ty;
    }
}

// ---- float32 arithmetic of filterTooCloseCandidates (A5); operation order matters ----
B2A_HD float f_mul(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
B2A_HD float f_add(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
B2A_HD float f_sqrt(float a)
{
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(a);
#else
    return sqrtf(a);
#endif
}
B2A_HD float f_div(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
B2A_HD float norm2f(float dx, float dy) { return f_add(f_mul(dx, dx), f_mul(dy, dy)); }

B2A_HD float quad_perimeter(const float *c)
{
    float p = 0.f;
    for (int i = 0; i < 4; ++i) {
        int j = (i + 1) & 3;
        p = f_add(p, f_sqrt(norm2f(c[2 * i] - c[2 * j], c[2 * i + 1] - c[2 * j + 1])));
    }
    return p;
}
B2A_HD float quad_avg_distance(const float *a, const float *b)
{
    float minsq = FLT_MAX;
    for (int fc = 0; fc < 4; ++fc) {
        float dsq = 0.f;
        for (int c = 0; c < 4; ++c) {
            int mc = (c + fc) & 3;
            dsq = f_add(dsq, norm2f(a[2 * mc] - b[2 * c], a[2 * mc + 1] - b[2 * c + 1]));
        }
        dsq = f_div(dsq, 4.f);
        if (dsq < minsq) minsq = dsq;
    }
    return f_sqrt(minsq);
}
B2A_HD float quad_module_size(const float *c, int markerSize, int borderBits)
{
    float a = quad_perimeter(c);
    int nm = markerSize + borderBits * 2;
    return f_div(a, 4.f * (float)nm);
}
B2A_HD bool quad_near_border(const float *c, int W, int H, int mdb)
{
    for (int j = 0; j < 4; ++j) {
        float x = c[2 * j], y = c[2 * j + 1];
        if (x < (float)mdb || y < (float)mdb || x > (float)(W - 1 - mdb) || y > (float)(H - 1 - mdb)) return true;
    }
    return false;
}
// pointPolygonTest(measureDist = false) >= 0 for a float quad
B2A_HD int point_in_quad(const float *poly, float ptx, float pty)
{
    int counter = 0;
    float vx = poly[6], vy = poly[7];
    for (int i = 0; i < 4; ++i) {
        float v0x = vx, v0y = vy;
        vx = poly[2 * i]; vy = poly[2 * i + 1];
        if ((v0y <= pty && vy <= pty) || (v0y > pty && vy > pty) || (v0x < ptx && vx < ptx)) {
            if (pty == vy && (ptx == vx || (pty == v0y && ((v0x <= ptx && ptx <= vx) || (vx <= ptx && ptx <= v0x)))))
                return 0;
            continue;
        }
        // coordinates are integers below 2^13: the products are exact in double
        double dist = (double)(pty - v0y) * (double)(vx - v0x) - (double)(ptx - v0x) * (double)(vy - v0y);
        if (dist == 0) return 0;
        if (vy < v0y) dist = -dist;
        counter += dist > 0;
    }
    return (counter % 2 == 0) ? -1 : 1;
}
B2A_HD bool quad_inside_quad(const float *inner, const float *outer)
{
    return point_in_quad(outer, inner[0], inner[1]) >= 0 && point_in_quad(outer, inner[2], inner[3]) >= 0 &&
           point_in_quad(outer, inner[4], inner[5]) >= 0 && point_in_quad(outer, inner[6], inner[7]) >= 0;
}

// ---------------------------------------------------------------------------------------------
// A7 arithmetic: getPerspectiveTransform (8x8 LU, partial pivoting, double, no FMA contraction),
// closed-form 3x3 inverse, nearest-neighbour source coordinates, Otsu scan.
// ---------------------------------------------------------------------------------------------
B2A_HD double d_mul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
B2A_HD double d_add(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
B2A_HD double d_sub(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
// correctly rounded reciprocal (== 1.0 / a in IEEE arithmetic, cheaper than a general division on the device)
B2A_HD double d_rcp(double a)
{
#if defined(__CUDA_ARCH__)
    return __drcp_rn(a);
#else
    return 1.0 / a;
#endif
}
B2A_HD double d_fma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
B2A_HD double d_div(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}

// src quad -> [0,S-1]^2 ; writes the INVERSE map M (dst -> src) used by the warp
B2A_HD void perspective_inverse(const float *src, int S, double *M)
{
    double A[8][8], b[8];
    const float fs = (float)S - 1.f;
    const float dstx[4] = {0.f, fs, fs, 0.f}, dsty[4] = {0.f, 0.f, fs, fs};
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) A[i][j] = 0.0;
    for (int i = 0; i < 4; ++i) {
        double sx = src[2 * i], sy = src[2 * i + 1], dx = dstx[i], dy = dsty[i];
        A[i][0] = A[i + 4][3] = sx;
        A[i][1] = A[i + 4][4] = sy;
        A[i][2] = A[i + 4][5] = 1.0;
        A[i][6] = d_mul(-sx, dx); A[i][7] = d_mul(-sy, dx);
        A[i + 4][6] = d_mul(-sx, dy); A[i + 4][7] = d_mul(-sy, dy);
        b[i] = dx; b[i + 4] = dy;
    }
    // cv::LU with partial pivoting; compile-time indices only (the row exchange is a predicated swap
    // against every candidate row) so that the 8x8 system stays in registers on the device
    B2A_UNROLL
    for (int i = 0; i < 8; ++i) {
        int k = i;
        double best = fabs(A[i][i]);
    B2A_UNROLL
        for (int j = i + 1; j < 8; ++j) { const double v = fabs(A[j][i]); if (v > best) { best = v; k = j; } }
    B2A_UNROLL
        for (int j = i + 1; j < 8; ++j) {
            const bool sw = (k == j);
    B2A_UNROLL
            for (int c = 0; c < 8; ++c) { const double x = A[i][c], y = A[j][c]; A[i][c] = sw ? y : x; A[j][c] = sw ? x : y; }
            const double x = b[i], y = b[j]; b[i] = sw ? y : x; b[j] = sw ? x : y;
        }
        double d = d_div(-1.0, A[i][i]);
    B2A_UNROLL
        for (int j = i + 1; j < 8; ++j) {
            double alpha = d_mul(A[j][i], d);
    B2A_UNROLL
            for (int kk = i + 1; kk < 8; ++kk) A[j][kk] = d_add(A[j][kk], d_mul(alpha, A[i][kk]));
            b[j] = d_add(b[j], d_mul(alpha, b[i]));
        }
    }
    B2A_UNROLL
    for (int i = 7; i >= 0; --i) {
        double s = b[i];
    B2A_UNROLL
        for (int k = i + 1; k < 8; ++k) s = d_sub(s, d_mul(A[i][k], b[k]));
        b[i] = d_div(s, A[i][i]);
    }
    const double a0 = b[0], a1 = b[1], a2 = b[2], a3 = b[3], a4 = b[4], a5 = b[5], a6 = b[6], a7 = b[7], a8 = 1.0;
    double det = d_add(d_sub(d_mul(a0, d_sub(d_mul(a4, a8), d_mul(a5, a7))), d_mul(a1, d_sub(d_mul(a3, a8), d_mul(a5, a6)))),
                       d_mul(a2, d_sub(d_mul(a3, a7), d_mul(a4, a6))));
    det = d_div(1.0, det);
    M[0] = d_mul(d_sub(d_mul(a4, a8), d_mul(a5, a7)), det);
    M[1] = d_mul(d_sub(d_mul(a2, a7), d_mul(a1, a8)), det);
    M[2] = d_mul(d_sub(d_mul(a1, a5), d_mul(a2, a4)), det);
    M[3] = d_mul(d_sub(d_mul(a5, a6), d_mul(a3, a8)), det);
    M[4] = d_mul(d_sub(d_mul(a0, a8), d_mul(a2, a6)), det);
    M[5] = d_mul(d_sub(d_mul(a2, a3), d_mul(a0, a5)), det);
    M[6] = d_mul(d_sub(d_mul(a3, a7), d_mul(a4, a6)), det);
    M[7] = d_mul(d_sub(d_mul(a1, a6), d_mul(a0, a7)), det);
    M[8] = d_mul(d_sub(d_mul(a0, a4), d_mul(a1, a3)), det);
}
// source pixel of destination (x,y) under INTER_NEAREST (round half to even); returns 0 outside
// offset of the source pixel inside the frame, or -1 when it falls outside (INTER_NEAREST reads 0 there)
B2A_HD long long warp_source(int W, int H, size_t pitch, const double *M, int x, int y)
{
    const double X0 = d_add(d_mul(M[1], (double)y), M[2]);
    const double Y0 = d_add(d_mul(M[4], (double)y), M[5]);
    const double W0 = d_add(d_mul(M[7], (double)y), M[8]);
    double w = d_add(W0, d_mul(M[6], (double)x));
    w = (w != 0.0) ? d_rcp(w) : 0.0;
    double fx = d_mul(d_add(X0, d_mul(M[0], (double)x)), w);
    double fy = d_mul(d_add(Y0, d_mul(M[3], (double)x)), w);
    fx = fx < -2147483648.0 ? -2147483648.0 : (fx > 2147483647.0 ? 2147483647.0 : fx);
    fy = fy < -2147483648.0 ? -2147483648.0 : (fy > 2147483647.0 ? 2147483647.0 : fy);
    const long long X = (long long)rint(fx), Y = (long long)rint(fy);
    return (X >= 0 && X < W && Y >= 0 && Y < H) ? (long long)((size_t)Y * pitch + (size_t)X) : -1;
}
B2A_HD unsigned warp_sample(const uint8_t *__restrict__ gray, int W, int H, size_t pitch, const double *M, int x, int y)
{
    const long long o = warp_source(W, H, pitch, M, x, y);
    return o >= 0 ? gray[o] : 0u;
}
// OpenCV's Otsu between-class-variance scan over a 256-bin histogram of n samples (getThreshVal_Otsu_8u):
//     mu = sum(i h[i]) / n;  per bin:  p = h[i] / n;  mu1 *= q1;  q1 += p;  q2 = 1 - q1;
//     if (min(q1,q2) < FLT_EPSILON || max(q1,q2) > 1 - FLT_EPSILON) continue;
//     mu1 = (mu1 + i p) / q1;  mu2 = (mu - q1 mu1) / q2;  sigma = q1 q2 (mu1 - mu2)^2;  first strict maximum wins
// restated so that the results are bit-identical but the loop-carried part is as short as possible:
//   * bins below the first / above the last non-empty bin cannot change anything (mu1 = q1 = 0 before; every
//     later bin `continue`s), and sum(i h[i]) is an exact integer  ->  only [lo, hi] is walked;
//   * q1 does not depend on mu1: a first one-lane pass leaves the running sums (otsu_prefix);
//   * validity, 1 - q1 and the correctly rounded reciprocal y = RN(1 / q1) are per-bin work for any lane (otsu_bin);
//   * the mu1 recurrence (otsu_chain) then needs no division: with y = RN(1/b), q = RN(a y), r = a - b q (exact, fma),
//     RN(q + r y) is the correctly rounded quotient a / b (Markstein's theorem; no overflow / underflow here);
//   * the variances are again per-bin work (otsu_sigma).
B2A_HD double otsu_mu(long long isum, int n) { return d_mul((double)isum, d_div(1.0, (double)n)); }
B2A_HD void otsu_bin_inputs(int i, int hv, int n, double &p_i, double &ip_i)
{
    p_i = d_mul((double)hv, d_div(1.0, (double)n));
    ip_i = d_mul((double)i, p_i);
}
// in: q1s[i] = p_i for lo <= i <= hi;  out: q1s[i] = q1 after bin i
B2A_HD void otsu_prefix(int lo, int hi, double *q1s)
{
    double q1 = 0;
    for (int i = lo; i <= hi; ++i) { q1 = d_add(q1, q1s[i]); q1s[i] = q1; }
}
// per bin: y = RN(1 / q1), or -1 where OpenCV skips the bin
B2A_HD double otsu_bin(double q1)
{
    const double q2 = d_sub(1.0, q1);
    const double mn = q1 < q2 ? q1 : q2, mx = q1 > q2 ? q1 : q2;
    if (mn < (double)FLT_EPSILON || mx > 1.0 - (double)FLT_EPSILON) return -1.0;
    return d_rcp(q1);
}
// in: q1s = running sums, ys = otsu_bin(q1s), mu1s[i] = i * p_i;  out: mu1s[i] = mu1 after bin i (valid bins)
B2A_HD void otsu_chain(int lo, int hi, const double *q1s, const double *ys, double *mu1s)
{
    double mu1 = 0, q1prev = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
    for (int i = lo; i <= hi; ++i) {
        const double b = q1s[i], y = ys[i];
        const double t = d_mul(mu1, q1prev);
        q1prev = b;
        if (y < 0) { mu1 = t; continue; }
        const double a = d_add(t, mu1s[i]);
        const double q = d_mul(a, y);
        const double r = d_fma(-b, q, a);
        mu1 = d_fma(r, y, q);
        mu1s[i] = mu1;
    }
}
// variance of bin i (0 for skipped bins: never beats max_sigma's initial 0 under strict '>')
B2A_HD double otsu_sigma(double mu, double q1, double y, double mu1)
{
    if (y < 0) return 0.0;
    const double q2 = d_sub(1.0, q1);
    const double mu2 = d_div(d_sub(mu, d_mul(q1, mu1)), q2);
    const double dm = d_sub(mu1, mu2);
    return d_mul(d_mul(d_mul(q1, q2), dm), dm);
}
B2A_HD int otsu_threshold(const int *h, int n)
{
    double q1s[256], ys[256], mu1s[256];
    long long isum = 0;
    int lo = 256, hi = -1;
    for (int i = 0; i < 256; ++i) { ys[i] = -1.0; q1s[i] = 0.0; mu1s[i] = 0.0; isum += (long long)i * h[i]; if (h[i]) { if (lo == 256) lo = i; hi = i; } }
    const double mu = otsu_mu(isum, n);
    for (int i = lo; i <= hi; ++i) otsu_bin_inputs(i, h[i], n, q1s[i], mu1s[i]);
    otsu_prefix(lo, hi, q1s);
    for (int i = lo; i <= hi; ++i) ys[i] = otsu_bin(q1s[i]);
    otsu_chain(lo, hi, q1s, ys, mu1s);
    double max_sigma = 0;
    int max_val = 0;
    for (int i = 0; i < 256; ++i) {
        const double sigma = otsu_sigma(mu, q1s[i], ys[i], mu1s[i]);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    return max_val;
}
// the textbook loop (test reference for the restatement above)
B2A_HD int otsu_threshold_sequential(const int *h, int n)
{
    double mu = 0, scale = d_div(1.0, (double)n);
    for (int i = 0; i < 256; ++i) mu = d_add(mu, d_mul((double)i, (double)h[i]));
    mu = d_mul(mu, scale);
    double mu1 = 0, q1 = 0, max_sigma = 0;
    int max_val = 0;
    for (int i = 0; i < 256; ++i) {
        const double p_i = d_mul((double)h[i], scale);
        mu1 = d_mul(mu1, q1);
        q1 = d_add(q1, p_i);
        const double q2 = d_sub(1.0, q1);
        const double mn = q1 < q2 ? q1 : q2, mx = q1 > q2 ? q1 : q2;
        if (mn < (double)FLT_EPSILON || mx > 1.0 - (double)FLT_EPSILON) continue;
        mu1 = d_div(d_add(mu1, d_mul((double)i, p_i)), q1);
        const double mu2 = d_div(d_sub(mu, d_mul(q1, mu1)), q2);
        const double dm = d_sub(mu1, mu2);
        const double sigma = d_mul(d_mul(d_mul(q1, q2), dm), dm);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    return max_val;
}

// ---- A7 glue shared by k_identify and the host emulation ----
// mode: 0 / 1 = every bit is 0 / 1 (flat patch), 2 = Otsu threshold needed
B2A_HD int ident_mode(long long sum, long long sq, int S, int m0, double minOtsuStdDev)
{
    const int cnt = (S - 2 * m0) * (S - 2 * m0);
    const double scale = d_div(1.0, (double)cnt);
    const double mean = d_mul((double)sum, scale);
    double var = d_sub(d_mul((double)sq, scale), d_mul(mean, mean));
    if (var < 0) var = 0;
    const double sd = sqrt(var);
    if (sd < minOtsuStdDev) return (mean > 127.0) ? 1 : 0;
    return 2;
}
B2A_HD void ident_decide(long long sum, long long sq, int S, int m0, double minOtsuStdDev, const int *hist, int &mode, int &thr)
{
    mode = ident_mode(sum, sq, S, m0, minOtsuStdDev);
    thr = (mode == 2) ? otsu_threshold(hist, S * S) : 0;
}
B2A_HD int ident_cell_bit(const uint8_t *patch, int S, int cellSize, int cellMargin, int cy, int cx, int thr)
{
    const int cw = cellSize - 2 * cellMargin;
    int nz = 0;
    for (int yy = 0; yy < cw; ++yy)
        for (int xx = 0; xx < cw; ++xx)
            nz += patch[(cy * cellSize + cellMargin + yy) * S + cx * cellSize + cellMargin + xx] > thr;
    return nz > (cw * cw) / 2;
}
// one cell of the bit matrix: a border cell adds to the error count, an inner cell to the code word (byte k of the cv2 byte
// list in bits 8k..8k+7); the cells are independent, so the kernel spreads them over the lanes and reduces
B2A_HD void ident_cell_accumulate(int cidx, unsigned bit, int markerSize, int bb, int &err, unsigned long long &code)
{
    const int nb = markerSize + 2 * bb;
    const int cy = cidx / nb, cx = cidx - cy * nb;
    if (cy < bb || cy >= nb - bb || cx < bb || cx >= nb - bb) { err += bit != 0; return; }
    const int ms = markerSize, nbits = ms * ms, nby = (nbits + 7) / 8;
    const int i = (cy - bb) * ms + (cx - bb), byte = i >> 3;
    const int shift = (byte == nby - 1 && (nbits & 7)) ? ((nbits & 7) - 1 - (i & 7)) : (7 - (i & 7));
    code |= (unsigned long long)(bit != 0) << (8 * byte + shift);
}
// detectInvertedMarker (cv2 _identifyOneCandidate): a white marker on black is the same matrix with every bit flipped; the
// flipped matrix is taken when its border has fewer errors.  Flipping all cells = (border cells - err) errors, code ^ all-ones.
B2A_HD void ident_choose_inverted(int markerSize, int bb, int &err, unsigned long long &code)
{
    const int nb = markerSize + 2 * bb, nborder = nb * nb - markerSize * markerSize;
    const int inv = nborder - err;
    if (inv >= err) return;
    err = inv;
    int e2 = 0;
    unsigned long long all = 0;
    for (int c = 0; c < nb * nb; ++c) ident_cell_accumulate(c, 1u, markerSize, bb, e2, all);     // every inner cell's bit position
    code ^= all;
}
// border check + code word; false = border wrong
B2A_HD bool ident_border_code(const uint8_t *bits, int markerSize, int bb, int maxBorderErr, unsigned long long &code, bool detectInverted = false)
{
    const int nb = markerSize + 2 * bb;
    int err = 0;
    code = 0;
    for (int c = 0; c < nb * nb; ++c) ident_cell_accumulate(c, bits[c], markerSize, bb, err, code);
    if (detectInverted) ident_choose_inverted(markerSize, bb, err, code);
    return err <= maxBorderErr;
}
// smallest Hamming distance of marker m over its 4 rotations (first minimum wins)
B2A_HD int ident_marker_distance(const unsigned long long *dict4, unsigned long long code, int markerSize, int &rot)
{
    int best = markerSize * markerSize + 1;
    rot = -1;
    for (int r = 0; r < 4; ++r) { const int hd = popc64(dict4[r] ^ code); if (hd < best) { best = hd; rot = r; } }
    return best;
}

}  // namespace b2a

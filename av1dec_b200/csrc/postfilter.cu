// postfilter.cu -- whole-frame in-loop filter passes: deblocking, CDEF, loop restoration.
//
// Behaviour restated from the reference:
//   deblock  decoder/LoopFilter.cpp:40-370   (edge decisions :85-126, masks :206-289,
//            narrow/wide filters :145-205, level derivation :301-359)
//   CDEF     decoder/Cdef.cpp:41-261         (skip test :72-82, direction :203-261,
//            strengths :84-99, constrained filter :158-198)
//   LR       decoder/LoopRestoration.cpp:49-479 (unit/stripe geometry :49-134, source fetch
//            :234-246, Wiener :247-277, self-guided :353-479)
// Each pass is embarrassingly parallel.
#include "dev.h"
#include "av1_tables.h"
#include "kernels.h"

// ==========================================================================================
// Deblocking
// ==========================================================================================
namespace {

struct LfLevel {
    int lvl, limit, blimit, thresh;
};

AV1B_DEV LfLevel lf_strength(const Av1bLoopFilterParams& lf, const Av1bLfMi& mi, int plane, int pass)
{
    const int i = (plane == 0) ? pass : (plane + 1);
    const int delta_lf = lf.delta_lf_multi ? mi.delta_lf[i] : mi.delta_lf[0];
    const int ref = (mi.flags >> 2) & 7;
    const int mode_type = (mi.flags >> 1) & 1;
    // int8 arithmetic as in the reference (getLvl, LoopFilter.cpp:327-353)
    int8_t lvl_seg = (int8_t)clip3(0, 63, delta_lf + lf.level[i]);
    if (lf.delta_enabled) {
        const int n_shift = lvl_seg >> 5;
        if (ref == 0) lvl_seg = (int8_t)(lvl_seg + (lf.ref_deltas[0] << n_shift));
        else lvl_seg = (int8_t)(lvl_seg + (lf.ref_deltas[ref] << n_shift) + (lf.mode_deltas[mode_type] << n_shift));
        lvl_seg = (int8_t)clip3(0, 63, lvl_seg);
    }
    LfLevel L;
    L.lvl = lvl_seg;
    const int shift = lf.sharpness > 4 ? 2 : (lf.sharpness > 0 ? 1 : 0);
    L.limit = lf.sharpness > 0 ? clip3(1, 9 - lf.sharpness, L.lvl >> shift) : max(1, L.lvl >> shift);
    L.blimit = 2 * (L.lvl + 2) + L.limit;
    L.thresh = L.lvl >> 4;
    return L;
}

AV1B_DEV int f4clamp(int v) { return clip3(-128, 127, v); }

// Filter one sample line across an edge, in registers.  v[k] = sample at position k-8 relative
// to the edge (v[8] = q0, v[7] = p0, ... v[1] = p6, v[14] = q6).  Returns the half-width of the
// modified span n (samples v[8-n .. 8+n-1] changed) or 0.  (LoopFilter.cpp:127-289)
AV1B_DEV int lf_line(int* v, int plane, int limit, int blimit, int thresh, int filter_size)
{
    const int q0 = v[8], q1 = v[9], q2 = v[10], q3 = v[11];
    const int p0 = v[7], p1 = v[6], p2 = v[5], p3 = v[4];
    const int hev = (iabs(p1 - p0) > thresh) | (iabs(q1 - q0) > thresh);
    const int filter_len = (filter_size == 4) ? 4 : (plane ? 6 : (filter_size == 8 ? 8 : 16));
    int mask = (iabs(p1 - p0) > limit) | (iabs(q1 - q0) > limit) | ((iabs(p0 - q0) * 2 + (iabs(p1 - q1) >> 1)) > blimit);
    if (filter_len >= 6) mask |= (iabs(p2 - p1) > limit) | (iabs(q2 - q1) > limit);
    if (filter_len >= 8) mask |= (iabs(p3 - p2) > limit) | (iabs(q3 - q2) > limit);
    if (mask) return 0;
    int flat = 0, flat2 = 0;
    if (filter_size >= 8) {
        int m = (iabs(p1 - p0) > 1) | (iabs(q1 - q0) > 1) | (iabs(p2 - p0) > 1) | (iabs(q2 - q0) > 1);
        if (filter_len >= 8) m |= (iabs(p3 - p0) > 1) | (iabs(q3 - q0) > 1);
        flat = !m;
    }
    if (filter_size >= 16 && flat) {
        const int m = (iabs(v[1] - p0) > 1) | (iabs(v[14] - q0) > 1) | (iabs(v[2] - p0) > 1) | (iabs(v[13] - q0) > 1)
            | (iabs(v[3] - p0) > 1) | (iabs(v[12] - q0) > 1);
        flat2 = !m;
    }
    if (filter_size == 4 || !flat) {
        const int ps0 = p0 - 128, ps1 = p1 - 128, qs0 = q0 - 128, qs1 = q1 - 128;
        int f = hev ? f4clamp(ps1 - qs1) : 0;
        f = f4clamp(f + 3 * (qs0 - ps0));
        const int f1 = f4clamp(f + 4) >> 3, f2 = f4clamp(f + 3) >> 3;
        v[8] = f4clamp(qs0 - f1) + 128;
        v[7] = f4clamp(ps0 + f2) + 128;
        if (hev) return 1;
        const int f3 = (f1 + 1) >> 1;
        v[9] = f4clamp(qs1 - f3) + 128;
        v[6] = f4clamp(ps1 + f3) + 128;
        return 2;
    }
    if (filter_size == 8 || !flat2) {
        if (!plane) {
            // 8-tap luma (n = 3): sum of 7 neighbours (index clamped to p3..q3) + centre again
            // F[i] = (sum_{j=-3..3} v[clamp(i+j)] + v[i]) >> 3, i = -3..2 ; slide the 7-window
            int w = p3 * 3 + p2 + p1 + p0 + q0;              // window for i = -3: positions -6..0 -> clamp(-4)=p3 x3
            const int o0 = (w + p2 + 4) >> 3;                 // i=-3 centre p2 (pos -3)
            w += q1 - p3;                                     // i=-2: positions -5..1
            const int o1 = (w + p1 + 4) >> 3;
            w += q2 - p3;                                     // i=-1: positions -4..2
            const int o2 = (w + p0 + 4) >> 3;
            w += q3 - p3;                                     // i=0: positions -3..3
            const int o3 = (w + q0 + 4) >> 3;
            w += q3 - p2;                                     // i=1: positions -2..4 (clamp 4 -> q3)
            const int o4 = (w + q1 + 4) >> 3;
            w += q3 - p1;                                     // i=2: positions -1..5
            const int o5 = (w + q2 + 4) >> 3;
            v[5] = o0; v[6] = o1; v[7] = o2; v[8] = o3; v[9] = o4; v[10] = o5;
            return 3;
        }
        // 6-tap chroma (n = 2): window of 5 (clamped to p2..q2), weights 2 for |j| <= 1
        const int o0 = (p2 * 3 + p1 * 2 + p0 * 2 + q0 + 4) >> 3;           // i=-2: p2(x1 clamp + x2 w) ...
        const int o1 = (p2 + p1 * 2 + p0 * 2 + q0 * 2 + q1 + 4) >> 3;      // i=-1
        const int o2 = (p1 + p0 * 2 + q0 * 2 + q1 * 2 + q2 + 4) >> 3;      // i=0
        const int o3 = (p0 + q0 * 2 + q1 * 2 + q2 * 3 + 4) >> 3;           // i=1
        v[6] = o0; v[7] = o1; v[8] = o2; v[9] = o3;
        return 2;
    }
    // 14-tap luma (n = 6): F[i] = (sum_{j=-6..6} v[clamp(i+j)] + v[i-1] + v[i] + v[i+1] + 8) >> 4, i = -6..5
    {
        int o[12];
        const int p6 = v[1], q6 = v[14];
        // window sum for i = -6: positions -12..0 clamped to >= -7 (p6): p6 x6 + p5 + p4 + p3 + p2 + p1 + p0 + q0
        int w = p6 * 6 + v[2] + v[3] + p3 + p2 + p1 + p0 + q0;
        (void)q6;
        AV1B_UNROLL
        for (int i = -6; i < 6; i++) {
            // centre extra weights: v[i-1] + v[i] + v[i+1] (positions relative to the edge, index = pos + 8)
            const int c = v[clip3(1, 14, i - 1 + 8)] + v[i + 8] + v[clip3(1, 14, i + 1 + 8)];
            o[i + 6] = (w + c + 8) >> 4;
            // slide: drop position i-6, add position i+7 (both clamped to [-7, 6])
            w += v[clip3(1, 14, i + 7 + 8)] - v[clip3(1, 14, i - 6 + 8)];
        }
        AV1B_UNROLL
        for (int i = 0; i < 12; i++) v[2 + i] = o[i];
        return 6;
    }
}

}  // namespace

// Deblocking, one pass per launch (PASS 0: vertical edges, filter along x; PASS 1: horizontal).
// The work item is a 4-sample EDGE UNIT.  Only units that lie on a transform edge with a non-zero
// level do anything (one in two for 8x8 transforms, one in sixteen for 64x64), so a thread per
// unit would leave most lanes idle through the filter arithmetic.  Each warp therefore first TESTS
// 128 consecutive units (a thread per unit, four rounds: cheap, mostly metadata loads), queues the
// live ones in shared memory, and then FILTERS the queue 32 units at a time with full warps.
namespace {

enum { LF_WARPS = 4, LF_CHUNK = 128 };

struct LfJob {
    uint32_t unit;   // edge unit index inside the plane
    uint32_t params; // limit | blimit << 8 | thresh << 16 | filter_size << 24
};

// Is unit t of `plane` a live edge?  Fills the job.
template <int PASS>
AV1B_DEV bool lf_test(const Av1bFrameHdr* hdr, const Av1bLfMi* mis, const Av1bLoopFilterParams& lf, int plane, int t, LfJob& job)
{
    const int sub = plane ? 1 : 0;
    const int mi_cols = hdr->mi_cols;
    const int ucols = mi_cols >> sub;
    const int ur = t / ucols, uc = t - ur * ucols;
    int row = ur << sub, col = uc << sub;
    const int x = col * 4, y = row * 4;
    if (x >= hdr->frame_w || y >= hdr->frame_h) return false;
    if (PASS == 0 ? (x == 0) : (y == 0)) return false;
    row |= sub;
    col |= sub;
    const int xp = x >> sub, yp = y >> sub;
    const Av1bLfMi mi = mis[row * mi_cols + col];
    const int tx = (mi.tx >> (5 * plane)) & 31;
    // Tx_Width / Block_Width are powers of two: edge tests are masks
    if (PASS == 0 ? (xp & (k_tx_w[tx] - 1)) : (yp & (k_tx_h[tx] - 1))) return false;
    const int bw = max(4, k_block_w[mi.mi_size] >> sub), bh = max(4, k_block_h[mi.mi_size] >> sub);
    const bool skip = mi.flags & 1;
    const bool is_intra = ((mi.flags >> 2) & 7) == 0;
    const bool block_edge = PASS == 0 ? !(xp & (bw - 1)) : !(yp & (bh - 1));
    if (!(block_edge || !skip || is_intra)) return false;
    const int prev_row = row - (PASS == 1 ? (1 << sub) : 0);
    const int prev_col = col - (PASS == 0 ? (1 << sub) : 0);
    const Av1bLfMi pm = mis[prev_row * mi_cols + prev_col];
    const int ptx = (pm.tx >> (5 * plane)) & 31;
    const int base = PASS == 0 ? min(k_tx_w[ptx], k_tx_w[tx]) : min(k_tx_h[ptx], k_tx_h[tx]);
    const int filter_size = plane ? min(8, base) : min(16, base);
    LfLevel L = lf_strength(lf, mi, plane, PASS);
    if (!L.lvl) L = lf_strength(lf, pm, plane, PASS);
    if (L.lvl <= 0) return false;
    job.unit = (uint32_t)t;
    job.params = (uint32_t)L.limit | ((uint32_t)L.blimit << 8) | ((uint32_t)L.thresh << 16) | ((uint32_t)filter_size << 24);
    return true;
}

// Filter the four sample lines of one live edge unit.
template <int PASS>
AV1B_DEV void lf_apply(const PlaneView& pv, int plane, int ucols, const LfJob& job)
{
    const int ur = (int)job.unit / ucols, uc = (int)job.unit - ur * ucols;
    const int xp = uc * 4, yp = ur * 4;
    const int limit = job.params & 0xFF, blimit = (job.params >> 8) & 0xFF, thresh = (job.params >> 16) & 0xFF;
    const int filter_size = job.params >> 24;
    uint8_t* p = pv.p + (size_t)yp * pv.stride + xp;
    if (PASS == 0) {
        // rows yp..yp+3; per row the 16 samples x-8 .. x+7 come in as four aligned words.  All
        // four rows are requested before the first is filtered: the stores below may alias the
        // loads as far as the compiler knows, so it would not hoist them itself.
        uint32_t wa[4][4];
        AV1B_UNROLL
        for (int i = 0; i < 4; i++) {
            const uint8_t* r = p + (size_t)i * pv.stride;
            wa[i][1] = *(const uint32_t*)(r - 4);
            wa[i][2] = *(const uint32_t*)r;
            wa[i][0] = wa[i][3] = 0;
            if (filter_size == 16) {
                wa[i][0] = *(const uint32_t*)(r - 8);
                wa[i][3] = *(const uint32_t*)(r + 4);
            }
        }
        AV1B_UNROLL
        for (int i = 0; i < 4; i++) {
            uint8_t* r = p + (size_t)i * pv.stride;
            int v[16];
            AV1B_UNROLL
            for (int k = 0; k < 4; k++) {
                v[k] = (wa[i][0] >> (8 * k)) & 0xFF;
                v[4 + k] = (wa[i][1] >> (8 * k)) & 0xFF;
                v[8 + k] = (wa[i][2] >> (8 * k)) & 0xFF;
                v[12 + k] = (wa[i][3] >> (8 * k)) & 0xFF;
            }
            const int n = lf_line(v, plane, limit, blimit, thresh, filter_size);
            // byte stores: the neighbouring edge units own the other bytes of these words
            AV1B_UNROLL
            for (int k = 1; k <= 6; k++) {
                if (k <= n) {
                    r[-k] = (uint8_t)v[8 - k];
                    r[k - 1] = (uint8_t)v[7 + k];
                }
            }
        }
    } else {
        // columns xp..xp+3 live in the byte lanes of one word per row; rows yp-8 .. yp+7
        uint32_t w[16];
        const int lo = filter_size == 16 ? 0 : 4, hi = filter_size == 16 ? 16 : 12;
        AV1B_UNROLL
        for (int k = 0; k < 16; k++) w[k] = (k >= lo && k < hi) ? *(const uint32_t*)(p + (ptrdiff_t)(k - 8) * pv.stride) : 0u;
        int nmax = 0;
        AV1B_UNROLL
        for (int cidx = 0; cidx < 4; cidx++) {
            int v[16];
            AV1B_UNROLL
            for (int k = 0; k < 16; k++) v[k] = (w[k] >> (8 * cidx)) & 0xFF;
            const int n = lf_line(v, plane, limit, blimit, thresh, filter_size);
            nmax = max(nmax, n);
            AV1B_UNROLL
            for (int k = 2; k < 14; k++) w[k] = (w[k] & ~(0xFFu << (8 * cidx))) | ((uint32_t)v[k] << (8 * cidx));
        }
        AV1B_UNROLL
        for (int k = 1; k <= 6; k++) {
            if (k <= nmax) {
                *(uint32_t*)(p - (ptrdiff_t)k * pv.stride) = w[8 - k];
                *(uint32_t*)(p + (ptrdiff_t)(k - 1) * pv.stride) = w[7 + k];
            }
        }
    }
}

}  // namespace

template <int PASS> __global__ void __launch_bounds__(LF_WARPS * 32) deblock_kernel(PostCtx c)
{
    __shared__ LfJob queue[LF_WARPS][LF_CHUNK];
    const Av1bFrameHdr* hdr = (const Av1bFrameHdr*)c.cmd;
    const Av1bLfMi* mis = (const Av1bLfMi*)(c.cmd + hdr->off_lfmi);
    const Av1bLoopFilterParams lf = hdr->lf;
    const int plane = blockIdx.z;
    if (plane > 0 && !lf.level[1 + plane]) return;
    const int sub = plane ? 1 : 0;
    const int ucols = hdr->mi_cols >> sub, urows = hdr->mi_rows >> sub; // edge units of this plane
    const int total = ucols * urows;
    const PlaneView pv = c.src.pl[plane];
    const int nl = min(32u, blockDim.x), nw = max(1u, blockDim.x / 32);
    const int lane = threadIdx.x % nl, warp = threadIdx.x / nl;
    LfJob* q = queue[warp];
    const unsigned FULL = 0xFFFFFFFFu;
    for (int t0 = (blockIdx.x * nw + warp) * LF_CHUNK; t0 < total; t0 += gridDim.x * nw * LF_CHUNK) {
        // ---- test LF_CHUNK units, queue the live ones (order inside the queue does not matter:
        // within a pass no two edges touch the same samples)
        int n = 0;
        for (int r = 0; r < LF_CHUNK; r += nl) {
            const int t = t0 + r + lane;
            LfJob job;
            const bool live = t < total && lf_test<PASS>(hdr, mis, lf, plane, t, job);
            const unsigned m = __ballot_sync(FULL, live);
            if (live) q[n + __popc(m & ((1u << lane) - 1))] = job;
            n += __popc(m);
        }
        __syncwarp();
        // ---- filter them, a full warp at a time
        for (int k = lane; k < n; k += nl) lf_apply<PASS>(pv, plane, ucols, q[k]);
        __syncwarp();
    }
}

// ==========================================================================================
// CDEF
// ==========================================================================================
// One CTA filters a 64x64 luma area (8x8 CDEF blocks) and the matching 32x32 chroma areas.
//   1. the tile + 2-sample halo of each plane is staged in shared memory as 16-bit samples;
//      samples outside the MI-aligned frame become CDEF_LARGE so that their constrained
//      difference is 0 and they never win the min/max (they are "unavailable", Cdef.cpp:140-156)
//   2. direction search: one thread per (block, direction), bins in registers
//   3. filter: one warp per 8x8 block, one lane per horizontal sample PAIR, all arithmetic on
//      packed 16x2 lanes (VIMNMX.U16x2 / VIADD) -- the stage is issue-bound, not HBM-bound
namespace {

enum {
    CDEF_LARGE = 0x4000,
    CY_PITCH = 72,  // halfwords per luma tile row (36 words: conflict-free for 8 rows x 4 words)
    CY_ROWS = 68,
    CC_PITCH = 40,  // halfwords per chroma tile row
    CC_ROWS = 36,
};

struct CdefBlk {
    uint8_t idx;      // preset or 0xFF
    uint8_t pri[2];   // [0] luma (variance adjusted), [1] chroma
    uint8_t sec[2];
    uint8_t adjp[2];  // damping adjustment shifts
    uint8_t adjs[2];
    uint8_t dir[2];
    uint8_t pad;
};

struct CdefShared {
    uint16_t ya[CY_ROWS * CY_PITCH];      // luma tile, ya[r*P + c] = sample (x0 - 2 + c, y0 - 2 + r)
    uint16_t yb[CY_ROWS * CY_PITCH];      // same, shifted left by one sample (odd tap offsets stay word aligned)
    uint16_t ca[2][CC_ROWS * CC_PITCH];
    uint16_t cb[2][CC_ROWS * CC_PITCH];
    int cost[64][8];
    CdefBlk blk[64];
};

template <int D> AV1B_DEV constexpr int cdef_bin(int i, int j)
{
    return D == 0 ? i + j : D == 1 ? i + j / 2 : D == 2 ? i : D == 3 ? 3 + i - j / 2 : D == 4 ? 7 + i - j : D == 5 ? 3 - i / 2 + j
        : D == 6 ? j : i / 2 + j;
}

// cost of direction D for the 8x8 block whose top-left sample is at `blk` (reference cdefDirection)
template <int D> AV1B_DEV int cdef_cost(const uint16_t* blk)
{
    int part[15];
    AV1B_UNROLL
    for (int k = 0; k < 15; k++) part[k] = 0;
    AV1B_UNROLL
    for (int i = 0; i < 8; i++) {
        AV1B_UNROLL
        for (int j = 0; j < 8; j++) part[cdef_bin<D>(i, j)] += (int)blk[i * CY_PITCH + j] - 128;
    }
    int cost = 0;
    if (D == 2 || D == 6) {
        AV1B_UNROLL
        for (int k = 0; k < 8; k++) cost += part[k] * part[k];
        cost *= 105;
    } else if (D == 0 || D == 4) {
        AV1B_UNROLL
        for (int k = 0; k < 7; k++) cost += (part[k] * part[k] + part[14 - k] * part[14 - k]) * k_cdef_div_table[k + 1];
        cost += part[7] * part[7] * 105;
    } else {
        AV1B_UNROLL
        for (int k = 0; k < 5; k++) cost += part[3 + k] * part[3 + k];
        cost *= 105;
        AV1B_UNROLL
        for (int k = 0; k < 3; k++) cost += (part[k] * part[k] + part[10 - k] * part[10 - k]) * k_cdef_div_table[2 * k + 2];
    }
    return cost;
}

AV1B_DEV int cdef_cost_dyn(int d, const uint16_t* blk)
{
    switch (d) {
    case 0: return cdef_cost<0>(blk);
    case 1: return cdef_cost<1>(blk);
    case 2: return cdef_cost<2>(blk);
    case 3: return cdef_cost<3>(blk);
    case 4: return cdef_cost<4>(blk);
    case 5: return cdef_cost<5>(blk);
    case 6: return cdef_cost<6>(blk);
    default: return cdef_cost<7>(blk);
    }
}

// One constrained tap on two samples at once.  x2/p2: centre / tap sample pairs (16x2).
AV1B_DEV void cdef_tap(uint32_t p2, uint32_t x2, uint32_t thr2, int adj, uint32_t amask, uint32_t w, uint32_t emask, uint32_t& T,
    uint32_t& P, uint32_t& mx, uint32_t& mn)
{
    const uint32_t hi = __vmaxu2(p2, x2), lo = __vminu2(p2, x2);
    const uint32_t a = hi - lo;                       // |p - x| per half
    const uint32_t s = (a >> adj) & amask;            // |d| >> dampingAdj
    const uint32_t t = __vmaxu2(thr2, s) - s;         // max(0, thr - s)
    const uint32_t c = __vminu2(a, t);                // constrained magnitude
    const uint32_t cp = __vminu2(c, hi - x2);         // ... of the positive differences only
    T += w * c;
    P += w * cp;
    // CDEF_LARGE & 0xFF == 0: an unavailable sample never wins the max
    mx = __vmaxu2(mx, p2 & emask);
    mn = __vminu2(mn, lo);
}

// Filter the sample pair at (even) halfword index `ctr`.  `offs` holds the twelve tap offsets
// of the block (halfwords relative to the pair; already redirected into the shifted tile copy
// for odd displacements, so every load is an aligned 32-bit LDS):
//   [0..3] primary k=0 +/-, k=1 +/-   [4..7] secondary (dir+2) k=0 +/-, k=1 +/-   [8..11] (dir-2)
// Returns the two output bytes.  (reference cdefFilter, Cdef.cpp:158-198)
AV1B_DEV uint32_t cdef_filter_pair(const uint16_t* tile, int ctr, const int* offs, int pri, int sec, int adjp, int adjs)
{
    const uint16_t* q = tile + ctr;
    const uint32_t x2 = *(const uint32_t*)q;
    uint32_t T = 0, P = 0, mx = x2, mn = x2;
    const uint32_t pri2 = (uint32_t)pri * 0x00010001u, sec2 = (uint32_t)sec * 0x00010001u;
    const uint32_t maskp = (0xFFFFu >> adjp) * 0x00010001u, masks = (0xFFFFu >> adjs) * 0x00010001u;
    const uint32_t wp0 = (pri & 1) ? 3u : 4u, wp1 = (pri & 1) ? 3u : 2u;
    const uint32_t emask = 0x00FF00FFu;
    AV1B_UNROLL
    for (int k = 0; k < 4; k++)
        cdef_tap(*(const uint32_t*)(q + offs[k]), x2, pri2, adjp, maskp, k < 2 ? wp0 : wp1, emask, T, P, mx, mn);
    AV1B_UNROLL
    for (int k = 4; k < 12; k++)
        cdef_tap(*(const uint32_t*)(q + offs[k]), x2, sec2, adjs, masks, (k & 2) ? 1u : 2u, emask, T, P, mx, mn);
    // y = clip3(min, max, x + sign(sum) * ((|sum| + 8) >> 4)), sum = P - N, both halves at once
    // ((8 + sum - (sum < 0)) >> 4 rounds half away from zero)
    const uint32_t N = T - P;
    const uint32_t big = __vmaxu2(P, N);
    const uint32_t mpos = (((big - N) + 0x00080008u) >> 4) & 0x0FFF0FFFu;
    const uint32_t mneg = (((big - P) + 0x00080008u) >> 4) & 0x0FFF0FFFu;
    uint32_t y = __vminu2(x2 + mpos, mx);
    y = __vmaxu2(y, mneg + mn) - mneg;
    return (y & 0xFFu) | ((y >> 8) & 0xFF00u);
}

// The twelve redirected tap offsets for direction `dir` (see cdef_filter_pair).  `copy` is the
// halfword distance from the tile to its shifted copy.
AV1B_DEV void cdef_offsets(int dir, int pitch, int copy, int* offs)
{
    AV1B_UNROLL
    for (int g = 0; g < 3; g++) {
        const int d = g == 0 ? dir : (g == 1 ? ((dir + 2) & 7) : ((dir + 6) & 7));
        AV1B_UNROLL
        for (int k = 0; k < 2; k++) {
            const int o = k_cdef_directions[d][k][0] * pitch + k_cdef_directions[d][k][1];
            const int redirect = (o & 1) ? copy - 1 : 0;
            offs[g * 4 + k * 2] = o + redirect;
            offs[g * 4 + k * 2 + 1] = -o + redirect;
        }
    }
}

// Stage rows [y0-2, y0-2+rows) x tile columns [0, 4*words-2) of a plane into the 16-bit tiles
// (tile column c <-> frame column x0 - 2 + c).  Each work item converts four samples from two
// aligned 32-bit loads into two words of the tile and two words of the shifted copy.
// INTERIOR: the whole staged area lies inside the MI-aligned frame (no availability tests).
template <bool INTERIOR>
AV1B_DEV void cdef_stage(const PlaneView& src, int x0, int y0, int pw, int ph, int rows, int words, int pitch, uint16_t* ta,
    uint16_t* tb, int tid, int nt)
{
    // item (r, wi): frame columns xb = x0 - 2 + 4*wi .. xb+3 (+ xb+4 for the shifted copy); x0 - 2 is
    // 2 (mod 4), so the samples straddle two aligned words
    const unsigned magic = (unsigned)((0x100000000ull + words - 1) / words);
    for (int e = tid; e < rows * words; e += nt) {
        const int r = (int)__umulhi((unsigned)e, magic), wi = e - r * words;
        const int y = y0 - 2 + r, xa = x0 - 4 + wi * 4; // aligned word holding columns xa..xa+3
        const uint8_t* rowp = src.p + (ptrdiff_t)y * src.stride + xa;
        uint32_t lo = __ldg((const uint32_t*)rowp), hi = __ldg((const uint32_t*)rowp + 1);
        // samples s0..s4 = frame columns xa+2 .. xa+6
        uint32_t s01, s23, s12, s34;
        if (INTERIOR) {
            s01 = __byte_perm(lo, 0, 0x4342); // (lo.b2, 0, lo.b3, 0)
            s23 = __byte_perm(hi, 0, 0x4140); // (hi.b0, 0, hi.b1, 0)
            s12 = (lo >> 24) | ((hi & 0xFFu) << 16); // (lo.b3, 0, hi.b0, 0)
            s34 = __byte_perm(hi, 0, 0x4241); // (hi.b1, 0, hi.b2, 0)
        } else {
            const bool yok = y >= 0 && y < ph;
            uint32_t v[5];
            AV1B_UNROLL
            for (int k = 0; k < 5; k++) {
                const int x = xa + 2 + k;
                const uint32_t byte = k < 2 ? ((lo >> (16 + 8 * k)) & 0xFF) : ((hi >> (8 * (k - 2))) & 0xFF);
                v[k] = (yok && x >= 0 && x < pw) ? byte : (uint32_t)CDEF_LARGE;
            }
            s01 = v[0] | (v[1] << 16);
            s23 = v[2] | (v[3] << 16);
            s12 = v[1] | (v[2] << 16);
            s34 = v[3] | (v[4] << 16);
        }
        uint32_t* da = (uint32_t*)(ta + r * pitch + wi * 4);
        uint32_t* db = (uint32_t*)(tb + r * pitch + wi * 4);
        da[0] = s01;
        da[1] = s23;
        db[0] = s12;
        db[1] = s34;
    }
}

}  // namespace

__global__ void __launch_bounds__(256) cdef_kernel(PostCtx c)
{
    __shared__ CdefShared S;
    const Av1bFrameHdr* hdr = (const Av1bFrameHdr*)c.cmd;
    const uint8_t* cdef8 = c.cmd + hdr->off_cdef8;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int c8 = hdr->mi_cols >> 1, r8 = hdr->mi_rows >> 1; // 8x8 blocks in the frame
    const int fbx = blockIdx.x * 8, fby = blockIdx.y * 8;      // first 8x8 block of this CTA
    const int pw = hdr->mi_cols * 4, ph = hdr->mi_rows * 4;
    // ---- 1. stage (17 / 9 items per row: 68 / 36 tile columns)
    const bool interior = fbx > 0 && fby > 0 && fbx * 8 + 68 <= pw && fby * 8 + 66 <= ph;
    if (interior) {
        cdef_stage<true>(c.src.pl[0], fbx * 8, fby * 8, pw, ph, CY_ROWS, 17, CY_PITCH, S.ya, S.yb, tid, nt);
        cdef_stage<true>(c.src.pl[1], fbx * 4, fby * 4, pw >> 1, ph >> 1, CC_ROWS, 9, CC_PITCH, S.ca[0], S.cb[0], tid, nt);
        cdef_stage<true>(c.src.pl[2], fbx * 4, fby * 4, pw >> 1, ph >> 1, CC_ROWS, 9, CC_PITCH, S.ca[1], S.cb[1], tid, nt);
    } else {
        cdef_stage<false>(c.src.pl[0], fbx * 8, fby * 8, pw, ph, CY_ROWS, 17, CY_PITCH, S.ya, S.yb, tid, nt);
        cdef_stage<false>(c.src.pl[1], fbx * 4, fby * 4, pw >> 1, ph >> 1, CC_ROWS, 9, CC_PITCH, S.ca[0], S.cb[0], tid, nt);
        cdef_stage<false>(c.src.pl[2], fbx * 4, fby * 4, pw >> 1, ph >> 1, CC_ROWS, 9, CC_PITCH, S.ca[1], S.cb[1], tid, nt);
    }
    for (int e = tid; e < 64; e += nt) {
        const int by = fby + (e >> 3), bx = fbx + (e & 7);
        S.blk[e].idx = (by < r8 && bx < c8) ? cdef8[by * c8 + bx] : 0xFF;
    }
    __syncthreads();
    // ---- 2. direction search
    for (int e = tid; e < 512; e += nt) {
        const int b = e & 63, d = e >> 6; // 64 consecutive threads share a direction: no divergence inside a warp
        if (S.blk[b].idx == 0xFF) continue;
        S.cost[b][d] = cdef_cost_dyn(d, S.ya + ((b >> 3) * 8 + 2) * CY_PITCH + (b & 7) * 8 + 2);
    }
    __syncthreads();
    const Av1bCdefParams& cp = hdr->cdef;
    for (int e = tid; e < 64; e += nt) {
        CdefBlk& B = S.blk[e];
        if (B.idx == 0xFF) continue;
        int best = 0, dir = 0;
        for (int d = 0; d < 8; d++)
            if (S.cost[e][d] > best) {
                best = S.cost[e][d];
                dir = d;
            }
        const int var = (best - S.cost[e][(dir + 4) & 7]) >> 10;
        int pri = cp.y_pri[B.idx];
        const int dir_y = pri == 0 ? 0 : dir;
        const int var_str = (var >> 6) ? min(floor_log2((unsigned)(var >> 6)), 12) : 0;
        pri = var ? ((pri * (4 + var_str) + 8) >> 4) : 0;
        const int sec = cp.y_sec[B.idx];
        const int pri_uv = cp.uv_pri[B.idx], sec_uv = cp.uv_sec[B.idx];
        const int damp = cp.damping;
        B.pri[0] = (uint8_t)pri;
        B.sec[0] = (uint8_t)sec;
        B.adjp[0] = (uint8_t)(pri ? max(0, damp - floor_log2((unsigned)pri)) : 0);
        B.adjs[0] = (uint8_t)(sec ? max(0, damp - floor_log2((unsigned)sec)) : 0);
        B.dir[0] = (uint8_t)dir_y;
        B.pri[1] = (uint8_t)pri_uv;
        B.sec[1] = (uint8_t)sec_uv;
        B.adjp[1] = (uint8_t)(pri_uv ? max(0, damp - 1 - floor_log2((unsigned)pri_uv)) : 0);
        B.adjs[1] = (uint8_t)(sec_uv ? max(0, damp - 1 - floor_log2((unsigned)sec_uv)) : 0);
        B.dir[1] = (uint8_t)(pri_uv == 0 ? 0 : k_cdef_uv_dir[1][1][dir]);
    }
    __syncthreads();
    // ---- 3. filter.  Luma: 16 lanes per 8x8 block (two sample pairs each); chroma: 8 lanes per
    //         4x4 block.  Tap offsets are computed once per lane and block.
    const int nl = min(32u, blockDim.x), nw = max(1u, blockDim.x / 32);
    const int lane = tid % nl, warp = tid / nl;
    {
        const int sub = min(16, nl), per = max(1, nl / sub); // lanes per block, blocks per warp pass
        const int sl = lane % sub, sg = lane / sub;
        const int copy = (int)(S.yb - S.ya);
        for (int b = warp * per + sg; b < 64; b += nw * per) {
            const CdefBlk B = S.blk[b];
            const int bx = (fbx + (b & 7)) * 8, by = (fby + (b >> 3)) * 8;
            if (bx >= pw || by >= ph) continue;
            uint8_t* dst = c.cdef.pl[0].p + (size_t)by * c.cdef.pl[0].stride + bx;
            const bool active = B.idx != 0xFF && (B.pri[0] | B.sec[0]);
            int offs[12];
            if (active) cdef_offsets(B.dir[0], CY_PITCH, copy, offs);
            for (int pr = sl; pr < 32; pr += sub) {
                const int r = pr >> 2, cpair = pr & 3;
                const int ctr = ((b >> 3) * 8 + r + 2) * CY_PITCH + (b & 7) * 8 + cpair * 2 + 2;
                uint32_t out;
                if (active) out = cdef_filter_pair(S.ya, ctr, offs, B.pri[0], B.sec[0], B.adjp[0], B.adjs[0]);
                else out = (uint32_t)S.ya[ctr] | ((uint32_t)S.ya[ctr + 1] << 8);
                *(uint16_t*)(dst + (size_t)r * c.cdef.pl[0].stride + cpair * 2) = (uint16_t)out;
            }
        }
    }
    {
        const int cpw = pw >> 1, cph = ph >> 1;
        const int sub = min(8, nl), per = max(1, nl / sub);
        const int sl = lane % sub, sg = lane / sub;
        for (int it = warp * per + sg; it < 128; it += nw * per) {
            const int plane = 1 + (it >> 6), b = it & 63;
            const CdefBlk B = S.blk[b];
            const int bx = (fbx + (b & 7)) * 4, by = (fby + (b >> 3)) * 4;
            if (bx >= cpw || by >= cph) continue;
            const uint16_t* ta = S.ca[plane - 1];
            const int copy = (int)(S.cb[plane - 1] - ta);
            const PlaneView dv = c.cdef.pl[plane];
            const bool active = B.idx != 0xFF && (B.pri[1] | B.sec[1]);
            int offs[12];
            if (active) cdef_offsets(B.dir[1], CC_PITCH, copy, offs);
            for (int pr = sl; pr < 8; pr += sub) {
                const int r = pr >> 1, cpair = pr & 1;
                const int ctr = ((b >> 3) * 4 + r + 2) * CC_PITCH + (b & 7) * 4 + cpair * 2 + 2;
                uint32_t out;
                if (active) out = cdef_filter_pair(ta, ctr, offs, B.pri[1], B.sec[1], B.adjp[1], B.adjs[1]);
                else out = (uint32_t)ta[ctr] | ((uint32_t)ta[ctr + 1] << 8);
                *(uint16_t*)(dv.p + (size_t)(by + r) * dv.stride + bx + cpair * 2) = (uint16_t)out;
            }
        }
    }
}

// ==========================================================================================
// Loop restoration
// ==========================================================================================
namespace {

enum { LR_TW = 32, LR_MAXH = 64, LR_SW = 40, LR_SH = LR_MAXH + 6, LR_AW = LR_TW + 2, LR_AH = LR_MAXH + 2 };

struct LrShared {
    uint8_t src[LR_SH * LR_SW];        // source samples, 3-sample halo; src[r*LR_SW + c] = sample (x0-3+c-1, y0-3+r)
    union {
        int16_t wien[LR_SH * LR_TW];   // Wiener horizontal pass
        struct {
            uint16_t h1[LR_SH * LR_AW]; // horizontal box sums of x
            uint32_t h2[LR_SH * LR_AW]; // horizontal box sums of x^2
        } box;
    };
    uint16_t a[LR_AH * LR_AW];         // SGR A (a2)
    uint32_t b[LR_AH * LR_AW];         // SGR B (b2)
    uint16_t flt[2][LR_MAXH * LR_TW];  // SGR filtered planes
    uint16_t xdiv[256];                // ((z << 8) + z/2) / (z + 1)
};

// Row of the frame that get_source_sample() reads for tile row `y` (LoopRestoration.cpp:234-246
// + extendBorder, VideoFrame.cpp:81-101); *from_deblocked tells which frame.
AV1B_DEV int lr_source_row(int y, int start, int end, int ph, bool* from_deblocked)
{
    *from_deblocked = false;
    if (y < start) {
        y = max(start - 2, y);
        *from_deblocked = true;
    } else if (y >= end) {
        y = min(end + 1, y);
        *from_deblocked = true;
    }
    return clip3(0, ph - 1, y);
}

// 2-D decomposition of the CTA's threads: tx walks columns (up to 32 wide), ty walks rows.
struct Lane2D {
    int tx, ty, ntx, nty;
};
AV1B_DEV Lane2D lane2d(int tid, int nt)
{
    Lane2D l;
    l.ntx = min(nt, 32);
    l.nty = max(1, nt / l.ntx);
    l.tx = tid % l.ntx;
    l.ty = tid / l.ntx;
    return l;
}

// One self-guided pass.  Source samples sit at S.src[(i + 3) * LR_SW + (j + 4)] for tile sample (i, j).
AV1B_DEV void sgr_pass(LrShared& S, int w, int h, int set, int pass, int r, const Lane2D& L, int nt)
{
    const int eps = k_sgr_params[set][pass * 2 + 1];
    const int n = (2 * r + 1) * (2 * r + 1);
    const int n2e = n * n * eps;
    const unsigned s = (unsigned)(((1 << 20) + n2e / 2) / n2e);
    const int one_over_n = ((1 << 12) + (n / 2)) / n;
    const int aw = w + 2;
    // horizontal box sums for rows -1-r .. h+r, columns -1 .. w
    const int hr0 = -1 - r, hrows = h + 2 + 2 * r;
    const unsigned aw_magic = (unsigned)((0x100000000ull + aw - 1) / aw); // e / aw == umulhi(e, magic) for e < 2^16
    for (int e = L.tx + L.ty * L.ntx; e < hrows * aw; e += L.ntx * L.nty) {
        {
            const int rr = (int)__umulhi((unsigned)e, aw_magic), jj = e - rr * aw; // jj = j + 1
            const uint8_t* p = S.src + (hr0 + rr + 3) * LR_SW + (jj + 3);
            int s1, s2;
            if (r == 2) {
                const int v0 = p[-2], v1 = p[-1], v2 = p[0], v3 = p[1], v4 = p[2];
                s1 = v0 + v1 + v2 + v3 + v4;
                s2 = v0 * v0 + v1 * v1 + v2 * v2 + v3 * v3 + v4 * v4;
            } else {
                const int v1 = p[-1], v2 = p[0], v3 = p[1];
                s1 = v1 + v2 + v3;
                s2 = v1 * v1 + v2 * v2 + v3 * v3;
            }
            S.box.h1[rr * LR_AW + jj] = (uint16_t)s1;
            S.box.h2[rr * LR_AW + jj] = (uint32_t)s2;
        }
    }
    __syncthreads();
    // vertical sums -> a2 / b2.  Pass 0 only ever reads A/B on rows whose index is odd.
    const int istep = pass == 0 ? 2 : 1;
    const int nrows = pass == 0 ? (h + 3) / 2 : h + 2;
    for (int e = L.tx + L.ty * L.ntx; e < nrows * aw; e += L.ntx * L.nty) {
        {
            const int ri = (int)__umulhi((unsigned)e, aw_magic), jj = e - ri * aw;
            const int i = -1 + ri * istep;
            const int base = (i - r - hr0) * LR_AW + jj;
            int a, b;
            if (r == 2) {
                b = S.box.h1[base] + S.box.h1[base + LR_AW] + S.box.h1[base + 2 * LR_AW] + S.box.h1[base + 3 * LR_AW] + S.box.h1[base + 4 * LR_AW];
                a = S.box.h2[base] + S.box.h2[base + LR_AW] + S.box.h2[base + 2 * LR_AW] + S.box.h2[base + 3 * LR_AW] + S.box.h2[base + 4 * LR_AW];
            } else {
                b = S.box.h1[base] + S.box.h1[base + LR_AW] + S.box.h1[base + 2 * LR_AW];
                a = S.box.h2[base] + S.box.h2[base + LR_AW] + S.box.h2[base + 2 * LR_AW];
            }
            const unsigned p = (unsigned)max(0, a * n - b * b);
            const unsigned z = (p * s + (1u << 19)) >> 20;
            const int a2 = z >= 255 ? 256 : (z == 0 ? 1 : S.xdiv[z]);
            const int b2 = (256 - a2) * b * one_over_n;
            S.a[(i + 1) * LR_AW + jj] = (uint16_t)a2;
            S.b[(i + 1) * LR_AW + jj] = (uint32_t)((b2 + (1 << 11)) >> 12);
        }
    }
    __syncthreads();
    for (int i = L.ty; i < h; i += L.nty) {
        for (int j = L.tx; j < w; j += L.ntx) {
            const uint16_t* A = S.a + (i + 1) * LR_AW + (j + 1);
            const uint32_t* B = S.b + (i + 1) * LR_AW + (j + 1);
            int a, b, shift;
            if (pass == 0) {
                if (i & 1) {
                    a = 6 * A[0] + 5 * (A[-1] + A[1]);
                    b = 6 * (int)B[0] + 5 * (int)(B[-1] + B[1]);
                    shift = 4;
                } else {
                    a = 6 * (A[-LR_AW] + A[LR_AW]) + 5 * (A[-LR_AW - 1] + A[-LR_AW + 1] + A[LR_AW - 1] + A[LR_AW + 1]);
                    b = 6 * (int)(B[-LR_AW] + B[LR_AW]) + 5 * (int)(B[-LR_AW - 1] + B[-LR_AW + 1] + B[LR_AW - 1] + B[LR_AW + 1]);
                    shift = 5;
                }
            } else {
                a = 4 * (A[0] + A[-1] + A[1] + A[-LR_AW] + A[LR_AW]) + 3 * (A[-LR_AW - 1] + A[-LR_AW + 1] + A[LR_AW - 1] + A[LR_AW + 1]);
                b = 4 * (int)(B[0] + B[-1] + B[1] + B[-LR_AW] + B[LR_AW]) + 3 * (int)(B[-LR_AW - 1] + B[-LR_AW + 1] + B[LR_AW - 1] + B[LR_AW + 1]);
                shift = 5;
            }
            const int v = a * S.src[(i + 3) * LR_SW + (j + 4)] + b;
            S.flt[pass][i * LR_TW + j] = (uint16_t)round2(v, 8 + shift - 4);
        }
    }
    __syncthreads();
}

}  // namespace

// grid: (tiles_x, stripes, plane). One CTA = 32 columns x one 64-luma-row stripe of one plane.
__global__ void __launch_bounds__(256) lr_kernel(PostCtx c)
{
    __shared__ LrShared S;
    const Av1bFrameHdr* hdr = (const Av1bFrameHdr*)c.cmd;
    const Av1bLrParams lp = hdr->lr;
    const int plane = blockIdx.z, sub = plane ? 1 : 0;
    const int pw = (hdr->frame_w + sub) >> sub, ph = (hdr->frame_h + sub) >> sub;
    const int tid = threadIdx.x, nt = blockDim.x;
    const Lane2D L = lane2d(tid, nt);
    const int x0 = blockIdx.x * LR_TW;
    const int start = (-8 + (int)blockIdx.y * 64) >> sub, end = start + (64 >> sub);
    const int y0 = max(start, 0), y1 = min(end, ph);
    if (x0 >= pw || y0 >= y1) return;
    const int w = min((int)LR_TW, pw - x0), h = y1 - y0;
    const PlaneView cdef = c.cdef.pl[plane], deb = c.src.pl[plane], out = c.lr.pl[plane];
    int type = 0;
    Av1bLrUnit unit;
    if (lp.frame_type[plane]) {
        const int us = lp.unit_size[plane];
        const int urow = min((int)lp.unit_rows[plane] - 1, (y0 + (8 >> sub)) / us);
        const int ucol = min((int)lp.unit_cols[plane] - 1, x0 / us);
        unit = ((const Av1bLrUnit*)(c.cmd + hdr->off_lru))[lp.unit_first[plane] + urow * lp.unit_cols[plane] + ucol];
        type = unit.type;
    }
    if (type == 0) {
        // RESTORE_NONE: the LR frame is a copy of the CDEF frame.  Word copies, byte tail.
        const int ww = w >> 2;
        for (int i = L.ty; i < h; i += L.nty) {
            const uint8_t* srow = cdef.p + (size_t)(y0 + i) * cdef.stride + x0;
            uint8_t* drow = out.p + (size_t)(y0 + i) * out.stride + x0;
            for (int j = L.tx; j < ww; j += L.ntx) ((uint32_t*)drow)[j] = __ldg((const uint32_t*)srow + j);
            for (int j = (w & ~3) + L.tx; j < w; j += L.ntx) drow[j] = __ldg(srow + j);
        }
        return;
    }
    // ---- stage source: rows y0-3 .. y0+h+2, columns x0-4 .. x0+35 (10 aligned words per row)
    {
        const bool interior = x0 >= 4 && x0 + 36 <= pw;
        for (int r = L.ty; r < h + 6; r += L.nty) {
            bool fd;
            const int sy = lr_source_row(y0 - 3 + r, start, end, ph, &fd);
            const uint8_t* rowp = (fd ? deb.p : cdef.p) + (size_t)sy * (fd ? deb.stride : cdef.stride);
            for (int wi = L.tx; wi < 10; wi += L.ntx) {
                uint32_t v;
                if (interior) {
                    v = __ldg((const uint32_t*)(rowp + x0 - 4 + wi * 4));
                } else {
                    v = 0;
                    AV1B_UNROLL
                    for (int k = 0; k < 4; k++) v |= (uint32_t)__ldg(rowp + clip3(0, pw - 1, x0 - 4 + wi * 4 + k)) << (8 * k);
                }
                *(uint32_t*)(S.src + r * LR_SW + wi * 4) = v;
            }
        }
    }
    if (type == 2) {
        for (int z = tid; z < 256; z += nt) S.xdiv[z] = (uint16_t)(((z << 8) + (z >> 1)) / (z + 1));
    }
    __syncthreads();
    if (type == 1) {
        int vf[4], hf[4];
        vf[3] = 128;
        hf[3] = 128;
        for (int k = 0; k < 3; k++) {
            vf[k] = unit.wiener[0][k];
            hf[k] = unit.wiener[1][k];
            vf[3] -= 2 * unit.wiener[0][k];
            hf[3] -= 2 * unit.wiener[1][k];
        }
        for (int r = L.ty; r < h + 6; r += L.nty) {
            for (int cc = L.tx; cc < w; cc += L.ntx) {
                const uint8_t* p = S.src + r * LR_SW + cc + 1; // sample (x0 + cc - 3) sits at column cc + 1
                const int s = hf[0] * (p[0] + p[6]) + hf[1] * (p[1] + p[5]) + hf[2] * (p[2] + p[4]) + hf[3] * p[3];
                S.wien[r * LR_TW + cc] = (int16_t)clip3(-2048, 6143, (s + 4) >> 3);
            }
        }
        __syncthreads();
        for (int r = L.ty; r < h; r += L.nty) {
            for (int cc = L.tx; cc < w; cc += L.ntx) {
                const int16_t* q = S.wien + r * LR_TW + cc;
                const int s = vf[0] * (q[0] + q[6 * LR_TW]) + vf[1] * (q[LR_TW] + q[5 * LR_TW]) + vf[2] * (q[2 * LR_TW] + q[4 * LR_TW])
                    + vf[3] * q[3 * LR_TW];
                out.p[(size_t)(y0 + r) * out.stride + x0 + cc] = (uint8_t)clip_u8((s + 1024) >> 11);
            }
        }
    } else {
        const int set = unit.sgr_set;
        const int r0 = k_sgr_params[set][0], r1 = k_sgr_params[set][2];
        if (r0) sgr_pass(S, w, h, set, 0, r0, L, nt);
        if (r1) sgr_pass(S, w, h, set, 1, r1, L, nt);
        const int w0 = unit.sgr_xqd[0], w1 = unit.sgr_xqd[1], w2 = 128 - w0 - w1;
        for (int i = L.ty; i < h; i += L.nty) {
            for (int j = L.tx; j < w; j += L.ntx) {
                const int u = S.src[(i + 3) * LR_SW + (j + 4)] << 4;
                int v = w1 * u;
                v += w0 * (r0 ? (int)S.flt[0][i * LR_TW + j] : u);
                v += w2 * (r1 ? (int)S.flt[1][i * LR_TW + j] : u);
                out.p[(size_t)(y0 + i) * out.stride + x0 + j] = (uint8_t)clip_u8(round2(v, 11));
            }
        }
    }
}

// ==========================================================================================
// launchers
// ==========================================================================================
void launch_deblock(const PostCtx& c, const Av1bFrameHdr& h, av1b_stream_t st)
{
    if (!h.lf.level[0] && !h.lf.level[1]) return;
    const long long total = (long long)h.mi_cols * h.mi_rows;
    const int per_cta = 4 * 128; // LF_WARPS warps x LF_CHUNK units
    int grid = (int)((total + per_cta - 1) / per_cta);
    if (grid > 148 * 64) grid = 148 * 64;
    AV1B_LAUNCH(deblock_kernel<0>, (grid, 1, 3), (128), st, c);
    AV1B_LAUNCH(deblock_kernel<1>, (grid, 1, 3), (128), st, c);
}

void launch_cdef(const PostCtx& c, const Av1bFrameHdr& h, av1b_stream_t st)
{
    if (!h.cdef.enabled) return;
    const int gx = (h.mi_cols * 4 + 63) / 64, gy = (h.mi_rows * 4 + 63) / 64;
    AV1B_LAUNCH(cdef_kernel, (gx, gy, 1), (256), st, c);
}

void launch_lr(const PostCtx& c, const Av1bFrameHdr& h, av1b_stream_t st)
{
    if (!h.lr.uses_lr) return;
    const int gx = (h.frame_w + LR_TW - 1) / LR_TW;
    const int gy = (h.frame_h + 8 + 63) / 64;
    AV1B_LAUNCH(lr_kernel, (gx, gy, 3), (256), st, c);
}

// ==========================================================================================
// output conversion
// ==========================================================================================
// Planar 4:2:0 to NV12: luma rows copied, chroma rows interleaved U0 V0 U1 V1 ...  One thread per
// 4 luma samples / 2 chroma pairs; pure streaming (3 * w * h bytes moved).
__global__ void __launch_bounds__(256) nv12_kernel(FrameView src, uint8_t* dst_y, int pitch_y, uint8_t* dst_uv, int pitch_uv, int w, int h)
{
    const int cw = w >> 1, ch = h >> 1;
    const int lw4 = (w + 3) >> 2, cw2 = (cw + 1) >> 1;
    const int n_luma = lw4 * h, n_chroma = cw2 * ch;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_luma + n_chroma; t += gridDim.x * blockDim.x) {
        if (t < n_luma) {
            const int y = t / lw4, x = (t - y * lw4) * 4;
            const uint8_t* s = src.pl[0].p + (size_t)y * src.pl[0].stride + x;
            uint8_t* d = dst_y + (size_t)y * pitch_y + x;
            for (int k = 0; k < 4 && x + k < w; k++) d[k] = s[k];
        } else {
            const int u = t - n_luma;
            const int y = u / cw2, x = (u - y * cw2) * 2;
            const uint8_t* su = src.pl[1].p + (size_t)y * src.pl[1].stride + x;
            const uint8_t* sv = src.pl[2].p + (size_t)y * src.pl[2].stride + x;
            uint8_t* d = dst_uv + (size_t)y * pitch_uv + 2 * x;
            for (int k = 0; k < 2 && x + k < cw; k++) {
                d[2 * k] = su[k];
                d[2 * k + 1] = sv[k];
            }
        }
    }
}

void launch_to_nv12(const FrameView& src, uint8_t* dst_y, int pitch_y, uint8_t* dst_uv, int pitch_uv, int w, int h, av1b_stream_t st)
{
    const long long n = (long long)((w + 3) >> 2) * h + (long long)(((w >> 1) + 1) >> 1) * (h >> 1);
    if (n <= 0) return;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    AV1B_LAUNCH(nv12_kernel, (grid), (256), st, src, dst_y, pitch_y, dst_uv, pitch_uv, w, h);
}

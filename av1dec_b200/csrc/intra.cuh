// intra.cuh -- AV1 intra prediction for one transform block, executed by one group of lanes (a
// warp in the superblock wavefront, a CTA on the global-memory path).
//
// Behaviour restated from decoder/IntraPredict.cpp of the reference (edge assembly :563-611,
// DC :485, Paeth :151, smooth :526-561, directional with edge filter / upsample :269-469,
// recursive filter-intra :112-149, chroma-from-luma :632-667).
//
// Shape of the computation (what makes the dependent pass fast is the LENGTH of this chain):
//   phase 1  both edges are assembled word-wise into the group's scratch (aligned 32-bit loads of
//            the row above, four gathered bytes per word of the left column);
//   phase 2  directional modes only: corner filter + 5-tap edge filter in ONE step, written to a
//            second pair of arrays (no staging copy), every tap read straight from phase 1;
//   phase 3  only where the spec upsamples an edge: second pair -> first pair;
//   phase 4  the prediction, FOUR horizontally adjacent samples per work item, with the residual
//            add, the CfL term and the clip fused, one aligned 32-bit store per item.
// DC / smooth / Paeth / V / H blocks run phases 1 and 4 only.
//
// Reference quirks that are reproduced on purpose:
//   * directional numPx uses maxX WITHOUT the -1 (IntraPredict.cpp:385,401)
//   * the frame-edge clamp of the edge fetch uses ((MiCols*4)>>subX)-1 (:568-569)
#pragma once
#include "dev.h"
#include "av1_tables.h"

namespace intra {

enum { EDGE_OFF = 32, EDGE_LEN = 192 };

struct Scratch {
    // edge[0] above / edge[1] left as assembled (AboveRow[i] at edge[0][EDGE_OFF + i]);
    // edge[2] / edge[3] the same after the edge filter
    alignas(16) uint8_t edge[4][EDGE_LEN];
    alignas(16) uint8_t pred[32 * 32]; // inter-intra only: the intra half of the blend (<= 32x32)
    int acc;                           // CfL sum (groups larger than a warp)
};

struct Args {
    int x, y, log2w, log2h; // block position / size in plane samples
    int max_x, max_y;       // ((MiCols*4)>>subX)-1, ((MiRows*4)>>subY)-1
    int plane_idx;
    int mode;               // PREDICTION_MODE (0..12)
    int angle_delta;
    bool have_left, have_above, have_above_right, have_below_left;
    bool edge_filter_enabled; // sequence enable_intra_edge_filter
    bool edge_smooth;         // get_filter_type()
    bool filter_intra;
    int fi_mode;
    int strip;                // row strip of the block this op predicts: log2(strips) | index << 2 (0 = all of it)
    bool cfl;                 // chroma-from-luma on top of the DC prediction
    int cfl_alpha;
    int max_luma_w, max_luma_h;
};

enum { K_DC = 0, K_V, K_H, K_PAETH, K_SMOOTH, K_SMOOTH_V, K_SMOOTH_H, K_DIR_LT90, K_DIR_MID, K_DIR_GT180, K_FILTER_INTRA };

// Everything about one block's prediction that does NOT depend on sample values: derived once,
// off the dependency chain (the wavefront decodes a superblock's ops lane-parallel while it still
// waits for its neighbours and keeps them bit-packed in shared memory), so that the chain itself
// is only loads, arithmetic and stores.
struct Prep {
    int lw, lh;
    bool have_left, have_above;
    int above_n, left_n; // samples of the row above from x / of the left column from y that exist (then replicate)
    int kind;            // K_*
    // directional
    int dx, dy, up_above, up_left;
    bool filt, corner;   // phase 2 needed; the corner sample is filtered too
    int sA, sL, szA, szL;
    int fi_mode;
    int strip;
    bool cfl;
    int alpha, lim_w, lim_h; // CfL: luma clamp limits relative to the block's luma origin
};

// Where the block lives.
struct Io {
    const uint8_t* blk;   // block origin in the plane (edges: blk - stride, blk - 1)
    int stride;
    uint8_t* P;           // where the prediction goes (== blk for in-place), rows 4-byte aligned
    int pp;
    const int16_t* res;   // residual of this block (sample (0,0) of the block) added on the way, or null
    int rpitch;
    const uint8_t* luma;  // CfL: luma sample under the block origin
    int luma_stride;
};

// Plane samples live in the superblock tile in shared memory (SMEM) or in the frame in global
// memory (frames with intrabc): there they were written by other CTAs moments ago, so the loads
// bypass L1.
template <bool SMEM> AV1B_DEV int ld8(const uint8_t* p) { return SMEM ? (int)*p : (int)__ldcg(p); }
template <bool SMEM> AV1B_DEV uint32_t ld32(const uint8_t* p) { return SMEM ? *(const uint32_t*)p : __ldcg((const uint32_t*)p); }

AV1B_DEV int edge_filter_strength(int w, int h, bool smooth, int delta)
{
    int d = iabs(delta), wh = w + h, s = 0;
    if (!smooth) {
        if (wh <= 8) { if (d >= 56) s = 1; }
        else if (wh <= 12) { if (d >= 40) s = 1; }
        else if (wh <= 16) { if (d >= 40) s = 1; }
        else if (wh <= 24) { if (d >= 8) s = 1; if (d >= 16) s = 2; if (d >= 32) s = 3; }
        else if (wh <= 32) { s = 1; if (d >= 4) s = 2; if (d >= 32) s = 3; }
        else s = 3;
    } else {
        if (wh <= 8) { if (d >= 40) s = 1; if (d >= 64) s = 2; }
        else if (wh <= 16) { if (d >= 20) s = 1; if (d >= 48) s = 2; }
        else if (wh <= 24) { if (d >= 4) s = 3; }
        else s = 3;
    }
    return s;
}

AV1B_DEV int edge_upsample(int w, int h, bool smooth, int delta)
{
    int d = iabs(delta), wh = w + h;
    if (d <= 0 || d >= 40) return 0;
    return smooth ? (wh <= 8) : (wh <= 16);
}

// Sum of `v` over the nt (<= 32) lanes of the group, in every lane.
AV1B_DEV unsigned warp_sum(unsigned v, int nt)
{
#ifdef AV1B_EMU
    (void)nt;
    return v;
#else
    return nt == 32 ? __reduce_add_sync(0xFFFFFFFFu, v) : v; // nt < 32 only in single-lane test configurations
#endif
}

// Four predicted samples (packed, sample j in byte j) of row i, columns 4q .. 4q+3: residual add
// and clip fused, one aligned 32-bit store.
AV1B_DEV void put4(const Io& o, int i, int q, uint32_t v)
{
    if (o.res) v = add_res4(v, *(const uint2*)(o.res + i * o.rpitch + 4 * q));
    *(uint32_t*)(o.P + i * o.pp + 4 * q) = v;
}

AV1B_DEV uint32_t pack4(int v0, int v1, int v2, int v3) { return (uint32_t)v0 | ((uint32_t)v1 << 8) | ((uint32_t)v2 << 16) | ((uint32_t)v3 << 24); }

// Sub-sampled luma under four chroma samples (row i, columns j0 .. j0+3 of the block), each the
// sum of a 2x2 luma quad times two (IntraPredict.cpp:640-650), as two packed 16x2 words.
template <bool SMEM> AV1B_DEV uint2 cfl_luma4(int lim_w, int lim_h, const Io& o, int i, int j0)
{
    const int ly = min(2 * i, lim_h);
    const int lx0 = 2 * j0;
    const uint8_t* q = o.luma + (ptrdiff_t)ly * o.luma_stride;
    if (lx0 + 6 <= lim_w) {
        const uint32_t a0 = ld32<SMEM>(q + lx0), a1 = ld32<SMEM>(q + lx0 + 4);
        const uint32_t b0 = ld32<SMEM>(q + o.luma_stride + lx0), b1 = ld32<SMEM>(q + o.luma_stride + lx0 + 4);
        // bytes (0,1) and (2,3) of a word are the quads of two adjacent chroma samples
        const uint32_t s0 = (a0 & 0x00FF00FFu) + ((a0 >> 8) & 0x00FF00FFu) + (b0 & 0x00FF00FFu) + ((b0 >> 8) & 0x00FF00FFu);
        const uint32_t s1 = (a1 & 0x00FF00FFu) + ((a1 >> 8) & 0x00FF00FFu) + (b1 & 0x00FF00FFu) + ((b1 >> 8) & 0x00FF00FFu);
        return make_uint2(s0 << 1, s1 << 1);
    }
    uint32_t v[4];
    AV1B_UNROLL
    for (int m = 0; m < 4; m++) {
        const int lx = min(lx0 + 2 * m, lim_w);
        v[m] = (uint32_t)(ld8<SMEM>(q + lx) + ld8<SMEM>(q + lx + 1) + ld8<SMEM>(q + o.luma_stride + lx) + ld8<SMEM>(q + o.luma_stride + lx + 1)) << 1;
    }
    return make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
}

// The sample-independent part (IntraPredict.cpp:379-411 for the directional set-up).
AV1B_DEV Prep prepare(const Args& a)
{
    Prep p;
    const int w = 1 << a.log2w, h = 1 << a.log2h;
    p.lw = a.log2w, p.lh = a.log2h;
    p.have_left = a.have_left, p.have_above = a.have_above;
    p.above_n = min(w + h, min(a.max_x, a.x + (a.have_above_right ? 2 * w : w) - 1) - a.x + 1);
    p.left_n = min(w + h, min(a.max_y, a.y + (a.have_below_left ? 2 * h : h) - 1) - a.y + 1);
    p.dx = p.dy = p.up_above = p.up_left = 0;
    p.filt = p.corner = false;
    p.sA = p.sL = p.szA = p.szL = 0;
    p.fi_mode = a.fi_mode;
    p.strip = a.strip;
    p.cfl = a.cfl;
    p.alpha = a.cfl_alpha;
    p.lim_w = min(255, max(0, a.max_luma_w - 2 - 2 * a.x));
    p.lim_h = min(255, max(0, a.max_luma_h - 2 - 2 * a.y));
    const int mode = a.mode;
    if (a.plane_idx == 0 && a.filter_intra) p.kind = K_FILTER_INTRA;
    else if (mode >= 1 && mode <= 8) {
        const int p_angle = k_mode_to_angle[mode] + a.angle_delta * 3;
        if (p_angle == 90) p.kind = K_V;
        else if (p_angle == 180) p.kind = K_H;
        else {
            if (a.edge_filter_enabled) {
                p.corner = p_angle > 90 && p_angle < 180 && (w + h) >= 24;
                const int maxx_q = a.max_x + 1, maxy_q = a.max_y + 1; // quirk: no -1
                if (a.have_above) {
                    p.sA = edge_filter_strength(w, h, a.edge_smooth, p_angle - 90);
                    p.szA = min(w, maxx_q - a.x + 1) + (p_angle < 90 ? h : 0) + 1;
                }
                if (a.have_left) {
                    p.sL = edge_filter_strength(w, h, a.edge_smooth, p_angle - 180);
                    p.szL = min(h, maxy_q - a.y + 1) + (p_angle > 180 ? w : 0) + 1;
                }
                p.up_above = edge_upsample(w, h, a.edge_smooth, p_angle - 90);
                p.up_left = edge_upsample(w, h, a.edge_smooth, p_angle - 180);
                p.filt = p.corner || p.sA || p.sL || p.up_above || p.up_left;
            }
            if (p_angle < 90) {
                p.kind = K_DIR_LT90;
                p.dx = k_dr_intra_derivative[p_angle];
            } else if (p_angle < 180) {
                p.kind = K_DIR_MID;
                p.dx = k_dr_intra_derivative[180 - p_angle];
                p.dy = k_dr_intra_derivative[p_angle - 90];
            } else {
                p.kind = K_DIR_GT180;
                p.dy = k_dr_intra_derivative[270 - p_angle];
            }
        }
    } else p.kind = mode == 0 ? K_DC : (mode == 12 ? K_PAETH : (mode == 9 ? K_SMOOTH : (mode == 10 ? K_SMOOTH_V : K_SMOOTH_H)));
    return p;
}

// Bit-packed Prep (four words).  run() works from this form and extracts a field where it needs
// it, so a block only pays for the fields of its own mode.
struct Packed {
    uint32_t w[4];
    AV1B_DEV_M int lw() const { return (int)(w[0] & 7); }
    AV1B_DEV_M int lh() const { return (int)((w[0] >> 3) & 7); }
    AV1B_DEV_M bool have_left() const { return (w[0] >> 6) & 1; }
    AV1B_DEV_M bool have_above() const { return (w[0] >> 7) & 1; }
    AV1B_DEV_M int kind() const { return (int)((w[0] >> 8) & 15); }
    AV1B_DEV_M bool cfl() const { return (w[0] >> 12) & 1; }
    AV1B_DEV_M int up_above() const { return (int)((w[0] >> 13) & 1); }
    AV1B_DEV_M int up_left() const { return (int)((w[0] >> 14) & 1); }
    AV1B_DEV_M bool corner() const { return (w[0] >> 15) & 1; }
    AV1B_DEV_M bool filt() const { return (w[0] >> 16) & 1; }
    AV1B_DEV_M int fi_mode() const { return (int)((w[0] >> 17) & 7); }
    AV1B_DEV_M int sA() const { return (int)((w[0] >> 20) & 3); }
    AV1B_DEV_M int sL() const { return (int)((w[0] >> 22) & 3); }
    AV1B_DEV_M int above_n() const { return (int)(w[0] >> 24); }
    AV1B_DEV_M int left_n() const { return (int)(w[1] & 0xFF); }
    AV1B_DEV_M int szA() const { return (int)((w[1] >> 8) & 0xFF); }
    AV1B_DEV_M int szL() const { return (int)((w[1] >> 16) & 0xFF); }
    AV1B_DEV_M int alpha() const { return (int)(int8_t)(w[1] >> 24); }
    AV1B_DEV_M int dx() const { return (int)(w[2] & 0x7FF); }
    AV1B_DEV_M int dy() const { return (int)((w[2] >> 11) & 0x7FF); }
    AV1B_DEV_M int lim_w() const { return (int)(w[3] & 0xFF); }
    AV1B_DEV_M int lim_h() const { return (int)((w[3] >> 8) & 0xFF); }
    AV1B_DEV_M int strip_lg() const { return (int)((w[3] >> 24) & 3); }
    AV1B_DEV_M int strip_idx() const { return (int)((w[3] >> 26) & 7); }
};
AV1B_DEV Packed pack(const Prep& p)
{
    Packed k;
    k.w[0] = (uint32_t)p.lw | ((uint32_t)p.lh << 3) | ((uint32_t)p.have_left << 6) | ((uint32_t)p.have_above << 7) | ((uint32_t)p.kind << 8)
        | ((uint32_t)p.cfl << 12) | ((uint32_t)p.up_above << 13) | ((uint32_t)p.up_left << 14) | ((uint32_t)p.corner << 15) | ((uint32_t)p.filt << 16)
        | ((uint32_t)p.fi_mode << 17) | ((uint32_t)p.sA << 20) | ((uint32_t)p.sL << 22) | ((uint32_t)p.above_n << 24);
    k.w[1] = (uint32_t)p.left_n | ((uint32_t)p.szA << 8) | ((uint32_t)p.szL << 16) | ((uint32_t)(p.alpha & 0xFF) << 24);
    k.w[2] = (uint32_t)p.dx | ((uint32_t)p.dy << 11);
    k.w[3] = (uint32_t)p.lim_w | ((uint32_t)p.lim_h << 8) | ((uint32_t)p.strip << 24); // (bit 16: the wavefront's residual flag)
    return k;
}

// Predict one block into o.P (rows 4-byte aligned) and add the residual / CfL term on the way.
// o.P may be the block's own position in the plane: the edges are copied out first and nothing
// else of the plane is read afterwards (CfL reads the LUMA plane).  All threads of the group must
// call it.
// NTC: the group size when it is a compile-time constant (32 = one warp per op), 0 = use nt_rt.
template <bool SMEM, int NTC>
AV1B_DEV void run(const Packed& p, const Io& o, Scratch& S, int tid, int nt_rt)
{
    const int nt = NTC ? NTC : nt_rt;
    const int lw = p.lw(), lh = p.lh();
    const int w = 1 << lw, h = 1 << lh;
    const int lq = lw - 2, nq = w >> 2;  // groups of four columns
    const int items = h << lq;
    // A large block may be shared by several warps, each predicting a strip of rows (the emitter
    // splits the op, Av1bOp::fi_mode): every warp prepares the edges for itself, the prediction
    // of a sample depends on nothing but the edges.  (Never a filter-intra block.)
    const int strip_rows = h >> p.strip_lg();
    const int e_lo = (p.strip_idx() * strip_rows) << lq, e_hi = e_lo + (strip_rows << lq);
    uint8_t* const A = S.edge[0] + EDGE_OFF;
    uint8_t* const L = S.edge[1] + EDGE_OFF;
    uint8_t* const P = o.P;
    const int pp = o.pp;
    const int kind = p.kind();
    const bool hl = p.have_left(), ha = p.have_above();
    const uint8_t* const row_above = o.blk - o.stride;
    const uint8_t* const col_left = o.blk - 1;
    // DC / V / H / Paeth / smooth blocks whose edges all exist read them where they lie (tile in
    // shared memory): no assembly, no barrier.  EA[i] = AboveRow[i], EL[i * es] = LeftCol[i].
    const bool direct = SMEM && kind <= K_SMOOTH_H && ha && hl && p.above_n() >= w && p.left_n() >= h;
    const uint8_t* const EA = direct ? row_above : A;
    const uint8_t* const EL = direct ? col_left : L;
    const int es = direct ? o.stride : 1;
    // ---- phase 1: edge assembly (IntraPredict.cpp:579-611).  Every item gathers four bytes at
    // base + min(n - 1, i) * step; the three kinds of item (word of the row above, word of the
    // left column, corner) differ only in those operands, so the lanes do not diverge.
    // Filter-intra with every edge sample there reads the tile as well: the row above and the
    // column to the left of a 4x2 sub-block are then the same expression whether they are the
    // block's edge or samples predicted a step earlier.
    const bool fi_inplace = SMEM && kind == K_FILTER_INTRA && ha && hl && o.P == o.blk && p.above_n() >= w && p.left_n() >= h;
    if (!direct && !fi_inplace) {
        const int an = p.above_n(), ln = p.left_n();
        const int nw4 = (w + h) >> 2;
        const bool none = !ha && !hl;
        AV1B_NOUNROLL
        for (int e = tid; e <= 2 * nw4; e += nt) {
            const bool is_a = e < nw4, is_c = e == 2 * nw4;
            const int i0 = is_a ? 4 * e : 4 * (e - nw4);
            const uint8_t* base;
            int n, step;
            if (is_c) base = ha ? (hl ? row_above - 1 : row_above) : col_left, n = 1, step = 0; // (x-1,y-1) | (x,y-1) | (x-1,y)
            else if (is_a) base = ha ? row_above : col_left, n = ha ? an : 1, step = 1;            // row above | (x-1,y) replicated
            else base = hl ? col_left : row_above, n = hl ? ln : 1, step = o.stride;              // left column | (x,y-1) replicated
            uint32_t v = pack4(ld8<SMEM>(base + (ptrdiff_t)min(n - 1, i0) * step), ld8<SMEM>(base + (ptrdiff_t)min(n - 1, i0 + 1) * step),
                ld8<SMEM>(base + (ptrdiff_t)min(n - 1, i0 + 2) * step), ld8<SMEM>(base + (ptrdiff_t)min(n - 1, i0 + 3) * step));
            if (none) v = is_c ? 0x80808080u : (is_a ? 0x7F7F7F7Fu : 0x81818181u);
            if (is_c) {
                A[-1] = (uint8_t)v;
                L[-1] = (uint8_t)v;
            } else *(uint32_t*)((is_a ? A : L) + i0) = v;
        }
        block_sync(nt);
    }
    if (kind == K_FILTER_INTRA) {
        // ---- recursive filter-intra: 4x2 sub-blocks, anti-diagonal wavefront.  The recursion
        // feeds on PREDICTED samples, so the residual is added in a pass of its own afterwards.
        const int w4 = w >> 2, h2 = h >> 1;
        // a lane keeps the same position inside the 4x2 sub-block on every step when the group
        // size is a multiple of 8: its seven taps are loaded once, not on every diagonal
        const bool fixed_k = (nt & 7) == 0;
        int taps[7];
        AV1B_UNROLL
        for (int i = 0; i < 7; i++) taps[i] = k_intra_filter_taps[p.fi_mode()][tid & 7][i];
        for (int d = 0; d < w4 + h2 - 1; d++) {
            int j_lo = max(0, d - (h2 - 1)), j_hi = min(w4 - 1, d);
            int nsb = j_hi - j_lo + 1;
            AV1B_NOUNROLL
            for (int e = tid; e < nsb * 8; e += nt) {
                int j4 = j_lo + (e >> 3), i2 = d - j4, k = e & 7;
                int i1 = k >> 2, j1 = k & 3;
                int px[7];
                if (fi_inplace) {
                    const uint8_t* t = P + ((i2 << 1) - 1) * pp + (j4 << 2) - 1; // above-left of the sub-block
                    AV1B_UNROLL
                    for (int i = 0; i < 5; i++) px[i] = t[i];
                    px[5] = t[pp];
                    px[6] = t[2 * pp];
                } else {
                    AV1B_UNROLL
                    for (int i = 0; i < 5; i++) {
                        if (!i2) px[i] = A[(j4 << 2) + i - 1];
                        else if (!j4 && !i) px[i] = L[(i2 << 1) - 1];
                        else px[i] = P[((i2 << 1) - 1) * pp + (j4 << 2) + i - 1];
                    }
                    AV1B_UNROLL
                    for (int i = 5; i < 7; i++) {
                        if (!j4) px[i] = L[(i2 << 1) + i - 5];
                        else px[i] = P[((i2 << 1) + i - 5) * pp + (j4 << 2) - 1];
                    }
                }
                int pr = 0;
                AV1B_UNROLL
                for (int i = 0; i < 7; i++) pr += (fixed_k ? taps[i] : (int)k_intra_filter_taps[p.fi_mode()][k][i]) * px[i];
                P[((i2 << 1) + i1) * pp + (j4 << 2) + j1] = (uint8_t)clip_u8(round2s(pr, 4));
            }
            block_sync(nt);
        }
        if (o.res) {
            AV1B_NOUNROLL
            for (int e = tid; e < items; e += nt) {
                const int i = e >> lq, q = e & (nq - 1);
                put4(o, i, q, *(const uint32_t*)(P + i * pp + 4 * q));
            }
            block_sync(nt);
        }
        return;
    }
    if (kind >= K_DIR_LT90 && kind <= K_DIR_GT180) {
        // ---- directional (IntraPredict.cpp:379-469)
        const int up_above = p.up_above(), up_left = p.up_left();
        const uint8_t* DA = A;
        const uint8_t* DL = L;
        if (p.filt()) {
            uint8_t* const A2 = S.edge[2] + EDGE_OFF;
            uint8_t* const L2 = S.edge[3] + EDGE_OFF;
            // phase 2: new[m] = sum_j kern[j] * old[clip(-1, sz-2, m-2+j)] for m = 0 .. sz-2 (the
            // reference filters a copy that starts at the corner), old[-1] being the FILTERED
            // corner where the spec filters it; other entries are carried over
            const int n = w + h;
            const int cf = p.corner() ? ((L[0] * 5 + A[-1] * 6 + A[0] * 5 + 8) >> 4) : (int)A[-1];
            AV1B_NOUNROLL
            for (int e = tid; e < 2 * n; e += nt) {
                const bool above = e < n;
                const int m = above ? e : e - n;
                const uint8_t* old = above ? A : L;
                const int s = above ? p.sA() : p.sL(), sz = above ? p.szA() : p.szL();
                int v;
                if (s && m <= sz - 2) {
                    // taps {0,4,8,4,0}, {0,5,6,5,0}, {2,4,4,4,2}
                    const int k0 = s == 3 ? 2 : 0, k1 = s == 2 ? 5 : 4, k2 = s == 1 ? 8 : (s == 2 ? 6 : 4);
                    const int hi = sz - 2;
                    const int t0 = max(m - 2, -1), t1 = m - 1, t3 = min(m + 1, hi), t4 = min(m + 2, hi);
                    const int v0 = t0 < 0 ? cf : (int)old[t0], v1 = t1 < 0 ? cf : (int)old[t1];
                    v = (k0 * (v0 + (int)old[t4]) + k1 * (v1 + (int)old[t3]) + k2 * (int)old[m] + 8) >> 4;
                } else v = old[m];
                uint8_t* nw = above ? A2 : L2;
                nw[m] = (uint8_t)v;
                if (m == 0) nw[-1] = (uint8_t)cf;
            }
            block_sync(nt);
            DA = A2;
            DL = L2;
            if (up_above | up_left) {
                // phase 3: 2x upsampling of edge[-1 .. n-1] into edge[-2 .. 2n-2] (reference
                // intraEdgeUpsample), second pair -> first pair
                const int nA = up_above ? w + (kind == K_DIR_LT90 ? h : 0) : 0;
                const int nL = up_left ? h + (kind == K_DIR_GT180 ? w : 0) : 0;
                AV1B_NOUNROLL
                for (int e = tid; e < nA + nL; e += nt) {
                    const bool above = e < nA;
                    const int i = above ? e : e - nA, nn = above ? nA : nL;
                    const uint8_t* old = above ? A2 : L2;
                    uint8_t* nw = above ? A : L;
                    const int s = -(int)old[max(i - 2, -1)] + 9 * (int)old[i - 1] + 9 * (int)old[i] - (int)old[min(i + 1, nn - 1)];
                    nw[2 * i - 1] = (uint8_t)clip_u8((s + 8) >> 4);
                    nw[2 * i] = old[i];
                    if (i == 0) nw[-2] = old[-1];
                }
                block_sync(nt);
                if (up_above) DA = A;
                if (up_left) DL = L;
            }
        }
        if (kind == K_DIR_LT90) {
            const int dx = p.dx();
            const int max_base = (w + h - 1) << up_above;
            const int top = DA[max_base];
            AV1B_NOUNROLL
            for (int e = e_lo + tid; e < e_hi; e += nt) {
                const int i = e >> lq, q = e & (nq - 1);
                const int idx = (i + 1) * dx;
                const int shift = ((idx << up_above) >> 1) & 31;
                const int b0 = (idx >> (6 - up_above)) + ((4 * q) << up_above);
                int v[4];
                AV1B_UNROLL
                for (int m = 0; m < 4; m++) {
                    const int base = b0 + (m << up_above);
                    v[m] = base < max_base ? ((DA[base] * (32 - shift) + DA[base + 1] * shift + 16) >> 5) : top;
                }
                put4(o, i, q, pack4(v[0], v[1], v[2], v[3]));
            }
        } else if (kind == K_DIR_MID) {
            const int dx = p.dx(), dy = p.dy();
            AV1B_NOUNROLL
            for (int e = e_lo + tid; e < e_hi; e += nt) {
                const int i = e >> lq, q = e & (nq - 1);
                int v[4];
                AV1B_UNROLL
                for (int m = 0; m < 4; m++) {
                    const int j = 4 * q + m;
                    int idx = (j << 6) - (i + 1) * dx;
                    int base = idx >> (6 - up_above);
                    if (base >= -(1 << up_above)) {
                        const int shift = ((idx << up_above) >> 1) & 31;
                        v[m] = (DA[base] * (32 - shift) + DA[base + 1] * shift + 16) >> 5;
                    } else {
                        idx = (i << 6) - (j + 1) * dy;
                        base = idx >> (6 - up_left);
                        const int shift = ((idx << up_left) >> 1) & 31;
                        v[m] = (DL[base] * (32 - shift) + DL[base + 1] * shift + 16) >> 5;
                    }
                }
                put4(o, i, q, pack4(v[0], v[1], v[2], v[3]));
            }
        } else {
            const int dy = p.dy();
            AV1B_NOUNROLL
            for (int e = e_lo + tid; e < e_hi; e += nt) {
                const int i = e >> lq, q = e & (nq - 1);
                int v[4];
                AV1B_UNROLL
                for (int m = 0; m < 4; m++) {
                    const int idx = (4 * q + m + 1) * dy;
                    const int base = (idx >> (6 - up_left)) + (i << up_left);
                    const int shift = ((idx << up_left) >> 1) & 31;
                    v[m] = (DL[base] * (32 - shift) + DL[base + 1] * shift + 16) >> 5;
                }
                put4(o, i, q, pack4(v[0], v[1], v[2], v[3]));
            }
        }
    } else if (kind == K_V) {
        AV1B_NOUNROLL
        for (int e = e_lo + tid; e < e_hi; e += nt) {
            const int i = e >> lq, q = e & (nq - 1);
            put4(o, i, q, *(const uint32_t*)(EA + 4 * q));
        }
    } else if (kind == K_H) {
        AV1B_NOUNROLL
        for (int e = e_lo + tid; e < e_hi; e += nt) {
            const int i = e >> lq, q = e & (nq - 1);
            put4(o, i, q, (uint32_t)EL[i * es] * 0x01010101u);
        }
    } else if (kind == K_PAETH) {
        const int tl = EA[-1];
        AV1B_NOUNROLL
        for (int e = e_lo + tid; e < e_hi; e += nt) {
            const int i = e >> lq, q = e & (nq - 1);
            const uint32_t aw = *(const uint32_t*)(EA + 4 * q);
            const int l = EL[i * es];
            int v[4];
            AV1B_UNROLL
            for (int m = 0; m < 4; m++) {
                const int t = (int)((aw >> (8 * m)) & 0xFF);
                const int base = t + l - tl;
                const int pl = iabs(base - l), pt = iabs(base - t), ptl = iabs(base - tl);
                v[m] = (pl <= pt && pl <= ptl) ? l : (pt <= ptl ? t : tl);
            }
            put4(o, i, q, pack4(v[0], v[1], v[2], v[3]));
        }
    } else if (kind == K_DC) {
        // ---- DC.  Both edge sums travel in one word (each <= 64 * 255) through one warp
        // reduction; the divisor w + h is 2^k, 3 * 2^k or 5 * 2^k, so the division is a shift and
        // an exact multiply-high.
        int sl = 0, sa = 0;
        if (nt <= 32) {
            unsigned part = 0;
            const int wq = w >> 2, hq = h >> 2;
            AV1B_NOUNROLL
            for (int k = tid; k < wq + hq; k += nt) {
                if (k < wq) part = av1b_dp4a_uu(*(const uint32_t*)(EA + 4 * k), 0x01010101u, part);
                else {
                    const uint8_t* l4 = EL + 4 * (k - wq) * es;
                    part += (uint32_t)(l4[0] + l4[es] + l4[2 * es] + l4[3 * es]) << 16;
                }
            }
            part = warp_sum(part, nt);
            sl = (int)(part >> 16);
            sa = (int)(part & 0xFFFF);
        } else {
            for (int k = 0; k < h; k++) sl += EL[k * es];
            for (int k = 0; k < w; k++) sa += EA[k];
        }
        int avg;
        if (hl && ha) {
            const int lmin = min(lw, lh), ratio = iabs(lw - lh); // w + h = (1 + 2^ratio) << lmin
            const unsigned q = (unsigned)(sl + sa + ((w + h) >> 1)) >> lmin;
            avg = ratio == 0 ? (int)(q >> 1) : (ratio == 1 ? (int)(__umulhi(q, 0xAAAAAAABu) >> 1) : (int)(__umulhi(q, 0xCCCCCCCDu) >> 2));
        } else if (hl) avg = clip_u8((sl + (h >> 1)) >> lh);
        else if (ha) avg = clip_u8((sa + (w >> 1)) >> lw);
        else avg = 128;
        if (!p.cfl()) {
            const uint32_t word = (uint32_t)avg * 0x01010101u;
            AV1B_NOUNROLL
            for (int e = e_lo + tid; e < e_hi; e += nt) put4(o, e >> lq, e & (nq - 1), word);
        } else {
            // ---- chroma-from-luma on top of the DC value (IntraPredict.cpp:632-667): two passes over
            // the sub-sampled luma (sum, then apply) instead of a staging buffer
            const int lim_w = p.lim_w(), lim_h = p.lim_h();
            int local = 0;
            AV1B_NOUNROLL
            for (int e = tid; e < items; e += nt) {
                const uint2 l4 = cfl_luma4<SMEM>(lim_w, lim_h, o, e >> lq, 4 * (e & (nq - 1)));
                local += (int)((l4.x & 0xFFFF) + (l4.x >> 16) + (l4.y & 0xFFFF) + (l4.y >> 16));
            }
            int total;
            if (nt <= 32) total = (int)warp_sum((unsigned)local, nt);
            else {
                if (tid == 0) S.acc = 0;
                block_sync(nt);
                atomicAdd(&S.acc, local);
                block_sync(nt);
                total = S.acc;
            }
            const int lavg = round2(total, lw + lh);
            const int alpha = p.alpha();
            AV1B_NOUNROLL
            for (int e = e_lo + tid; e < e_hi; e += nt) {
                const int i = e >> lq, q = e & (nq - 1);
                const uint2 l4 = cfl_luma4<SMEM>(lim_w, lim_h, o, i, 4 * q);
                const int v0 = clip_u8(avg + round2s(alpha * ((int)(l4.x & 0xFFFF) - lavg), 6));
                const int v1 = clip_u8(avg + round2s(alpha * ((int)(l4.x >> 16) - lavg), 6));
                const int v2 = clip_u8(avg + round2s(alpha * ((int)(l4.y & 0xFFFF) - lavg), 6));
                const int v3 = clip_u8(avg + round2s(alpha * ((int)(l4.y >> 16) - lavg), 6));
                put4(o, i, q, pack4(v0, v1, v2, v3));
            }
        }
    } else if (kind == K_SMOOTH) {
        const uint8_t* wx = k_sm_weights + (w - 4);
        const uint8_t* wy = k_sm_weights + (h - 4);
        const int bl = EL[(h - 1) * es], tr = EA[w - 1];
        AV1B_NOUNROLL
        for (int e = e_lo + tid; e < e_hi; e += nt) {
            const int i = e >> lq, q = e & (nq - 1);
            const uint32_t aw = *(const uint32_t*)(EA + 4 * q);
            const int l = EL[i * es], wyi = wy[i];
            const int rowc = (256 - wyi) * bl + 256;
            int v[4];
            AV1B_UNROLL
            for (int m = 0; m < 4; m++) {
                const int wxj = wx[4 * q + m];
                v[m] = (wyi * (int)((aw >> (8 * m)) & 0xFF) + rowc + wxj * l + (256 - wxj) * tr) >> 9;
            }
            put4(o, i, q, pack4(v[0], v[1], v[2], v[3]));
        }
    } else if (kind == K_SMOOTH_V) {
        const uint8_t* wy = k_sm_weights + (h - 4);
        const int bl = EL[(h - 1) * es];
        AV1B_NOUNROLL
        for (int e = e_lo + tid; e < e_hi; e += nt) {
            const int i = e >> lq, q = e & (nq - 1);
            const uint32_t aw = *(const uint32_t*)(EA + 4 * q);
            const int wyi = wy[i];
            const int rowc = (256 - wyi) * bl + 128;
            int v[4];
            AV1B_UNROLL
            for (int m = 0; m < 4; m++) v[m] = (wyi * (int)((aw >> (8 * m)) & 0xFF) + rowc) >> 8;
            put4(o, i, q, pack4(v[0], v[1], v[2], v[3]));
        }
    } else { // K_SMOOTH_H
        const uint8_t* wx = k_sm_weights + (w - 4);
        const int tr = EA[w - 1];
        AV1B_NOUNROLL
        for (int e = e_lo + tid; e < e_hi; e += nt) {
            const int i = e >> lq, q = e & (nq - 1);
            const int l = EL[i * es];
            int v[4];
            AV1B_UNROLL
            for (int m = 0; m < 4; m++) {
                const int wxj = wx[4 * q + m];
                v[m] = (wxj * l + (256 - wxj) * tr + 128) >> 8;
            }
            put4(o, i, q, pack4(v[0], v[1], v[2], v[3]));
        }
    }
    block_sync(nt);
}

}  // namespace intra

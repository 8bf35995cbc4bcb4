// intra.cuh -- AV1 intra prediction for one transform block, executed by one CTA.
//
// Restates decoder/IntraPredict.cpp of the reference (edge assembly :563-611, DC :485,
// Paeth :151, smooth :526-561, directional with edge filter / upsample :269-469,
// recursive filter-intra :112-149, chroma-from-luma :632-667) with the CTA's threads
// striding over samples.  The prediction lands in shared memory (`pred`, pitch = w) so the
// caller can add the residual / blend and store once.
//
// Reference quirks that are reproduced on purpose:
//   * directional numPx uses maxX WITHOUT the -1 (IntraPredict.cpp:385,401)
//   * the frame-edge clamp of the edge fetch uses ((MiCols*4)>>subX)-1 (:568-569)
#pragma once
#include "dev.h"
#include "av1_tables.h"

namespace intra {

enum { EDGE_OFF = 32, EDGE_LEN = 320 };

struct Scratch {
    uint8_t above[EDGE_LEN]; // AboveRow[i] at above[EDGE_OFF + i]
    uint8_t left[EDGE_LEN];
    uint8_t tmp[EDGE_LEN];   // filter / upsample staging
    uint8_t pred[32 * 32];   // inter-intra only: the intra half of the blend (<= 32x32)
    int acc;                 // CfL sum
};

struct Args {
    const uint8_t* plane; // current-frame plane, sample (0,0)
    int stride;
    int x, y, log2w, log2h;
    int max_x, max_y; // ((MiCols*4)>>subX)-1, ((MiRows*4)>>subY)-1
    int plane_idx;
    int mode;         // PREDICTION_MODE (0..12)
    int angle_delta;
    bool have_left, have_above, have_above_right, have_below_left;
    bool edge_filter_enabled; // sequence enable_intra_edge_filter
    bool edge_smooth;         // get_filter_type()
    bool filter_intra;
    int fi_mode;
};

// The plane pointer may address the superblock tile in shared memory (wave_kernel) or the frame
// in global memory (legacy path for intrabc frames): a volatile generic load is correct for both
// (it is never served from a stale L1 line).
AV1B_DEV int px(const Args& a, int x, int y) { return *(const volatile uint8_t*)(a.plane + (ptrdiff_t)y * a.stride + x); }

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

// In-place 5-tap smoothing of edge[-1 .. sz-2] -> edge[0 .. sz-2] (reference intraEdgeFilter), for
// the above edge (A, szA, strength sA) and the left edge (L, szL, sL) in one pass over both: they
// are independent, and a pass costs a staging copy and two group barriers whatever its length.
// A strength of 0 skips that edge.  tmp: EDGE_LEN bytes, the left edge uses its second half.
template <int NTC>
AV1B_DEV void filter_edges(uint8_t* A, int szA, int sA, uint8_t* L, int szL, int sL, uint8_t* tmp, int tid, int nt_rt)
{
    const int nt = NTC ? NTC : nt_rt;
    if (!sA) szA = 0;
    if (!sL) szL = 0;
    if (!(szA | szL)) return;
    uint8_t* const tA = tmp;
    uint8_t* const tL = tmp + EDGE_LEN / 2;
    AV1B_NOUNROLL
    for (int k = tid; k < szA + szL; k += nt) {
        if (k < szA) tA[k] = A[k - 1];
        else tL[k - szA] = L[k - szA - 1];
    }
    block_sync(nt);
    AV1B_NOUNROLL
    for (int k = tid; k < szA + szL; k += nt) {
        const bool above = k < szA;
        const int i = above ? k : k - szA, sz = above ? szA : szL;
        if (i < 1) continue;
        const uint8_t* t = above ? tA : tL;
        const uint8_t* kern = k_intra_edge_kernel[(above ? sA : sL) - 1];
        int s = 0;
        AV1B_UNROLL
        for (int j = 0; j < 5; j++) s += kern[j] * t[clip3(0, sz - 1, i - 2 + j)];
        (above ? A : L)[i - 1] = (uint8_t)((s + 8) >> 4);
    }
    block_sync(nt);
}

// 2x upsampling of edge[-1 .. n-1] into edge[-2 .. 2n-2] (reference intraEdgeUpsample), above edge
// (nA samples) and left edge (nL samples) in one pass; n = 0 skips that edge.
template <int NTC>
AV1B_DEV void upsample_edges(uint8_t* A, int nA, uint8_t* L, int nL, uint8_t* tmp, int tid, int nt_rt)
{
    const int nt = NTC ? NTC : nt_rt;
    if (!(nA | nL)) return;
    uint8_t* const tA = tmp;
    uint8_t* const tL = tmp + EDGE_LEN / 2;
    const int cA = nA ? nA + 3 : 0, cL = nL ? nL + 3 : 0;
    // t[k] = dup[k] = edge[clip(-1, n-1, k-2)], k = 0 .. n+2
    AV1B_NOUNROLL
    for (int k = tid; k < cA + cL; k += nt) {
        if (k < cA) tA[k] = A[clip3(-1, nA - 1, k - 2)];
        else tL[k - cA] = L[clip3(-1, nL - 1, k - cA - 2)];
    }
    block_sync(nt);
    if (tid == 0) {
        if (nA) A[-2] = tA[0];
        if (nL) L[-2] = tL[0];
    }
    AV1B_NOUNROLL
    for (int k = tid; k < nA + nL; k += nt) {
        const bool above = k < nA;
        const int i = above ? k : k - nA;
        const uint8_t* t = above ? tA : tL;
        uint8_t* edge = above ? A : L;
        const int s = -t[i] + 9 * t[i + 1] + 9 * t[i + 2] - t[i + 3];
        edge[2 * i - 1] = (uint8_t)clip_u8((s + 8) >> 4);
        edge[2 * i] = t[i + 2];
    }
    block_sync(nt);
}

// Predict one block into P (row pitch pp).  P may be the block's own position in the plane: the
// edges are copied out first and nothing else of the plane is read afterwards.  All threads of
// the group must call it.
// NTC: the group size when it is a compile-time constant (32 = one warp per op), 0 = use nt_rt.
template <int NTC>
AV1B_DEV void predict(const Args& a, Scratch& S, uint8_t* P, int pp, int tid, int nt_rt)
{
    const int nt = NTC ? NTC : nt_rt;
    const int w = 1 << a.log2w, h = 1 << a.log2h;
    uint8_t* A = S.above + EDGE_OFF;
    uint8_t* L = S.left + EDGE_OFF;
    const int x = a.x, y = a.y;
    // ---- edge assembly (IntraPredict.cpp:579-611)
    {
        const bool hl = a.have_left, ha = a.have_above;
        int above_const = -1, left_const = -1;
        if (!ha && hl) above_const = px(a, x - 1, y);
        else if (!ha && !hl) above_const = 127;
        if (!hl && ha) left_const = px(a, x, y - 1);
        else if (!ha && !hl) left_const = 129;
        const int above_limit = min(a.max_x, x + (a.have_above_right ? 2 * w : w) - 1);
        const int left_limit = min(a.max_y, y + (a.have_below_left ? 2 * h : h) - 1);
        AV1B_NOUNROLL
        for (int i = tid; i < w + h; i += nt) {
            A[i] = (uint8_t)(above_const >= 0 ? above_const : px(a, min(above_limit, x + i), y - 1));
            L[i] = (uint8_t)(left_const >= 0 ? left_const : px(a, x - 1, min(left_limit, y + i)));
        }
        if (tid == 0) {
            int c;
            if (ha && hl) c = px(a, x - 1, y - 1);
            else if (ha) c = px(a, x, y - 1);
            else if (hl) c = px(a, x - 1, y);
            else c = 128;
            A[-1] = (uint8_t)c;
            L[-1] = (uint8_t)c;
        }
        block_sync(nt);
    }
    const int lw = a.log2w;
    if (a.plane_idx == 0 && a.filter_intra) {
        // ---- recursive filter-intra: 4x2 sub-blocks, anti-diagonal wavefront
        const int w4 = w >> 2, h2 = h >> 1;
        // a lane keeps the same position inside the 4x2 sub-block on every step when the group
        // size is a multiple of 8: its seven taps are loaded once, not on every diagonal
        const bool fixed_k = (nt & 7) == 0;
        int taps[7];
        AV1B_UNROLL
        for (int i = 0; i < 7; i++) taps[i] = k_intra_filter_taps[a.fi_mode][tid & 7][i];
        for (int d = 0; d < w4 + h2 - 1; d++) {
            int j_lo = max(0, d - (h2 - 1)), j_hi = min(w4 - 1, d);
            int nsb = j_hi - j_lo + 1;
            AV1B_NOUNROLL
            for (int e = tid; e < nsb * 8; e += nt) {
                int j4 = j_lo + (e >> 3), i2 = d - j4, k = e & 7;
                int i1 = k >> 2, j1 = k & 3;
                int p[7];
                AV1B_UNROLL
                for (int i = 0; i < 5; i++) {
                    if (!i2) p[i] = A[(j4 << 2) + i - 1];
                    else if (!j4 && !i) p[i] = L[(i2 << 1) - 1];
                    else p[i] = P[((i2 << 1) - 1) * pp + (j4 << 2) + i - 1];
                }
                AV1B_UNROLL
                for (int i = 5; i < 7; i++) {
                    if (!j4) p[i] = L[(i2 << 1) + i - 5];
                    else p[i] = P[((i2 << 1) + i - 5) * pp + (j4 << 2) - 1];
                }
                int pr = 0;
                AV1B_UNROLL
                for (int i = 0; i < 7; i++) pr += (fixed_k ? taps[i] : (int)k_intra_filter_taps[a.fi_mode][k][i]) * p[i];
                P[((i2 << 1) + i1) * pp + (j4 << 2) + j1] = (uint8_t)clip_u8(round2s(pr, 4));
            }
            block_sync(nt);
        }
        return;
    }
    const int mode = a.mode;
    if (mode >= 1 && mode <= 8) {
        // ---- directional (IntraPredict.cpp:379-469)
        const int p_angle = k_mode_to_angle[mode] + a.angle_delta * 3;
        int up_above = 0, up_left = 0;
        if (a.edge_filter_enabled && p_angle != 90 && p_angle != 180) {
            if (p_angle > 90 && p_angle < 180 && (w + h) >= 24) {
                if (tid == 0) {
                    int s = (L[0] * 5 + A[-1] * 6 + A[0] * 5 + 8) >> 4;
                    L[-1] = (uint8_t)s;
                    A[-1] = (uint8_t)s;
                }
                block_sync(nt);
            }
            const int maxx_q = a.max_x + 1, maxy_q = a.max_y + 1; // quirk: no -1
            int sA = 0, numA = 0, sL = 0, numL = 0;
            if (a.have_above) {
                sA = edge_filter_strength(w, h, a.edge_smooth, p_angle - 90);
                numA = min(w, maxx_q - x + 1) + (p_angle < 90 ? h : 0) + 1;
            }
            if (a.have_left) {
                sL = edge_filter_strength(w, h, a.edge_smooth, p_angle - 180);
                numL = min(h, maxy_q - y + 1) + (p_angle > 180 ? w : 0) + 1;
            }
            filter_edges<NTC>(A, numA, sA, L, numL, sL, S.tmp, tid, nt);
            up_above = edge_upsample(w, h, a.edge_smooth, p_angle - 90);
            up_left = edge_upsample(w, h, a.edge_smooth, p_angle - 180);
            upsample_edges<NTC>(A, up_above ? w + (p_angle < 90 ? h : 0) : 0, L, up_left ? h + (p_angle > 180 ? w : 0) : 0, S.tmp, tid, nt);
        }
        if (p_angle < 90) {
            const int dx = k_dr_intra_derivative[p_angle];
            const int max_base = (w + h - 1) << up_above;
            AV1B_NOUNROLL
            for (int e = tid; e < w * h; e += nt) {
                int i = e >> a.log2w, j = e & (w - 1);
                int idx = (i + 1) * dx;
                int base = (idx >> (6 - up_above)) + (j << up_above);
                int shift = ((idx << up_above) >> 1) & 31;
                P[(e >> lw) * pp + (e & (w - 1))] = (uint8_t)(base < max_base ? ((A[base] * (32 - shift) + A[base + 1] * shift + 16) >> 5) : A[max_base]);
            }
        } else if (p_angle > 90 && p_angle < 180) {
            const int dx = k_dr_intra_derivative[180 - p_angle];
            const int dy = k_dr_intra_derivative[p_angle - 90];
            AV1B_NOUNROLL
            for (int e = tid; e < w * h; e += nt) {
                int i = e >> a.log2w, j = e & (w - 1);
                int idx = (j << 6) - (i + 1) * dx;
                int base = idx >> (6 - up_above);
                int v;
                if (base >= -(1 << up_above)) {
                    int shift = ((idx << up_above) >> 1) & 31;
                    v = (A[base] * (32 - shift) + A[base + 1] * shift + 16) >> 5;
                } else {
                    idx = (i << 6) - (j + 1) * dy;
                    base = idx >> (6 - up_left);
                    int shift = ((idx << up_left) >> 1) & 31;
                    v = (L[base] * (32 - shift) + L[base + 1] * shift + 16) >> 5;
                }
                P[(e >> lw) * pp + (e & (w - 1))] = (uint8_t)v;
            }
        } else if (p_angle > 180) {
            const int dy = k_dr_intra_derivative[270 - p_angle];
            AV1B_NOUNROLL
            for (int e = tid; e < w * h; e += nt) {
                int i = e >> a.log2w, j = e & (w - 1);
                int idx = (j + 1) * dy;
                int base = (idx >> (6 - up_left)) + (i << up_left);
                int shift = ((idx << up_left) >> 1) & 31;
                P[(e >> lw) * pp + (e & (w - 1))] = (uint8_t)((L[base] * (32 - shift) + L[base + 1] * shift + 16) >> 5);
            }
        } else if (p_angle == 90) {
            AV1B_NOUNROLL
            for (int e = tid; e < w * h; e += nt) P[(e >> lw) * pp + (e & (w - 1))] = A[e & (w - 1)];
        } else {
            AV1B_NOUNROLL
            for (int e = tid; e < w * h; e += nt) P[(e >> lw) * pp + (e & (w - 1))] = L[e >> a.log2w];
        }
    } else if (mode == 12) {
        // ---- Paeth
        const int tl = A[-1];
        AV1B_NOUNROLL
        for (int e = tid; e < w * h; e += nt) {
            int i = e >> a.log2w, j = e & (w - 1);
            int base = A[j] + L[i] - tl;
            int pl = iabs(base - L[i]), pt = iabs(base - A[j]), ptl = iabs(base - tl);
            P[(e >> lw) * pp + (e & (w - 1))] = (pl <= pt && pl <= ptl) ? L[i] : (pt <= ptl ? A[j] : (uint8_t)tl);
        }
    } else if (mode == 0) {
        // ---- DC.  Both edge sums travel in one word (each <= 64 * 255) through one warp
        // reduction; the divisor w + h is 2^k, 3 * 2^k or 5 * 2^k, so the division is a shift and
        // an exact multiply-high.
        int sl = 0, sa = 0;
        if (nt <= 32) {
            unsigned part = 0;
            AV1B_NOUNROLL
            for (int k = tid; k < h; k += nt) part += (unsigned)L[k] << 16;
            AV1B_NOUNROLL
            for (int k = tid; k < w; k += nt) part += A[k];
            part = warp_sum(part, nt);
            sl = (int)(part >> 16);
            sa = (int)(part & 0xFFFF);
        } else {
            for (int k = 0; k < h; k++) sl += L[k];
            for (int k = 0; k < w; k++) sa += A[k];
        }
        int avg;
        if (a.have_left && a.have_above) {
            const int lmin = min(a.log2w, a.log2h), ratio = iabs(a.log2w - a.log2h); // w + h = (1 + 2^ratio) << lmin
            const unsigned q = (unsigned)(sl + sa + ((w + h) >> 1)) >> lmin;
            avg = ratio == 0 ? (int)(q >> 1) : (ratio == 1 ? (int)(__umulhi(q, 0xAAAAAAABu) >> 1) : (int)(__umulhi(q, 0xCCCCCCCDu) >> 2));
        } else if (a.have_left) avg = clip_u8((sl + (h >> 1)) >> a.log2h);
        else if (a.have_above) avg = clip_u8((sa + (w >> 1)) >> a.log2w);
        else avg = 128;
        AV1B_NOUNROLL
        for (int e = tid; e < w * h; e += nt) P[(e >> lw) * pp + (e & (w - 1))] = (uint8_t)avg;
    } else if (mode == 9) {
        const uint8_t* wx = k_sm_weights + (w - 4);
        const uint8_t* wy = k_sm_weights + (h - 4);
        const int bl = L[h - 1], tr = A[w - 1];
        AV1B_NOUNROLL
        for (int e = tid; e < w * h; e += nt) {
            int i = e >> a.log2w, j = e & (w - 1);
            int v = wy[i] * A[j] + (256 - wy[i]) * bl + wx[j] * L[i] + (256 - wx[j]) * tr;
            P[(e >> lw) * pp + (e & (w - 1))] = (uint8_t)((v + 256) >> 9);
        }
    } else if (mode == 10) {
        const uint8_t* wy = k_sm_weights + (h - 4);
        const int bl = L[h - 1];
        AV1B_NOUNROLL
        for (int e = tid; e < w * h; e += nt) {
            int i = e >> a.log2w, j = e & (w - 1);
            P[(e >> lw) * pp + (e & (w - 1))] = (uint8_t)((wy[i] * A[j] + (256 - wy[i]) * bl + 128) >> 8);
        }
    } else { // mode == 11, SMOOTH_H
        const uint8_t* wx = k_sm_weights + (w - 4);
        const int tr = A[w - 1];
        AV1B_NOUNROLL
        for (int e = tid; e < w * h; e += nt) {
            int i = e >> a.log2w, j = e & (w - 1);
            P[(e >> lw) * pp + (e & (w - 1))] = (uint8_t)((wx[j] * L[i] + (256 - wx[j]) * tr + 128) >> 8);
        }
    }
    block_sync(nt);
}

// Chroma-from-luma on top of the DC prediction already in P (IntraPredict.cpp:632-667).
// `luma` is plane 0 of the current frame (already reconstructed for this block).  Two passes over
// the sub-sampled luma (sum, then apply) instead of a staging buffer.
AV1B_DEV int cfl_luma(const Args& a, const uint8_t* luma, int luma_stride, int max_luma_w, int max_luma_h, int i, int j)
{
    const int ly = min((a.y + i) << 1, max_luma_h - 2);
    const int lx = min((a.x + j) << 1, max_luma_w - 2);
    const volatile uint8_t* q = luma + (ptrdiff_t)ly * luma_stride + lx;
    return (q[0] + q[1] + q[luma_stride] + q[luma_stride + 1]) << 1;
}

template <int NTC>
AV1B_DEV void apply_cfl(const Args& a, const uint8_t* luma, int luma_stride, int alpha, int max_luma_w, int max_luma_h,
    Scratch& S, uint8_t* P, int pp, int tid, int nt_rt)
{
    const int nt = NTC ? NTC : nt_rt;
    const int w = 1 << a.log2w, h = 1 << a.log2h;
    if (tid == 0) S.acc = 0;
    block_sync(nt);
    int local = 0;
    AV1B_NOUNROLL
    for (int e = tid; e < w * h; e += nt) local += cfl_luma(a, luma, luma_stride, max_luma_w, max_luma_h, e >> a.log2w, e & (w - 1));
    atomicAdd(&S.acc, local);
    block_sync(nt);
    const int avg = round2(S.acc, a.log2w + a.log2h);
    AV1B_NOUNROLL
    for (int e = tid; e < w * h; e += nt) {
        const int i = e >> a.log2w, j = e & (w - 1);
        const int scaled = round2s(alpha * (cfl_luma(a, luma, luma_stride, max_luma_w, max_luma_h, i, j) - avg), 6);
        uint8_t* d = P + i * pp + j;
        *d = (uint8_t)clip_u8(*d + scaled);
    }
    block_sync(nt);
}

}  // namespace intra

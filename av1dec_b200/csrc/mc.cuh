// mc.cuh -- motion compensation: 8-tap sub-pel convolve, warped motion, compound / masked /
// distance-weighted blending and OBMC strips, for one "inter prediction unit" (Av1bIpu).
//
// Restates the arithmetic of the reference's InterPredict (decoder/InterPredict.cpp):
//   motionVectorScaling :66-83 (unscaled references only, as the reference asserts :395)
//   blockPixelPredict :319-331, blockSubPixelPredict :333-383, getFilterIdx :85-97
//   blockWarp :507-553, maskBlend :584-609, predict_overlap :611-628
//   wedgeMask :888-899, differenceWeightMask :901-915, final writers :1023-1045
// but is organised for a CTA: the unit is cut into <=32x32 tiles, each tile's horizontal pass
// lands in shared memory (int16), the vertical pass and the blend run one thread per sample.
#pragma once
#include "dev.h"
#include "av1_tables.h"
#include "../../include/av1b200_format.h"

namespace mc {

// Tiles of TILE_W x TILE_H samples.  Scratch is per executing group (a warp in inter_kernel, the
// CTA on the intrabc path): ~5 KB.
// INTER_PITCH: row pitch of the transposed intermediate (odd word count: conflict-free both ways).
enum { TILE_W = 32, TILE_H = 16, WIN_PITCH = TILE_W + 8, WIN_ROWS = TILE_H + 7, INTER_PITCH = 26, INTER_ELEMS = 8 * 15 * 8 };

struct Scratch {
    alignas(8) uint8_t win[WIN_ROWS * WIN_PITCH + 8]; // reference window of the tile (+7 samples each way), clamped
    alignas(8) int16_t inter[INTER_ELEMS];            // horizontal-pass intermediates (>= TILE_W * INTER_PITCH)
    alignas(8) int16_t pred[2][TILE_W * TILE_H];      // per-list predictions of the current tile
};

struct RefPlane {
    const uint8_t* p;
    int stride;
    int last_x, last_y;
    bool coherent; // true: read through L2 (current frame, intrabc)
};

AV1B_DEV int ref_px(const RefPlane& r, int x, int y)
{
    x = clip3(0, r.last_x, x);
    y = clip3(0, r.last_y, y);
    const uint8_t* q = r.p + (size_t)y * r.stride + x;
    return r.coherent ? (int)__ldcg(q) : (int)__ldg(q);
}

// InterpFilter -> row of k_subpel_filters (reference getFilterIdx)
AV1B_DEV int filter_row(int size, int interp)
{
    if (size <= 4) {
        if (interp == 0 || interp == 2) return 4;
        if (interp == 1) return 5;
    }
    return interp;
}

// Wedge mask table: [9 block sizes][2 signs][16 wedges][32*32], built by the engine at start-up.
AV1B_DEV int wedge_size_index(int mi_size)
{
    // BLOCK_8X8=3 8X16=4 16X8=5 16X16=6 16X32=7 32X16=8 32X32=9 8X32=18 32X8=19
    if (mi_size >= 3 && mi_size <= 9) return mi_size - 3;
    if (mi_size == 18) return 7;
    if (mi_size == 19) return 8;
    return -1;
}
AV1B_DEV const uint8_t* wedge_mask_ptr(const uint8_t* table, int mi_size, int sign, int index)
{
    return table + ((size_t)((wedge_size_index(mi_size) * 2 + sign) * 16 + index)) * 1024;
}

struct Params {
    const Av1bFrameHdr* hdr;
    const Av1bBlkAux* aux;   // may be null when the unit needs none
    const uint8_t* wedge;    // wedge table
    uint8_t* mask;           // luma-resolution mask of the block (diff-weighted compound): address of
    int mask_pitch;          //   the block's top-left sample, bytes per row
    PlaneView dst;           // destination plane of the current frame
    const int16_t* res;      // residual plane to add while writing (independent units), or null
    int rpitch;
    RefPlane ref[2];
};

AV1B_DEV int ilog2_pow2(int v) { return 31 - __clz(v); }

// Translational prediction of one tile: the tw x th samples whose integer reference position
// starts at (px0, py0), filtered with `taps` (packed halved taps: .x/.y horizontal, .z/.w
// vertical) when `subpel`, into pred (pitch TILE_W).  tw is a power of two (2..32), th <= TILE_H.
AV1B_DEV void convolve_tile(const RefPlane& R, int px0, int py0, bool subpel, uint4 taps, int tw, int th, int round1,
    Scratch& s, int16_t* pred, int tid, int nt)
{
    const int ltw = ilog2_pow2(tw);
    // word path: aligned plane, window fully inside the reference (the padding right of each
    // row absorbs the <= 7 bytes a word read may run past the window)
    const bool words = ((((uintptr_t)R.p) | (unsigned)R.stride) & 3) == 0;
    if (!subpel) {
        const int sh = 14 - 3 - round1;
        const bool inside = px0 >= 0 && px0 + tw - 1 <= R.last_x && py0 >= 0 && py0 + th - 1 <= R.last_y;
        if (words && inside && tw >= 4) {
            const uint8_t* base = R.p + (ptrdiff_t)py0 * R.stride + px0;
            const unsigned mis8 = ((unsigned)(uintptr_t)base & 3) * 8;
            base -= mis8 >> 3;
            const int lq = ltw - 2;
            for (int e = tid; e < (th << lq); e += nt) {
                const int r = e >> lq, q = e & ((1 << lq) - 1);
                const uint32_t* g = (const uint32_t*)(base + (ptrdiff_t)r * R.stride) + q;
                const uint32_t lo = R.coherent ? __ldcg(g) : __ldg(g);
                const uint32_t hi = R.coherent ? __ldcg(g + 1) : __ldg(g + 1);
                const uint32_t w = __funnelshift_r(lo, hi, mis8);
                uint2 o;
                o.x = (__byte_perm(w, 0, 0x4140)) << sh;
                o.y = (__byte_perm(w, 0, 0x4342)) << sh;
                *(uint2*)(pred + r * TILE_W + 4 * q) = o;
            }
        } else {
            for (int e = tid; e < (th << ltw); e += nt) {
                const int r = e >> ltw, c = e & (tw - 1);
                pred[r * TILE_W + c] = (int16_t)(ref_px(R, px0 + c, py0 + r) << sh);
            }
        }
        block_sync(nt);
        return;
    }
    // stage the clamped reference window: rows py0-3 .. py0+th+3, columns px0-3 .. px0+tw+3
    {
        const int ww = tw + 7, wh = th + 7;
        const bool inside = px0 - 3 >= 0 && px0 + tw + 3 <= R.last_x && py0 - 3 >= 0 && py0 + th + 3 <= R.last_y;
        if (words && inside) {
            const uint8_t* base = R.p + (ptrdiff_t)(py0 - 3) * R.stride + (px0 - 3);
            const unsigned mis8 = ((unsigned)(uintptr_t)base & 3) * 8;
            base -= mis8 >> 3;
            const int nwr = (ww + 3) >> 2;              // window words per row: 3, 4, 6, 10
            const unsigned inv = nwr == 3 ? 21846u : nwr == 4 ? 16385u : nwr == 6 ? 10923u : 6554u; // 65536 / nwr + 1
            uint32_t* win32 = (uint32_t*)s.win;
            AV1B_UNROLL4
            for (int e = tid; e < nwr * wh; e += nt) {
                const int r = (int)(((unsigned)e * inv) >> 16), k = e - r * nwr;
                const uint32_t* g = (const uint32_t*)(base + (ptrdiff_t)r * R.stride) + k;
                const uint32_t lo = R.coherent ? __ldcg(g) : __ldg(g);
                const uint32_t hi = R.coherent ? __ldcg(g + 1) : __ldg(g + 1);
                win32[r * (WIN_PITCH / 4) + k] = __funnelshift_r(lo, hi, mis8);
            }
        } else {
            const unsigned magic = ww == 39 ? 110127367u : ww == 23 ? 186737709u : ww == 15 ? 286331154u : ww == 11 ? 390451573u : 477218589u;
            for (int e = tid; e < ww * wh; e += nt) {
                const int r = (int)__umulhi((unsigned)e, magic), c = e - r * ww;
                s.win[r * WIN_PITCH + c] = (uint8_t)ref_px(R, px0 - 3 + c, py0 - 3 + r);
            }
        }
    }
    block_sync(nt);
    // Every AV1 sub-pel tap is even (reference table InterPredict.cpp:99-); halved taps fit a
    // signed byte, so the 8-tap sums run as packed dot products and the roundings drop one bit:
    //   (2s + 4) >> 3 == (s + 2) >> 2,   round2(2s, n) == (s + (1 << (n - 2))) >> (n - 1).
    const uint32_t fha = taps.x, fhb = taps.y, fva = taps.z, fvb = taps.w;
    // horizontal pass, four outputs per lane; intermediates are stored TRANSPOSED
    // (inter[c * INTER_PITCH + r]) so that the vertical pass reads row pairs as one word.
    {
        const int lq = ltw >= 2 ? ltw - 2 : 0;
        const uint32_t* win32 = (const uint32_t*)s.win;
        for (int e = tid; e < ((th + 7) << lq); e += nt) {
            const int r = e >> lq, q = e & ((1 << lq) - 1);
            const uint32_t* wp = win32 + r * (WIN_PITCH / 4) + q;
            const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
            int16_t* o = s.inter + (4 * q) * INTER_PITCH + r;
            int sum = av1b_dp4a_us(w1, fhb, av1b_dp4a_us(w0, fha, 2));
            o[0] = (int16_t)(sum >> 2);
            sum = av1b_dp4a_us(__byte_perm(w1, w2, 0x4321), fhb, av1b_dp4a_us(__byte_perm(w0, w1, 0x4321), fha, 2));
            o[INTER_PITCH] = (int16_t)(sum >> 2);
            if (tw > 2) {
                sum = av1b_dp4a_us(__byte_perm(w1, w2, 0x5432), fhb, av1b_dp4a_us(__byte_perm(w0, w1, 0x5432), fha, 2));
                o[2 * INTER_PITCH] = (int16_t)(sum >> 2);
                sum = av1b_dp4a_us(__byte_perm(w1, w2, 0x6543), fhb, av1b_dp4a_us(__byte_perm(w0, w1, 0x6543), fha, 2));
                o[3 * INTER_PITCH] = (int16_t)(sum >> 2);
            }
        }
    }
    block_sync(nt);
    // vertical pass: one lane = one column, two output rows
    {
        const int rnd = 1 << (round1 - 2), shv = round1 - 1;
        for (int e = tid; e < ((th >> 1) << ltw); e += nt) {
            const int k = e >> ltw, c = e & (tw - 1);
            const uint32_t* q = (const uint32_t*)(s.inter + c * INTER_PITCH) + k;
            const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3], w4 = q[4];
            int sum = av1b_dp2a_lo(w0, fva, rnd);
            sum = av1b_dp2a_hi(w1, fva, sum);
            sum = av1b_dp2a_lo(w2, fvb, sum);
            sum = av1b_dp2a_hi(w3, fvb, sum);
            pred[(2 * k) * TILE_W + c] = (int16_t)(sum >> shv);
            sum = av1b_dp2a_lo(__funnelshift_r(w0, w1, 16), fva, rnd);
            sum = av1b_dp2a_hi(__funnelshift_r(w1, w2, 16), fva, sum);
            sum = av1b_dp2a_lo(__funnelshift_r(w2, w3, 16), fvb, sum);
            sum = av1b_dp2a_hi(__funnelshift_r(w3, w4, 16), fvb, sum);
            pred[(2 * k + 1) * TILE_W + c] = (int16_t)(sum >> shv);
        }
    }
    block_sync(nt);
}

// ---- fast translational path (inter_fast_kernel): no staged window, no prediction buffer ----
// Horizontal pass straight from global memory into the transposed intermediate, then one pass
// that filters vertically, blends the lists, adds the residual and stores, 4x2 samples a lane.
struct FastScratch {
    alignas(8) int16_t inter[2][TILE_W * INTER_PITCH];
};

// True when every sample an 8-tap window around the tile touches lies inside the reference and
// the plane can be read by aligned words.
AV1B_DEV bool fast_tile_ok(const RefPlane& R, int px0, int py0, int tw, int th)
{
    return px0 - 3 >= 0 && px0 + tw + 3 <= R.last_x && py0 - 3 >= 0 && py0 + th + 3 <= R.last_y
        && ((((uintptr_t)R.p) | (unsigned)R.stride) & 3) == 0;
}

// Horizontal pass of one list: rows py0-3 .. py0+th+3 (sub-pel) or py0 .. py0+th-1 (integer
// position; the value stored is then already the prediction, sample << sh).
AV1B_DEV void fast_h(const RefPlane& R, int px0, int py0, bool subpel, uint32_t fha, uint32_t fhb, int ltw, int th, int sh,
    int16_t* inter, int tid, int nt)
{
    const int lq = ltw - 2;
    if (subpel) {
        const uint8_t* g0 = R.p + (ptrdiff_t)(py0 - 3) * R.stride + (px0 - 3);
        const unsigned mis8 = ((unsigned)(uintptr_t)g0 & 3) * 8;
        g0 -= mis8 >> 3;
        for (int e = tid; e < ((th + 7) << lq); e += nt) {
            const int r = e >> lq, q = e & ((1 << lq) - 1);
            const uint32_t* g = (const uint32_t*)(g0 + (ptrdiff_t)r * R.stride) + q;
            const uint32_t a0 = __ldg(g), a1 = __ldg(g + 1), a2 = __ldg(g + 2), a3 = __ldg(g + 3);
            const uint32_t w0 = __funnelshift_r(a0, a1, mis8), w1 = __funnelshift_r(a1, a2, mis8), w2 = __funnelshift_r(a2, a3, mis8);
            int16_t* o = inter + (4 * q) * INTER_PITCH + r;
            int sum = av1b_dp4a_us(w1, fhb, av1b_dp4a_us(w0, fha, 2));
            o[0] = (int16_t)(sum >> 2);
            sum = av1b_dp4a_us(__byte_perm(w1, w2, 0x4321), fhb, av1b_dp4a_us(__byte_perm(w0, w1, 0x4321), fha, 2));
            o[INTER_PITCH] = (int16_t)(sum >> 2);
            sum = av1b_dp4a_us(__byte_perm(w1, w2, 0x5432), fhb, av1b_dp4a_us(__byte_perm(w0, w1, 0x5432), fha, 2));
            o[2 * INTER_PITCH] = (int16_t)(sum >> 2);
            sum = av1b_dp4a_us(__byte_perm(w1, w2, 0x6543), fhb, av1b_dp4a_us(__byte_perm(w0, w1, 0x6543), fha, 2));
            o[3 * INTER_PITCH] = (int16_t)(sum >> 2);
        }
    } else {
        const uint8_t* g0 = R.p + (ptrdiff_t)py0 * R.stride + px0;
        const unsigned mis8 = ((unsigned)(uintptr_t)g0 & 3) * 8;
        g0 -= mis8 >> 3;
        for (int e = tid; e < (th << lq); e += nt) {
            const int r = e >> lq, q = e & ((1 << lq) - 1);
            const uint32_t* g = (const uint32_t*)(g0 + (ptrdiff_t)r * R.stride) + q;
            const uint32_t w = __funnelshift_r(__ldg(g), __ldg(g + 1), mis8);
            int16_t* o = inter + (4 * q) * INTER_PITCH + r;
            o[0] = (int16_t)((w & 0xFF) << sh);
            o[INTER_PITCH] = (int16_t)(((w >> 8) & 0xFF) << sh);
            o[2 * INTER_PITCH] = (int16_t)(((w >> 16) & 0xFF) << sh);
            o[3 * INTER_PITCH] = (int16_t)((w >> 24) << sh);
        }
    }
}

// Vertical pass for column c, output rows 2k and 2k+1 (v0, v1: predictions before blending).
AV1B_DEV void fast_v(const int16_t* inter, int c, int k, bool subpel, uint32_t fva, uint32_t fvb, int rnd, int shv, int& v0, int& v1)
{
    const uint32_t* q = (const uint32_t*)(inter + c * INTER_PITCH) + k;
    if (!subpel) {
        const uint32_t w = q[0];
        v0 = (int)(int16_t)(w & 0xFFFF);
        v1 = (int)w >> 16;
        return;
    }
    const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3], w4 = q[4];
    int sum = av1b_dp2a_lo(w0, fva, rnd);
    sum = av1b_dp2a_hi(w1, fva, sum);
    sum = av1b_dp2a_lo(w2, fvb, sum);
    sum = av1b_dp2a_hi(w3, fvb, sum);
    v0 = sum >> shv;
    sum = av1b_dp2a_lo(__funnelshift_r(w0, w1, 16), fva, rnd);
    sum = av1b_dp2a_hi(__funnelshift_r(w1, w2, 16), fva, sum);
    sum = av1b_dp2a_lo(__funnelshift_r(w2, w3, 16), fvb, sum);
    sum = av1b_dp2a_hi(__funnelshift_r(w3, w4, 16), fvb, sum);
    v1 = sum >> shv;
}

// Prediction of list `l` for the tile at (tx,ty) size (tw,th) into s.pred[l] (pitch TILE_W).
// tw is a power of two (2..32), th <= TILE_H.
AV1B_DEV void predict_tile(const Params& P, const Av1bIpu& u, int l, int tx, int ty, int tw, int th,
    Scratch& s, int tid, int nt)
{
    const int subx = u.plane ? 1 : 0, suby = u.plane ? 1 : 0;
    const bool compound = (u.flags & AV1B_IPUF_COMPOUND) != 0;
    const int round1 = compound ? 7 : 11;
    const RefPlane& R = P.ref[l];
    const int ltw = ilog2_pow2(tw);
    int use_warp = u.warp[l];
    if (u.w < 8 || u.h < 8) use_warp = 0;
    int16_t* pred = s.pred[l];
    if (use_warp) {
        const int32_t* wp = (use_warp == 1) ? P.aux->warp_params : P.hdr->gm_params[u.ref_frame[l]];
        const int16_t* ab = (use_warp == 1) ? P.aux->warp_abgd : P.hdr->gm_abgd[u.ref_frame[l]];
        const int alpha = ab[0], beta = ab[1], gamma = ab[2], delta = ab[3];
        const int lnux = ltw - 3, nu = (tw >> 3) * (th >> 3);
        // horizontal pass: 15 x 8 per 8x8 unit
        for (int e = tid; e < nu * 120; e += nt) {
            const int unit = e / 120, k = e - unit * 120;
            const int i1 = (k >> 3) - 7, i2 = (k & 7) - 4;
            const int uy = unit >> lnux, ux = unit & ((1 << lnux) - 1);
            const int srcx = (u.x + tx + ux * 8 + 4) << subx;
            const int srcy = (u.y + ty + uy * 8 + 4) << suby;
            const int dstx = wp[2] * srcx + wp[3] * srcy + wp[0];
            const int dsty = wp[4] * srcx + wp[5] * srcy + wp[1];
            const int x4 = dstx >> subx, y4 = dsty >> suby;
            const int ix4 = x4 >> 16, sx4 = x4 & 0xFFFF, iy4 = y4 >> 16;
            const int sx = sx4 + alpha * i2 + beta * i1;
            const int offs = ((sx + 512) >> 10) + 64;
            const int16_t* f = k_warped_filters[offs];
            int sum = 0;
            AV1B_UNROLL
            for (int t = 0; t < 8; t++) sum += f[t] * ref_px(R, ix4 + i2 - 3 + t, iy4 + i1);
            s.inter[e] = (int16_t)((sum + 4) >> 3);
        }
        block_sync(nt);
        for (int e = tid; e < nu * 64; e += nt) {
            const int unit = e >> 6, k = e & 63;
            const int i1 = (k >> 3) - 4, i2 = (k & 7) - 4;
            const int uy = unit >> lnux, ux = unit & ((1 << lnux) - 1);
            const int srcx = (u.x + tx + ux * 8 + 4) << subx;
            const int srcy = (u.y + ty + uy * 8 + 4) << suby;
            const int dsty = wp[4] * srcx + wp[5] * srcy + wp[1];
            const int y4 = dsty >> suby;
            const int sy4 = y4 & 0xFFFF;
            const int sy = sy4 + gamma * i2 + delta * i1;
            const int offs = ((sy + 512) >> 10) + 64;
            const int16_t* f = k_warped_filters[offs];
            const int16_t* in = s.inter + unit * 120 + (i2 + 4);
            int sum = 0;
            AV1B_UNROLL
            for (int t = 0; t < 8; t++) sum += f[t] * in[(i1 + t + 4) * 8];
            pred[(uy * 8 + i1 + 4) * TILE_W + ux * 8 + i2 + 4] = (int16_t)round2(sum, round1);
        }
        block_sync(nt);
        return;
    }
    // translational prediction
    const int mvx = (2 * u.mv[l][1]) >> subx; // 1/16 sample
    const int mvy = (2 * u.mv[l][0]) >> suby;
    const int fx = mvx & 15, fy = mvy & 15;
    uint4 taps;
    taps.x = k_subpel_packed[filter_row(u.w, u.filt[1])][fx][0];
    taps.y = k_subpel_packed[filter_row(u.w, u.filt[1])][fx][1];
    taps.z = k_subpel_packed[filter_row(u.h, u.filt[0])][fy][0];
    taps.w = k_subpel_packed[filter_row(u.h, u.filt[0])][fy][1];
    convolve_tile(R, u.x + tx + (mvx >> 4), u.y + ty + (mvy >> 4), (fx | fy) != 0, taps, tw, th, round1, s, pred, tid, nt);
}

// Execute one unit: predict every tile and write / blend it into the destination plane.
AV1B_DEV void run_ipu(const Params& P, const Av1bIpu& u, Scratch& s, int tid, int nt)
{
    const bool compound = (u.flags & AV1B_IPUF_COMPOUND) != 0;
    const int plane = u.plane;
    for (int ty = 0; ty < u.h; ty += TILE_H) {
        const int th = min((int)TILE_H, u.h - ty);
        for (int tx = 0; tx < u.w; tx += TILE_W) {
            const int tw = min((int)TILE_W, u.w - tx);
            const int ltw = ilog2_pow2(tw);
            predict_tile(P, u, 0, tx, ty, tw, th, s, tid, nt);
            if (compound) predict_tile(P, u, 1, tx, ty, tw, th, s, tid, nt);
            const bool simple = u.kind == AV1B_IPU_PRED && (!compound || u.comp_type == AV1B_COMP_AVERAGE || u.comp_type == AV1B_COMP_DISTANCE);
            if (simple && tw >= 4 && !((u.x | P.dst.stride | (int)(uintptr_t)P.dst.p) & 3)) {
                // four samples per lane, one aligned word store
                const int lq = ltw - 2;
                const int w0 = !compound ? 1 : (u.comp_type == AV1B_COMP_AVERAGE ? 8 : u.fwd_w);
                const int w1 = !compound ? 0 : (u.comp_type == AV1B_COMP_AVERAGE ? 8 : u.bck_w);
                const int sh = compound ? 8 : 0;
                for (int e = tid; e < (th << lq); e += nt) {
                    const int r = e >> lq, q = e & ((1 << lq) - 1);
                    const uint2 a = *(const uint2*)(s.pred[0] + r * TILE_W + 4 * q);
                    uint2 b = make_uint2(0, 0);
                    if (compound) b = *(const uint2*)(s.pred[1] + r * TILE_W + 4 * q);
                    const int o0 = clip_u8(round2(w0 * (int)(int16_t)(a.x & 0xFFFF) + w1 * (int)(int16_t)(b.x & 0xFFFF), sh));
                    const int o1 = clip_u8(round2(w0 * ((int)a.x >> 16) + w1 * ((int)b.x >> 16), sh));
                    const int o2 = clip_u8(round2(w0 * (int)(int16_t)(a.y & 0xFFFF) + w1 * (int)(int16_t)(b.y & 0xFFFF), sh));
                    const int o3 = clip_u8(round2(w0 * ((int)a.y >> 16) + w1 * ((int)b.y >> 16), sh));
                    uint32_t out = (uint32_t)o0 | ((uint32_t)o1 << 8) | ((uint32_t)o2 << 16) | ((uint32_t)o3 << 24);
                    if (P.res) out = add_res4(out, *(const uint2*)(P.res + (size_t)(u.y + ty + r) * P.rpitch + (u.x + tx + 4 * q)));
                    *(uint32_t*)(P.dst.p + (size_t)(u.y + ty + r) * P.dst.stride + (u.x + tx + 4 * q)) = out;
                }
                block_sync(nt);
                continue;
            }
            for (int e = tid; e < (th << ltw); e += nt) {
                const int r = e >> ltw, c = e & (tw - 1);
                const int i = ty + r, j = tx + c; // position inside the unit
                uint8_t* d = P.dst.p + (size_t)(u.y + i) * P.dst.stride + (u.x + j);
                const int p0 = s.pred[0][r * TILE_W + c];
                int out;
                if (u.kind != AV1B_IPU_PRED) {
                    // OBMC strip (reference predict_overlap): mask runs along rows for the
                    // above pass, along columns for the left pass.
                    const int len = (u.kind == AV1B_IPU_OBMC_ABOVE) ? u.h : u.w;
                    const int m = k_obmc_mask[len - 2 + ((u.kind == AV1B_IPU_OBMC_ABOVE) ? i : j)];
                    const int cur = *(volatile uint8_t*)d;
                    out = clip_u8(round2(m * cur + (64 - m) * clip_u8(p0), 6));
                } else if (!compound) {
                    out = clip_u8(p0);
                } else {
                    const int p1 = s.pred[1][r * TILE_W + c];
                    if (u.comp_type == AV1B_COMP_AVERAGE) {
                        out = clip_u8(round2(p0 + p1, 5));
                    } else if (u.comp_type == AV1B_COMP_DISTANCE) {
                        out = clip_u8(round2(u.fwd_w * p0 + u.bck_w * p1, 8));
                    } else {
                        int m;
                        if (u.comp_type == AV1B_COMP_WEDGE) {
                            const uint8_t* W = wedge_mask_ptr(P.wedge, P.aux->mi_size, P.aux->wedge_sign, P.aux->wedge_index);
                            if (!plane) m = W[i * 32 + j];
                            else
                                m = (W[(2 * i) * 32 + 2 * j] + W[(2 * i) * 32 + 2 * j + 1] + W[(2 * i + 1) * 32 + 2 * j]
                                        + W[(2 * i + 1) * 32 + 2 * j + 1] + 2)
                                    >> 2;
                        } else { // DIFFWTD
                            if (!plane) {
                                int diff = iabs(p0 - p1);
                                diff = (diff + 8) >> 4;
                                m = clip3(0, 64, 38 + diff / 16);
                                if (P.aux->mask_type) m = 64 - m;
                                P.mask[i * P.mask_pitch + j] = (uint8_t)m;
                            } else {
                                const volatile uint8_t* M = P.mask;
                                const int mp = P.mask_pitch;
                                m = (M[(2 * i) * mp + 2 * j] + M[(2 * i) * mp + 2 * j + 1] + M[(2 * i + 1) * mp + 2 * j]
                                        + M[(2 * i + 1) * mp + 2 * j + 1] + 2)
                                    >> 2;
                            }
                        }
                        out = clip_u8(round2(m * p0 + (64 - m) * p1, 10));
                    }
                }
                if (P.res) out = clip_u8(out + P.res[(size_t)(u.y + i) * P.rpitch + (u.x + j)]);
                *d = (uint8_t)out;
            }
            block_sync(nt);
        }
    }
}

}  // namespace mc

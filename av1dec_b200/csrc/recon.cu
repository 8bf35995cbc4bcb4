// recon.cu -- block reconstruction kernels for sm_100a.
//
//   itx_kernel<CLS>    batched inverse transform, one launch per size class: dequantised int16
//                      coefficients -> int16 residual planes (frame layout), 32/max(w,h) transform
//                      blocks per warp, independent and fully parallel.
//   inter_fast_kernel  motion compensation of plain translational blocks (the bulk): walks the
//                      prediction-unit list, lane-parallel set-up, (unit, tile) jobs per warp,
//                      residual add fused into the store.
//   inter_kernel       every other inter block (warp, OBMC, masks, frame-edge windows), a warp per
//                      block.
//   wave_kernel        the dependent pass: intra prediction (+CfL, filter-intra, palette), inter-
//                      intra blend and residual add on a shared-memory superblock tile, ops sorted by
//                      dependency level, superblocks in wavefront order; neighbouring superblocks overlap
//                      (per-superblock progress words, Av1bSb hints; all-zero hints = classic 2-SB lag).
//   wave_kernel_global the same ops on global memory for frames with intrabc.
//
// Reference for the behaviour: decoder/TransformBlock.cpp:2376-2456 (TransformBlock::decode),
// decoder/Block.cpp:100-174,1600-1608 (compute_prediction / Block::decode),
// decoder/Tile.cpp:172-181 (SB raster order).
#include "dev.h"
#include "av1_tables.h"
#include "itx.cuh"
#include "intra.cuh"
#include "mc.cuh"
#include "kernels.h"
#include <mutex>
#include <algorithm>
#include <cstdio>
#include <cstdlib>

// ------------------------------------------------------------------------------------------
// inverse transform
// ------------------------------------------------------------------------------------------
namespace {

enum { ITX_WARPS = 4, ITX_TMP_STRIDE = 66, ITX_TMP_ROWS = 32 };

}  // namespace

struct ResPlanes {
    int16_t* p[3]; // frame-layout residual planes, or p[0] == null: compact arena mode
    int pitch[3];
};

namespace {

// One transform block handled by the G lanes [lane_in_group = 0..G-1] of a warp.  MAXLOG bounds
// log2 of both dimensions for this size class, so small classes compile to small, low-register
// kernels (the 64-point butterflies alone need ~250 registers).
template <int MAXLOG>
AV1B_DEV void itx_block(const Av1bOp& op, const int16_t* __restrict__ coef, int16_t* __restrict__ res, const ResPlanes& rp,
    int16_t* tmp, int tstride, int gl, int G)
{
    const int txs = op.tx_size;
    const int lw = (int)((AV1T_TX_WLOG2_PACKED >> (3 * txs)) & 7), lh = (int)((AV1T_TX_HLOG2_PACKED >> (3 * txs)) & 7);
    const int w = 1 << lw, h = 1 << lh;
    const int tw = min(w, 32);
    const bool lossless = op.lossless != 0;
    const int rk = lossless ? itx::K_WHT : itx::row_kind(op.tx_type);
    const int ck = lossless ? itx::K_WHT : itx::col_kind(op.tx_type);
    const bool rect = (lw - lh == 1) || (lh - lw == 1);
    const int row_shift = lossless ? 0 : k_tx_row_shift[txs];
    const int col_shift = lossless ? 0 : 4;
    const int nz_rows = min((int)op.nz_rows, min(h, 32));
    const int nz_cols = lossless ? tw : min(max((int)op.nz_cols, 1), tw);
    const int16_t* c = coef + op.coef_off;
    int16_t* out;
    int out_stride;
    if (rp.p[0]) {
        out = rp.p[op.plane] + (size_t)op.y * rp.pitch[op.plane] + op.x;
        out_stride = rp.pitch[op.plane];
    } else {
        out = res + op.res_off;
        out_stride = w;
    }
    for (int i = gl; i < nz_rows; i += G) {
        int16_t* trow = tmp + i * tstride;
        if (MAXLOG >= 6 && lw == 6) itx::row_pass<6>(c + i * tw, tw, nz_cols, trow, rk, rect, row_shift);
        else if (MAXLOG >= 5 && lw == 5) itx::row_pass<5>(c + i * tw, tw, nz_cols, trow, rk, rect, row_shift);
        else if (MAXLOG >= 4 && lw == 4) itx::row_pass<4>(c + i * tw, tw, nz_cols, trow, rk, rect, row_shift);
        else if (MAXLOG >= 3 && lw == 3) itx::row_pass<3>(c + i * tw, tw, nz_cols, trow, rk, rect, row_shift);
        else itx::row_pass<2>(c + i * tw, tw, nz_cols, trow, rk, rect, row_shift);
    }
    __syncwarp();
    const bool fud = itx::flip_ud(op.tx_type), flr = itx::flip_lr(op.tx_type);
    for (int j = gl; j < w; j += G) {
        const int jo = flr ? (w - 1 - j) : j;
        if (MAXLOG >= 6 && lh == 6) itx::col_pass<6>(tmp + j, tstride, nz_rows, out + jo, out_stride, fud, ck, col_shift);
        else if (MAXLOG >= 5 && lh == 5) itx::col_pass<5>(tmp + j, tstride, nz_rows, out + jo, out_stride, fud, ck, col_shift);
        else if (MAXLOG >= 4 && lh == 4) itx::col_pass<4>(tmp + j, tstride, nz_rows, out + jo, out_stride, fud, ck, col_shift);
        else if (MAXLOG >= 3 && lh == 3) itx::col_pass<3>(tmp + j, tstride, nz_rows, out + jo, out_stride, fud, ck, col_shift);
        else itx::col_pass<2>(tmp + j, tstride, nz_rows, out + jo, out_stride, fud, ck, col_shift);
    }
}

}  // namespace

// Class 3 (a dimension of 32 or 64): the whole CTA works on ONE transform block at a time.  A 32- or
// 64-point DCT is shared by two lanes of different warps (even / odd half of the flow graph, results
// exchanged through shared memory, last Hadamard stage in the combine step), so the row pass of a
// 64x64 block keeps 64 lanes busy and the column pass all 128 -- the one-lane-per-transform layout
// of the small classes would leave a handful of 250-register warps per SM to do all the work.
AV1B_DEV void itx_cta_big(const Av1bOp* __restrict__ ops, const uint32_t* __restrict__ list, uint32_t n, const int16_t* __restrict__ coef,
    int16_t* __restrict__ res, const ResPlanes& rp, unsigned bid, unsigned nblocks)
{
    enum { XP = 33, TP = 66 };
    __shared__ int X[2][64][XP];     // [half][row | column][k]: the two halves before the last stage
    __shared__ int16_t tmp[32][TP];  // row-pass output (rows < nz_rows)
    const int tid = threadIdx.x, nt = blockDim.x;
    for (uint32_t t = bid; t < n; t += nblocks) {
        const Av1bOp op = ops[list[t]];
        const int txs = op.tx_size;
        const int lw = (int)((AV1T_TX_WLOG2_PACKED >> (3 * txs)) & 7), lh = (int)((AV1T_TX_HLOG2_PACKED >> (3 * txs)) & 7);
        const int w = 1 << lw, h = 1 << lh;
        const int tw = min(w, 32);
        const int rk = itx::row_kind(op.tx_type), ck = itx::col_kind(op.tx_type);
        const bool rect = (lw - lh == 1) || (lh - lw == 1);
        const int row_shift = k_tx_row_shift[txs];
        const int col_shift = 4;
        const int nz_rows = min(max((int)op.nz_rows, 1), min(h, 32));
        const int nz_cols = min(max((int)op.nz_cols, 1), tw);
        const int16_t* c = coef + op.coef_off;
        int16_t* out;
        int out_stride;
        if (rp.p[0]) {
            out = rp.p[op.plane] + (size_t)op.y * rp.pitch[op.plane] + op.x;
            out_stride = rp.pitch[op.plane];
        } else {
            out = res + op.res_off;
            out_stride = w;
        }
        const bool fud = itx::flip_ud(op.tx_type), flr = itx::flip_lr(op.tx_type);
        // ---- row pass
        if (rk == itx::K_DCT && lw >= 5) {
            for (int e = tid; e < 64; e += nt) { // lanes 0..31: even halves of the rows, 32..63: odd halves
                const int i = e & 31, half = e >> 5;
                if (i >= nz_rows) continue;
                if (lw == 6) itx::row_half<6>(c + i * tw, nz_cols, rect, half, X[half][i]);
                else itx::row_half<5>(c + i * tw, nz_cols, rect, half, X[half][i]);
            }
            __syncthreads();
            const int lm = lw - 1, M = 1 << lm;
            for (int e = tid; e < (nz_rows << lm); e += nt) {
                const int i = e >> lm, k = e & (M - 1);
                const int E = X[0][i][k], O = X[1][i][M - 1 - k];
                tmp[i][k] = (int16_t)clip3(-32768, 32767, round2(clip3(-32768, 32767, E + O), row_shift));
                tmp[i][w - 1 - k] = (int16_t)clip3(-32768, 32767, round2(clip3(-32768, 32767, E - O), row_shift));
            }
        } else {
            for (int i = tid; i < nz_rows; i += nt) {
                if (lw == 5) itx::row_identity32(c + i * tw, nz_cols, tmp[i], rect, row_shift);
                else if (lw == 4) itx::row_pass<4>(c + i * tw, tw, nz_cols, tmp[i], rk, rect, row_shift);
                else itx::row_pass<3>(c + i * tw, tw, nz_cols, tmp[i], rk, rect, row_shift);
            }
        }
        __syncthreads();
        // ---- column pass
        if (ck == itx::K_DCT && lh >= 5) {
            for (int e = tid; e < 2 * w; e += nt) { // first w lanes: even halves of the columns, next w: odd halves
                const int half = e >= w ? 1 : 0, j = e - half * w;
                if (lh == 6) itx::col_half<6>(&tmp[0][j], TP, nz_rows, half, X[half][j]);
                else itx::col_half<5>(&tmp[0][j], TP, nz_rows, half, X[half][j]);
            }
            __syncthreads();
            const int M = h >> 1;
            for (int e = tid; e < (M << lw); e += nt) {
                const int k = e >> lw, j = e & (w - 1); // j fastest: coalesced residual stores
                const int E = X[0][j][k], O = X[1][j][M - 1 - k];
                const int v0 = clip3(-32768, 32767, round2(clip3(-32768, 32767, E + O), col_shift));
                const int v1 = clip3(-32768, 32767, round2(clip3(-32768, 32767, E - O), col_shift));
                const int jo = flr ? (w - 1 - j) : j;
                const int r0 = fud ? (h - 1 - k) : k, r1 = fud ? k : (h - 1 - k);
                out[(size_t)r0 * out_stride + jo] = (int16_t)v0;
                out[(size_t)r1 * out_stride + jo] = (int16_t)v1;
            }
        } else {
            for (int j = tid; j < w; j += nt) {
                const int jo = flr ? (w - 1 - j) : j;
                if (lh == 5) itx::col_identity32(&tmp[0][j], TP, nz_rows, out + jo, out_stride, fud, col_shift);
                else if (lh == 4) itx::col_pass<4>(&tmp[0][j], TP, nz_rows, out + jo, out_stride, fud, ck, col_shift);
                else itx::col_pass<3>(&tmp[0][j], TP, nz_rows, out + jo, out_stride, fud, ck, col_shift);
            }
        }
        __syncthreads(); // tmp / X are reused by the next block
    }
}

// The transform blocks of one size class CLS (0: 4x4, 1: max dim 8, 2: max dim 16, 3: 32 and 64),
// CTA `bid` of `nblocks`.
template <int CLS>
AV1B_DEV void itx_cta(const Av1bOp* __restrict__ ops, const uint32_t* __restrict__ list, uint32_t n, const int16_t* __restrict__ coef,
    int16_t* __restrict__ res, const ResPlanes& rp, unsigned bid, unsigned nblocks)
{
    if (CLS == 3) {
        itx_cta_big(ops, list, n, coef, res, rp, bid, nblocks);
        return;
    }
    constexpr int GMAX = 4 << (CLS == 3 ? 2 : CLS); // (class 3 never gets here: keep its big transforms out of this code)
    constexpr int TSTRIDE = GMAX + 2;
    constexpr int REGION = GMAX * (GMAX + 2);
    __shared__ int16_t tmp_all[ITX_WARPS][(32 / GMAX) * REGION];
    const int nl = min(32u, blockDim.x);
    const int nw = max(1u, blockDim.x / 32);
    const int lane = threadIdx.x % nl;
    const int warp = threadIdx.x / nl;
    const int G = min(GMAX, nl);    // lanes per transform block
    const int per = max(1, nl / G); // transform blocks per warp pass
    const int gl = lane % G, grp = lane / G;
    int16_t* tmp = tmp_all[warp] + grp * REGION;
    for (uint32_t t0 = (bid * nw + warp) * per; t0 < n; t0 += nblocks * nw * per) {
        const uint32_t t = t0 + grp;
        if (t < n) {
            const Av1bOp op = ops[list[t]];
            itx_block<(CLS == 3) ? 4 : CLS + 2>(op, coef, res, rp, tmp, TSTRIDE, gl, G);
        }
        __syncwarp();
    }
}

// One launch per size class: small transforms compile to few registers and run at full occupancy.
template <int CLS>
__global__ void __launch_bounds__(ITX_WARPS * 32)
    itx_kernel(const Av1bOp* __restrict__ ops, const uint32_t* __restrict__ list, uint32_t n,
        const int16_t* __restrict__ coef, int16_t* __restrict__ res, ResPlanes rp)
{
    itx_cta<CLS>(ops, list, n, coef, res, rp, blockIdx.x, gridDim.x);
}

// Small frames: all four classes in ONE launch (CTA ranges per class) -- a CIF frame's inverse
// transform is a few hundred CTAs, and four launches of a few microseconds each cost more than the
// occupancy the per-class kernels buy.
struct ItxPlan {
    uint32_t first[4], cnt[4], cta_end[4];
};
__global__ void __launch_bounds__(ITX_WARPS * 32)
    itx_kernel_all(const Av1bOp* __restrict__ ops, const uint32_t* __restrict__ list, const int16_t* __restrict__ coef,
        int16_t* __restrict__ res, ResPlanes rp, ItxPlan p)
{
    const unsigned b = blockIdx.x;
    if (b < p.cta_end[0]) itx_cta<0>(ops, list + p.first[0], p.cnt[0], coef, res, rp, b, p.cta_end[0]);
    else if (b < p.cta_end[1]) itx_cta<1>(ops, list + p.first[1], p.cnt[1], coef, res, rp, b - p.cta_end[0], p.cta_end[1] - p.cta_end[0]);
    else if (b < p.cta_end[2]) itx_cta<2>(ops, list + p.first[2], p.cnt[2], coef, res, rp, b - p.cta_end[1], p.cta_end[2] - p.cta_end[1]);
    else itx_cta<3>(ops, list + p.first[3], p.cnt[3], coef, res, rp, b - p.cta_end[2], p.cta_end[3] - p.cta_end[2]);
}

// ------------------------------------------------------------------------------------------
// inter prediction (independent pass)
// ------------------------------------------------------------------------------------------
AV1B_DEV void setup_mc_params(const ReconCtx& c, const Av1bFrameHdr* hdr, const Av1bIpu& u, mc::Params& P)
{
    const int plane = u.plane, sub = plane ? 1 : 0;
    P.hdr = hdr;
    P.aux = (u.aux != 0xFFFFFFFFu) ? ((const Av1bBlkAux*)(c.cmd + hdr->off_aux) + u.aux) : nullptr;
    P.wedge = c.wedge;
    // mask of the block this unit belongs to (only compound diff-weighted blocks touch it: they
    // are >= 8x8, so the luma origin is exactly the unit origin scaled back to luma)
    P.mask = c.mask ? c.mask + (size_t)(u.y << sub) * c.mask_pitch + (u.x << sub) : nullptr;
    P.mask_pitch = c.mask_pitch;
    P.dst = c.cur.pl[plane];
    P.res = ((u.flags & AV1B_IPUF_ADD_RES) && c.rp[0]) ? c.rp[plane] : nullptr;
    P.rpitch = c.rpitch[plane];
    const int nl = (u.flags & AV1B_IPUF_COMPOUND) ? 2 : 1;
    for (int l = 0; l < nl; l++) {
        mc::RefPlane& R = P.ref[l];
        if (u.flags & AV1B_IPUF_INTRABC) {
            R.p = c.cur.pl[plane].p;
            R.stride = c.cur.pl[plane].stride;
            R.last_x = ((hdr->ref_w[0] + sub) >> sub) - 1;
            R.last_y = ((hdr->ref_h[0] + sub) >> sub) - 1;
            R.coherent = true;
        } else {
            const PlaneView& pv = c.ref[u.ref_slot[l]].pl[plane];
            R.p = pv.p;
            R.stride = pv.stride;
            R.last_x = ((hdr->ref_w[u.ref_frame[l]] + sub) >> sub) - 1;
            R.last_y = ((hdr->ref_h[u.ref_frame[l]] + sub) >> sub) - 1;
            R.coherent = false;
        }
    }
}

// One WARP per inter block: its prediction units run back to back (prediction, then OBMC strips,
// in emission order), then the residual is added.  No CTA-wide barrier: everything is ordered by
// __syncwarp, each warp owns a private scratch area, and 8x8 blocks do not idle a whole CTA.
enum { INTER_WARPS = 8 };

static __device__ void inter_block_body(const ReconCtx& c, mc::Scratch* M, unsigned bx, unsigned gx)
{
    const Av1bFrameHdr* hdr = (const Av1bFrameHdr*)c.cmd;
    const Av1bInterBlk* blks = (const Av1bInterBlk*)(c.cmd + hdr->off_iblk);
    const Av1bIpu* ipus = (const Av1bIpu*)(c.cmd + hdr->off_ipu);
    const int nl = min(32u, blockDim.x), nw = max(1u, blockDim.x / 32);
    const int lane = threadIdx.x % nl, warp = threadIdx.x / nl;
    mc::Scratch& S = M[warp];
    for (unsigned b = bx * nw + warp; b < hdr->n_iblk; b += gx * nw) {
        const Av1bInterBlk blk = blks[b];
        if (blk.flags & AV1B_IBF_FAST) continue; // inter_fast_kernel's
        for (unsigned k = 0; k < blk.n_ipu; k++) {
            const Av1bIpu u = ipus[blk.first_ipu + k];
            mc::Params P;
            setup_mc_params(c, hdr, u, P);
            mc::run_ipu(P, u, S, lane, nl);
        }
        if ((blk.flags & AV1B_IBF_ADD_RESIDUAL) && c.rp[0]) {
            // plain inter block: reconstruction = prediction + residual, no ordering constraint
            const int np = (blk.flags & AV1B_IBF_HAS_CHROMA) ? 3 : 1;
            for (int plane = 0; plane < np; plane++) {
                const int bx = plane ? blk.cx : blk.x, by = plane ? blk.cy : blk.y;
                const int bw = plane ? blk.cw : blk.bw, bh = plane ? blk.ch : blk.bh;
                const int lbw = mc::ilog2_pow2(bw);
                const PlaneView dst = c.cur.pl[plane];
                const int16_t* rp = c.rp[plane];
                const int rpitch = c.rpitch[plane];
                for (int e = lane; e < (bh << lbw); e += nl) {
                    const int i = e >> lbw, j = e & (bw - 1);
                    const int r = rp[(size_t)(by + i) * rpitch + bx + j];
                    if (r) {
                        volatile uint8_t* d = dst.p + (size_t)(by + i) * dst.stride + bx + j;
                        *d = (uint8_t)clip_u8((int)*d + r);
                    }
                }
            }
        }
        block_sync(nl);
    }
}

// Units of blocks whose units are independent of each other (no OBMC strips, no diff-weighted mask
// shared between planes) but need the general predictor -- warped motion, wedge masks, 2-sample-wide
// chroma, inter-intra's inter half: one WARP PER UNIT instead of per block (three times the
// parallelism, a third of the serial chain), residual added in the same store.
static __device__ void inter_unit_body(const ReconCtx& c, mc::Scratch* M, unsigned bx, unsigned gx)
{
    const Av1bFrameHdr* hdr = (const Av1bFrameHdr*)c.cmd;
    const Av1bIpu* ipus = (const Av1bIpu*)(c.cmd + hdr->off_ipu);
    const unsigned n_ipu = hdr->n_ipu;
    const int nl = min(32u, blockDim.x), nw = max(1u, blockDim.x / 32);
    const int lane = threadIdx.x % nl, warp = threadIdx.x / nl;
    mc::Scratch& S = M[warp];
    // a CTA takes 32 consecutive units; every warp scans them (a lane per unit) and the flagged
    // ones are dealt round-robin to the warps
    for (unsigned chunk = bx * nl; chunk < n_ipu; chunk += gx * nl) {
        const unsigned idx = chunk + lane;
        const bool mine = idx < n_ipu && (ipus[idx].flags & AV1B_IPUF_INDEP);
        unsigned todo = __ballot_sync(0xFFFFFFFFu, mine);
        for (int turn = 0; todo; turn++) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            if (turn % nw != warp) continue;
            const Av1bIpu u = ipus[chunk + j];
            mc::Params P;
            setup_mc_params(c, hdr, u, P);
            mc::run_ipu(P, u, S, lane, nl);
        }
    }
}

// Fast path for the units of AV1B_IBF_FAST blocks (plain translational prediction, the bulk of
// any inter frame).  Units are independent, so the kernel walks the UNIT list, a warp taking 32
// consecutive units at a time: each lane derives one unit's parameters (reference plane, clamp
// limits, integer position, packed filter taps), then the warp runs the flagged units one after
// the other with the parameters broadcast by shuffle -- the per-unit scalar set-up, which would
// otherwise cost a full warp instruction per value per unit, is paid once per 32 units.  The
// residual of plain inter blocks is added in the same store (AV1B_IPUF_ADD_RES).
#ifdef AV1B_EMU
enum { FAST_WARPS = 4, FAST_SLICES = 4, FAST_JOBS_PER_SLICE = 2 }; // one unit per chunk under emulation: still exercise the slicing
#else
enum { FAST_WARPS = 4, FAST_SLICES = 4, FAST_JOBS_PER_SLICE = 48 };
#endif

static __device__ void inter_fast_body(const ReconCtx& c, mc::Scratch* M, unsigned bx, unsigned gx, unsigned by, unsigned gy)
{
    const Av1bFrameHdr* hdr = (const Av1bFrameHdr*)c.cmd;
    const uint4* ipus = (const uint4*)(c.cmd + hdr->off_ipu);
    const unsigned n_ipu = hdr->n_ipu;
    const int nl = min(32u, blockDim.x), nw = max(1u, blockDim.x / 32);
    const int lane = threadIdx.x % nl, warp = threadIdx.x / nl;
    const unsigned FULL = 0xFFFFFFFFu;
    mc::Scratch& S = M[warp];
    for (unsigned chunk = bx * nl; chunk < n_ipu; chunk += gx * nl) {
        // ---- lane-parallel set-up: lane i <-> unit chunk + i
        const unsigned idx = chunk + lane;
        uint32_t u_xy = 0, u_dim = 0, u_fl = 0;   // x|y<<16, w|h<<8|plane<<16|kind<<24, flags|comp|fwd|bck
        uint32_t r_lo[2] = { 0, 0 }, r_hi[2] = { 0, 0 }, r_stride[2] = { 0, 0 }, r_last[2] = { 0, 0 };
        int r_px[2] = { 0, 0 }, r_py[2] = { 0, 0 };
        uint4 r_taps[2] = {};
        uint32_t subpel = 0;
        bool fast = false;
        uint4 a = make_uint4(0, 0, 0, 0), b = a;
        if (idx < n_ipu) {
            a = __ldg(ipus + 2 * idx), b = __ldg(ipus + 2 * idx + 1);
            u_xy = a.x, u_dim = a.y, u_fl = b.z;
            fast = (u_fl & AV1B_IPUF_FAST) != 0;
        }
        // ---- jobs = (unit, tile) pairs in unit order.  A chunk of large units holds up to 8x the
        // jobs of a chunk of small ones: it is shared by up to FAST_SLICES CTAs (by), each
        // repeating the set-up and taking every n_slices-th group of jobs; the CTAs a light chunk
        // does not need leave here.
        const int my_w = u_dim & 0xFF, my_h = (u_dim >> 8) & 0xFF;
        const int my_tiles = fast ? ((my_w + mc::TILE_W - 1) / mc::TILE_W) * ((my_h + mc::TILE_H - 1) / mc::TILE_H) : 0;
        int incl = my_tiles;
        for (int d = 1; d < nl; d <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        const int excl = incl - my_tiles;
        const int total = __shfl_sync(FULL, incl, nl - 1);
        const int n_slices = min((int)gy, (total + FAST_JOBS_PER_SLICE - 1) / FAST_JOBS_PER_SLICE);
        if ((int)by >= n_slices) continue;
        {
            if (fast) {
                const int x = u_xy & 0xFFFF, y = u_xy >> 16, w = u_dim & 0xFF, h = (u_dim >> 8) & 0xFF;
                const int plane = (u_dim >> 16) & 0xFF, sub = plane ? 1 : 0;
                const int lists = (u_fl & AV1B_IPUF_COMPOUND) ? 2 : 1;
                const int filt_v = b.y & 0xFF, filt_h = (b.y >> 8) & 0xFF;
                for (int l = 0; l < lists; l++) {
                    const int slot = (int8_t)((b.x >> (8 * l)) & 0xFF), rf = (b.x >> (16 + 8 * l)) & 0xFF;
                    const PlaneView& pv = c.ref[slot & 7].pl[plane];
                    r_lo[l] = (uint32_t)(uintptr_t)pv.p;
                    r_hi[l] = (uint32_t)((uintptr_t)pv.p >> 32);
                    r_stride[l] = (uint32_t)pv.stride;
                    r_last[l] = (uint32_t)(((hdr->ref_w[rf] + sub) >> sub) - 1) | ((uint32_t)(((hdr->ref_h[rf] + sub) >> sub) - 1) << 16);
                    const uint32_t mvw = l ? a.w : a.z;
                    const int mvx = (2 * (int)(int16_t)(mvw >> 16)) >> sub, mvy = (2 * (int)(int16_t)(mvw & 0xFFFF)) >> sub;
                    const int fx = mvx & 15, fy = mvy & 15;
                    r_px[l] = x + (mvx >> 4);
                    r_py[l] = y + (mvy >> 4);
                    if (fx | fy) subpel |= 1u << l;
                    const uint32_t* th = k_subpel_packed[mc::filter_row(w, filt_h)][fx];
                    const uint32_t* tv = k_subpel_packed[mc::filter_row(h, filt_v)][fy];
                    r_taps[l] = make_uint4(th[0], th[1], tv[0], tv[1]);
                }
            }
        }
        // the slice's jobs, dealt round-robin to its warps; parameters are broadcast from the unit's lane
        for (int job = (int)by * nw + warp; job < total; job += n_slices * nw) {
            const int j = 31 - __clz(__ballot_sync(FULL, excl <= job));
            const int tile = job - __shfl_sync(FULL, excl, j);
            const uint32_t xy = __shfl_sync(FULL, u_xy, j), dim = __shfl_sync(FULL, u_dim, j), fl = __shfl_sync(FULL, u_fl, j);
            const uint32_t sp = __shfl_sync(FULL, subpel, j);
            const int x = xy & 0xFFFF, y = xy >> 16, w = dim & 0xFF, h = (dim >> 8) & 0xFF, plane = (dim >> 16) & 0xFF;
            const bool compound = (fl & AV1B_IPUF_COMPOUND) != 0;
            const int comp = (fl >> 8) & 0xFF;
            const int round1 = compound ? 7 : 11;
            mc::RefPlane R[2];
            int px[2], py[2];
            uint4 taps[2];
            AV1B_UNROLL
            for (int l = 0; l < 2; l++) {
                if (l && !compound) break;
                R[l].p = (const uint8_t*)(uintptr_t)((uint64_t)__shfl_sync(FULL, r_lo[l], j) | ((uint64_t)__shfl_sync(FULL, r_hi[l], j) << 32));
                R[l].stride = (int)__shfl_sync(FULL, r_stride[l], j);
                const uint32_t last = __shfl_sync(FULL, r_last[l], j);
                R[l].last_x = last & 0xFFFF;
                R[l].last_y = last >> 16;
                R[l].coherent = false;
                px[l] = __shfl_sync(FULL, r_px[l], j);
                py[l] = __shfl_sync(FULL, r_py[l], j);
                taps[l].x = __shfl_sync(FULL, r_taps[l].x, j);
                taps[l].y = __shfl_sync(FULL, r_taps[l].y, j);
                taps[l].z = __shfl_sync(FULL, r_taps[l].z, j);
                taps[l].w = __shfl_sync(FULL, r_taps[l].w, j);
            }
            const PlaneView dst = c.cur.pl[plane];
            const int16_t* res = ((fl & AV1B_IPUF_ADD_RES) && c.rp[0]) ? c.rp[plane] : nullptr;
            const int rpitch = c.rpitch[plane];
            // blend weights: single (1,0,>>0), average (8,8,>>8 == (p0+p1+16)>>5), distance (fwd,bck,>>8)
            const int w0 = !compound ? 1 : (comp == AV1B_COMP_AVERAGE ? 8 : (int)((fl >> 16) & 0xFF));
            const int w1 = !compound ? 0 : (comp == AV1B_COMP_AVERAGE ? 8 : (int)(fl >> 24));
            const int sh = compound ? 8 : 0;
            {
                // w is a power of two: tiles per row is 1, 2 or 4
                const int ltr = w > 64 ? 2 : (w > 32 ? 1 : 0);
                const int ty = (tile >> ltr) * mc::TILE_H, tx = (tile & ((1 << ltr) - 1)) * mc::TILE_W;
                const int th = min((int)mc::TILE_H, h - ty);
                const int tw = min((int)mc::TILE_W, w - tx);
                const int ltw = mc::ilog2_pow2(tw), lq = ltw - 2;
                if (mc::fast_tile_ok(R[0], px[0] + tx, py[0] + ty, tw, th) && (!compound || mc::fast_tile_ok(R[1], px[1] + tx, py[1] + ty, tw, th))) {
                    mc::FastScratch& F = *(mc::FastScratch*)&S;
                    const int shi = 14 - 3 - round1; // integer-position prediction = sample << shi
                    mc::fast_h(R[0], px[0] + tx, py[0] + ty, (sp & 1) != 0, taps[0].x, taps[0].y, ltw, th, shi, F.inter[0], lane, nl);
                    if (compound) mc::fast_h(R[1], px[1] + tx, py[1] + ty, (sp & 2) != 0, taps[1].x, taps[1].y, ltw, th, shi, F.inter[1], lane, nl);
                    block_sync(nl);
                    const int rnd = 1 << (round1 - 2), shv = round1 - 1;
                    for (int e = lane; e < ((th >> 1) << lq); e += nl) {
                        const int k = e >> lq, q = e & ((1 << lq) - 1);
                        uint32_t row0 = 0, row1 = 0;
                        AV1B_UNROLL
                        for (int i = 0; i < 4; i++) {
                            int a0, a1;
                            mc::fast_v(F.inter[0], 4 * q + i, k, (sp & 1) != 0, taps[0].z, taps[0].w, rnd, shv, a0, a1);
                            if (compound) {
                                int b0, b1;
                                mc::fast_v(F.inter[1], 4 * q + i, k, (sp & 2) != 0, taps[1].z, taps[1].w, rnd, shv, b0, b1);
                                a0 = round2(w0 * a0 + w1 * b0, 8);
                                a1 = round2(w0 * a1 + w1 * b1, 8);
                            }
                            row0 |= (uint32_t)clip_u8(a0) << (8 * i);
                            row1 |= (uint32_t)clip_u8(a1) << (8 * i);
                        }
                        const int yy = y + ty + 2 * k, xx = x + tx + 4 * q;
                        if (res) {
                            const uint2 r0 = *(const uint2*)(res + (size_t)yy * rpitch + xx), r1 = *(const uint2*)(res + (size_t)(yy + 1) * rpitch + xx);
                            row0 = add_res4(row0, r0);
                            row1 = add_res4(row1, r1);
                        }
                        uint8_t* d = dst.p + (size_t)yy * dst.stride + xx;
                        *(uint32_t*)d = row0;
                        *(uint32_t*)(d + dst.stride) = row1;
                    }
                    block_sync(nl);
                    continue;
                }
                // tile touching the reference border: staged, clamped window
                {
                    mc::convolve_tile(R[0], px[0] + tx, py[0] + ty, (sp & 1) != 0, taps[0], tw, th, round1, S, S.pred[0], lane, nl);
                    if (compound) mc::convolve_tile(R[1], px[1] + tx, py[1] + ty, (sp & 2) != 0, taps[1], tw, th, round1, S, S.pred[1], lane, nl);
                    for (int e = lane; e < (th << lq); e += nl) {
                        const int r = e >> lq, q = e & ((1 << lq) - 1);
                        const uint2 p0 = *(const uint2*)(S.pred[0] + r * mc::TILE_W + 4 * q);
                        uint2 p1 = make_uint2(0, 0);
                        if (compound) p1 = *(const uint2*)(S.pred[1] + r * mc::TILE_W + 4 * q);
                        int o0 = clip_u8(round2(w0 * (int)(int16_t)(p0.x & 0xFFFF) + w1 * (int)(int16_t)(p1.x & 0xFFFF), sh));
                        int o1 = clip_u8(round2(w0 * ((int)p0.x >> 16) + w1 * ((int)p1.x >> 16), sh));
                        int o2 = clip_u8(round2(w0 * (int)(int16_t)(p0.y & 0xFFFF) + w1 * (int)(int16_t)(p1.y & 0xFFFF), sh));
                        int o3 = clip_u8(round2(w0 * ((int)p0.y >> 16) + w1 * ((int)p1.y >> 16), sh));
                        const int yy = y + ty + r, xx = x + tx + 4 * q;
                        if (res) {
                            const uint2 rr = *(const uint2*)(res + (size_t)yy * rpitch + xx);
                            o0 = clip_u8(o0 + (int)(int16_t)(rr.x & 0xFFFF));
                            o1 = clip_u8(o1 + ((int)rr.x >> 16));
                            o2 = clip_u8(o2 + (int)(int16_t)(rr.y & 0xFFFF));
                            o3 = clip_u8(o3 + ((int)rr.y >> 16));
                        }
                        *(uint32_t*)(dst.p + (size_t)yy * dst.stride + xx) = (uint32_t)o0 | ((uint32_t)o1 << 8) | ((uint32_t)o2 << 16) | ((uint32_t)o3 << 24);
                    }
                    block_sync(nl);
                }
            }
        }
    }
}

// One launch for the whole inter pass: the three kinds of work touch disjoint blocks, so their CTAs
// run side by side instead of one kernel after the other (each of them alone is a set of serial
// chains that leaves most of the GPU idle).  The CTAs with the long chains come first.
__global__ void __launch_bounds__(INTER_WARPS * 32, 3) inter_kernel(ReconCtx c, unsigned g_block, unsigned g_unit, unsigned g_fast)
{
    __shared__ mc::Scratch M[INTER_WARPS];
    unsigned b = blockIdx.x;
    if (b < g_block) {
        inter_block_body(c, M, b, g_block);
        return;
    }
    b -= g_block;
    if (b < g_unit) {
        inter_unit_body(c, M, b, g_unit);
        return;
    }
    b -= g_unit;
    inter_fast_body(c, M, b % g_fast, g_fast, b / g_fast, FAST_SLICES);
}

// ------------------------------------------------------------------------------------------
// dependent pass (superblock wavefront)
// ------------------------------------------------------------------------------------------
namespace {

// Where the op executor reads and writes samples / residuals of one plane.
struct PlaneIo {
    uint8_t* pix;        // address of sample (0,0) in frame coordinates (may point into the smem tile)
    int pitch;
    const int16_t* res;  // address of residual (0,0) in frame coordinates, or null
    int rpitch;
};

// The three planes of a frame (or of a superblock tile) as scalars picked by select: no array, so
// the pointers stay in registers and the compiler can see that a tile lives in shared memory
// (LDS / STS instead of generic accesses).
struct PlaneSet {
    uint8_t *pix0, *pix1, *pix2;
    int pitch0, pitch12;
    const int16_t *res0, *res1, *res2;
    int rpitch0, rpitch12;
};

enum { WAVE_OP_CHUNK = 128 };
#ifdef AV1B_EMU
enum { WAVE_NT = 0 }; // the emulation runs every group with one thread
#else
enum { WAVE_NT = 32 };
#endif

struct OpScratch {
    intra::Scratch I;
};

// Frame constants the op executor needs, read from the header once per kernel.
struct FrameConst {
    int max_x[2], max_y[2]; // ((MiCols*4)>>sub)-1, ((MiRows*4)>>sub)-1 for luma / chroma
    bool edge_filter;
};

AV1B_DEV FrameConst frame_const(const Av1bFrameHdr* hdr)
{
    FrameConst f;
    for (int sub = 0; sub < 2; sub++) {
        f.max_x[sub] = ((hdr->mi_cols * 4) >> sub) - 1;
        f.max_y[sub] = ((hdr->mi_rows * 4) >> sub) - 1;
    }
    f.edge_filter = hdr->enable_intra_edge_filter != 0;
    return f;
}

// The reference's view of an intra transform block, from the op record.
AV1B_DEV intra::Args intra_args(const Av1bOp& op, const FrameConst& fc, int lw, int lh)
{
    const int sub = op.plane ? 1 : 0;
    intra::Args a;
    a.x = op.x;
    a.y = op.y;
    a.log2w = lw;
    a.log2h = lh;
    a.max_x = sub ? fc.max_x[1] : fc.max_x[0];
    a.max_y = sub ? fc.max_y[1] : fc.max_y[0];
    a.plane_idx = op.plane;
    a.mode = op.mode;
    a.angle_delta = op.angle_delta;
    a.have_left = (op.flags & AV1B_OPF_HAVE_LEFT) != 0;
    a.have_above = (op.flags & AV1B_OPF_HAVE_ABOVE) != 0;
    a.have_above_right = (op.flags & AV1B_OPF_HAVE_ABOVE_RIGHT) != 0;
    a.have_below_left = (op.flags & AV1B_OPF_HAVE_BELOW_LEFT) != 0;
    a.edge_filter_enabled = fc.edge_filter;
    a.edge_smooth = (op.flags & AV1B_OPF_EDGE_SMOOTH) != 0;
    a.filter_intra = (op.flags & AV1B_OPF_FILTER_INTRA) != 0;
    a.fi_mode = op.fi_mode & 7;
    a.strip = (op.flags & AV1B_OPF_FILTER_INTRA) ? 0 : op.fi_mode >> 3;
    a.cfl = op.kind == AV1B_OP_INTRA && (op.flags & AV1B_OPF_CFL) != 0;
    a.cfl_alpha = op.cfl_alpha;
    a.max_luma_w = op.max_luma_w;
    a.max_luma_h = op.max_luma_h;
    return a;
}

// Intra ops of the superblock wavefront are DECODED in shared memory ahead of their turn (lane-
// parallel, while the CTA still waits for its neighbours): the 32-byte record keeps x, y, plane
// and the level word and carries intra::Packed in place of the raw fields; kind gets bit 7.
enum { WAVE_OP_DECODED = 0x80 };
AV1B_DEV void decode_intra_op(Av1bOp* slot, const FrameConst& fc)
{
    const Av1bOp op = *slot;
    if (op.kind != AV1B_OP_INTRA) return;
    const int lw = (int)((AV1T_TX_WLOG2_PACKED >> (3 * op.tx_size)) & 7), lh = (int)((AV1T_TX_HLOG2_PACKED >> (3 * op.tx_size)) & 7);
    const intra::Packed k = intra::pack(intra::prepare(intra_args(op, fc, lw, lh)));
    uint32_t* w = (uint32_t*)slot;
    w[1] = (uint32_t)op.plane | ((uint32_t)(AV1B_OP_INTRA | WAVE_OP_DECODED) << 8);
    w[2] = k.w[0];
    w[3] = k.w[1];
    w[4] = k.w[2];
    w[6] = k.w[3] | ((op.flags & AV1B_OPF_HAS_RESID) ? 0x10000u : 0u); // lim_w, lim_h in the low half
}

// An op fetched from shared memory and, for decoded intra ops, the addresses it works on -- all a
// warp can know about its NEXT op before the level barrier opens.
struct StagedOp {
    uint4 wa, wb;
    intra::Io o;
    int state; // 0 none, 1 decoded intra op (o is valid), 2 any other op
};
AV1B_DEV StagedOp stage_op(const Av1bOp* slot, const PlaneSet& io)
{
    StagedOp s;
    s.wa = ((const uint4*)slot)[0];
    s.wb = ((const uint4*)slot)[1];
    s.state = 2;
    if (s.wa.y & (WAVE_OP_DECODED << 8)) {
        const int x = (int)(s.wa.x & 0xFFFF), y = (int)(s.wa.x >> 16), plane = (int)(s.wa.y & 0xFF);
        uint8_t* const pix = plane == 0 ? io.pix0 : (plane == 1 ? io.pix1 : io.pix2);
        const int pitch = plane == 0 ? io.pitch0 : io.pitch12;
        const int16_t* const rbase = plane == 0 ? io.res0 : (plane == 1 ? io.res1 : io.res2);
        const int rpitch = plane == 0 ? io.rpitch0 : io.rpitch12;
        s.o.P = pix + (ptrdiff_t)y * pitch + x;
        s.o.blk = s.o.P;
        s.o.stride = s.o.pp = pitch;
        s.o.res = ((s.wb.z & 0x10000u) && rbase) ? rbase + (ptrdiff_t)y * rpitch + x : nullptr;
        s.o.rpitch = rpitch;
        s.o.luma = io.pix0 + (ptrdiff_t)(2 * y) * io.pitch0 + 2 * x;
        s.o.luma_stride = io.pitch0;
        s.state = 1;
    }
    return s;
}

// SMEM: the planes in `io` are the superblock tile in shared memory (wave_kernel) -- said to the
// compiler so that sample accesses become shared-memory instructions instead of generic ones.
// NTC: compile-time group size (32: one warp per op) or 0 = runtime nt_rt.
template <bool SMEM, int NTC>
AV1B_DEV void exec_op(const ReconCtx& c, const Av1bFrameHdr* hdr, const FrameConst& fc, const Av1bOp& op, const PlaneSet& io, OpScratch& S,
    mc::Scratch* M, int tid, int nt_rt)
{
    const int nt = NTC ? NTC : nt_rt;
    const int plane = op.plane;
    PlaneIo D;
    D.pix = plane == 0 ? io.pix0 : (plane == 1 ? io.pix1 : io.pix2);
    D.pitch = plane == 0 ? io.pitch0 : io.pitch12;
    D.res = plane == 0 ? io.res0 : (plane == 1 ? io.res1 : io.res2);
    D.rpitch = plane == 0 ? io.rpitch0 : io.rpitch12;
    uint8_t* const pix = D.pix;
    uint8_t* const luma_pix = io.pix0;
    int lw, lh;
    if (op.kind == AV1B_OP_INTERINTRA || op.kind == AV1B_OP_INTRABC) {
        lw = op.tx_size & 15;
        lh = op.tx_size >> 4;
    } else {
        lw = (int)((AV1T_TX_WLOG2_PACKED >> (3 * op.tx_size)) & 7);
        lh = (int)((AV1T_TX_HLOG2_PACKED >> (3 * op.tx_size)) & 7);
    }
    const int w = 1 << lw, h = 1 << lh;
    const bool has_res = (op.flags & AV1B_OPF_HAS_RESID) && D.res;
    uint8_t* dst = pix + (ptrdiff_t)op.y * D.pitch + op.x;
    const int16_t* res = has_res ? (D.res + (ptrdiff_t)op.y * D.rpitch + op.x) : nullptr;
    switch (op.kind) {
    case AV1B_OP_INTER_RES: {
        if (!res) break;
        const int lq = lw - 2;
        for (int e = tid; e < (h << lq); e += nt) {
            const int i = e >> lq, q = e & ((1 << lq) - 1);
            uint32_t* d = (uint32_t*)(dst + i * D.pitch + 4 * q);
            *d = add_res4(SMEM ? *d : __ldcg(d), *(const uint2*)(res + i * D.rpitch + 4 * q));
        }
        break;
    }
    case AV1B_OP_INTRA:
    case AV1B_OP_INTERINTRA: {
        // Intra blocks are predicted straight into their place, residual and CfL fused into the
        // store.  Inter-intra blocks predict into scratch (the place holds the inter half).
        const bool in_place = op.kind == AV1B_OP_INTRA;
        const intra::Packed p = intra::pack(intra::prepare(intra_args(op, fc, lw, lh)));
        intra::Io o;
        o.blk = dst;
        o.stride = D.pitch;
        o.P = in_place ? dst : S.I.pred;
        o.pp = in_place ? D.pitch : w;
        o.res = in_place ? res : nullptr;
        o.rpitch = D.rpitch;
        o.luma = luma_pix + (ptrdiff_t)(2 * op.y) * io.pitch0 + 2 * op.x;
        o.luma_stride = io.pitch0;
        intra::run<SMEM, NTC>(p, o, S.I, tid, nt);
        if (!in_place) {
            // inter-intra blend over the inter prediction already in place
            // (reference maskBlend, InterPredict.cpp:584-609; masks :555-582, :888-899)
            const Av1bBlkAux* aux = (const Av1bBlkAux*)(c.cmd + hdr->off_aux) + op.aux;
            const uint8_t* W = aux->wedge_interintra ? mc::wedge_mask_ptr(c.wedge, aux->mi_size, aux->wedge_sign, aux->wedge_index) : nullptr;
            const int scale = 128 / max(w, h);
            const int iim = aux->interintra_mode;
            for (int e = tid; e < w * h; e += nt) {
                const int i = e >> lw, j = e & (w - 1);
                int m;
                if (W) {
                    if (!plane) m = W[i * 32 + j];
                    else
                        m = (W[(2 * i) * 32 + 2 * j] + W[(2 * i) * 32 + 2 * j + 1] + W[(2 * i + 1) * 32 + 2 * j]
                                + W[(2 * i + 1) * 32 + 2 * j + 1] + 2)
                            >> 2;
                } else if (iim == 1) m = k_ii_weights_1d[i * scale];
                else if (iim == 2) m = k_ii_weights_1d[j * scale];
                else if (iim == 3) m = k_ii_weights_1d[min(i, j) * scale];
                else m = 32;
                volatile uint8_t* d = dst + i * D.pitch + j;
                const int inter = *d;
                *d = (uint8_t)clip_u8(round2(m * S.I.pred[e] + (64 - m) * inter, 6));
            }
        }
        break;
    }
    case AV1B_OP_PALETTE: {
        const Av1bBlkAux* aux = (const Av1bBlkAux*)(c.cmd + hdr->off_aux) + op.aux;
        const int pi = plane ? 1 : 0;
        const uint8_t* map = c.cmd + hdr->off_pal + aux->pal_map_off[pi];
        const int ms = aux->pal_map_stride[pi];
        const int ox = op.x - aux->base_x[pi], oy = op.y - aux->base_y[pi];
        const uint8_t* colors = aux->pal_colors[plane];
        for (int e = tid; e < w * h; e += nt) {
            const int i = e >> lw, j = e & (w - 1);
            int v = colors[map[(oy + i) * ms + ox + j]];
            if (res) v = clip_u8(v + res[i * D.rpitch + j]);
            dst[i * D.pitch + j] = (uint8_t)v;
        }
        break;
    }
    case AV1B_OP_INTRABC: {
        // only reached on the global-memory path (frames with allow_intrabc)
        if (SMEM) break;
        const Av1bIpu u = ((const Av1bIpu*)(c.cmd + hdr->off_ipu))[op.aux];
        mc::Params P;
        setup_mc_params(c, hdr, u, P);
        mc::run_ipu(P, u, *M, tid, nt);
        break;
    }
    }
    block_sync(nt);
}

// Superblock scheduling.  A superblock (r, c) may start once its left neighbour and the
// superblocks of the row above up to column c + lag - 1 are finished (lag = 2: the above-right
// neighbour, all an intra edge can reach; frames with intrabc use the reach of the spec's block
// vector constraint, see wave_kernel_global).  CTAs draw tickets from an atomic counter; ticket t
// is the t-th superblock in WAVEFRONT order -- sorted by d = c + lag * r, then by row -- so the
// in-flight tickets are the superblocks that can actually run together.  (Raster order would keep
// a window of consecutive tickets inside one or two rows: a 4K frame then runs ~1.5 superblocks at
// a time instead of ~30.)  Both dependencies have a smaller d, hence a smaller ticket, hence are
// already owned by a running CTA: no deadlock whatever the number of CTAs.
AV1B_DEV void sb_from_ticket(int t, int rows, int cols, int lag, int& r, int& c)
{
    for (int d = 0;; d++) {
        // superblocks on diagonal d: rows r with 0 <= d - lag * r <= cols - 1
        const int r_hi = min(rows - 1, d / lag);
        const int r_lo = d > cols - 1 ? (d - (cols - 1) + lag - 1) / lag : 0;
        const int cnt = r_hi - r_lo + 1;
        if (cnt <= 0) continue;
        if (t < cnt) {
            r = r_lo + t;
            c = d - lag * r;
            return;
        }
        t -= cnt;
    }
}

AV1B_DEV void wave_wait(int* progress, int r, int col, int sb_cols, int lag, int tid, int nt)
{
    if (tid == 0) {
        // poll with plain L2 reads; one acquire (which also drops this SM's L1 lines) once the
        // counters are there
        if (r > 0) {
            const int need = min(col + lag, sb_cols);
            while (av1b_ld_relaxed(progress + r - 1) < need) av1b_nanosleep(32);
        }
        if (col > 0) {
            while (av1b_ld_relaxed(progress + r) < col) av1b_nanosleep(32);
        }
        if (r > 0 || col > 0) (void)av1b_ld_acquire(progress + r - (r > 0 ? 1 : 0));
    }
    block_sync(nt);
}

// The CTA barrier orders every thread's stores before thread 0's release (the pattern of a
// cooperative grid barrier): one fence by one thread instead of one per thread.
AV1B_DEV void wave_signal(int* progress, int r, int col, int tid, int nt)
{
    block_sync(nt);
    if (tid == 0) av1b_st_release(progress + r, col + 1);
}


// Leaving the kernel: the last CTA out puts the counters back to zero for the next launch (every
// other CTA has drawn a ticket past the end and polls nothing any more), so no memset node per
// frame is needed.
AV1B_DEV void wave_leave(int* sync, int n_progress, int tid, int nt, int* s_flag)
{
    block_sync(nt); // s_flag is the shared ticket word: every thread has read its last ticket before it is reused
    if (tid == 0) {
        __threadfence();
        *s_flag = atomicAdd(sync + 1, 1) == (int)gridDim.x - 1;
    }
    block_sync(nt);
    if (*s_flag) {
        for (int k = tid; k < n_progress; k += nt) sync[2 + k] = 0;
        if (tid == 0) {
            sync[0] = 0;
            sync[1] = 0;
        }
    }
}

// Shared-memory footprint of one superblock (bytes) for SB size `sb` (64 or 128): the sample tile
// -- luma (sb+1) rows of 2sb+8 bytes (a halo row above, long enough for above-right reads, and a
// halo column to the left; sample (x0, y0) sits at byte 4 of row 1 so that every block row is
// word aligned), two chroma planes likewise -- and the int16 residuals of the superblock.
#define WAVE_TILE_BYTES(sb) (((((sb) + 1) * (2 * (sb) + 8) + 2 * ((sb) / 2 + 1) * ((sb) + 8)) + 15) & ~15)
#define WAVE_RES_BYTES(sb) (3 * (sb) * (sb))
#define WAVE_SMEM_BYTES(sb, warps) (WAVE_TILE_BYTES(sb) + WAVE_RES_BYTES(sb) + (int)sizeof(OpScratch) * (warps))

}  // namespace

int wave_tile_bytes(int sb) { return WAVE_TILE_BYTES(sb); }

// Superblock-in-shared-memory wavefront: the SB's samples (all planes) live in a smem tile with a
// one-sample halo row above (long enough for above-right reads) and halo column to the left, and
// so do its residuals, so the chain of intra predictions never waits on L2.  Inside the SB the ops
// come sorted by dependency level (host/emitter.cpp scheduleSb): all ops of one level are
// independent, each is executed by ONE warp, and the CTA only synchronises between levels -- a
// 128x128 SB of 4x4 blocks needs ~100 level steps instead of ~650 sequential ops.  Frames without
// intrabc only.
template <int WARPS, int MIN_CTAS>
__global__ void __launch_bounds__(WARPS * 32, MIN_CTAS) wave_kernel(ReconCtx c)
{
    alignas(16) __shared__ Av1bOp s_ops[2][WAVE_OP_CHUNK];
    __shared__ int s_sb;
    __shared__ unsigned long long s_polled;
#ifdef AV1B_EMU
    alignas(16) static uint8_t dyn[WAVE_SMEM_BYTES(128, WARPS)];
#else
    extern __shared__ __align__(16) uint8_t dyn[];
#endif
    const Av1bFrameHdr* hdr = (const Av1bFrameHdr*)c.cmd;
    const Av1bSb* sbs = (const Av1bSb*)(c.cmd + hdr->off_sb);
    const Av1bOp* ops = (const Av1bOp*)(c.cmd + hdr->off_ops);
    const FrameConst fc = frame_const(hdr);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nl = min(32u, blockDim.x), nw = max(1u, blockDim.x / 32);
    (void)nw;
    const int lane = tid % nl, warp = tid / nl;
    const int sb_cols = hdr->sb_cols, n_sb = hdr->n_sb;
    const int sbs_y = 1 << hdr->sb_log2;
    const bool load_pred = !hdr->frame_is_intra; // inter prediction already sits in the frame
    const bool have_res = c.rp[0] != nullptr && hdr->n_itx != 0;
    int* ticket = c.sync;
    int* progress = c.sync + 2;
    const int n0 = sbs_y, n1 = sbs_y >> 1;
    const int pitch0 = 2 * n0 + 8, pitch1 = 2 * n1 + 8;
    uint8_t* const t0 = dyn;
    uint8_t* const t1 = t0 + (n0 + 1) * pitch0;
    uint8_t* const t2 = t1 + (n1 + 1) * pitch1;
    int16_t* const q0 = (int16_t*)(dyn + WAVE_TILE_BYTES(n0));
    int16_t* const q1 = q0 + n0 * n0;
    int16_t* const q2 = q1 + n1 * n1;
    OpScratch* const scratch = (OpScratch*)(dyn + WAVE_TILE_BYTES(n0) + WAVE_RES_BYTES(n0)) + warp;
    // progress[sb]: bit 0 = the upper half of superblock sb's right column is final and in the frame,
    // bit 1 = the left half of its bottom row, bit 2 = all of it
    auto poll = [&](int idx, int mask) {
#ifdef AV1B_WAVE_WATCHDOG
        unsigned spins = 0;
        while (!(av1b_ld_relaxed(progress + idx) & mask)) {
            av1b_nanosleep(32);
            if (++spins == (1u << 22)) {
                printf("wave watchdog: sb %d waits for sb %d mask %d, sees %d (ticket %d, grid %d)\n", s_sb, idx, mask, av1b_ld_relaxed(progress + idx),
                    av1b_ld_relaxed(ticket), (int)gridDim.x);
                break;
            }
        }
#else
        if (c.trace) { // trace: time spent polling after the first wait, summed per superblock
            const unsigned long long t_in = av1b_gtime();
            while (!(av1b_ld_relaxed(progress + idx) & mask)) av1b_nanosleep(32);
            atomicAdd(&s_polled, (unsigned long long)(av1b_gtime() - t_in));
            return;
        }
        while (!(av1b_ld_relaxed(progress + idx) & mask)) av1b_nanosleep(32);
#endif
    };
    for (;;) {
        block_sync(nt);
        if (tid == 0) {
            const int t = atomicAdd(ticket, 1);
            int rr = 0, cc = 0;
            if (t < n_sb) sb_from_ticket(t, hdr->sb_rows, sb_cols, 2, rr, cc);
            s_sb = t < n_sb ? rr * sb_cols + cc : n_sb;
        }
        block_sync(nt);
        const int sb = s_sb;
        if (sb >= n_sb) break;
        const int r = sb / sb_cols, col = sb - r * sb_cols;
        const Av1bSb e = sbs[sb];
        unsigned long long* const tr = (c.trace && tid == 0 && (unsigned)sb < c.trace_cap) ? c.trace + 8 * (size_t)sb : nullptr;
        if (tr) tr[0] = (unsigned long long)sb, tr[1] = av1b_smid(), tr[2] = av1b_gtime();
        if (e.n_ops == 0) {
            // nothing to reconstruct: the samples the neighbours read are in the frame already
            if (tid == 0) av1b_st_release(progress + sb, 7);
            if (tr) tr[3] = tr[4] = tr[5] = tr[6] = tr[7] = av1b_gtime();
            continue;
        }
        // ---- everything that does not depend on the neighbours is requested BEFORE the wait (a
        // CTA standing by on the next diagonal has it done when its dependencies finish): the first
        // chunk of the op list, the superblock's residuals and, for inter frames, its own samples.
        {
            const uint4* src = (const uint4*)(ops + e.first_op);
            const unsigned n0c = min((unsigned)WAVE_OP_CHUNK, e.n_ops);
            uint4* dstv = (uint4*)s_ops[0];
            for (unsigned q = tid; q < n0c * 2; q += nt) dstv[q] = __ldg(src + q);
        }
        const int lg0 = mc::ilog2_pow2(n0); // 6 or 7
        auto load_own = [&](int pl, uint8_t* t, int n, int lg, int pitch) {
            // tile sample (x, y) in frame coordinates lives at t[(y - y0 + 1) * pitch + (x - x0 + 4)]
            const int x0 = col * n, y0 = r * n;
            const PlaneView g = c.cur.pl[pl];
            const int lwords = lg - 2;
            AV1B_NOUNROLL
            for (int k = tid; k < (n << lwords); k += nt) {
                const int i = k >> lwords, j = k & ((1 << lwords) - 1);
                *(uint32_t*)(t + (i + 1) * pitch + 4 + 4 * j) = __ldcg((const uint32_t*)(g.p + (size_t)(y0 + i) * g.stride + x0) + j);
            }
        };
        auto load_res = [&](int pl, int16_t* q, int n, int lg) {
            const int x0 = col * n, y0 = r * n;
            const int16_t* g = c.rp[pl] + (size_t)y0 * c.rpitch[pl] + x0;
            const int lch = lg - 3; // 16-byte chunks per row
            AV1B_NOUNROLL
            for (int k = tid; k < (n << lch); k += nt) {
                const int i = k >> lch, j = k & ((1 << lch) - 1);
                *((uint4*)(q + i * n) + j) = __ldcg((const uint4*)(g + (size_t)i * c.rpitch[pl]) + j);
            }
        };
        if (have_res) {
            load_res(0, q0, n0, lg0);
            load_res(1, q1, n1, lg0 - 1);
            load_res(2, q2, n1, lg0 - 1);
        }
        if (load_pred) {
            load_own(0, t0, n0, lg0, pitch0);
            load_own(1, t1, n1, lg0 - 1, pitch1);
            load_own(2, t2, n1, lg0 - 1, pitch1);
        }
        // the sample-independent half of every intra op of the chunk, one op per thread
        block_sync(nt);
        for (unsigned q = tid; q < min((unsigned)WAVE_OP_CHUNK, e.n_ops); q += nt) decode_intra_op(&s_ops[0][q], fc);
        block_sync(nt); // ops are staged (stage_op) by other threads than the ones that decoded them
        // ---- halo: the row above (x0-4 .. x0+2n-1 as words) and / or the column to the left, of all
        // three planes as ONE list of items so that every load is in flight before the first store waits
        // `ahalf`: only the part of the row above that belongs to the above-right superblock; `lhalf`:
        // only the lower half of the column; (id, cnt): the threads that share the work (the whole
        // CTA, or one warp on its own while the others run ops).
        auto load_halo = [&](bool above, bool ahalf, bool left, bool lhalf, int id, int cnt) {
            const int aw0 = (n0 >> 1) + 1, aw1 = (n1 >> 1) + 1; // words of a row above
            const int a0 = ahalf ? (n0 >> 2) + 1 : 0, a1 = ahalf ? (n1 >> 2) + 1 : 0;
            const int ca0 = (above && r > 0) ? aw0 - a0 : 0, ca1 = (above && r > 0) ? aw1 - a1 : 0;
            const int l0 = lhalf ? n0 >> 1 : 0, l1 = lhalf ? n1 >> 1 : 0;
            const int cl0 = (left && col > 0) ? n0 - l0 : 0, cl1 = (left && col > 0) ? n1 - l1 : 0;
            const int n_above = ca0 + 2 * ca1, total = n_above + cl0 + 2 * cl1;
            AV1B_NOUNROLL
            for (int b0 = 0; b0 < total; b0 += 2 * cnt) {
                uint32_t v[2];
                uint8_t* dsts[2];
                bool word[2];
                AV1B_UNROLL
                for (int u = 0; u < 2; u++) {
                    int k = b0 + u * cnt + id;
                    dsts[u] = nullptr;
                    word[u] = false;
                    v[u] = 0;
                    if (k >= total) continue;
                    if (k < n_above) {
                        const int pl = k < ca0 ? 0 : (k < ca0 + ca1 ? 1 : 2);
                        k -= pl == 0 ? 0 : (pl == 1 ? ca0 : ca0 + ca1);
                        k += pl ? a1 : a0;
                        const int n = pl ? n1 : n0;
                        const PlaneView g = c.cur.pl[pl];
                        uint8_t* t = pl == 0 ? t0 : (pl == 1 ? t1 : t2);
                        v[u] = __ldcg((const uint32_t*)(g.p + (size_t)(r * n - 1) * g.stride + col * n - 4) + k);
                        dsts[u] = t + 4 * k;
                        word[u] = true;
                    } else {
                        k -= n_above;
                        const int pl = k < cl0 ? 0 : (k < cl0 + cl1 ? 1 : 2);
                        k -= pl == 0 ? 0 : (pl == 1 ? cl0 : cl0 + cl1);
                        k += pl ? l1 : l0;
                        const int n = pl ? n1 : n0, pitch = pl ? pitch1 : pitch0;
                        const PlaneView g = c.cur.pl[pl];
                        uint8_t* t = pl == 0 ? t0 : (pl == 1 ? t1 : t2);
                        v[u] = __ldcg(g.p + (size_t)(r * n + k) * g.stride + col * n - 1);
                        dsts[u] = t + (k + 1) * pitch + 3;
                    }
                }
                AV1B_UNROLL
                for (int u = 0; u < 2; u++) {
                    if (!dsts[u]) continue;
                    if (word[u]) *(uint32_t*)dsts[u] = v[u];
                    else *dsts[u] = (uint8_t)v[u];
                }
            }
        };
        PlaneSet io;
        io.pix0 = t0 + (ptrdiff_t)(1 - r * n0) * pitch0 + (4 - col * n0);
        io.pix1 = t1 + (ptrdiff_t)(1 - r * n1) * pitch1 + (4 - col * n1);
        io.pix2 = t2 + (ptrdiff_t)(1 - r * n1) * pitch1 + (4 - col * n1);
        io.pitch0 = pitch0;
        io.pitch12 = pitch1;
        // residual (x, y) in frame coordinates lives at q[(y - y0) * n + (x - x0)]
        io.res0 = have_res ? q0 - (ptrdiff_t)(r * n0) * n0 - col * n0 : nullptr;
        io.res1 = have_res ? q1 - (ptrdiff_t)(r * n1) * n1 - col * n1 : nullptr;
        io.res2 = have_res ? q2 - (ptrdiff_t)(r * n1) * n1 - col * n1 : nullptr;
        io.rpitch0 = n0;
        io.rpitch12 = n1;
        auto exec_staged = [&](const StagedOp& st) {
            if (st.state == 1) {
                intra::Packed k;
                k.w[0] = st.wa.z, k.w[1] = st.wa.w, k.w[2] = st.wb.x, k.w[3] = st.wb.z;
                intra::run<true, WAVE_NT>(k, st.o, scratch->I, lane, nl);
                return;
            }
            union {
                uint4 v[2];
                Av1bOp op;
            } u;
            u.v[0] = st.wa, u.v[1] = st.wb;
            exec_op<true, WAVE_NT>(c, hdr, fc, u.op, io, *scratch, nullptr, lane, nl);
        };
        // ---- hand-off.  The neighbours only read this superblock's right column and bottom row, so
        // those go out ahead of the rest of the tile -- each half as soon as no later op writes it
        // (Av1bSb::pub_r1 / pub_b1), everything at the end -- and the progress word moves as soon as
        // they are on their way (one fence by one thread).
        // An early hand-off is the work of ONE warp (`solo` >= 0: that warp), the others go on with
        // the next level: what it copies is final, and its own lane 0 releases the progress word
        // after the warp's stores (__syncwarp orders them before the release).
        int pubbits = 0;
        auto publish = [&](int bits, int solo) {
            const bool all = (bits & 4) != 0;
            pubbits |= all ? 7 : bits;
            if (solo >= 0 && warp != solo) return;
            const int id = solo >= 0 ? lane : tid, cnt = solo >= 0 ? nl : nt;
            for (int pl = 0; pl < 3; pl++) {
                const int sub = pl ? 1 : 0;
                const int n = pl ? n1 : n0, pitch = pl ? pitch1 : pitch0;
                const uint8_t* t = pl == 0 ? t0 : (pl == 1 ? t1 : t2);
                const int x0 = col * n, y0 = r * n;
                const int pw = (hdr->mi_cols * 4) >> sub, ph = (hdr->mi_rows * 4) >> sub;
                const int cw = min(n, pw - x0), chh = min(n, ph - y0);
                const PlaneView g = c.cur.pl[pl];
                if (chh == n && (all || (bits & 2))) { // a superblock row below exists
                    uint32_t* d = (uint32_t*)(g.p + (size_t)(y0 + n - 1) * g.stride + x0);
                    const int words = (all ? cw : min(cw, n >> 1)) >> 2;
                    AV1B_NOUNROLL
                    for (int k = id; k < words; k += cnt) d[k] = *(const uint32_t*)(t + n * pitch + 4 + 4 * k);
                }
                if (cw == n && (all || (bits & 1))) { // a superblock to the right exists
                    uint8_t* d = g.p + (size_t)y0 * g.stride + x0 + n - 1;
                    const int rows = all ? chh : min(chh, n >> 1);
                    AV1B_NOUNROLL
                    for (int k = id; k < rows; k += cnt) d[(size_t)k * g.stride] = t[(k + 1) * pitch + 4 + n - 1];
                }
            }
            if (solo >= 0) {
                __syncwarp();
                if (lane == 0) av1b_st_release(progress + sb, pubbits);
            } else {
                block_sync(nt);
                if (tid == 0) av1b_st_release(progress + sb, pubbits);
            }
        };
        // ---- waiting, level by level: before a level runs, the halves of the neighbours' borders it
        // reads (Av1bSb::wait_*) must be final; the superblocks above and above-left always are.
        // The first wait of a superblock is the whole CTA's; a later one is normally done one level
        // AHEAD by one warp on its own (`solo` >= 0) while the others run the ops of the level before
        // -- the level barrier then makes the halo it loaded visible -- so a neighbour that is
        // already there costs nothing.  Up to four progress words are polled by four lanes at once.
        int waited_l = -1, waited_a = -1; // -1 nothing loaded yet, 0 no wait, 1 the first half, 2 all
        // the last level at which a wait can still come up (0xFF = never): past it, and with nothing
        // left to announce early, a level costs none of this bookkeeping
        auto lv_or0 = [](unsigned v) { return v == 0xFFu ? 0u : v; };
        const unsigned busy_until = max(max(lv_or0(e.wait_l1), lv_or0(e.wait_l2)), max(lv_or0(e.wait_a1), lv_or0(e.wait_a2)));
        const int want_bits = (e.pub_r1 ? 1 : 0) | (e.pub_b1 ? 2 : 0);
        bool waits_done = false;
        auto open_level = [&](unsigned level, int solo) {
            const int need_l = (col == 0) ? 0 : (level >= e.wait_l2 ? 2 : (level >= e.wait_l1 ? 1 : 0));
            const int need_a = (r == 0 || col + 1 >= sb_cols) ? 0 : (level >= e.wait_a2 ? 2 : (level >= e.wait_a1 ? 1 : 0));
            waits_done = level >= busy_until; // (the first call never returns early: waited_l is -1)
            if (need_l <= waited_l && need_a <= waited_a) return;
            const bool first = waited_l < 0;
            const bool more_l = need_l > max(waited_l, 0), more_a = need_a > max(waited_a, 0);
            const bool lhalf = waited_l >= 1; // the upper half of the column was final when it was loaded
            if (first) solo = -1;
            if (solo < 0 || warp == solo) {
                const int id = solo >= 0 ? lane : tid;
                if (id < 4) {
                    int idx = -1, mask = 4;
                    if (id == 0 && first && r > 0) idx = sb - sb_cols;
                    if (id == 1 && first && r > 0 && col > 0) idx = sb - sb_cols - 1;
                    if (id == 2 && more_l) idx = sb - 1, mask = need_l == 2 ? 4 : 5;
                    if (id == 3 && more_a) idx = sb - sb_cols + 1, mask = need_a == 2 ? 4 : 6;
                    if (idx >= 0) {
                        poll(idx, mask);
                        (void)av1b_ld_acquire(progress + idx);
                    }
                }
                if (solo >= 0) {
                    __syncwarp();
                    load_halo(more_a, true, more_l, lhalf, lane, nl);
                }
            }
            if (solo < 0) {
                block_sync(nt);
                if (tr && first) tr[3] = av1b_gtime();
                load_halo(first || more_a, !first, first || more_l, !first && lhalf, tid, nt);
                block_sync(nt);
                if (tr && first) tr[4] = av1b_gtime(), s_polled = 0;
            }
            waited_l = max(waited_l, need_l);
            waited_a = max(waited_a, need_a);
        };
        // ---- the ops, level by level, one warp per op.  The op list streams through a double
        // buffer: the next chunk is requested before the current one runs, so its L2 latency hides
        // behind the levels in between.
        int buf = 0;
        for (unsigned k0 = 0; k0 < e.n_ops; k0 += WAVE_OP_CHUNK, buf ^= 1) {
            const unsigned nk = min((unsigned)WAVE_OP_CHUNK, e.n_ops - k0);
            const Av1bOp* cur_ops = s_ops[buf];
            const unsigned k1 = k0 + WAVE_OP_CHUNK;
            const unsigned nn = k1 < e.n_ops ? min((unsigned)WAVE_OP_CHUNK, e.n_ops - k1) : 0;
            constexpr int PRE = (WAVE_OP_CHUNK * 2 + WARPS * 32 - 1) / (WARPS * 32); // uint4s of the next chunk per thread
            uint4 pre[PRE];
            AV1B_UNROLL
            for (int u = 0; u < PRE; u++) {
                pre[u] = make_uint4(0, 0, 0, 0);
                if ((unsigned)(tid + u * nt) < nn * 2) pre[u] = __ldg((const uint4*)(ops + e.first_op + k1) + tid + u * nt);
            }
            // res_off = level | (ops left in this level) << 16 (emitter scheduleSb); a producer that
            // leaves the count zero gets one op per step, still a valid order
            // (the word of the group that starts the NEXT level is read before the barrier and carried
            // over: nothing but the op itself sits between a barrier and the next)
            unsigned ro0 = cur_ops[0].res_off;
            unsigned g0 = 0, g1 = min(nk, max(1u, ro0 >> 16));
#ifndef AV1B_EMU
            // Op j of a level goes to warp (j + parity * nw/2) mod nw, the parity flipping with
            // every level: a level rarely has more than nw/2 ops, so the warps that work in one
            // level idle in the next and have fetched their next op and worked out its addresses
            // BEFORE the barrier opens -- the chain between two barriers is loads, arithmetic,
            // stores.
            unsigned par = 0;
            unsigned mine = g0 + ((warp + par * (nw >> 1)) & (nw - 1));
            StagedOp st;
            st.state = 0;
            if (mine < g1) st = stage_op(cur_ops + mine, io);
#endif
            while (g0 < nk) {
                const unsigned level = ro0 & 0xFFFFu;
                // level log of ONE superblock (trace capacity beyond the per-superblock records):
                // per level 16 words -- barrier exit, then (start, end) of the op of warps 0..6
                // (compiled in with -DAV1B_WAVE_LEVEL_LOG only: the test costs every level ~8 %)
#ifdef AV1B_WAVE_LEVEL_LOG
                unsigned long long* const lvlog = (c.trace && c.trace_cap >= (unsigned)n_sb + 64 && sb == (n_sb >> 1) + (sb_cols >> 1) && lane == 0 && level < 32)
                    ? c.trace + 8 * (size_t)n_sb + 16 * level : nullptr;
#else
                unsigned long long* const lvlog = nullptr;
#endif
                if (lvlog && warp == 0) lvlog[0] = av1b_gtime();
                if (lvlog && warp < 7) lvlog[1 + 2 * warp] = lvlog[2 + 2 * warp] = 0;
                // early hand-off of the border halves the levels before this one made final (a
                // level may straddle two chunks, so it only counts as over once a later one starts)
#ifdef AV1B_EMU
                const int solo = -1;
#else
                const int solo = (int)((nw - 1 + par * (nw >> 1)) & (nw - 1)); // the warp least likely to hold an op of this level
#endif
                if ((pubbits & want_bits) != want_bits && k0 + g0 > 0) {
                    int bits = 0;
                    if (e.pub_r1 && level > e.pub_r1 && !(pubbits & 1)) bits |= 1;
                    if (e.pub_b1 && level > e.pub_b1 && !(pubbits & 2)) bits |= 2;
                    if (bits) publish(bits, solo);
                }
                if (!waits_done) open_level(level, -1); // (no-op when the look-ahead of the level before did it)
#ifdef AV1B_EMU
                // the emulation runs the ops of a level in REVERSE order: if the level analysis
                // missed a dependency, the conformance MD5s under emulation break
                for (unsigned k = g1; k-- > g0;) exec_staged(stage_op(cur_ops + k, io));
                const unsigned ro1 = g1 < nk ? cur_ops[g1].res_off : 0u;
                const unsigned g2 = g1 < nk ? min(nk, g1 + max(1u, ro1 >> 16)) : g1;
#else
                if (mine < g1) {
                    if (lvlog && warp < 7) lvlog[1 + 2 * warp] = av1b_gtime();
                    exec_staged(st);
                    if (lvlog && warp < 7) lvlog[2 + 2 * warp] = av1b_gtime();
                    for (unsigned k = mine + nw; k < g1; k += nw) exec_staged(stage_op(cur_ops + k, io));
                }
                const unsigned ro1 = g1 < nk ? cur_ops[g1].res_off : 0u;
                const unsigned g2 = g1 < nk ? min(nk, g1 + max(1u, ro1 >> 16)) : g1;
                if (g1 < nk && !waits_done) open_level(ro1 & 0xFFFFu, solo);
                par ^= 1;
                mine = g1 + ((warp + par * (nw >> 1)) & (nw - 1));
                if (mine < g2) st = stage_op(cur_ops + mine, io);
#endif
                if (lvlog && warp == 7) lvlog[15] = av1b_gtime(); // warp 7 reaches the level barrier
                block_sync(nt);
                g0 = g1;
                g1 = g2;
                ro0 = ro1;
            }
            if (nn) {
#ifdef AV1B_EMU
                for (unsigned q = 0; q < nn * 2; q++) ((uint4*)s_ops[buf ^ 1])[q] = ((const uint4*)(ops + e.first_op + k1))[q];
#else
                AV1B_UNROLL
                for (int u = 0; u < PRE; u++)
                    if ((unsigned)(tid + u * nt) < nn * 2) ((uint4*)s_ops[buf ^ 1])[tid + u * nt] = pre[u];
#endif
                block_sync(nt);
                for (unsigned q = tid; q < nn; q += nt) decode_intra_op(&s_ops[buf ^ 1][q], fc);
                block_sync(nt);
            }
        }
        publish(4, -1);
        if (tr) tr[5] = av1b_gtime(), tr[6] = tr[5] + s_polled; // "signal" column of the trace = mid-superblock polling
        // ---- flush the tile (MI-aligned area only)
        auto flush_plane = [&](int pl, const uint8_t* t, int n, int pitch) {
            const int sub = pl ? 1 : 0;
            const int x0 = col * n, y0 = r * n;
            const int pw = (hdr->mi_cols * 4) >> sub, ph = (hdr->mi_rows * 4) >> sub;
            const int cw = min(n, pw - x0), chh = min(n, ph - y0);
            const PlaneView g = c.cur.pl[pl];
            const int words = cw >> 2; // MI-aligned widths are multiples of 4 in every plane
            if (words == (n >> 2)) {
                const int lwords = mc::ilog2_pow2(n) - 2;
                AV1B_NOUNROLL
                for (int k = tid; k < (chh << lwords); k += nt) {
                    const int i = k >> lwords, j = k & ((1 << lwords) - 1);
                    *((uint32_t*)(g.p + (size_t)(y0 + i) * g.stride + x0) + j) = *(const uint32_t*)(t + (i + 1) * pitch + 4 + 4 * j);
                }
            } else {
                AV1B_NOUNROLL
                for (int k = tid; k < chh * words; k += nt) {
                    const int i = k / words, j = k - i * words;
                    *((uint32_t*)(g.p + (size_t)(y0 + i) * g.stride + x0) + j) = *(const uint32_t*)(t + (i + 1) * pitch + 4 + 4 * j);
                }
            }
        };
        flush_plane(0, t0, n0, pitch0);
        flush_plane(1, t1, n1, pitch1);
        flush_plane(2, t2, n1, pitch1);
        if (tr) tr[7] = av1b_gtime();
    }
    wave_leave(c.sync, n_sb, tid, nt, &s_sb);
}

// Global-memory variant (frames with allow_intrabc: block copies read arbitrary earlier parts of
// the frame being reconstructed).  Same ops, samples read through L2.
__global__ void __launch_bounds__(256) wave_kernel_global(ReconCtx c)
{
    __shared__ OpScratch S;
    __shared__ mc::Scratch M;
    __shared__ int s_sb;
    const Av1bFrameHdr* hdr = (const Av1bFrameHdr*)c.cmd;
    const Av1bSb* sbs = (const Av1bSb*)(c.cmd + hdr->off_sb);
    const Av1bOp* ops = (const Av1bOp*)(c.cmd + hdr->off_ops);
    const FrameConst fc = frame_const(hdr);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int sb_cols = hdr->sb_cols, n_sb = hdr->n_sb;
    int* ticket = c.sync;
    int* progress = c.sync + 2;
    // An intrabc block vector may point into the row k superblock rows above up to 5k - 4 columns
    // (64-sample units) to the right of the current one (spec 7.11.3.2 / libaom av1_is_dv_valid:
    // gradient 1 + INTRABC_DELAY_SB64 [+ 1 for 128x128]); waiting for the row above through column
    // c + lag - 1 with lag 6 (SB64) / 4 (SB128) covers every row by induction.
    const int lag = hdr->sb_log2 == 6 ? 6 : 4;
    const bool have_res = c.rp[0] && hdr->n_itx;
    PlaneSet io;
    io.pix0 = c.cur.pl[0].p, io.pix1 = c.cur.pl[1].p, io.pix2 = c.cur.pl[2].p;
    io.pitch0 = c.cur.pl[0].stride, io.pitch12 = c.cur.pl[1].stride;
    io.res0 = have_res ? c.rp[0] : nullptr, io.res1 = have_res ? c.rp[1] : nullptr, io.res2 = have_res ? c.rp[2] : nullptr;
    io.rpitch0 = c.rpitch[0], io.rpitch12 = c.rpitch[1];
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const int t = atomicAdd(ticket, 1);
            int rr = 0, cc = 0;
            if (t < n_sb) sb_from_ticket(t, hdr->sb_rows, sb_cols, lag, rr, cc);
            s_sb = t < n_sb ? rr * sb_cols + cc : n_sb;
        }
        __syncthreads();
        const int sb = s_sb;
        if (sb >= n_sb) break;
        const int r = sb / sb_cols, col = sb - r * sb_cols;
        wave_wait(progress, r, col, sb_cols, lag, tid, nt);
        const Av1bSb e = sbs[sb];
        for (unsigned k = 0; k < e.n_ops; k++) {
            const Av1bOp op = ops[e.first_op + k];
            exec_op<false, 0>(c, hdr, fc, op, io, S, &M, tid, nt);
        }
        wave_signal(progress, r, col, tid, nt);
    }
    wave_leave(c.sync, n_sb, tid, nt, &s_sb);
}

// ------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------
int launch_itx(const ReconCtx& c, const Av1bFrameHdr& h, av1b_stream_t st)
{
    if (!h.n_itx) return 0;
    const Av1bOp* ops = (const Av1bOp*)(c.cmd + h.off_ops);
    const uint32_t* list = (const uint32_t*)(c.cmd + h.off_itx);
    const int16_t* coef = (const int16_t*)(c.cmd + h.off_coef);
    ResPlanes rp;
    for (int i = 0; i < 3; i++) {
        rp.p[i] = c.rp[i];
        rp.pitch[i] = c.rpitch[i];
    }
    ItxPlan plan;
    uint32_t prev = 0, ctas = 0;
    int classes = 0;
    for (int k = 0; k < 4; k++) {
        const uint32_t end = h.itx_class_end[k] > h.n_itx ? h.n_itx : h.itx_class_end[k];
        plan.first[k] = prev;
        plan.cnt[k] = end > prev ? end - prev : 0;
        prev = end > prev ? end : prev;
        const uint32_t per_cta = k == 3 ? 1 : ITX_WARPS * (32 / (4 << k)); // transform blocks per CTA pass (class 3: the CTA shares one)
        uint32_t nb = (plan.cnt[k] + per_cta - 1) / per_cta;
        if (nb > 148 * 16) nb = 148 * 16;
        ctas += nb;
        plan.cta_end[k] = ctas;
        classes += plan.cnt[k] != 0;
    }
    if (classes > 1 && h.n_itx <= 8192) {
        AV1B_LAUNCH(itx_kernel_all, (ctas), (ITX_WARPS * 32), st, ops, list, coef, c.res, rp, plan);
        return 1;
    }
    for (int k = 0; k < 4; k++) {
        const uint32_t cnt = plan.cnt[k];
        if (!cnt) continue;
        const uint32_t* lst = list + plan.first[k];
        const int nb = (int)(plan.cta_end[k] - (k ? plan.cta_end[k - 1] : 0));
        switch (k) {
        case 0: AV1B_LAUNCH(itx_kernel<0>, (nb), (ITX_WARPS * 32), st, ops, lst, cnt, coef, c.res, rp); break;
        case 1: AV1B_LAUNCH(itx_kernel<1>, (nb), (ITX_WARPS * 32), st, ops, lst, cnt, coef, c.res, rp); break;
        case 2: AV1B_LAUNCH(itx_kernel<2>, (nb), (ITX_WARPS * 32), st, ops, lst, cnt, coef, c.res, rp); break;
        default: AV1B_LAUNCH(itx_kernel<3>, (nb), (ITX_WARPS * 32), st, ops, lst, cnt, coef, c.res, rp); break;
        }
    }
    return classes;
}

void launch_inter(const ReconCtx& c, const Av1bFrameHdr& h, av1b_stream_t st)
{
    if (!h.n_iblk) return;
    unsigned g_block = (h.n_iblk + INTER_WARPS - 1) / INTER_WARPS;   // a warp per block
    if (g_block > 148 * 6) g_block = 148 * 6;
    unsigned g_unit = (h.n_ipu + 31) / 32;                            // a 32-unit chunk per CTA pass
    if (g_unit > 148 * 6) g_unit = 148 * 6;
    unsigned g_fast = (h.n_ipu + 31) / 32;
    if (g_fast > 148 * 24) g_fast = 148 * 24;
    if (!g_fast) g_fast = 1;
    AV1B_LAUNCH(inter_kernel, (g_block + g_unit + g_fast * FAST_SLICES), (INTER_WARPS * 32), st, c, g_block, g_unit, g_fast);
}

void launch_wave(const ReconCtx& c, const Av1bFrameHdr& h, av1b_stream_t st)
{
    if (!h.n_ops) return;
    // At most min(rows, ceil(cols / lag)) superblocks can be in progress at once (see
    // sb_from_ticket); half as many CTAs again stand by on the next diagonal so that a finished
    // dependency is picked up at once.  More would only spin on SM slots other streams need.
    const int lag = h.allow_intrabc ? (h.sb_log2 == 6 ? 6 : 4) : 2;
    const int width = std::max(1, std::min<int>((int)h.sb_rows, ((int)h.sb_cols + lag - 1) / lag));
    int grid = std::min<int>((int)h.n_sb, width + width / 2 + 1);
    if (grid > 148 * 2) grid = 148 * 2;
    if (!h.allow_intrabc) {
        // quadrant-level dependencies let the superblocks of TWO consecutive diagonals run together
        // (a superblock starts when its left neighbour is half done): twice the active set
        const char* genv = getenv("AV1B200_WAVE_GRID");
        const char* qenv = getenv("AV1B200_WAVE_GRIDQ"); // CTAs per superblock of the classic wavefront width, in quarters
        const int gq = qenv && atoi(qenv) > 0 ? atoi(qenv) : 10;
        grid = std::min<int>((int)h.n_sb, genv && atoi(genv) > 0 ? atoi(genv) : (gq * width + 3) / 4 + 1);
        if (grid > 148) grid = 148;
    }
    if (h.allow_intrabc) {
        AV1B_LAUNCH(wave_kernel_global, (grid), (256), st, c);
        return;
    }
    // Three builds.  The op code wants ~230 registers: at 16 warps per CTA it is held to 128 and pays
    // for it in every op (spills, rematerialised addresses), so the default is 8 warps with the full
    // register budget (a level rarely has more than 8 ops); 16 warps win only on frames of very many
    // small, simple blocks; 4 warps at two CTAs per SM are for a device crowded with streams.
    // AV1B200_WAVE_WARPS picks (measured: profiles/r02_wave_variants.md).
    const char* wenv = getenv("AV1B200_WAVE_WARPS");
    const int wv = wenv ? atoi(wenv) : 8;
    const int warps = (wv == 16 || wv == 4) ? wv : 8;
    const int smem = WAVE_SMEM_BYTES(1 << h.sb_log2, warps);
#ifndef AV1B_EMU
    {
        // the attribute is per DEVICE: a process may run decoders on several GPUs
        static std::mutex mu;
        static uint64_t done[4] = { 0, 0, 0, 0 }; // bit per device ordinal (<= 256 devices)
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        if (!(done[(dev >> 6) & 3] & (1ull << (dev & 63)))) {
            cudaFuncSetAttribute(wave_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WAVE_SMEM_BYTES(128, 16));
            cudaFuncSetAttribute(wave_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WAVE_SMEM_BYTES(128, 8));
            cudaFuncSetAttribute(wave_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WAVE_SMEM_BYTES(128, 4));
            done[(dev >> 6) & 3] |= 1ull << (dev & 63);
        }
    }
    if (warps == 16) wave_kernel<16, 1><<<dim3(grid), dim3(512), smem, st>>>(c);
    else if (warps == 8) wave_kernel<8, 1><<<dim3(grid), dim3(256), smem, st>>>(c);
    else wave_kernel<4, 2><<<dim3(grid), dim3(128), smem, st>>>(c);
#else
    (void)smem;
    AV1B_LAUNCH((wave_kernel<8, 1>), (grid), (256), st, c);
#endif
}

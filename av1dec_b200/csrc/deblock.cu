// deblock.cu -- deblocking filter: both edge passes of a plane in ONE launch, out of place.
//
// Behaviour restated from the reference: decoder/LoopFilter.cpp:40-370 (edge decisions :85-126,
// masks :206-289, narrow / wide filters :145-205, level derivation :301-359).  The reference
// filters every vertical edge of a plane, then every horizontal edge, in place; inside one pass no
// two edges touch the same samples.
//
// One CTA owns a 128 x 64 tile of one plane:
//   1. the tile plus an 8-sample halo is staged in shared memory with 64-bit coalesced loads;
//   2. every 4-sample EDGE UNIT of both passes that can modify the tile is tested (a thread per
//      unit: mode-info loads only) and the live ones are queued in shared memory -- the vertical
//      edges x0 .. x0+128 on all staged rows, the horizontal edges y0 .. y0+64 on the tile's own
//      columns.  Edges on the tile border are filtered by both neighbours; each keeps its side;
//   3. vertical edges are filtered in shared memory, a thread per sample LINE (4 per unit), then
//      the horizontal edges on the result -- the halo rows carry the vertically filtered samples
//      the horizontal pass reads (up to 7 rows beyond an edge);
//   4. the tile's own 128 x 64 samples are written to the OUTPUT frame with 64-bit stores.
// The input frame is never modified, so tiles are independent and the pass needs 2*S bytes of
// traffic for both edge directions together (it was 4*S as two in-place launches).
#include "dev.h"
#include "av1_tables.h"
#include "kernels.h"

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


enum {
    LF_TW = 128, LF_THGT = 64, LF_HALO = 8,
    LF_PITCH = LF_TW + 2 * LF_HALO + 8,          // 152 bytes = 38 words: 16 rows hit 16 different banks
    LF_ROWS = LF_THGT + 2 * LF_HALO,
    LF_VCOLS = LF_TW / 4 + 1,                    // vertical-edge unit columns x0/4 .. (x0+128)/4
    LF_VROWS = LF_ROWS / 4,
    LF_HCOLS = LF_TW / 4,
    LF_HROWS = LF_THGT / 4 + 1,                  // horizontal-edge unit rows y0/4 .. (y0+64)/4
    LF_THREADS = 256,
};

// A live edge unit, queued by filter-size class (4 / 8 / 16) so that the lanes of a warp run
// the same filter: bits 0..13 byte offset in LfSmem::px of q0 of the unit's first line,
// 14..19 limit, 20..25 level (blimit = 2 * (level + 2) + limit, thresh = level >> 4).
typedef uint32_t LfJob;

struct LfSmem {
    alignas(16) uint8_t px[LF_ROWS * LF_PITCH]; // px[(r + 8) * LF_PITCH + (c + 8)] = tile sample (r, c)
    LfJob vq[3][LF_VCOLS * LF_VROWS];
    LfJob hq[3][LF_VCOLS * LF_VROWS]; // (same capacity as vq so that one lf_run_pass serves both)
    int nv[3], nh[3];
    Av1bLoopFilterParams lf; // frame parameters (indexed by reference / plane: kept out of local memory)
};

// Is the edge unit at plane unit coordinates (uc, ur) (4-sample units) live in PASS?  Fills the
// filter parameters.  (loop_filter_edge, LoopFilter.cpp:85-126)
template <int PASS>
AV1B_DEV bool lf_test(const PostHdr* hdr, const Av1bLfMi* mis, const Av1bLoopFilterParams& lf, int plane, int uc, int ur, LfJob& job, int& cls)
{
    const int sub = plane ? 1 : 0;
    const int mi_cols = hdr->mi_cols;
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
    job = ((uint32_t)L.limit << 14) | ((uint32_t)L.lvl << 20);
    cls = filter_size == 4 ? 0 : (filter_size == 8 ? 1 : 2);
    return true;
}

// Append the live units of one pass to the queue of their filter-size class: a thread per unit,
// one shared-memory atomic per warp and class (ballot-compacted).  Units are (col, row) in the
// ncols x nrows window starting at unit (uc0, ur0) of the plane.
template <int PASS, int QCAP>
AV1B_DEV void lf_collect(const PostHdr* hdr, const Av1bLfMi* mis, const Av1bLoopFilterParams& lf, int plane, int uc0, int ur0, int ncols,
    int nrows, int x0, int y0, LfJob (*queue)[QCAP], int* count, int tid, int nt)
{
    const int sub = plane ? 1 : 0;
    const int ucols = hdr->mi_cols >> sub, urows = hdr->mi_rows >> sub;
    const int total = ncols * nrows;
    const int lane = tid & 31;
    for (int e0 = 0; e0 < total; e0 += nt) {
        const int e = e0 + tid;
        LfJob job = 0;
        int cls = -1;
        if (e < total) {
            const int j = e / ncols, i = e - j * ncols;
            const int uc = uc0 + i, ur = ur0 + j;
            if (uc >= 0 && ur >= 0 && uc < ucols && ur < urows && lf_test<PASS>(hdr, mis, lf, plane, uc, ur, job, cls))
                job |= (uint32_t)((4 * ur - y0 + LF_HALO) * LF_PITCH + (4 * uc - x0 + LF_HALO));
            else cls = -1;
        }
        AV1B_UNROLL
        for (int k = 0; k < 3; k++) {
            const unsigned m = __ballot_sync(0xFFFFFFFFu, cls == k);
            if (!m) continue;
            int base = 0;
            if (lane == (__ffs(m) - 1)) base = atomicAdd(count + k, __popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
            if (cls == k) queue[k][base + __popc(m & ((1u << lane) - 1))] = job;
        }
    }
}

// Filter the sample lines of the queued units of one size class: a thread per line.  `step` =
// distance between the samples of a line (1: vertical edge, LF_PITCH: horizontal edge), `next` =
// distance between the four lines of a unit.
template <int FS, bool CHROMA> AV1B_DEV void lf_run(uint8_t* px, const LfJob* queue, int n, int step, int next, int tid, int nt)
{
    // samples the masks and filters of this size read: 2 / 3 (chroma 6-tap) / 4 / 7 each side
    const int reach = FS == 4 ? 2 : (FS == 8 ? (CHROMA ? 3 : 4) : 7);
    for (int e = tid; e < 4 * n; e += nt) {
        const LfJob job = queue[e >> 2];
        uint8_t* q0 = px + (job & 0x3FFF) + (e & 3) * next;
        const int limit = (job >> 14) & 63, lvl = (job >> 20) & 63;
        int v[16];
        AV1B_UNROLL
        for (int k = 0; k < 16; k++) v[k] = (k >= 8 - reach && k < 8 + reach) ? q0[(k - 8) * step] : 0;
        const int nmod = lf_line(v, CHROMA ? 1 : 0, limit, 2 * (lvl + 2) + limit, lvl >> 4, FS);
        AV1B_UNROLL
        for (int k = 1; k <= (FS == 4 ? 2 : (FS == 8 ? 3 : 6)); k++) {
            if (k <= nmod) {
                q0[-k * step] = (uint8_t)v[8 - k];
                q0[(k - 1) * step] = (uint8_t)v[7 + k];
            }
        }
    }
}

template <bool CHROMA> AV1B_DEV void lf_run_pass(uint8_t* px, LfJob (*queue)[LF_VCOLS * LF_VROWS], const int* n, int step, int next, int tid, int nt)
{
    lf_run<4, CHROMA>(px, queue[0], n[0], step, next, tid, nt);
    lf_run<8, CHROMA>(px, queue[1], n[1], step, next, tid, nt);
    if (!CHROMA) lf_run<16, false>(px, queue[2], n[2], step, next, tid, nt);
}

}  // namespace

// grid: (tiles_x, tiles_y, plane)
__global__ void __launch_bounds__(LF_THREADS) deblock_kernel(PostCtx c)
{
    __shared__ LfSmem S;
    const PostHdr* hdr = &c.h;
    const Av1bLfMi* mis = (const Av1bLfMi*)(c.cmd + hdr->off_lfmi);
    const int plane = blockIdx.z, sub = plane ? 1 : 0;
    const int pw = (hdr->mi_cols * 4) >> sub, ph = (hdr->mi_rows * 4) >> sub; // MI-aligned plane
    const int x0 = blockIdx.x * LF_TW, y0 = blockIdx.y * LF_THGT;
    if (x0 >= pw || y0 >= ph) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    const PlaneView src = c.src.pl[plane], dst = c.deb.pl[plane];
    for (int k = tid; k < 3; k += nt) S.nv[k] = S.nh[k] = 0;
    for (int k = tid; k < (int)(sizeof(Av1bLoopFilterParams) / 4); k += nt) ((uint32_t*)&S.lf)[k] = ((const uint32_t*)&hdr->lf)[k];
    const Av1bLoopFilterParams& lf = S.lf;
    // ---- 1. stage rows y0-8 .. y0+71, columns x0-8 .. x0+135 (rows clamped into the padded plane)
    const int chunks = (LF_TW + 2 * LF_HALO) / 8;
    for (int e = tid; e < LF_ROWS * chunks; e += nt) {
        const int r = e / chunks, k = e - r * chunks;
        const int y = clip3(-LF_HALO, ph + LF_HALO - 1, y0 - LF_HALO + r);
        *(uint2*)(S.px + r * LF_PITCH + 8 * k) = __ldg((const uint2*)(src.p + (ptrdiff_t)y * src.stride + x0 - LF_HALO) + k);
    }
    __syncthreads();
    const bool filtered = plane == 0 || hdr->lf.level[1 + plane];
    if (filtered) {
        // ---- 2. collect the live edge units of both passes
        lf_collect<0>(hdr, mis, lf, plane, x0 / 4, (y0 - LF_HALO) / 4, LF_VCOLS, LF_VROWS, x0, y0, S.vq, S.nv, tid, nt);
        lf_collect<1>(hdr, mis, lf, plane, x0 / 4, y0 / 4, LF_HCOLS, LF_HROWS, x0, y0, S.hq, S.nh, tid, nt);
        __syncthreads();
        // ---- 3. vertical edges, then horizontal edges on the result
        if (plane) {
            lf_run_pass<true>(S.px, S.vq, S.nv, 1, LF_PITCH, tid, nt);
            __syncthreads();
            lf_run_pass<true>(S.px, S.hq, S.nh, LF_PITCH, 1, tid, nt);
        } else {
            lf_run_pass<false>(S.px, S.vq, S.nv, 1, LF_PITCH, tid, nt);
            __syncthreads();
            lf_run_pass<false>(S.px, S.hq, S.nh, LF_PITCH, 1, tid, nt);
        }
        __syncthreads();
    }
    // ---- 4. write the tile's own samples
    const int ochunks = min((int)LF_TW, pw - x0 + 7) / 8, orows = min((int)LF_THGT, ph - y0);
    for (int e = tid; e < orows * (LF_TW / 8); e += nt) {
        const int r = e / (LF_TW / 8), k = e - r * (LF_TW / 8);
        if (k >= ochunks) continue;
        *(uint2*)(dst.p + (size_t)(y0 + r) * dst.stride + x0 + 8 * k) = *(const uint2*)(S.px + (r + LF_HALO) * LF_PITCH + LF_HALO + 8 * k);
    }
}

void launch_deblock(const PostCtx& c, const Av1bFrameHdr& h, av1b_stream_t st)
{
    if (!h.lf.level[0] && !h.lf.level[1]) return;
    const int gx = (h.mi_cols * 4 + LF_TW - 1) / LF_TW, gy = (h.mi_rows * 4 + LF_THGT - 1) / LF_THGT;
    AV1B_LAUNCH(deblock_kernel, (gx, gy, 3), (LF_THREADS), st, c);
}

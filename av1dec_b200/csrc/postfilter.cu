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

// Filter one sample line across an edge. q0 at `p`, samples along the filter axis `step` apart.
AV1B_DEV void lf_sample(uint8_t* p, int step, int plane, int limit, int blimit, int thresh, int filter_size)
{
    const int q0 = p[0], q1 = p[step], q2 = p[2 * step], q3 = p[3 * step];
    const int p0 = p[-step], p1 = p[-2 * step], p2 = p[-3 * step], p3 = p[-4 * step];
    const int hev = (iabs(p1 - p0) > thresh) | (iabs(q1 - q0) > thresh);
    const int filter_len = (filter_size == 4) ? 4 : (plane ? 6 : (filter_size == 8 ? 8 : 16));
    int mask = (iabs(p1 - p0) > limit) | (iabs(q1 - q0) > limit) | ((iabs(p0 - q0) * 2 + iabs(p1 - q1) / 2) > blimit);
    if (filter_len >= 6) mask |= (iabs(p2 - p1) > limit) | (iabs(q2 - q1) > limit);
    if (filter_len >= 8) mask |= (iabs(p3 - p2) > limit) | (iabs(q3 - q2) > limit);
    if (mask) return;
    int flat = 0, flat2 = 0;
    if (filter_size >= 8) {
        int m = (iabs(p1 - p0) > 1) | (iabs(q1 - q0) > 1) | (iabs(p2 - p0) > 1) | (iabs(q2 - q0) > 1);
        if (filter_len >= 8) m |= (iabs(p3 - p0) > 1) | (iabs(q3 - q0) > 1);
        flat = !m;
    }
    int q4 = 0, q5 = 0, q6 = 0, p4 = 0, p5 = 0, p6 = 0;
    if (filter_size >= 16) {
        q4 = p[4 * step]; q5 = p[5 * step]; q6 = p[6 * step];
        p4 = p[-5 * step]; p5 = p[-6 * step]; p6 = p[-7 * step];
        int m = (iabs(p6 - p0) > 1) | (iabs(q6 - q0) > 1) | (iabs(p5 - p0) > 1) | (iabs(q5 - q0) > 1) | (iabs(p4 - p0) > 1)
            | (iabs(q4 - q0) > 1);
        flat2 = !m;
    }
    if (filter_size == 4 || !flat) {
        const int ps0 = p0 - 128, ps1 = p1 - 128, qs0 = q0 - 128, qs1 = q1 - 128;
        int f = hev ? f4clamp(ps1 - qs1) : 0;
        f = f4clamp(f + 3 * (qs0 - ps0));
        const int f1 = f4clamp(f + 4) >> 3, f2 = f4clamp(f + 3) >> 3;
        p[0] = (uint8_t)(f4clamp(qs0 - f1) + 128);
        p[-step] = (uint8_t)(f4clamp(ps0 + f2) + 128);
        if (!hev) {
            const int f3 = (f1 + 1) >> 1;
            p[step] = (uint8_t)(f4clamp(qs1 - f3) + 128);
            p[-2 * step] = (uint8_t)(f4clamp(ps1 + f3) + 128);
        }
    } else if (filter_size == 8 || !flat2) {
        if (!plane) {
            // 8-tap luma: n = 3, centre weight 2
            const int v[8] = { p3, p2, p1, p0, q0, q1, q2, q3 }; // index = pos + 4
            int F[6];
            AV1B_UNROLL
            for (int i = -3; i < 3; i++) {
                int t = 0;
                AV1B_UNROLL
                for (int j = -3; j <= 3; j++) t += v[clip3(-4, 3, i + j) + 4] * (j == 0 ? 2 : 1);
                F[i + 3] = (t + 4) >> 3;
            }
            AV1B_UNROLL
            for (int i = -3; i < 3; i++) p[i * step] = (uint8_t)F[i + 3];
        } else {
            // 6-tap chroma: n = 2, weights 2 for |j| <= 1
            const int v[6] = { p2, p1, p0, q0, q1, q2 }; // index = pos + 3
            int F[4];
            AV1B_UNROLL
            for (int i = -2; i < 2; i++) {
                int t = 0;
                AV1B_UNROLL
                for (int j = -2; j <= 2; j++) t += v[clip3(-3, 2, i + j) + 3] * ((j >= -1 && j <= 1) ? 2 : 1);
                F[i + 2] = (t + 4) >> 3;
            }
            AV1B_UNROLL
            for (int i = -2; i < 2; i++) p[i * step] = (uint8_t)F[i + 2];
        }
    } else {
        // 14-tap luma: n = 6, weights 2 for |j| <= 1
        const int v[14] = { p6, p5, p4, p3, p2, p1, p0, q0, q1, q2, q3, q4, q5, q6 }; // index = pos + 7
        int F[12];
        AV1B_UNROLL
        for (int i = -6; i < 6; i++) {
            int t = 0;
            AV1B_UNROLL
            for (int j = -6; j <= 6; j++) t += v[clip3(-7, 6, i + j) + 7] * ((j >= -1 && j <= 1) ? 2 : 1);
            F[i + 6] = (t + 8) >> 4;
        }
        AV1B_UNROLL
        for (int i = -6; i < 6; i++) p[i * step] = (uint8_t)F[i + 6];
    }
}

}  // namespace

// One thread per (edge unit, sample line): 4 consecutive threads share one 4-sample edge unit.
__global__ void __launch_bounds__(256) deblock_kernel(PostCtx c, int pass)
{
    const Av1bFrameHdr* hdr = (const Av1bFrameHdr*)c.cmd;
    const Av1bLfMi* mis = (const Av1bLfMi*)(c.cmd + hdr->off_lfmi);
    const Av1bLoopFilterParams lf = hdr->lf;
    const int plane = blockIdx.z;
    if (plane > 0 && !lf.level[1 + plane]) return;
    const int sub = plane ? 1 : 0;
    const int mi_cols = hdr->mi_cols, mi_rows = hdr->mi_rows;
    const int ucols = mi_cols >> sub, urows = mi_rows >> sub; // edge units of this plane
    const long long total = (long long)ucols * urows * 4;
    const PlaneView pv = c.src.pl[plane];
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        // pass 0 (vertical edges): consecutive threads walk down rows of one unit column slowly;
        // pass 1 (horizontal edges): consecutive threads walk along x (coalesced).
        int ur, uc, i;
        if (pass == 0) {
            // lines of the same edge column are 'stride' apart; order threads x-fastest over units
            const long long line = t / ucols; // sample row index inside plane (units*4)
            uc = (int)(t - line * ucols);
            ur = (int)(line >> 2);
            i = (int)(line & 3);
        } else {
            const long long urow = t / (ucols * 4);
            const int rem = (int)(t - urow * (ucols * 4));
            ur = (int)urow;
            uc = rem >> 2;
            i = rem & 3;
        }
        int row = (ur << sub), col = (uc << sub);
        const int x = col * 4, y = row * 4;
        if (x >= hdr->frame_w || y >= hdr->frame_h) continue;
        if (pass == 0 ? (x == 0) : (y == 0)) continue;
        row |= sub;
        col |= sub;
        const int xp = x >> sub, yp = y >> sub;
        const int prev_row = row - ((pass == 1) ? (1 << sub) : 0);
        const int prev_col = col - ((pass == 0) ? (1 << sub) : 0);
        const Av1bLfMi mi = mis[row * mi_cols + col];
        const Av1bLfMi pm = mis[prev_row * mi_cols + prev_col];
        const int tx = (mi.tx >> (5 * plane)) & 31, ptx = (pm.tx >> (5 * plane)) & 31;
        const int bw = max(4, k_block_w[mi.mi_size] >> sub), bh = max(4, k_block_h[mi.mi_size] >> sub);
        const bool skip = mi.flags & 1;
        const bool is_intra = ((mi.flags >> 2) & 7) == 0;
        const bool block_edge = pass == 0 ? (xp % bw == 0) : (yp % bh == 0);
        const bool tx_edge = pass == 0 ? (xp % k_tx_w[tx] == 0) : (yp % k_tx_h[tx] == 0);
        if (!(tx_edge && (block_edge || !skip || is_intra))) continue;
        const int base = pass == 0 ? min(k_tx_w[ptx], k_tx_w[tx]) : min(k_tx_h[ptx], k_tx_h[tx]);
        const int filter_size = plane ? min(8, base) : min(16, base);
        LfLevel L = lf_strength(lf, mi, plane, pass);
        if (!L.lvl) L = lf_strength(lf, pm, plane, pass);
        if (L.lvl <= 0) continue;
        uint8_t* p = pv.p + (size_t)(yp + (pass == 0 ? i : 0)) * pv.stride + xp + (pass == 1 ? i : 0);
        lf_sample(p, pass == 0 ? 1 : pv.stride, plane, L.limit, L.blimit, L.thresh, filter_size);
    }
}

// ==========================================================================================
// CDEF
// ==========================================================================================
namespace {

AV1B_DEV int cdef_constrain(int diff, int threshold, int damping)
{
    if (!threshold) return 0;
    const int adj = max(0, damping - floor_log2((unsigned)threshold));
    const int ad = iabs(diff);
    const int v = clip3(0, ad, threshold - (ad >> adj));
    return diff < 0 ? -v : v;
}

// Direction search for the 8x8 luma block at (x0,y0). (reference cdefDirection)
AV1B_DEV void cdef_direction(const PlaneView& y, int x0, int y0, int& dir, int& var)
{
    int partial[8][15];
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 15; j++) partial[i][j] = 0;
    for (int i = 0; i < 8; i++) {
        for (int j = 0; j < 8; j++) {
            const int x = (int)__ldg(y.p + (size_t)(y0 + i) * y.stride + x0 + j) - 128;
            partial[0][i + j] += x;
            partial[1][i + j / 2] += x;
            partial[2][i] += x;
            partial[3][3 + i - j / 2] += x;
            partial[4][7 + i - j] += x;
            partial[5][3 - i / 2 + j] += x;
            partial[6][j] += x;
            partial[7][i / 2 + j] += x;
        }
    }
    int cost[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    for (int i = 0; i < 8; i++) {
        cost[2] += partial[2][i] * partial[2][i];
        cost[6] += partial[6][i] * partial[6][i];
    }
    cost[2] *= 105;
    cost[6] *= 105;
    for (int i = 0; i < 7; i++) {
        cost[0] += (partial[0][i] * partial[0][i] + partial[0][14 - i] * partial[0][14 - i]) * k_cdef_div_table[i + 1];
        cost[4] += (partial[4][i] * partial[4][i] + partial[4][14 - i] * partial[4][14 - i]) * k_cdef_div_table[i + 1];
    }
    cost[0] += partial[0][7] * partial[0][7] * 105;
    cost[4] += partial[4][7] * partial[4][7] * 105;
    for (int i = 1; i < 8; i += 2) {
        for (int j = 0; j < 5; j++) cost[i] += partial[i][3 + j] * partial[i][3 + j];
        cost[i] *= 105;
        for (int j = 0; j < 3; j++)
            cost[i] += (partial[i][j] * partial[i][j] + partial[i][10 - j] * partial[i][10 - j]) * k_cdef_div_table[2 * j + 2];
    }
    int best = 0;
    dir = 0;
    for (int i = 0; i < 8; i++) {
        if (cost[i] > best) {
            best = cost[i];
            dir = i;
        }
    }
    var = (best - cost[(dir + 4) & 7]) >> 10;
}

}  // namespace

// One CTA per 64x64 luma area (8x8 CDEF blocks), all three planes.
__global__ void __launch_bounds__(256) cdef_kernel(PostCtx c)
{
    __shared__ uint8_t s_idx[64];  // preset index or 0xFF
    __shared__ uint8_t s_dir[64];
    __shared__ int s_var[64];
    const Av1bFrameHdr* hdr = (const Av1bFrameHdr*)c.cmd;
    const uint8_t* cdef8 = c.cmd + hdr->off_cdef8;
    const Av1bCdefParams cp = hdr->cdef;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int c8 = hdr->mi_cols >> 1, r8 = hdr->mi_rows >> 1; // 8x8 blocks in the frame
    const int fbx = blockIdx.x * 8, fby = blockIdx.y * 8;      // first 8x8 block of this CTA
    for (int e = tid; e < 64; e += nt) {
        const int by = fby + (e >> 3), bx = fbx + (e & 7);
        int idx = 0xFF, dir = 0, var = 0;
        if (by < r8 && bx < c8) {
            idx = cdef8[by * c8 + bx];
            if (idx != 0xFF) cdef_direction(c.src.pl[0], bx * 8, by * 8, dir, var);
        }
        s_idx[e] = (uint8_t)idx;
        s_dir[e] = (uint8_t)dir;
        s_var[e] = var;
    }
    __syncthreads();
    for (int plane = 0; plane < 3; plane++) {
        const int sub = plane ? 1 : 0;
        const int bs = 8 >> sub;                   // block size in this plane
        const int tile = 64 >> sub;                // CTA tile size in this plane
        const int pw = (hdr->mi_cols * 4) >> sub, ph = (hdr->mi_rows * 4) >> sub; // MI-aligned plane size
        const int x_base = (fbx * 8) >> sub, y_base = (fby * 8) >> sub;
        const PlaneView src = c.src.pl[plane], dst = c.cdef.pl[plane];
        for (int e = tid; e < tile * tile; e += nt) {
            const int ly = e / tile, lx = e - ly * tile;
            const int x = x_base + lx, y = y_base + ly;
            if (x >= pw || y >= ph) continue;
            const int b = (ly / bs) * 8 + (lx / bs);
            const int px = __ldg(src.p + (size_t)y * src.stride + x);
            const int idx = s_idx[b];
            int out = px;
            if (idx != 0xFF) {
                int pri, sec, dir, damping;
                const int ydir = s_dir[b];
                if (!plane) {
                    pri = cp.y_pri[idx];
                    sec = cp.y_sec[idx];
                    dir = pri == 0 ? 0 : ydir;
                    const int var = s_var[b];
                    const int var_str = (var >> 6) ? min(floor_log2((unsigned)(var >> 6)), 12) : 0;
                    pri = var ? ((pri * (4 + var_str) + 8) >> 4) : 0;
                    damping = cp.damping;
                } else {
                    pri = cp.uv_pri[idx];
                    sec = cp.uv_sec[idx];
                    dir = pri == 0 ? 0 : k_cdef_uv_dir[1][1][ydir];
                    damping = cp.damping - 1;
                }
                int sum = 0, mx = px, mn = px;
                AV1B_UNROLL
                for (int k = 0; k < 2; k++) {
                    AV1B_UNROLL
                    for (int sgn = -1; sgn <= 1; sgn += 2) {
                        {
                            const int yy = y + sgn * k_cdef_directions[dir][k][0];
                            const int xx = x + sgn * k_cdef_directions[dir][k][1];
                            if (xx >= 0 && xx < pw && yy >= 0 && yy < ph) {
                                const int p = __ldg(src.p + (size_t)yy * src.stride + xx);
                                sum += k_cdef_pri_taps[pri & 1][k] * cdef_constrain(p - px, pri, damping);
                                mx = max(mx, p);
                                mn = min(mn, p);
                            }
                        }
                        AV1B_UNROLL
                        for (int off = -2; off <= 2; off += 4) {
                            const int d2 = (dir + off) & 7;
                            const int yy = y + sgn * k_cdef_directions[d2][k][0];
                            const int xx = x + sgn * k_cdef_directions[d2][k][1];
                            if (xx >= 0 && xx < pw && yy >= 0 && yy < ph) {
                                const int p = __ldg(src.p + (size_t)yy * src.stride + xx);
                                sum += k_cdef_sec_taps[pri & 1][k] * cdef_constrain(p - px, sec, damping);
                                mx = max(mx, p);
                                mn = min(mn, p);
                            }
                        }
                    }
                }
                out = clip3(mn, mx, px + ((8 + sum - (sum < 0)) >> 4));
            }
            dst.p[(size_t)y * dst.stride + x] = (uint8_t)out;
        }
    }
}

// ==========================================================================================
// Loop restoration
// ==========================================================================================
namespace {

enum { LR_TW = 32, LR_MAXH = 64, LR_SW = LR_TW + 6, LR_SH = LR_MAXH + 6 };

struct LrShared {
    uint8_t src[LR_SH * LR_SW];          // source samples with 3-sample halo
    int16_t wien[LR_SH * LR_TW];         // Wiener horizontal pass
    int a[(LR_MAXH + 2) * (LR_TW + 2)];  // SGR A
    int b[(LR_MAXH + 2) * (LR_TW + 2)];  // SGR B
    int flt[2][LR_MAXH * LR_TW];         // SGR filtered planes
};

// get_source_sample + extendBorder(3) (LoopRestoration.cpp:234-246, VideoFrame.cpp:81-101)
AV1B_DEV int lr_source(const PlaneView& cdef, const PlaneView& deb, int x, int y, int start, int end, int pw, int ph)
{
    const PlaneView* s = &cdef;
    if (y < start) {
        y = max(start - 2, y);
        s = &deb;
    } else if (y >= end) {
        y = min(end + 1, y);
        s = &deb;
    }
    x = clip3(0, pw - 1, x);
    y = clip3(0, ph - 1, y);
    return __ldg(s->p + (size_t)y * s->stride + x);
}

AV1B_DEV void sgr_pass(LrShared& S, int w, int h, int set, int pass, int r, int tid, int nt)
{
    const int eps = k_sgr_params[set][pass * 2 + 1];
    const int n = (2 * r + 1) * (2 * r + 1);
    const int n2e = n * n * eps;
    const unsigned s = (unsigned)(((1 << 20) + n2e / 2) / n2e);
    const int one_over_n = ((1 << 12) + (n / 2)) / n;
    const int aw = w + 2;
    for (int e = tid; e < (h + 2) * aw; e += nt) {
        const int i = e / aw - 1, j = e - (i + 1) * aw - 1;
        int a = 0, b = 0;
        for (int dy = -r; dy <= r; dy++)
            for (int dx = -r; dx <= r; dx++) {
                const int cpx = S.src[(i + dy + 3) * LR_SW + (j + dx + 3)];
                a += cpx * cpx;
                b += cpx;
            }
        const unsigned p = (unsigned)max(0, a * n - b * b);
        const unsigned z = (p * s + (1u << 19)) >> 20;
        int a2;
        if (z >= 255) a2 = 256;
        else if (z == 0) a2 = 1;
        else a2 = (int)(((z << 8) + (z / 2)) / (z + 1));
        const int b2 = (256 - a2) * b * one_over_n;
        S.a[e] = a2;
        S.b[e] = (b2 + (1 << 11)) >> 12;
    }
    __syncthreads();
    for (int e = tid; e < w * h; e += nt) {
        const int i = e / w, j = e - i * w;
        const int shift = (pass == 0 && (i & 1)) ? 4 : 5;
        int a = 0, b = 0;
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                int weight;
                if (pass == 0) weight = ((i + dy) & 1) ? (dx == 0 ? 6 : 5) : 0;
                else weight = (dx == 0 || dy == 0) ? 4 : 3;
                a += weight * S.a[(i + dy + 1) * aw + (j + dx + 1)];
                b += weight * S.b[(i + dy + 1) * aw + (j + dx + 1)];
            }
        const int v = a * S.src[(i + 3) * LR_SW + (j + 3)] + b;
        S.flt[pass][i * LR_TW + j] = round2(v, 8 + shift - 4);
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
        for (int e = tid; e < w * h; e += nt) {
            const int i = e / w, j = e - i * w;
            out.p[(size_t)(y0 + i) * out.stride + x0 + j] = __ldg(cdef.p + (size_t)(y0 + i) * cdef.stride + x0 + j);
        }
        return;
    }
    for (int e = tid; e < (h + 6) * (w + 6); e += nt) {
        const int i = e / (w + 6), j = e - i * (w + 6);
        S.src[i * LR_SW + j] = (uint8_t)lr_source(cdef, deb, x0 + j - 3, y0 + i - 3, start, end, pw, ph);
    }
    __syncthreads();
    if (type == 1) {
        int vf[7], hf[7];
        vf[3] = 128;
        hf[3] = 128;
        for (int k = 0; k < 3; k++) {
            vf[k] = vf[6 - k] = unit.wiener[0][k];
            hf[k] = hf[6 - k] = unit.wiener[1][k];
            vf[3] -= 2 * unit.wiener[0][k];
            hf[3] -= 2 * unit.wiener[1][k];
        }
        for (int e = tid; e < (h + 6) * w; e += nt) {
            const int r = e / w, cc = e - r * w;
            int s = 0;
            AV1B_UNROLL
            for (int t = 0; t < 7; t++) s += hf[t] * S.src[r * LR_SW + cc + t];
            S.wien[r * LR_TW + cc] = (int16_t)clip3(-2048, 6143, (s + 4) >> 3);
        }
        __syncthreads();
        for (int e = tid; e < w * h; e += nt) {
            const int r = e / w, cc = e - r * w;
            int s = 0;
            AV1B_UNROLL
            for (int t = 0; t < 7; t++) s += vf[t] * S.wien[(r + t) * LR_TW + cc];
            out.p[(size_t)(y0 + r) * out.stride + x0 + cc] = (uint8_t)clip_u8((s + 1024) >> 11);
        }
    } else {
        const int set = unit.sgr_set;
        const int r0 = k_sgr_params[set][0], r1 = k_sgr_params[set][2];
        if (r0) sgr_pass(S, w, h, set, 0, r0, tid, nt);
        if (r1) sgr_pass(S, w, h, set, 1, r1, tid, nt);
        const int w0 = unit.sgr_xqd[0], w1 = unit.sgr_xqd[1], w2 = 128 - w0 - w1;
        for (int e = tid; e < w * h; e += nt) {
            const int i = e / w, j = e - i * w;
            const int u = S.src[(i + 3) * LR_SW + (j + 3)] << 4;
            int v = w1 * u;
            v += w0 * (r0 ? S.flt[0][i * LR_TW + j] : u);
            v += w2 * (r1 ? S.flt[1][i * LR_TW + j] : u);
            out.p[(size_t)(y0 + i) * out.stride + x0 + j] = (uint8_t)clip_u8(round2(v, 11));
        }
    }
}

// ==========================================================================================
// launchers
// ==========================================================================================
void launch_deblock(const PostCtx& c, const Av1bFrameHdr& h, av1b_stream_t st)
{
    if (!h.lf.level[0] && !h.lf.level[1]) return;
    const long long total = (long long)h.mi_cols * h.mi_rows * 4;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 32) grid = 148 * 32;
    for (int pass = 0; pass < 2; pass++) AV1B_LAUNCH(deblock_kernel, (grid, 1, 3), (256), st, c, pass);
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

// cdef.cu -- constrained directional enhancement filter, one whole-frame pass.
//
// Behaviour restated from the reference: decoder/Cdef.cpp:41-261 (skip test :72-82, direction
// search :203-261, strengths :84-99, constrained filter :158-198, availability :131-156).
//
// One CTA filters a 64x64 luma area (8x8 CDEF blocks of 8x8) and the matching 32x32 chroma areas.
//   1. stage   each plane's tile + 2-sample halo goes to shared memory ONCE, as bytes, with
//              64-bit loads (rows 8-byte aligned; 80-byte luma pitch = 20 words, so the eight rows
//              of a block sit in eight different bank groups).
//   2. search  one work item per (block, direction): the 64 samples come in as sixteen words and
//              every partial sum is a chain of IDP.4A with compile-time byte selectors
//              (sum of (px - 128) over the selected bytes in one instruction).
//   3. filter  one work item per block ROW (8 luma / 4 chroma samples).  A tap row is three
//              aligned words + two funnel shifts whatever the direction; samples are widened to
//              packed 16x2 and the constrain function runs on both halves at once:
//                 |d|           VABSDIFF4
//                 K - t         VIADDMNMX   t = max(0, thr - (|d| >> adj))
//                 max(d,-t)+K   VIADDMNMX
//                 min(.,t)+K    VIADDMNMX
//              and the weighted sum accumulates the BIASED value c + K >= 1 with one IMAD per
//              tap, so no lane ever goes negative and nothing carries between the halves.
//              Min / max tracking is one VIMNMX3 per tap pair.  Taps whose strength is zero are
//              skipped, and with only one strength active the final clamp cannot trigger
//              (|sum| / 16 <= 12/16 of the largest tap difference), so min / max are skipped too.
//   Frame-edge tiles run the same code with an availability mask per tap (EDGE = true).
#include "dev.h"
#include "av1_tables.h"
#include "kernels.h"

namespace {

enum {
    CD_X0 = 16,    // the tile row starts CD_X0 columns left of the tile (16-byte aligned rows for bulk copies)
    CD_YP = 96,    // bytes per luma tile row: tile columns -16 .. 79
    CD_YROWS = 68, // tile rows -2 .. 65
    CD_CP = 64,    // chroma: tile columns -16 .. 47
    CD_CROWS = 36,
    CD_K = 16,     // bias of the constrained difference (strengths are <= 15)
    CD_THREADS = 256,
};

struct CdefBlk {
    uint8_t idx;      // preset or 0xFF (block left untouched)
    uint8_t pri[2];   // [0] luma (variance adjusted), [1] chroma
    uint8_t sec[2];
    uint8_t adjp[2];  // damping adjustment shifts
    uint8_t adjs[2];
    uint8_t pad[3];
    // tap displacements [luma / chroma][group * 2 + k][dy, dx]; groups: primary (dir),
    // secondary (dir + 2), secondary (dir - 2)
    int8_t d[2][6][2];
};

struct alignas(128) CdefSmem {
    alignas(128) uint8_t y[CD_YROWS * CD_YP];        // y[(r + 2) * CD_YP + (c + CD_X0)] = luma tile sample (r, c)
    alignas(128) uint8_t c[2][CD_CROWS * CD_CP];   // (both tile sizes are multiples of 128 bytes: TMA destinations)
    int cost[64][8];
    CdefBlk blk[64];
};

template <int D> AV1B_DEV constexpr int cdef_bin(int i, int j)
{
    return D == 0 ? i + j : D == 1 ? i + j / 2 : D == 2 ? i : D == 3 ? 3 + i - j / 2 : D == 4 ? 7 + i - j : D == 5 ? 3 - i / 2 + j
        : D == 6 ? j : i / 2 + j;
}

// cost of direction D for the 8x8 block held in w[16] (row i = words 2i, 2i+1, samples XOR 0x80 so
// that the bytes read as px - 128).  Reference cdefDirection, Cdef.cpp:203-261.
template <int D> AV1B_DEV int cdef_cost(const uint32_t* w)
{
    int part[15];
    AV1B_UNROLL
    for (int k = 0; k < 15; k++) part[k] = 0;
    AV1B_UNROLL
    for (int i = 0; i < 8; i++) {
        AV1B_UNROLL
        for (int hw = 0; hw < 2; hw++) {
            AV1B_UNROLL
            for (int k = 0; k < 4; k++) {
                const int bin = cdef_bin<D>(i, 4 * hw + k);
                bool first = true;
                AV1B_UNROLL
                for (int kk = 0; kk < 4; kk++)
                    if (kk < k && cdef_bin<D>(i, 4 * hw + kk) == bin) first = false;
                if (!first) continue;
                uint32_t sel = 0;
                AV1B_UNROLL
                for (int kk = 0; kk < 4; kk++)
                    if (kk >= k && cdef_bin<D>(i, 4 * hw + kk) == bin) sel |= 1u << (8 * kk);
                part[bin] = av1b_dp4a_ss(w[2 * i + hw], sel, part[bin]);
            }
        }
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

AV1B_DEV int cdef_cost_dyn(int d, const uint32_t* w)
{
    switch (d) {
    case 0: return cdef_cost<0>(w);
    case 1: return cdef_cost<1>(w);
    case 2: return cdef_cost<2>(w);
    case 3: return cdef_cost<3>(w);
    case 4: return cdef_cost<4>(w);
    case 5: return cdef_cost<5>(w);
    case 6: return cdef_cost<6>(w);
    default: return cdef_cost<7>(w);
    }
}

// Four bytes -> two packed 16x2 sample pairs.
AV1B_DEV void widen4(uint32_t w, uint32_t& p01, uint32_t& p23)
{
    p01 = __byte_perm(w, 0, 0x4140);
    p23 = __byte_perm(w, 0, 0x4342);
}

// The constrained difference of two sample pairs, biased by CD_K (see the file comment).
//   p, x2   tap / centre pairs; nxk = K - x per half; kmthr = K - threshold per half
AV1B_DEV uint32_t constrain_k(uint32_t p, uint32_t x2, uint32_t nxk, uint32_t kmthr, int adj, uint32_t amask)
{
    const uint32_t K2 = CD_K * 0x00010001u;
    const uint32_t a = __vabsdiffu4(p, x2);                  // |d|; an unavailable sample (0x4000) reads as >= 0x4000
    const uint32_t s = (a >> adj) & amask;                   // |d| >> dampingAdj
    const uint32_t ntk = __viaddmin_s16x2(s, kmthr, K2);     // K - t,  t = max(0, thr - s)
    const uint32_t c1k = __viaddmax_s16x2(p, nxk, ntk);      // max(d, -t) + K
    return __vmins2(c1k, 2 * K2 - ntk);                      // min(., t) + K   (2K - ntk >= K per half: no borrow)
}

// Filter one row of a block: NP sample pairs (4 luma, 2 chroma) starting at tile byte `ctr`.
//   pitch     bytes per tile row
//   xg, yg    plane coordinates of the first sample (EDGE only), pw / ph: MI-aligned plane size
// Returns the output bytes in out[0 .. NP/2).  (reference cdefFilter, Cdef.cpp:158-198)
template <int NP, bool EDGE>
AV1B_DEV void cdef_filter_row(const uint8_t* tile, int ctr, int pitch, const int8_t (*disp)[2], int pri, int sec, int adjp, int adjs, int xg,
    int yg, int pw, int ph, uint32_t* out)
{
    uint32_t x[NP], nxk[NP], acc[NP], mx[NP], mn[NP];
    {
        const uint32_t* cw = (const uint32_t*)(tile + ctr);
        AV1B_UNROLL
        for (int n = 0; n < NP / 2; n++) widen4(cw[n], x[2 * n], x[2 * n + 1]);
    }
    AV1B_UNROLL
    for (int n = 0; n < NP; n++) {
        nxk[n] = __vadd2(~x[n], (CD_K + 1) * 0x00010001u); // K - x
        acc[n] = 7 * 0x00010001u;
        mx[n] = mn[n] = x[n];
    }
    const bool both = pri && sec;
    AV1B_UNROLL
    for (int g = 0; g < 6; g++) {
        const bool primary = g < 2;
        if (primary ? !pri : !sec) continue;
        const int thr = primary ? pri : sec, adj = primary ? adjp : adjs;
        const uint32_t kmthr = (uint32_t)(CD_K - thr) * 0x00010001u;
        const uint32_t amask = (0xFFFFu >> adj) * 0x00010001u;
        const uint32_t wgt = primary ? ((g & 1) ? ((pri & 1) ? 3u : 2u) : ((pri & 1) ? 3u : 4u)) : ((g & 1) ? 1u : 2u);
        const int dy = disp[g][0], dx = disp[g][1];
        const int off = dy * pitch + dx;
        uint32_t pp[2][NP];
        AV1B_UNROLL
        for (int sgn = 0; sgn < 2; sgn++) {
            const int a = ctr + (sgn ? -off : off);
            const uint32_t* wp = (const uint32_t*)(tile + (a & ~3));
            const uint32_t sh = (uint32_t)(a & 3) * 8;
            const uint32_t w0 = wp[0], w1 = wp[1];
            widen4(__funnelshift_r(w0, w1, sh), pp[sgn][0], pp[sgn][1]);
            if (NP == 4) {
                const uint32_t w2 = wp[2];
                widen4(__funnelshift_r(w1, w2, sh), pp[sgn][2], pp[sgn][3]);
            }
            if (EDGE) {
                // a tap outside the MI-aligned frame does not exist (Cdef.cpp:131-156): it becomes 0x4000,
                // which constrains to 0 and never wins the minimum; the maximum masks it out below
                const int ty = yg + (sgn ? -dy : dy), tx = xg + (sgn ? -dx : dx);
                const bool rowok = ty >= 0 && ty < ph;
                AV1B_UNROLL
                for (int n = 0; n < NP; n++) {
                    const int cx = tx + 2 * n;
                    uint32_t m = rowok ? 0u : 0xFFFFFFFFu;
                    if (cx < 0 || cx >= pw) m |= 0x0000FFFFu;
                    if (cx + 1 < 0 || cx + 1 >= pw) m |= 0xFFFF0000u;
                    pp[sgn][n] = (pp[sgn][n] & ~m) | (0x40004000u & m);
                }
            }
        }
        AV1B_UNROLL
        for (int n = 0; n < NP; n++) {
            const uint32_t ca = constrain_k(pp[0][n], x[n], nxk[n], kmthr, adj, amask);
            const uint32_t cb = constrain_k(pp[1][n], x[n], nxk[n], kmthr, adj, amask);
            acc[n] += wgt * ca;
            acc[n] += wgt * cb;
            if (both) {
                if (EDGE) mx[n] = __vimax3_u16x2(mx[n], pp[0][n] & 0x00FF00FFu, pp[1][n] & 0x00FF00FFu);
                else mx[n] = __vimax3_u16x2(mx[n], pp[0][n], pp[1][n]);
                mn[n] = __vimin3_u16x2(mn[n], pp[0][n], pp[1][n]);
            }
        }
    }
    // y = clip3(min, max, x + ((8 + sum - (sum < 0)) >> 4)) with acc = sum + 7 + B, B = K * (sum of the weights)
    const uint32_t B = (uint32_t)CD_K * ((pri ? 12u : 0u) + (sec ? 12u : 0u));
    const uint32_t nb6 = (uint32_t)(0x10000 - (int)(B + 6)) * 0x00010001u & 0xFFFFFFFFu; // -(B + 6) per half
    const uint32_t nb16 = (uint32_t)(0x10000 - (int)(B >> 4)) * 0x00010001u;               // -(B / 16) per half
    uint32_t yv[NP];
    AV1B_UNROLL
    for (int n = 0; n < NP; n++) {
        const uint32_t ge = __viaddmin_s16x2_relu(acc[n], nb6, 0x00010001u); // [sum >= 0]
        const uint32_t q = ((acc[n] + ge) >> 4) & 0x0FFF0FFFu;               // floor((sum + 7 + ge) / 16) + B / 16
        uint32_t y = __vadd2(x[n] + q, nb16);
        if (both) y = __vmins2(__vmaxs2(y, mn[n]), mx[n]);
        yv[n] = y;
    }
    AV1B_UNROLL
    for (int n = 0; n < NP / 2; n++) out[n] = __byte_perm(yv[2 * n], yv[2 * n + 1], 0x6420);
}

template <bool EDGE> AV1B_DEV void cdef_filter_tile(CdefSmem& S, const PostCtx& c, int fbx, int fby, int pw, int ph, int tid, int nt)
{
    // ---- luma: item = (block, row); a warp covers four adjacent blocks x eight rows
    for (int e = tid; e < 512; e += nt) {
        const int r = e & 7, bx = (e >> 3) & 7, by = e >> 6;
        const int b = by * 8 + bx;
        const int xg = (fbx + bx) * 8, yg = (fby + by) * 8 + r;
        if (xg >= pw || yg >= ph) continue;
        const CdefBlk& B = S.blk[b];
        const int ctr = (by * 8 + r + 2) * CD_YP + bx * 8 + CD_X0;
        uint32_t out[2];
        if (B.idx != 0xFF && (B.pri[0] | B.sec[0]))
            cdef_filter_row<4, EDGE>(S.y, ctr, CD_YP, B.d[0], B.pri[0], B.sec[0], B.adjp[0], B.adjs[0], xg, yg, pw, ph, out);
        else {
            out[0] = *(const uint32_t*)(S.y + ctr);
            out[1] = *(const uint32_t*)(S.y + ctr + 4);
        }
        *(uint2*)(c.cdef.pl[0].p + (size_t)yg * c.cdef.pl[0].stride + xg) = make_uint2(out[0], out[1]);
    }
    // ---- chroma: item = (plane, block, row of 4); a warp covers eight adjacent blocks x four rows
    const int cpw = pw >> 1, cph = ph >> 1;
    for (int e = tid; e < 512; e += nt) {
        const int r = e & 3, bx = (e >> 2) & 7, by = (e >> 5) & 7, plane = 1 + (e >> 8);
        const int b = by * 8 + bx;
        const int xg = (fbx + bx) * 4, yg = (fby + by) * 4 + r;
        if (xg >= cpw || yg >= cph) continue;
        const CdefBlk& B = S.blk[b];
        const uint8_t* tile = S.c[plane - 1];
        const int ctr = (by * 4 + r + 2) * CD_CP + bx * 4 + CD_X0;
        uint32_t out[1];
        if (B.idx != 0xFF && (B.pri[1] | B.sec[1]))
            cdef_filter_row<2, EDGE>(tile, ctr, CD_CP, B.d[1], B.pri[1], B.sec[1], B.adjp[1], B.adjs[1], xg, yg, cpw, cph, out);
        else out[0] = *(const uint32_t*)(tile + ctr);
        *(uint32_t*)(c.cdef.pl[plane].p + (size_t)yg * c.cdef.pl[plane].stride + xg) = out[0];
    }
}

// Load-based staging (no TMA descriptor: emulation, or a driver without the encoder): rows -2 ..
// rows-3 of a tile whose sample (0, 0) is plane sample (x0, y0), `pitch` bytes per row from column
// x0 - CD_X0 as 128-bit words.  Rows are clamped into the plane's padded area (their content is
// never used: taps there are unavailable).
AV1B_DEV void cdef_stage(const PlaneView& src, int x0, int y0, int ph, int rows, int pitch, uint8_t* tile, int tid, int nt)
{
    const int chunks = pitch >> 4;
    for (int e = tid; e < rows * chunks; e += nt) {
        const int r = e / chunks, k = e - r * chunks;
        const int y = clip3(-2, ph + 1, y0 - 2 + r);
        *(uint4*)(tile + r * pitch + 16 * k) = __ldg((const uint4*)(src.p + (ptrdiff_t)y * src.stride + x0 - CD_X0) + k);
    }
}

}  // namespace

void cdef_tile_box(int plane, int* box_w, int* box_h)
{
    *box_w = plane ? CD_CP : CD_YP;
    *box_h = plane ? CD_CROWS : CD_YROWS;
}

__global__ void __launch_bounds__(CD_THREADS, 5) cdef_kernel(const __grid_constant__ PostCtx c)
{
    __shared__ CdefSmem S;
    const PostHdr* hdr = &c.h;
    const uint8_t* cdef8 = c.cmd + hdr->off_cdef8;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int c8 = hdr->mi_cols >> 1, r8 = hdr->mi_rows >> 1; // 8x8 blocks in the frame
    const int fbx = blockIdx.x * 8, fby = blockIdx.y * 8;      // first 8x8 block of this CTA
    const int pw = hdr->mi_cols * 4, ph = hdr->mi_rows * 4;
    // ---- 1. stage.  With TMA descriptors the three tiles (halo included) are THREE requests to the
    // TMA unit, issued by one thread and counted in bytes on `bar`: nobody spends issue slots on
    // moving the bytes, and the preset look-up below overlaps the copies.  Tile rows past the
    // plane's padded area come back as zeros (never used: taps there are unavailable).
    alignas(8) __shared__ unsigned long long bar;
    const bool tma = c.tma_ok != 0;
    if (tma) {
        if (tid == 0) av1b_mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0) {
            av1b_mbar_expect_tx(&bar, CD_YROWS * CD_YP + 2 * CD_CROWS * CD_CP);
            av1b_tma_load_2d(S.y, &c.cdef_in[0], c.tma_x0 + fbx * 8 - CD_X0, c.tma_y0 + fby * 8 - 2, &bar);
            av1b_tma_load_2d(S.c[0], &c.cdef_in[1], c.tma_x0 + fbx * 4 - CD_X0, c.tma_y0 + fby * 4 - 2, &bar);
            av1b_tma_load_2d(S.c[1], &c.cdef_in[2], c.tma_x0 + fbx * 4 - CD_X0, c.tma_y0 + fby * 4 - 2, &bar);
        }
    } else {
        cdef_stage(c.deb.pl[0], fbx * 8, fby * 8, ph, CD_YROWS, CD_YP, S.y, tid, nt);
        cdef_stage(c.deb.pl[1], fbx * 4, fby * 4, ph >> 1, CD_CROWS, CD_CP, S.c[0], tid, nt);
        cdef_stage(c.deb.pl[2], fbx * 4, fby * 4, ph >> 1, CD_CROWS, CD_CP, S.c[1], tid, nt);
    }
    const Av1bCdefParams& cp = hdr->cdef;
    for (int e = tid; e < 64; e += nt) {
        const int by = fby + (e >> 3), bx = fbx + (e & 7);
        S.blk[e].idx = (by < r8 && bx < c8) ? cdef8[by * c8 + bx] : 0xFF;
    }
    if (tma) av1b_mbar_wait(&bar, 0);
    __syncthreads();
    // ---- 2. direction search (only blocks whose primary strengths are not both zero need one)
    for (int e = tid; e < 512; e += nt) {
        const int b = e & 63, d = e >> 6; // 64 consecutive threads share a direction: no divergence inside a warp
        const int idx = S.blk[b].idx;
        if (idx == 0xFF || !(cp.y_pri[idx] | cp.uv_pri[idx])) continue;
        uint32_t w[16];
        const uint8_t* bp = S.y + ((b >> 3) * 8 + 2) * CD_YP + (b & 7) * 8 + CD_X0;
        AV1B_UNROLL
        for (int i = 0; i < 8; i++) {
            const uint2 v = *(const uint2*)(bp + i * CD_YP);
            w[2 * i] = v.x ^ 0x80808080u;
            w[2 * i + 1] = v.y ^ 0x80808080u;
        }
        S.cost[b][d] = cdef_cost_dyn(d, w);
    }
    __syncthreads();
    for (int e = tid; e < 64; e += nt) {
        CdefBlk& B = S.blk[e];
        if (B.idx == 0xFF) continue;
        int best = 0, dir = 0, var = 0;
        if (cp.y_pri[B.idx] | cp.uv_pri[B.idx]) {
            for (int d = 0; d < 8; d++)
                if (S.cost[e][d] > best) {
                    best = S.cost[e][d];
                    dir = d;
                }
            var = (best - S.cost[e][(dir + 4) & 7]) >> 10;
        }
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
        B.pri[1] = (uint8_t)pri_uv;
        B.sec[1] = (uint8_t)sec_uv;
        B.adjp[1] = (uint8_t)(pri_uv ? max(0, damp - 1 - floor_log2((unsigned)pri_uv)) : 0);
        B.adjs[1] = (uint8_t)(sec_uv ? max(0, damp - 1 - floor_log2((unsigned)sec_uv)) : 0);
        const int dir_uv = pri_uv == 0 ? 0 : k_cdef_uv_dir[1][1][dir];
        for (int pl = 0; pl < 2; pl++) {
            const int dd = pl ? dir_uv : dir_y;
            for (int g = 0; g < 3; g++) {
                const int d = g == 0 ? dd : (g == 1 ? ((dd + 2) & 7) : ((dd + 6) & 7));
                for (int k = 0; k < 2; k++) {
                    B.d[pl][g * 2 + k][0] = k_cdef_directions[d][k][0];
                    B.d[pl][g * 2 + k][1] = k_cdef_directions[d][k][1];
                }
            }
        }
    }
    __syncthreads();
    // ---- 3. filter
    const bool edge = fbx == 0 || fby == 0 || fbx * 8 + 66 > pw || fby * 8 + 66 > ph;
    if (edge) cdef_filter_tile<true>(S, c, fbx, fby, pw, ph, tid, nt);
    else cdef_filter_tile<false>(S, c, fbx, fby, pw, ph, tid, nt);
}

void launch_cdef(const PostCtx& c, const Av1bFrameHdr& h, av1b_stream_t st)
{
    if (!h.cdef.enabled) return;
    const int gx = (h.mi_cols * 4 + 63) / 64, gy = (h.mi_rows * 4 + 63) / 64;
    AV1B_LAUNCH(cdef_kernel, (gx, gy, 1), (CD_THREADS), st, c);
}

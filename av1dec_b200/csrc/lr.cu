// lr.cu -- loop restoration (Wiener and self-guided), one whole-frame pass.
//
// Behaviour restated from the reference: decoder/LoopRestoration.cpp:49-479
//   unit / stripe geometry :49-134, source fetch (stripe rule + 3-sample replicated border)
//   :234-246 and VideoFrame.cpp:81-101, Wiener :247-277, self-guided :284-479.
//
// Work decomposition.  A CTA of 128 threads owns a TILE of TW x 32 samples of one plane
// (TW = 64, or 32 for planes whose restoration units are 32 wide); a tile never crosses a
// 64-luma-row stripe or a restoration unit, so its filter type and coefficients are uniform.
// The tile and its 3-sample halo are staged once in shared memory as bytes (64-bit loads, the
// stripe rule picks CDEF or deblocked rows per row), and every phase works on FOUR horizontally
// adjacent samples per work item out of aligned 32/64/128-bit shared-memory words:
//   Wiener   horizontal 7 taps = two IDP.4A per sample on funnel-shifted byte windows (the centre
//            tap 128 - 2*sum is split over both dot products so every tap fits int8), 32-bit
//            intermediates; vertical pass 4x4 samples per item from 128-bit rows; packed stores.
//   SGR      box sums by IDP.4A (sum and sum of squares of a 4-byte window in one instruction
//            each), A/B per 4 columns, the 3x3 weighting on packed 16x2 / 32-bit rows, final
//            blend fused into the last pass.
//   NONE     128-bit copy of the CDEF tile.
#include "dev.h"
#include "av1_tables.h"
#include "kernels.h"
#include <algorithm>
#include <cstring>

namespace {

enum { LR_TH = 32, LR_THREADS = 128, LR_ROWS = LR_TH + 6 };

// ((z << 8) + z / 2) / (z + 1) for z = 1..254; [0] = 1 and [255] = 256 are the two special
// cases of the reference's a2 derivation (LoopRestoration.cpp:389-397)
AV1T_CONST uint16_t k_sgr_xdiv[256] = {
    1, 128, 171, 192, 205, 213, 219, 224, 228, 230, 233, 235, 236, 238, 239, 240, 241, 242, 243, 243, 244, 244, 245, 245, 246, 246,
    247, 247, 247, 247, 248, 248, 248, 248, 249, 249, 249, 249, 249, 250, 250, 250, 250, 250, 250, 250, 251, 251, 251, 251, 251, 251,
    251, 251, 251, 251, 252, 252, 252, 252, 252, 252, 252, 252, 252, 252, 252, 252, 252, 252, 252, 252, 252, 253, 253, 253, 253, 253,
    253, 253, 253, 253, 253, 253, 253, 253, 253, 253, 253, 253, 253, 253, 253, 253, 253, 253, 253, 253, 253, 253, 253, 253, 254, 254,
    254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254,
    254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254,
    254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 254, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255,
    255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255,
    255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255,
    255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 256
};

template <int TW> struct LrGeom {
    enum {
        SP = TW + 16,      // bytes per staged source row: tile columns -8 .. TW+7
        OG = TW / 4,       // groups of four output columns
        AG = TW / 4 + 1,   // groups of four A/B columns (columns -1 .. TW, padded to a multiple of 4)
        AP = 4 * AG,       // elements per row of the box-sum and A/B arrays
    };
};

template <int TW> struct LrSmem {
    typedef LrGeom<TW> G;
    // src[(r + 3) * SP + (c + 8)] = source sample of tile position (r, c)
    alignas(16) uint8_t src[LR_ROWS * G::SP];
    union {
        alignas(16) int32_t wien[LR_ROWS * TW]; // Wiener horizontal pass, row r <-> tile row r - 3
        struct {
            alignas(16) uint32_t h2[LR_ROWS * G::AP]; // horizontal box sums of x^2
            alignas(16) uint16_t h1[LR_ROWS * G::AP]; // horizontal box sums of x
        } box;
    };
    alignas(16) uint32_t b[(LR_TH + 2) * G::AP]; // SGR B, row i + 1 <-> tile row i, column j + 1 <-> tile column j
    alignas(16) uint16_t a[(LR_TH + 2) * G::AP]; // SGR A
    alignas(16) uint16_t flt0[LR_TH * TW];       // first self-guided pass (r = 2)
    alignas(16) uint16_t xdiv[256];
};

// Row of the frame that get_source_sample() reads for plane row `y` (LoopRestoration.cpp:234-246
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

// ---------------------------------------------------------------------------------------------
// Self-guided filter phases.  `pass` 0: r = 2 (A/B only on odd tile rows -1, 1, 3, ...), 1: r = 1.
// ---------------------------------------------------------------------------------------------

// Horizontal box sums of x and x^2 for tile rows -1-r .. h+r, A/B columns 0 .. AP-1.
template <int TW, int R> AV1B_DEV void sgr_box_rows(LrSmem<TW>& S, int h, int tid, int nt)
{
    typedef LrGeom<TW> G;
    const int hrows = h + 2 + 2 * R; // box row rr <-> tile row rr - 1 - R <-> src row rr + 2 - R
    for (int e = tid; e < hrows * G::AG; e += nt) {
        const int rr = e / G::AG, g = e - rr * G::AG;
        const uint32_t* sp = (const uint32_t*)(S.src + (rr + 2 - R) * G::SP + 4 * g + 4);
        const uint32_t w0 = sp[0], w1 = sp[1], w2 = sp[2];
        // byte k of the run (w0, w1, w2) is src column 4g + 4 + k; A/B column 4g + m is centred on byte 3 + m
        uint32_t s1[4], s2[4];
        if (R == 2) {
            const uint32_t win[4] = { __funnelshift_r(w0, w1, 8), __funnelshift_r(w0, w1, 16), __funnelshift_r(w0, w1, 24), w1 };
            const uint32_t xb[4] = { (w1 >> 8) & 0xFF, (w1 >> 16) & 0xFF, w1 >> 24, w2 & 0xFF };
            AV1B_UNROLL
            for (int m = 0; m < 4; m++) {
                s1[m] = av1b_dp4a_uu(win[m], 0x01010101u, xb[m]);
                s2[m] = av1b_dp4a_uu(win[m], win[m], xb[m] * xb[m]);
            }
        } else {
            const uint32_t win[4] = { __funnelshift_r(w0, w1, 16) & 0xFFFFFFu, __funnelshift_r(w0, w1, 24) & 0xFFFFFFu, w1 & 0xFFFFFFu,
                __funnelshift_r(w1, w2, 8) & 0xFFFFFFu };
            AV1B_UNROLL
            for (int m = 0; m < 4; m++) {
                s1[m] = av1b_dp4a_uu(win[m], 0x01010101u, 0u);
                s2[m] = av1b_dp4a_uu(win[m], win[m], 0u);
            }
        }
        *(uint2*)(S.box.h1 + rr * G::AP + 4 * g) = make_uint2(s1[0] | (s1[1] << 16), s1[2] | (s1[3] << 16));
        *(uint4*)(S.box.h2 + rr * G::AP + 4 * g) = make_uint4(s2[0], s2[1], s2[2], s2[3]);
    }
}

// Vertical sums -> a2 / b2 (LoopRestoration.cpp:365-399).
template <int TW, int R> AV1B_DEV void sgr_ab(LrSmem<TW>& S, int h, int set, int tid, int nt)
{
    typedef LrGeom<TW> G;
    const int pass = R == 2 ? 0 : 1;
    const int eps = k_sgr_params[set][pass * 2 + 1];
    const int n = (2 * R + 1) * (2 * R + 1);
    const int n2e = n * n * eps;
    const unsigned s = (unsigned)(((1 << 20) + n2e / 2) / n2e);
    const int one_over_n = ((1 << 12) + (n / 2)) / n;
    const int istep = pass == 0 ? 2 : 1;
    const int nrows = pass == 0 ? (h + 3) / 2 : h + 2;
    for (int e = tid; e < nrows * G::AG; e += nt) {
        const int ri = e / G::AG, g = e - ri * G::AG;
        const int row = ri * istep; // A/B row index (tile row + 1) = first box row of the window
        uint32_t blo = 0, bhi = 0, a[4] = { 0, 0, 0, 0 };
        AV1B_UNROLL
        for (int k = 0; k < 2 * R + 1; k++) {
            const uint2 p1 = *(const uint2*)(S.box.h1 + (row + k) * G::AP + 4 * g);
            const uint4 p2 = *(const uint4*)(S.box.h2 + (row + k) * G::AP + 4 * g);
            blo += p1.x; // packed 16x2: column sums stay below 2^16
            bhi += p1.y;
            a[0] += p2.x;
            a[1] += p2.y;
            a[2] += p2.z;
            a[3] += p2.w;
        }
        const uint32_t bb[4] = { blo & 0xFFFFu, blo >> 16, bhi & 0xFFFFu, bhi >> 16 };
        uint32_t a2[4], b2[4];
        AV1B_UNROLL
        for (int m = 0; m < 4; m++) {
            const int b = (int)bb[m];
            const unsigned p = (unsigned)max(0, (int)a[m] * n - b * b);
            const unsigned z = (p * s + (1u << 19)) >> 20;
            a2[m] = S.xdiv[min(z, 255u)];
            b2[m] = (uint32_t)(((256 - (int)a2[m]) * b * one_over_n + (1 << 11)) >> 12);
        }
        *(uint2*)(S.a + row * G::AP + 4 * g) = make_uint2(a2[0] | (a2[1] << 16), a2[2] | (a2[3] << 16));
        *(uint4*)(S.b + row * G::AP + 4 * g) = make_uint4(b2[0], b2[1], b2[2], b2[3]);
    }
}

struct SgrOut {
    int w0, w1, w2;   // projection weights (w2 = 128 - w0 - w1)
    bool final;       // write the restored samples (else store the pass-0 plane for the next pass)
    bool have_flt0;   // final only: a pass-0 plane exists (r0 != 0)
    uint8_t* dst;     // output plane at tile origin
    int stride, w;    // output row stride, valid tile columns
};

// 3x3 weighting of A/B -> filtered plane (LoopRestoration.cpp:401-428), optionally fused with the
// final projection (:460-476).  One item = four adjacent samples of one row.
template <int TW, int PASS> AV1B_DEV void sgr_out(LrSmem<TW>& S, int h, const SgrOut& o, int tid, int nt)
{
    typedef LrGeom<TW> G;
    for (int e = tid; e < h * G::OG; e += nt) {
        const int i = e / G::OG, q = e - i * G::OG;
        uint32_t va[3], vb[6];
        int shift;
        {
            // rows of A/B feeding output row i: pass 0 even rows: i and i + 2; odd rows: i + 1; pass 1: i, i+1, i+2
            const int ra = (PASS == 0 && (i & 1)) ? i + 1 : i;
            const uint16_t* A = S.a + ra * G::AP + 4 * q;
            const uint32_t* B = S.b + ra * G::AP + 4 * q;
            const uint2 a0 = *(const uint2*)A;
            const uint32_t a1 = *(const uint32_t*)(A + 4);
            const uint4 b0 = *(const uint4*)B;
            const uint2 b1 = *(const uint2*)(B + 4);
            va[0] = a0.x, va[1] = a0.y, va[2] = a1;
            vb[0] = b0.x, vb[1] = b0.y, vb[2] = b0.z, vb[3] = b0.w, vb[4] = b1.x, vb[5] = b1.y;
            shift = 4;
            if (PASS == 1 || !(i & 1)) {
                const uint2 c0 = *(const uint2*)(A + 2 * G::AP);
                const uint32_t c1 = *(const uint32_t*)(A + 2 * G::AP + 4);
                const uint4 d0 = *(const uint4*)(B + 2 * G::AP);
                const uint2 d1 = *(const uint2*)(B + 2 * G::AP + 4);
                va[0] += c0.x, va[1] += c0.y, va[2] += c1; // packed 16x2, values <= 2 * 256
                vb[0] += d0.x, vb[1] += d0.y, vb[2] += d0.z, vb[3] += d0.w, vb[4] += d1.x, vb[5] += d1.y;
                shift = 5;
            }
        }
        int fa[4], fb[4];
        {
            const int v[6] = { (int)(va[0] & 0xFFFF), (int)(va[0] >> 16), (int)(va[1] & 0xFFFF), (int)(va[1] >> 16), (int)(va[2] & 0xFFFF),
                (int)(va[2] >> 16) };
            if (PASS == 0) {
                // 6 * centre column + 5 * side columns of the (summed) rows
                AV1B_UNROLL
                for (int m = 0; m < 4; m++) {
                    fa[m] = 5 * (v[m] + v[m + 1] + v[m + 2]) + v[m + 1];
                    fb[m] = 5 * (int)(vb[m] + vb[m + 1] + vb[m + 2]) + (int)vb[m + 1];
                }
            } else {
                // 4 * cross + 3 * corners = 3 * (outer rows 3-sum) + 4 * (middle row 3-sum) + outer rows centre
                const uint16_t* M = S.a + (i + 1) * G::AP + 4 * q;
                const uint32_t* N = S.b + (i + 1) * G::AP + 4 * q;
                const uint2 m0 = *(const uint2*)M;
                const uint32_t m1 = *(const uint32_t*)(M + 4);
                const uint4 n0 = *(const uint4*)N;
                const uint2 n1 = *(const uint2*)(N + 4);
                const int u[6] = { (int)(m0.x & 0xFFFF), (int)(m0.x >> 16), (int)(m0.y & 0xFFFF), (int)(m0.y >> 16), (int)(m1 & 0xFFFF),
                    (int)(m1 >> 16) };
                const uint32_t nb[6] = { n0.x, n0.y, n0.z, n0.w, n1.x, n1.y };
                AV1B_UNROLL
                for (int m = 0; m < 4; m++) {
                    fa[m] = 3 * (v[m] + v[m + 1] + v[m + 2]) + 4 * (u[m] + u[m + 1] + u[m + 2]) + v[m + 1];
                    fb[m] = 3 * (int)(vb[m] + vb[m + 1] + vb[m + 2]) + 4 * (int)(nb[m] + nb[m + 1] + nb[m + 2]) + (int)vb[m + 1];
                }
            }
        }
        const uint32_t xw = *(const uint32_t*)(S.src + (i + 3) * G::SP + 4 * q + 8);
        int f[4];
        AV1B_UNROLL
        for (int m = 0; m < 4; m++) {
            const int x = (int)((xw >> (8 * m)) & 0xFF);
            f[m] = (fa[m] * x + fb[m] + (1 << (3 + shift))) >> (4 + shift);
        }
        if (!o.final) {
            *(uint2*)(S.flt0 + i * TW + 4 * q) = make_uint2((uint32_t)f[0] | ((uint32_t)f[1] << 16), (uint32_t)f[2] | ((uint32_t)f[3] << 16));
            continue;
        }
        int f0[4];
        if (PASS == 1 && o.have_flt0) {
            const uint2 p = *(const uint2*)(S.flt0 + i * TW + 4 * q);
            f0[0] = (int)(p.x & 0xFFFF), f0[1] = (int)(p.x >> 16), f0[2] = (int)(p.y & 0xFFFF), f0[3] = (int)(p.y >> 16);
        }
        int y[4];
        AV1B_UNROLL
        for (int m = 0; m < 4; m++) {
            const int u = (int)((xw >> (8 * m)) & 0xFF) << 4;
            // v = w1 * u + w0 * flt0 + w2 * flt1, a missing pass contributes u (LoopRestoration.cpp:466-474)
            int v = o.w1 * u;
            if (PASS == 0) v += o.w0 * f[m] + o.w2 * u;
            else v += o.w0 * (o.have_flt0 ? f0[m] : u) + o.w2 * f[m];
            y[m] = (v + (1 << 10)) >> 11;
        }
        uint8_t* d = o.dst + (size_t)i * o.stride + 4 * q;
        const uint32_t word = pack_u8x4(y[0], y[1], y[2], y[3]);
        if (4 * q + 4 <= o.w) *(uint32_t*)d = word;
        else
            for (int m = 0; 4 * q + m < o.w; m++) d[m] = (uint8_t)(word >> (8 * m));
    }
}

struct LrGrid {
    uint8_t narrow[3]; // plane uses 32-wide tiles (restoration units of 32 samples)
    uint8_t pad;
};

// One CTA = one TW x 32 tile of one plane.
template <int TW> AV1B_DEV void lr_tile(const PostCtx& c, const int plane, uint8_t* smem_raw)
{
    typedef LrGeom<TW> G;
    LrSmem<TW>& S = *reinterpret_cast<LrSmem<TW>*>(smem_raw);
    const PostHdr* hdr = &c.h;
    const Av1bLrParams& lp = hdr->lr;
    // grid = (tiles per row, tile rows, planes of this launch); everything below is shifts: unit
    // sizes are powers of two and a stripe holds one (chroma) or two (luma) tile rows
    const int sub = plane ? 1 : 0;
    const int pw = (hdr->frame_w + sub) >> sub, ph = (hdr->frame_h + sub) >> sub;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int x0 = blockIdx.x * TW;
    const int ty = blockIdx.y;
    const int stripe = sub ? ty : ty >> 1, part = sub ? 0 : ty & 1;
    const int start = (-8 + stripe * 64) >> sub, end = start + (64 >> sub);
    const int ya = start + part * LR_TH;
    const int y0 = max(ya, 0), y1 = min(min(ya + LR_TH, end), ph);
    if (x0 >= pw || y0 >= y1) return;
    const int w = min((int)TW, pw - x0), h = y1 - y0;
    const PlaneView cdef = c.cdef.pl[plane], deb = c.deb.pl[plane], out = c.lr.pl[plane];
    int type = 0;
    uint32_t uw0 = 0, uw1 = 0, uw2 = 0; // the unit record as three words (kept in registers)
    if (lp.frame_type[plane]) {
        const int ush = 31 - __clz((int)lp.unit_size[plane]);
        const int urow = min((int)lp.unit_rows[plane] - 1, (y0 + (8 >> sub)) >> ush);
        const int ucol = min((int)lp.unit_cols[plane] - 1, x0 >> ush);
        const uint32_t* up = (const uint32_t*)((const Av1bLrUnit*)(c.cmd + hdr->off_lru) + lp.unit_first[plane] + urow * lp.unit_cols[plane] + ucol);
        uw0 = __ldg(up), uw1 = __ldg(up + 1), uw2 = __ldg(up + 2);
        type = (int)(uw0 & 0xFF);
    }
    // Av1bLrUnit: type, sgr_set, sgr_xqd[2] | wiener[0][0..2], wiener[1][0] | wiener[1][1..2], pad
    const int sgr_set = (int)((uw0 >> 8) & 0xFF);
    const int xqd0 = (int)(int8_t)(uw0 >> 16), xqd1 = (int)(int8_t)(uw0 >> 24);
    const int wv[3] = { (int)(int8_t)uw1, (int)(int8_t)(uw1 >> 8), (int)(int8_t)(uw1 >> 16) };
    const int wh[3] = { (int)(int8_t)(uw1 >> 24), (int)(int8_t)uw2, (int)(int8_t)(uw2 >> 8) };
    uint8_t* dst = out.p + (size_t)y0 * out.stride + x0;
    if (type == 0) {
        // RESTORE_NONE: the LR frame is a copy of the CDEF frame.  128-bit copies (rows are 16-byte
        // aligned and padded beyond the frame width).
        const int chunks = (w + 15) >> 4;
        for (int e = tid; e < h * (TW / 16); e += nt) {
            const int i = e / (TW / 16), k = e - i * (TW / 16);
            if (k >= chunks) continue;
            const uint8_t* srow = cdef.p + (size_t)(y0 + i) * cdef.stride + x0;
            *(uint4*)(dst + (size_t)i * out.stride + 16 * k) = __ldg((const uint4*)srow + k);
        }
        return;
    }
    // ---- stage source: tile rows -3 .. h+2, columns -8 .. TW+7 as 64-bit chunks.  A group of 16 (8)
    //      lanes takes one row: the row's source (stripe rule) is resolved once per lane and row
    {
        const bool interior = x0 >= 8 && x0 + TW + 8 <= pw;
        const int nch = G::SP / 8, lpr = TW == 64 ? 16 : 8; // chunks per row, lanes per row
        for (int e = tid; e < (h + 6) * lpr; e += nt) {
            const int r = e / lpr, k = e & (lpr - 1);
            if (k >= nch) continue;
            bool fd;
            const int sy = lr_source_row(y0 - 3 + r, start, end, ph, &fd);
            const uint8_t* rowp = (fd ? deb.p : cdef.p) + (size_t)sy * (fd ? deb.stride : cdef.stride);
            const int xs = x0 - 8 + 8 * k;
            uint2 v;
            if (interior) {
                v = __ldg((const uint2*)(rowp + xs));
            } else {
                v.x = v.y = 0;
                AV1B_UNROLL
                for (int b = 0; b < 4; b++) {
                    v.x |= (uint32_t)__ldg(rowp + clip3(0, pw - 1, xs + b)) << (8 * b);
                    v.y |= (uint32_t)__ldg(rowp + clip3(0, pw - 1, xs + 4 + b)) << (8 * b);
                }
            }
            *(uint2*)(S.src + r * G::SP + 8 * k) = v;
        }
    }
    if (type == 2) {
        for (int z = tid; z < 256; z += nt) S.xdiv[z] = k_sgr_xdiv[z];
    }
    __syncthreads();
    if (type == 1) {
        const int vf[4] = { wv[0], wv[1], wv[2], 128 - 2 * (wv[0] + wv[1] + wv[2]) };
        const int hf[4] = { wh[0], wh[1], wh[2], 128 - 2 * (wh[0] + wh[1] + wh[2]) };
        // horizontal: s = sum hf[t] * x[c - 3 + t]; the centre tap (0..218) is split in two so that
        // both 4-tap groups fit signed bytes: (h0,h1,h2,c3a) . x[c-3..c] + (c3b,h2,h1,h0) . x[c..c+3]
        const int c3a = hf[3] >> 1, c3b = hf[3] - c3a;
        const uint32_t ta = (uint32_t)(hf[0] & 0xFF) | ((uint32_t)(hf[1] & 0xFF) << 8) | ((uint32_t)(hf[2] & 0xFF) << 16) | ((uint32_t)c3a << 24);
        const uint32_t tb = (uint32_t)c3b | ((uint32_t)(hf[2] & 0xFF) << 8) | ((uint32_t)(hf[1] & 0xFF) << 16) | ((uint32_t)(hf[0] & 0xFF) << 24);
        for (int e = tid; e < (h + 6) * G::OG; e += nt) {
            const int r = e / G::OG, g = e - r * G::OG;
            const uint32_t* sp = (const uint32_t*)(S.src + r * G::SP + 4 * g + 4);
            const uint32_t w0 = sp[0], w1 = sp[1], w2 = sp[2]; // tile columns 4g-4 .. 4g+7
            const uint32_t wa[4] = { __funnelshift_r(w0, w1, 8), __funnelshift_r(w0, w1, 16), __funnelshift_r(w0, w1, 24), w1 };
            const uint32_t wb[4] = { w1, __funnelshift_r(w1, w2, 8), __funnelshift_r(w1, w2, 16), __funnelshift_r(w1, w2, 24) };
            int q[4];
            AV1B_UNROLL
            for (int m = 0; m < 4; m++) {
                const int s = av1b_dp4a_us(wa[m], ta, av1b_dp4a_us(wb[m], tb, 4));
                q[m] = clip3(-2048, 6143, s >> 3);
            }
            *(uint4*)(S.wien + r * TW + 4 * g) = make_uint4((uint32_t)q[0], (uint32_t)q[1], (uint32_t)q[2], (uint32_t)q[3]);
        }
        __syncthreads();
        // vertical: one item = 4 columns x 4 rows out of ten 128-bit intermediate rows
        for (int e = tid; e < ((h + 3) >> 2) * G::OG; e += nt) {
            const int rq = e / G::OG, g = e - rq * G::OG;
            int q[10][4];
            AV1B_UNROLL
            for (int k = 0; k < 10; k++) {
                const uint4 v = *(const uint4*)(S.wien + (4 * rq + k) * TW + 4 * g);
                q[k][0] = (int)v.x, q[k][1] = (int)v.y, q[k][2] = (int)v.z, q[k][3] = (int)v.w;
            }
            AV1B_UNROLL
            for (int i = 0; i < 4; i++) {
                if (4 * rq + i >= h) break;
                int y[4];
                AV1B_UNROLL
                for (int m = 0; m < 4; m++) {
                    const int s = vf[0] * (q[i][m] + q[i + 6][m]) + vf[1] * (q[i + 1][m] + q[i + 5][m]) + vf[2] * (q[i + 2][m] + q[i + 4][m])
                        + vf[3] * q[i + 3][m];
                    y[m] = (s + 1024) >> 11;
                }
                uint8_t* d = dst + (size_t)(4 * rq + i) * out.stride + 4 * g;
                const uint32_t word = pack_u8x4(y[0], y[1], y[2], y[3]);
                if (4 * g + 4 <= w) *(uint32_t*)d = word;
                else
                    for (int m = 0; 4 * g + m < w; m++) d[m] = (uint8_t)(word >> (8 * m));
            }
        }
    } else {
        const int set = sgr_set;
        const int r0 = k_sgr_params[set][0], r1 = k_sgr_params[set][2];
        SgrOut o;
        o.w0 = xqd0;
        o.w1 = xqd1;
        o.w2 = 128 - o.w0 - o.w1;
        o.dst = dst;
        o.stride = out.stride;
        o.w = w;
        o.have_flt0 = r0 != 0;
        if (r0) {
            sgr_box_rows<TW, 2>(S, h, tid, nt);
            __syncthreads();
            sgr_ab<TW, 2>(S, h, set, tid, nt);
            __syncthreads();
            o.final = r1 == 0;
            sgr_out<TW, 0>(S, h, o, tid, nt);
            __syncthreads();
        }
        if (r1) {
            sgr_box_rows<TW, 1>(S, h, tid, nt);
            __syncthreads();
            sgr_ab<TW, 1>(S, h, set, tid, nt);
            __syncthreads();
            o.final = true;
            sgr_out<TW, 1>(S, h, o, tid, nt);
        }
    }
}

}  // namespace

// grid = (tiles per row, tile rows, plane), both maxima over the planes.  Planes whose restoration
// units are at least 64 samples wide use 64-wide tiles, the others (chroma of a 64-unit luma with
// lr_uv_shift) 32-wide ones -- in the SAME launch: at 4K a launch of a few thousand short CTAs
// spends a third of its time ramping up and draining.
__global__ void __launch_bounds__(LR_THREADS) lr_kernel(PostCtx c, LrGrid grid)
{
    alignas(16) __shared__ uint8_t smem_raw[sizeof(LrSmem<64>)];
    const int plane = blockIdx.z;
    if (grid.narrow[plane]) lr_tile<32>(c, plane, smem_raw);
    else lr_tile<64>(c, plane, smem_raw);
}

int launch_lr(const PostCtx& c, const Av1bFrameHdr& h, av1b_stream_t st)
{
    if (!h.lr.uses_lr) return 0;
    LrGrid g;
    memset(&g, 0, sizeof(g));
    int gx = 0, gy = 0;
    for (int p = 0; p < 3; p++) {
        const int sub = p ? 1 : 0;
        const int pw = (h.frame_w + sub) >> sub, ph = (h.frame_h + sub) >> sub;
        g.narrow[p] = h.lr.frame_type[p] && h.lr.unit_size[p] < 64;
        const int tw = g.narrow[p] ? 32 : 64;
        const int stripes = (ph + (8 >> sub) + (64 >> sub) - 1) / (64 >> sub);
        gx = std::max(gx, (pw + tw - 1) / tw);
        gy = std::max(gy, stripes * ((64 >> sub) / LR_TH));
    }
    AV1B_LAUNCH(lr_kernel, (gx, gy, 3), (LR_THREADS), st, c, g);
    return 1;
}

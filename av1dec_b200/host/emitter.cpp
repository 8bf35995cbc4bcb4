// emitter.cpp -- see emitter.h.  Every function cites the reference code whose *reads* it
// replays; nothing here computes a pixel.
#include "ref_access.h"
#include "emitter.h"
#include "../csrc/av1_tables_host.h"

#include <cstdio>
#include <cstdlib>

using namespace YamiAv1;

namespace av1b200 {

static const uint32_t kNoAux = 0xFFFFFFFFu;

template <class T> static inline T clip3(T lo, T hi, T v) { return v < lo ? lo : (v > hi ? hi : v); }

void FrameEmitter::begin(FrameHeader& frame, const SequenceHeader& seq)
{
    m_frame = &frame;
    m_seq = &seq;
    memset(&m_hdr, 0, sizeof(m_hdr));
    Av1bFrameHdr& h = m_hdr;
    h.magic = AV1B_MAGIC;
    h.version = AV1B_FORMAT_VERSION;
    h.frame_w = (uint16_t)frame.FrameWidth;
    h.frame_h = (uint16_t)frame.FrameHeight;
    h.mi_cols = (uint16_t)frame.MiCols;
    h.mi_rows = (uint16_t)frame.MiRows;
    h.sb_log2 = seq.use_128x128_superblock ? 7 : 6;
    const int sb4 = 1 << (h.sb_log2 - 2);
    h.sb_cols = (uint16_t)((frame.MiCols + sb4 - 1) / sb4);
    h.sb_rows = (uint16_t)((frame.MiRows + sb4 - 1) / sb4);
    h.enable_intra_edge_filter = seq.enable_intra_edge_filter;
    h.frame_is_intra = frame.FrameIsIntra;
    h.allow_intrabc = frame.allow_intrabc;
    // reference slots and dimensions (InterPredict.cpp:389-394,982)
    for (int i = 0; i < 8; i++) h.ref_slot[i] = -1;
    h.ref_w[0] = (uint16_t)(frame.MiCols * MI_SIZE); // intrabc "reference" (InterPredict.cpp:990-995)
    h.ref_h[0] = (uint16_t)(frame.MiRows * MI_SIZE);
    if (!frame.FrameIsIntra) {
        for (int rf = LAST_FRAME; rf <= ALTREF_FRAME; rf++) {
            const int slot = frame.ref_frame_idx[rf - LAST_FRAME];
            const RefFrame& r = frame.m_refInfo.m_refs[slot];
            h.ref_slot[rf] = (int8_t)slot;
            h.ref_w[rf] = (uint16_t)r.RefUpscaledWidth;
            h.ref_h[rf] = (uint16_t)r.RefFrameHeight;
            for (int k = 0; k < 6; k++) h.gm_params[rf][k] = frame.gm_params[rf][k];
        }
    }
    m_gmReady = false;
    {
        Av1bSb zero;
        memset(&zero, 0, sizeof(zero));
        m_sbs.assign((size_t)h.sb_cols * h.sb_rows, zero);
        m_sbDepth.assign((size_t)h.sb_cols * h.sb_rows, 0);
    }
    m_ops.clear();
    m_itxOnly.clear();
    m_itx.clear();
    m_iblk.clear();
    m_ipu.clear();
    m_aux.clear();
    m_coef.clear();
    m_pal.clear();
    m_lru.clear();
    m_lftx.assign((size_t)3 * frame.MiRows * frame.MiCols, 0);
    m_nRes = 0;
    m_total = 0;
}

// AV1B200_WAVE_OVERLAP=0 leaves the overlap hints zero (the classic two-superblock-lag wavefront)
// Seeds (percent of a neighbour's depth, AV1B200_WAVE_SEED="L,A1,A2"): see scheduleSb.
struct WaveSeeds {
    int l2, a1, a2;
};
static const WaveSeeds& getenv_seeds()
{
    static const WaveSeeds s = [] {
        WaveSeeds v = { 50, 50, 80 };
        if (const char* e = getenv("AV1B200_WAVE_SEED")) sscanf(e, "%d,%d,%d", &v.l2, &v.a1, &v.a2);
        return v;
    }();
    return s;
}

// Intra blocks of at least this many samples are split into row strips of 256 samples, one op
// (one warp) per strip (AV1B200_WAVE_SPLIT, 0 = never).
static int getenv_split()
{
    static const int v = [] {
        const char* e = getenv("AV1B200_WAVE_SPLIT");
        return e ? atoi(e) : 512;
    }();
    return v;
}

static bool getenv_overlap()
{
    static const bool on = [] {
        const char* e = getenv("AV1B200_WAVE_OVERLAP");
        return !(e && atoi(e) == 0);
    }();
    return on;
}

void FrameEmitter::emitTile(Tile& tile)
{
    for (auto& sbp : tile.m_sbs) emitSb(tile, *sbp);
}

void FrameEmitter::emitSb(Tile& tile, SuperBlock& sb)
{
    const int sb4 = 1 << (m_hdr.sb_log2 - 2);
    // SuperBlock::decode (SuperBlock.cpp:46-51)
    tile.m_decoded.clear_block_decoded_flags(sb.m_r, sb.m_c, sb4);
    const size_t idx = (size_t)(sb.m_r / sb4) * m_hdr.sb_cols + (sb.m_c / sb4);
    const uint32_t first = (uint32_t)m_ops.size();
    const size_t firstItx = m_itx.size();
    walk(sb);
    m_sbs[idx].first_op = first;
    m_sbs[idx].n_ops = (uint32_t)m_ops.size() - first;
    scheduleSb(first, firstItx, sb.m_c * MI_SIZE, sb.m_r * MI_SIZE);
}

// Dependency levels inside one superblock.  Ops are emitted in bitstream order; two ops may run
// concurrently when neither reads nor overwrites samples the other writes.  Per plane, a 4x4-cell
// map remembers the level of the last op that wrote each cell; an op's level is one more than the
// largest level among the cells it reads (intra edges incl. above-right / below-left, CfL luma)
// or rewrites.  The ops are then stably sorted by level (still a valid sequential order) and the
// level (low half) and the run length to the end of the level (high half) are stored in
// Av1bOp::res_off, which frame submits do not otherwise use.
void FrameEmitter::scheduleSb(uint32_t first, size_t firstItx, int sbx, int sby)
{
    const uint32_t n = (uint32_t)m_ops.size() - first;
    if (!n) return;
    const int sbPix = 1 << m_hdr.sb_log2;
    enum { MAXC = 32 };
    static thread_local uint16_t cell[3][MAXC][MAXC];
    memset(cell, 0, sizeof(cell));
    m_levels.resize(n);
    uint32_t maxLevel = 0;
    // Ops that read the part of a neighbouring superblock that becomes final LATE -- the lower half
    // of the left superblock's right column, the above-right superblock's bottom row -- are seeded
    // with a level by which that part can be expected (the device starts a superblock when the one
    // above is finished: the left neighbour is then about two thirds through, the above-right one a
    // third; a level of theirs is taken to last as long as one of ours).  Without the seed such an
    // op sits in level 1 or 2 whenever the blocks before it are not intra, and the whole superblock
    // waits for the neighbour before its second level.  Their dependents follow by the usual rule.
    const uint32_t sbIdx = (uint32_t)((sby >> m_hdr.sb_log2) * m_hdr.sb_cols + (sbx >> m_hdr.sb_log2));
    const bool haveLeftSb = sbx > 0, haveAboveSb = sby > 0;
    const bool overlap = getenv_overlap();
    uint32_t seedL2 = 0, seedA1 = 0, seedA2 = 0; // levels BEFORE the op's own
    if (overlap) {
        const WaveSeeds& sd = getenv_seeds();
        if (haveLeftSb) seedL2 = m_sbDepth[sbIdx - 1] * sd.l2 / 100;
        if (haveAboveSb && (sbx >> m_hdr.sb_log2) + 1 < (int)m_hdr.sb_cols) {
            const uint32_t d = m_sbDepth[sbIdx - m_hdr.sb_cols + 1];
            seedA1 = d * sd.a1 / 100;
            seedA2 = d * sd.a2 / 100;
        }
    }
    for (uint32_t k = 0; k < n; k++) {
        Av1bOp& op = m_ops[first + k];
        const int pl = op.plane, sub = pl ? 1 : 0;
        const int nc = (sbPix >> sub) >> 2; // cells per row/column of this plane's SB tile
        int lw, lh;
        if (op.kind == AV1B_OP_INTERINTRA || op.kind == AV1B_OP_INTRABC) {
            lw = op.tx_size & 15;
            lh = op.tx_size >> 4;
        } else {
            lw = hk_tx_wlog2[op.tx_size];
            lh = hk_tx_hlog2[op.tx_size];
        }
        if (op.kind == AV1B_OP_INTRABC) {
            const Av1bIpu& u = m_ipu[op.aux];
            lw = 0;
            while ((1 << lw) < u.w) lw++;
            lh = 0;
            while ((1 << lh) < u.h) lh++;
        }
        if (k > 0 && op.kind == AV1B_OP_INTRA && !(op.flags & AV1B_OPF_FILTER_INTRA) && (op.fi_mode >> 5)) {
            m_levels[k] = m_levels[k - 1]; // a further strip of the block before: same reads, disjoint rows
            continue;
        }
        const int w = 1 << lw, h = 1 << lh;
        const int x = op.x - (sbx >> sub), y = op.y - (sby >> sub); // tile-relative
        const int cx0 = std::max(0, x >> 2), cy0 = std::max(0, y >> 2);
        const int cx1 = std::min(nc - 1, (x + w - 1) >> 2), cy1 = std::min(nc - 1, (y + h - 1) >> 2);
        uint32_t lvl = 0;
        auto rd = [&](int p, int cxa, int cxb, int cya, int cyb, int ncp) {
            cxa = std::max(cxa, 0);
            cya = std::max(cya, 0);
            cxb = std::min(cxb, ncp - 1);
            cyb = std::min(cyb, ncp - 1);
            for (int cy = cya; cy <= cyb; cy++)
                for (int cx = cxa; cx <= cxb; cx++) lvl = std::max<uint32_t>(lvl, cell[p][cy][cx]);
        };
        rd(pl, cx0, cx1, cy0, cy1, nc); // cells it (re)writes
        if (op.kind == AV1B_OP_INTRA || op.kind == AV1B_OP_INTERINTRA) {
            const int ar = (op.flags & AV1B_OPF_HAVE_ABOVE_RIGHT) ? 2 * w : w;
            const int bl = (op.flags & AV1B_OPF_HAVE_BELOW_LEFT) ? 2 * h : h;
            if (y > 0) rd(pl, (x - 1) >> 2, (x + ar - 1) >> 2, (y - 1) >> 2, (y - 1) >> 2, nc); // above row + corner
            if (x > 0) rd(pl, (x - 1) >> 2, (x - 1) >> 2, (y - 1) >> 2, (y + bl - 1) >> 2, nc); // left column + corner
            if (op.flags & AV1B_OPF_CFL) rd(0, (2 * x) >> 2, (2 * (x + w) - 1) >> 2, (2 * y) >> 2, (2 * (y + h) - 1) >> 2, nc * 2);
            const int np = sbPix >> sub, q = np >> 1;
            if (x <= 0 && y + bl > q) lvl = std::max(lvl, seedL2);
            if (y <= 0 && x + ar > np) lvl = std::max(lvl, x + ar - np > q ? seedA2 : seedA1);
        }
        lvl += 1;
        for (int cy = cy0; cy <= cy1; cy++)
            for (int cx = cx0; cx <= cx1; cx++) cell[pl][cy][cx] = (uint16_t)lvl;
        m_levels[k] = lvl;
        maxLevel = std::max(maxLevel, lvl);
    }
    // ---- overlap hints (av1b200_format.h, Av1bSb): the first level that reads each half of the left
    // superblock's right column and of the above-right superblock's bottom row (from the
    // availability flags), and the level after which the halves of this superblock's own border
    // that are announced early are final (the last op that writes them).
    m_sbDepth[sbIdx] = (uint16_t)std::min<uint32_t>(maxLevel, 0xFFFF);
    uint32_t wl1 = 0xFF, wl2 = 0xFF, wa1 = 0xFF, wa2 = 0xFF, pr1 = 1, pb1 = 1;
    bool hints = overlap && maxLevel < 0xFF;
    for (uint32_t k = 0; k < n && hints; k++) {
        const Av1bOp& op = m_ops[first + k];
        const int sub = op.plane ? 1 : 0;
        const int np = sbPix >> sub, q = np >> 1;
        int lw, lh;
        if (op.kind == AV1B_OP_INTERINTRA || op.kind == AV1B_OP_INTRABC) {
            lw = op.tx_size & 15;
            lh = op.tx_size >> 4;
        } else {
            lw = hk_tx_wlog2[op.tx_size];
            lh = hk_tx_hlog2[op.tx_size];
        }
        if (op.kind == AV1B_OP_INTRABC) {
            hints = false;
            break;
        }
        const int w = 1 << lw, h = 1 << lh;
        const int x = op.x - (sbx >> sub), y = op.y - (sby >> sub);
        const uint32_t lvl = m_levels[k];
        if (op.kind == AV1B_OP_INTRA || op.kind == AV1B_OP_INTERINTRA) {
            if (x <= 0 && haveLeftSb) { // the left superblock's right column, rows y-1 .. y+bl-1
                const int bl = (op.flags & AV1B_OPF_HAVE_BELOW_LEFT) ? 2 * h : h;
                if (y - 1 < q) wl1 = std::min(wl1, lvl);
                if (y + bl > q) wl2 = std::min(wl2, lvl);
            }
            if (y <= 0 && haveAboveSb) { // the row above, columns x-1 .. x+ar-1
                const int ar = (op.flags & AV1B_OPF_HAVE_ABOVE_RIGHT) ? 2 * w : w;
                if (x + ar > np) {
                    wa1 = std::min(wa1, lvl);
                    if (x + ar - np > q) wa2 = std::min(wa2, lvl);
                }
            }
        }
        if (x + w >= np && y < q) pr1 = std::max(pr1, lvl);
        if (y + h >= np && x < q) pb1 = std::max(pb1, lvl);
    }
    // stable counting sort by level
    m_count.assign(maxLevel + 2, 0);
    for (uint32_t k = 0; k < n; k++) m_count[m_levels[k] + 1]++;
    for (uint32_t l = 1; l < m_count.size(); l++) m_count[l] += m_count[l - 1];
    m_sorted.resize(n);
    m_perm.resize(n);
    for (uint32_t k = 0; k < n; k++) {
        const uint32_t pos = m_count[m_levels[k]]++;
        m_sorted[pos] = m_ops[first + k];
        m_sorted[pos].res_off = m_levels[k] & 0xFFFF;
        m_perm[k] = pos;
    }
    // res_off = level | (ops left in this level, this one included) << 16: the device finds the end
    // of a level without scanning
    for (uint32_t k = n, rem = 0; k-- > 0;) {
        rem = (k + 1 < n && (m_sorted[k + 1].res_off & 0xFFFF) == m_sorted[k].res_off) ? std::min<uint32_t>(rem + 1, 0xFFFF) : 1;
        m_sorted[k].res_off |= rem << 16;
    }
    std::copy(m_sorted.begin(), m_sorted.end(), m_ops.begin() + first);
    for (size_t i = firstItx; i < m_itx.size(); i++)
        if (!(m_itx[i] & 0x80000000u)) m_itx[i] = first + m_perm[m_itx[i] - first];
    Av1bSb& e = m_sbs[sbIdx];
    if (hints) {
        e.wait_l1 = (uint8_t)wl1, e.wait_l2 = (uint8_t)wl2, e.wait_a1 = (uint8_t)wa1, e.wait_a2 = (uint8_t)wa2;
        e.pub_r1 = (uint8_t)pr1, e.pub_b1 = (uint8_t)pb1;
    } else {
        e.wait_l1 = e.wait_l2 = e.wait_a1 = e.wait_a2 = e.pub_r1 = e.pub_b1 = 0; // the classic wavefront
    }
}

void FrameEmitter::walk(Partition& p)
{
    // Partition::decode (Partition.cpp:207-214): children in parse order
    for (auto& bt : p.m_blocks) {
        if (Block* b = dynamic_cast<Block*>(bt.get())) emitBlock(*b);
        else walk(*static_cast<Partition*>(bt.get()));
    }
}

void FrameEmitter::emitBlock(Block& b)
{
    // Block::decode (Block.cpp:1600-1608): prediction, then the transform blocks in order
    m_blockAux = kNoAux;
    if (b.is_inter) emitInter(b);
    for (auto& t : b.m_transformBlocks) emitTb(b, *t);
}

uint32_t FrameEmitter::auxFor(Block& b)
{
    if (m_blockAux != kNoAux) return m_blockAux;
    Av1bBlkAux a;
    memset(&a, 0, sizeof(a));
    a.mi_size = (uint8_t)b.MiSize;
    if (b.is_inter) {
        a.interintra_mode = (uint8_t)b.interintra_mode;
        a.wedge_interintra = b.interintra ? b.wedge_interintra : 0;
        a.wedge_index = b.wedge_index;
        a.wedge_sign = b.wedge_sign;
        a.mask_type = b.mask_type;
        if (b.motion_mode == LOCALWARP && b.m_localWarp.LocalValid) {
            int al, be, ga, de;
            b.m_localWarp.setupShear(b.m_localWarp.LocalWarpParams, al, be, ga, de);
            for (int k = 0; k < 6; k++) a.warp_params[k] = b.m_localWarp.LocalWarpParams[k];
            a.warp_abgd[0] = (int16_t)al;
            a.warp_abgd[1] = (int16_t)be;
            a.warp_abgd[2] = (int16_t)ga;
            a.warp_abgd[3] = (int16_t)de;
        }
    } else {
        // Palette::predict_palette inputs (Block.cpp:2279-2297)
        const auto& pal = b.m_palette;
        a.pal_size_y = b.PaletteSizeY;
        a.pal_size_uv = b.PaletteSizeUV;
        for (int i = 0; i < b.PaletteSizeY && i < 8; i++) a.pal_colors[0][i] = pal.palette_colors_y[i];
        for (int i = 0; i < b.PaletteSizeUV && i < 8; i++) {
            a.pal_colors[1][i] = pal.palette_colors_u[i];
            a.pal_colors[2][i] = pal.palette_colors_v[i];
        }
        const std::vector<std::vector<uint8_t>>* maps[2] = { &pal.ColorMapY, &pal.ColorMapUV };
        const int sizes[2] = { b.PaletteSizeY, b.PaletteSizeUV };
        for (int k = 0; k < 2; k++) {
            if (!sizes[k] || maps[k]->empty()) continue;
            const size_t rows = maps[k]->size(), cols = (*maps[k])[0].size();
            a.pal_map_off[k] = (uint32_t)m_pal.size();
            a.pal_map_stride[k] = (uint16_t)cols;
            for (size_t r = 0; r < rows; r++) m_pal.insert(m_pal.end(), (*maps[k])[r].begin(), (*maps[k])[r].end());
        }
        a.base_x[0] = (uint16_t)(b.MiCol * MI_SIZE);
        a.base_y[0] = (uint16_t)(b.MiRow * MI_SIZE);
        a.base_x[1] = (uint16_t)((b.MiCol >> b.subsampling_x) * MI_SIZE);
        a.base_y[1] = (uint16_t)((b.MiRow >> b.subsampling_y) * MI_SIZE);
    }
    m_aux.push_back(a);
    m_blockAux = (uint32_t)m_aux.size() - 1;
    return m_blockAux;
}

// get_filter_type (IntraPredict.cpp:211-267): does the left or above neighbour use a SMOOTH mode
bool FrameEmitter::edgeSmooth(const Block& b, int plane) const
{
    auto smooth = [&](int r, int c) -> bool {
        const ModeInfoBlock& info = m_frame->getModeInfo(r, c);
        int mode;
        if (!plane) mode = info.YMode;
        else {
            if (info.RefFrames[0] > INTRA_FRAME) return false;
            mode = info.UVMode;
        }
        return mode == SMOOTH_PRED || mode == SMOOTH_V_PRED || mode == SMOOTH_H_PRED;
    };
    bool above = false, left = false;
    if (plane ? b.AvailUChroma : b.AvailU) {
        int r = b.MiRow - 1, c = b.MiCol;
        if (plane > 0) {
            if (b.subsampling_x && !(b.MiCol & 1)) c++;
            if (b.subsampling_y && (b.MiRow & 1)) r--;
        }
        above = smooth(r, c);
    }
    if (plane ? b.AvailLChroma : b.AvailL) {
        int r = b.MiRow, c = b.MiCol - 1;
        if (plane > 0) {
            if (b.subsampling_x && (b.MiCol & 1)) c--;
            if (b.subsampling_y && !(b.MiRow & 1)) r++;
        }
        left = smooth(r, c);
    }
    return above || left;
}

// getDistanceWeights (InterPredict.cpp:917-960)
void FrameEmitter::distanceWeights(int candRow, int candCol, int& fwd, int& bck) const
{
    int dist[2];
    for (int l = 0; l < 2; l++) {
        const uint8_t ref = (uint8_t)m_frame->getModeInfo(candRow, candCol).RefFrames[l];
        dist[l] = clip3(0, (int)MAX_FRAME_DISTANCE, std::abs((int)m_frame->get_relative_dist(ref)));
    }
    const int d0 = dist[1], d1 = dist[0];
    const int order = d0 <= d1;
    int i = 3;
    if (d0 != 0 && d1 != 0) {
        for (i = 0; i < 3; i++) {
            const int c0 = hk_quant_dist_weight[i][order], c1 = hk_quant_dist_weight[i][1 - order];
            if (order ? (d0 * c0 > d1 * c1) : (d0 * c0 < d1 * c1)) break;
        }
    }
    fwd = hk_quant_dist_lookup[i][order];
    bck = hk_quant_dist_lookup[i][1 - order];
}

void FrameEmitter::emitInter(Block& b)
{
    FrameHeader& f = *m_frame;
    const SequenceHeader& seq = *m_seq;
    // Global-motion shear parameters once per frame (InterPredict.cpp:974-977); needs any block's
    // LocalWarp object to reach the reference's setupShear().
    if (!m_gmReady) {
        for (int rf = LAST_FRAME; rf <= ALTREF_FRAME; rf++) {
            m_hdr.gm_warp_ok[rf] = 0;
            if (!f.FrameIsIntra && f.GmType[rf] > TRANSLATION && !f.is_scaled(rf)) {
                int al, be, ga, de;
                const bool ok = b.m_localWarp.setupShear(f.gm_params[rf], al, be, ga, de);
                m_hdr.gm_abgd[rf][0] = (int16_t)al;
                m_hdr.gm_abgd[rf][1] = (int16_t)be;
                m_hdr.gm_abgd[rf][2] = (int16_t)ga;
                m_hdr.gm_abgd[rf][3] = (int16_t)de;
                m_hdr.gm_warp_ok[rf] = ok;
            }
        }
        m_gmReady = true;
    }
    const bool interintra = b.RefFrame[1] == INTRA_FRAME;
    const bool intrabc = b.use_intrabc;
    if (b.motion_mode == LOCALWARP) {
        // predict_inter, plane 0 (InterPredict.cpp:967-970)
        b.m_localWarp.warpEstimation();
        b.m_localWarp.setupShear();
    }
    Av1bInterBlk ib;
    memset(&ib, 0, sizeof(ib));
    ib.first_ipu = (uint32_t)m_ipu.size();
    ib.x = (uint16_t)(b.MiCol * MI_SIZE);
    ib.y = (uint16_t)(b.MiRow * MI_SIZE);
    ib.bw = (uint8_t)b.bw;
    ib.bh = (uint8_t)b.bh;
    if (b.HasChroma) {
        const int csz = seq.get_plane_residual_size(b.MiSize, 1);
        ib.flags |= AV1B_IBF_HAS_CHROMA;
        ib.cx = (uint16_t)((b.MiCol >> b.subsampling_x) * MI_SIZE);
        ib.cy = (uint16_t)((b.MiRow >> b.subsampling_y) * MI_SIZE);
        ib.cw = (uint8_t)(Num_4x4_Blocks_Wide[csz] * 4);
        ib.ch = (uint8_t)(Num_4x4_Blocks_High[csz] * 4);
    }
    // plain inter blocks: the inter pass adds the residual itself (no ordering constraint);
    // inter-intra blocks get it after the blend, inside the dependent pass
    if (!interintra && !intrabc) ib.flags |= AV1B_IBF_ADD_RESIDUAL;
    const int subBlockMiRow = b.MiRow & b.sbMask, subBlockMiCol = b.MiCol & b.sbMask;
    // Block::compute_prediction (Block.cpp:100-174)
    for (int plane = 0; plane < 1 + b.HasChroma * 2; plane++) {
        const int planeSz = seq.get_plane_residual_size(b.MiSize, plane);
        const int num4x4W = Num_4x4_Blocks_Wide[planeSz], num4x4H = Num_4x4_Blocks_High[planeSz];
        const int subX = plane ? b.subsampling_x : 0, subY = plane ? b.subsampling_y : 0;
        const int baseX = (b.MiCol >> subX) * MI_SIZE, baseY = (b.MiRow >> subY) * MI_SIZE;
        int candRow = (b.MiRow >> subY) << subY, candCol = (b.MiCol >> subX) << subX;
        if (interintra) {
            Av1bOp op;
            memset(&op, 0, sizeof(op));
            op.x = (uint16_t)baseX;
            op.y = (uint16_t)baseY;
            op.plane = (uint8_t)plane;
            op.kind = AV1B_OP_INTERINTRA;
            const int log2W = MI_SIZE_LOG2 + Mi_Width_Log2[planeSz], log2H = MI_SIZE_LOG2 + Mi_Height_Log2[planeSz];
            op.tx_size = (uint8_t)(log2W | (log2H << 4));
            switch (b.interintra_mode) {
            case II_DC_PRED: op.mode = DC_PRED; break;
            case II_V_PRED: op.mode = V_PRED; break;
            case II_H_PRED: op.mode = H_PRED; break;
            default: op.mode = SMOOTH_PRED; break;
            }
            uint8_t fl = 0;
            if (plane == 0 ? b.AvailL : b.AvailLChroma) fl |= AV1B_OPF_HAVE_LEFT;
            if (plane == 0 ? b.AvailU : b.AvailUChroma) fl |= AV1B_OPF_HAVE_ABOVE;
            if (b.m_decoded.getFlag(plane, (subBlockMiRow >> subY) - 1, (subBlockMiCol >> subX) + num4x4W)) fl |= AV1B_OPF_HAVE_ABOVE_RIGHT;
            if (b.m_decoded.getFlag(plane, (subBlockMiRow >> subY) + num4x4H, (subBlockMiCol >> subX) - 1)) fl |= AV1B_OPF_HAVE_BELOW_LEFT;
            op.flags = fl;
            op.aux = auxFor(b);
            m_ops.push_back(op);
        }
        int predW = b.bw >> subX, predH = b.bh >> subY;
        bool someUseIntra = false;
        for (int r = 0; r < (num4x4H << subY); r++)
            for (int c = 0; c < (num4x4W << subX); c++)
                if (f.getModeInfo(candRow + r, candCol + c).RefFrames[0] == INTRA_FRAME) someUseIntra = true;
        if (someUseIntra) {
            predH = num4x4H * 4;
            predW = num4x4W * 4;
            candRow = b.MiRow;
            candCol = b.MiCol;
        }
        int r = 0;
        for (int y = 0; y < num4x4H * 4; y += predH) {
            int c = 0;
            for (int x = 0; x < num4x4W * 4; x += predW) {
                // InterPredict::predict_inter(baseX + x, baseY + y, predW, predH, candRow + r, candCol + c)
                const ModeInfoBlock& info = f.getModeInfo(candRow + r, candCol + c);
                const bool isCompound = info.RefFrames[1] > INTRA_FRAME;
                Av1bIpu u;
                memset(&u, 0, sizeof(u));
                u.x = (uint16_t)(baseX + x);
                u.y = (uint16_t)(baseY + y);
                u.w = (uint8_t)predW;
                u.h = (uint8_t)predH;
                u.plane = (uint8_t)plane;
                u.kind = AV1B_IPU_PRED;
                u.aux = kNoAux;
                u.ref_slot[0] = u.ref_slot[1] = -1;
                bool needAux = false;
                for (int l = 0; l < 1 + isCompound; l++) {
                    const int refFrame = info.RefFrames[l];
                    u.mv[l][0] = info.Mvs[l].mv[0];
                    u.mv[l][1] = info.Mvs[l].mv[1];
                    if (!intrabc) {
                        u.ref_slot[l] = (int8_t)f.ref_frame_idx[refFrame - LAST_FRAME];
                        u.ref_frame[l] = (uint8_t)refFrame;
                        // getUseWarp (InterPredict.cpp:50-64); the w/h >= 8 test runs on the device
                        if (!f.force_integer_mv) {
                            if (b.motion_mode == LOCALWARP && b.m_localWarp.LocalValid) {
                                u.warp[l] = 1;
                                needAux = true;
                            } else if ((b.YMode == GLOBALMV || b.YMode == GLOBAL_GLOBALMV) && m_hdr.gm_warp_ok[refFrame]) {
                                u.warp[l] = 2;
                            }
                        }
                    }
                }
                u.filt[0] = (uint8_t)info.InterpFilters[0];
                u.filt[1] = (uint8_t)info.InterpFilters[1];
                u.flags = (isCompound ? AV1B_IPUF_COMPOUND : 0) | (interintra ? AV1B_IPUF_INTERINTRA : 0) | (intrabc ? AV1B_IPUF_INTRABC : 0);
                u.comp_type = (uint8_t)b.compound_type;
                if (isCompound) {
                    if (b.compound_type == COMPOUND_DISTANCE) {
                        int fw, bk;
                        distanceWeights(candRow + r, candCol + c, fw, bk);
                        u.fwd_w = (uint8_t)fw;
                        u.bck_w = (uint8_t)bk;
                    } else if (b.compound_type == COMPOUND_WEDGE || b.compound_type == COMPOUND_DIFFWTD) {
                        needAux = true;
                    }
                }
                if (needAux) u.aux = auxFor(b);
                m_ipu.push_back(u);
                if (intrabc) {
                    Av1bOp op;
                    memset(&op, 0, sizeof(op));
                    op.x = u.x;
                    op.y = u.y;
                    op.plane = (uint8_t)plane;
                    op.kind = AV1B_OP_INTRABC;
                    op.aux = (uint32_t)m_ipu.size() - 1;
                    m_ops.push_back(op);
                }
                c++;
            }
            r++;
        }
        if (b.motion_mode == OBMC_CAUSAL) emitObmc(b, plane, predW, predH);
    }
    ib.n_ipu = (uint16_t)(m_ipu.size() - ib.first_ipu);
    if (!intrabc && ib.n_ipu) {
        // blocks whose units do not depend on each other are handled unit by unit: by the fast
        // translational kernel where a unit qualifies, by the general predictor otherwise
        bool indep = true;
        for (size_t k = ib.first_ipu; k < m_ipu.size() && indep; k++) {
            const Av1bIpu& u = m_ipu[k];
            indep = u.kind == AV1B_IPU_PRED && !((u.flags & AV1B_IPUF_COMPOUND) && u.comp_type == AV1B_COMP_DIFFWTD);
        }
        if (indep) {
            ib.flags |= AV1B_IBF_FAST;
            const uint8_t add = (ib.flags & AV1B_IBF_ADD_RESIDUAL) ? AV1B_IPUF_ADD_RES : 0;
            for (size_t k = ib.first_ipu; k < m_ipu.size(); k++) {
                Av1bIpu& u = m_ipu[k];
                const bool big = u.w >= 8 && u.h >= 8;
                const bool simple = u.w >= 4 && !(big && (u.warp[0] || u.warp[1]))
                    && (!(u.flags & AV1B_IPUF_COMPOUND) || u.comp_type == AV1B_COMP_AVERAGE || u.comp_type == AV1B_COMP_DISTANCE);
                u.flags |= (simple ? AV1B_IPUF_FAST : AV1B_IPUF_INDEP) | add;
            }
        }
        m_iblk.push_back(ib);
    }
}

// overlappedMotionCompensation (InterPredict.cpp:658-709): which neighbours contribute strips
void FrameEmitter::emitObmc(Block& b, int plane, int w, int h)
{
    FrameHeader& f = *m_frame;
    const int subX = plane ? b.subsampling_x : 0, subY = plane ? b.subsampling_y : 0;
    auto strip = [&](int kind, int candRow, int candCol, int x4, int y4, int predW, int predH) {
        const ModeInfoBlock& info = f.getModeInfo(candRow, candCol);
        Av1bIpu u;
        memset(&u, 0, sizeof(u));
        u.x = (uint16_t)((x4 * 4) >> subX);
        u.y = (uint16_t)((y4 * 4) >> subY);
        u.w = (uint8_t)predW;
        u.h = (uint8_t)predH;
        u.plane = (uint8_t)plane;
        u.kind = (uint8_t)kind;
        u.mv[0][0] = info.Mvs[0].mv[0];
        u.mv[0][1] = info.Mvs[0].mv[1];
        u.ref_slot[0] = (int8_t)f.ref_frame_idx[info.RefFrames[0] - LAST_FRAME];
        u.ref_slot[1] = -1;
        u.ref_frame[0] = (uint8_t)info.RefFrames[0];
        u.filt[0] = (uint8_t)info.InterpFilters[0];
        u.filt[1] = (uint8_t)info.InterpFilters[1];
        u.aux = kNoAux;
        m_ipu.push_back(u);
    };
    if (b.AvailU && m_seq->get_plane_residual_size(b.MiSize, plane) >= BLOCK_8X8) {
        const int w4 = Num_4x4_Blocks_Wide[b.MiSize];
        int x4 = b.MiCol, nCount = 0;
        const int nLimit = std::min(4, (int)Mi_Width_Log2[b.MiSize]);
        while (nCount < nLimit && x4 < std::min(f.MiCols, b.MiCol + w4)) {
            const int candRow = b.MiRow - 1, candCol = x4 | 1;
            const ModeInfoBlock& info = f.getModeInfo(candRow, candCol);
            const int step4 = clip3(2, 16, Num_4x4_Blocks_Wide[info.MiSize]);
            if (info.RefFrames[0] > INTRA_FRAME) {
                nCount++;
                strip(AV1B_IPU_OBMC_ABOVE, candRow, candCol, x4, b.MiRow, std::min(w, (step4 * MI_SIZE) >> subX), std::min(h >> 1, 32 >> subY));
            }
            x4 += step4;
        }
    }
    if (b.AvailL) {
        const int h4 = Num_4x4_Blocks_High[b.MiSize];
        int y4 = b.MiRow, nCount = 0;
        const int nLimit = std::min(4, (int)Mi_Height_Log2[b.MiSize]);
        while (nCount < nLimit && y4 < std::min(f.MiRows, b.MiRow + h4)) {
            const int candCol = b.MiCol - 1, candRow = y4 | 1;
            const ModeInfoBlock& info = f.getModeInfo(candRow, candCol);
            const int step4 = clip3(2, 16, Num_4x4_Blocks_High[info.MiSize]);
            if (info.RefFrames[0] > INTRA_FRAME) {
                nCount++;
                strip(AV1B_IPU_OBMC_LEFT, candRow, candCol, b.MiCol, y4, std::min(w >> 1, 32 >> subX), std::min(h, (step4 * MI_SIZE) >> subY));
            }
            y4 += step4;
        }
    }
}

// TransformBlock::decode (TransformBlock.cpp:2376-2456) and reconstruct()'s dequantisation (:2255-2276)
void FrameEmitter::emitTb(Block& b, TransformBlock& t)
{
    FrameHeader& f = *m_frame;
    const int plane = t.plane;
    const int subX = plane ? b.subsampling_x : 0, subY = plane ? b.subsampling_y : 0;
    const int x = t.x, y = t.y;
    const TX_SIZE txSz = t.txSz;
    const int row = (y << subY) >> MI_SIZE_LOG2, col = (x << subX) >> MI_SIZE_LOG2;
    const int subBlockMiRow = row & b.sbMask, subBlockMiCol = col & b.sbMask;
    const int stepX = Tx_Width[txSz] >> MI_SIZE_LOG2, stepY = Tx_Height[txSz] >> MI_SIZE_LOG2;
    Av1bOp op;
    memset(&op, 0, sizeof(op));
    op.x = (uint16_t)x;
    op.y = (uint16_t)y;
    op.plane = (uint8_t)plane;
    op.tx_size = (uint8_t)txSz;
    if (!b.is_inter) {
        if (b.m_palette.isPalettePredict(plane)) {
            op.kind = AV1B_OP_PALETTE;
            op.aux = auxFor(b);
        } else {
            op.kind = AV1B_OP_INTRA;
            const bool isCfl = plane > 0 && b.UVMode == UV_CFL_PRED;
            const int mode = plane == 0 ? (int)b.YMode : (isCfl ? (int)DC_PRED : (int)b.UVMode);
            op.mode = (uint8_t)mode;
            op.angle_delta = plane == 0 ? b.AngleDeltaY : b.AngleDeltaUV;
            uint8_t fl = 0;
            if ((plane == 0 ? b.AvailL : b.AvailLChroma) || x > t.m_baseX) fl |= AV1B_OPF_HAVE_LEFT;
            if ((plane == 0 ? b.AvailU : b.AvailUChroma) || y > t.m_baseY) fl |= AV1B_OPF_HAVE_ABOVE;
            if (b.m_decoded.getFlag(plane, (subBlockMiRow >> subY) - 1, (subBlockMiCol >> subX) + stepX)) fl |= AV1B_OPF_HAVE_ABOVE_RIGHT;
            if (b.m_decoded.getFlag(plane, (subBlockMiRow >> subY) + stepY, (subBlockMiCol >> subX) - 1)) fl |= AV1B_OPF_HAVE_BELOW_LEFT;
            if (plane == 0 && b.use_filter_intra) {
                fl |= AV1B_OPF_FILTER_INTRA;
                op.fi_mode = (uint8_t)b.filter_intra_mode;
            } else if (is_directional_mode(mode) && m_seq->enable_intra_edge_filter) {
                if (edgeSmooth(b, plane)) fl |= AV1B_OPF_EDGE_SMOOTH;
            }
            if (isCfl) {
                fl |= AV1B_OPF_CFL;
                op.cfl_alpha = plane == 1 ? b.CflAlphaU : b.CflAlphaV;
                op.max_luma_w = (uint16_t)m_maxLumaW;
                op.max_luma_h = (uint16_t)m_maxLumaH;
            }
            op.flags = fl;
        }
        if (plane == 0) {
            m_maxLumaW = x + stepX * 4;
            m_maxLumaH = y + stepY * 4;
        }
    } else {
        op.kind = AV1B_OP_INTER_RES;
    }
    if (t.m_eob > 0) {
        const int tw = t.tw, th = t.th;
        const int dcq = t.get_dc_quant(), acq = t.get_ac_quant();
        const int denom = t.dqDenom;
        const size_t off = m_coef.size();
        m_coef.resize(off + (size_t)tw * th); // zero-filled: only non-zero levels are written below
        int16_t* out = &m_coef[off];
        int nzr = 0, nzc = 0;
        const int dshift = denom == 1 ? 0 : (denom == 2 ? 1 : 2); // dqDenom is 1, 2 or 4: magnitude division by shift
        // only the first eob positions of the scan can hold a level (TransformBlock::coeffs,
        // TransformBlock.cpp:1678-1697): walk those, not the whole 4 KB Quant array of the block
        const auto* quant = &t.Quant[0];
        const int16_t* scan = t.get_scan();
        int ltw = 0;
        while ((1 << ltw) < tw) ltw++;
        for (int cidx = 0; cidx < t.m_eob; cidx++) {
            const int pos = scan[cidx];
            const int q = quant[pos];
            if (!q) continue;
            const int i = pos >> ltw, j = pos & (tw - 1);
            const int dq = (int)((unsigned)q * (unsigned)(pos ? acq : dcq));
            const int mag = (int)((((dq < 0) ? (0u - (unsigned)dq) : (unsigned)dq) & 0xffffffu) >> dshift);
            out[pos] = (int16_t)clip3(-32768, 32767, dq < 0 ? -mag : mag);
            if (i + 1 > nzr) nzr = i + 1;
            if (j + 1 > nzc) nzc = j + 1;
        }
        op.flags |= AV1B_OPF_HAS_RESID;
        op.tx_type = (uint8_t)t.PlaneTxType;
        op.lossless = b.Lossless;
        op.nz_rows = (uint8_t)nzr;
        op.nz_cols = (uint8_t)nzc;
        op.coef_off = (uint32_t)off;
        op.res_off = m_nRes;
        m_nRes += (uint32_t)(Tx_Width[txSz] * Tx_Height[txSz]);
        m_itx.push_back((uint32_t)m_ops.size());
    }
    // ops of the dependent pass: intra / palette TBs always; residual-only TBs just for
    // inter-intra and intrabc blocks (plain inter blocks are finished by the inter pass).
    // The inverse-transform list refers to ops by index, so coded TBs of plain inter blocks go
    // to a separate op array that only the inverse transform reads.
    const bool plainInter = b.is_inter && !b.use_intrabc && b.RefFrame[1] != INTRA_FRAME;
    if (plainInter) {
        if (op.flags & AV1B_OPF_HAS_RESID) {
            m_itx.back() = 0x80000000u | (uint32_t)m_itxOnly.size();
            m_itxOnly.push_back(op);
        }
    } else if (op.kind != AV1B_OP_INTER_RES || (op.flags & AV1B_OPF_HAS_RESID)) {
        m_ops.push_back(op);
        // a large intra block is the work of several warps: one op per strip of rows (same block,
        // same edges, Av1bOp::fi_mode says which rows); the inverse transform knows the first only
        const int area = Tx_Width[txSz] * Tx_Height[txSz];
        if (op.kind == AV1B_OP_INTRA && !(op.flags & AV1B_OPF_FILTER_INTRA) && getenv_split() > 0 && area >= std::max(512, getenv_split())) {
            int lg = 1;
            while (lg < 3 && (area >> (8 + lg)) > 1) lg++;
            m_ops.back().fi_mode = (uint8_t)(lg << 3);
            for (int i = 1; i < (1 << lg); i++) {
                op.fi_mode = (uint8_t)((lg | (i << 2)) << 3);
                m_ops.push_back(op);
            }
        }
    }
    // LoopfilterTxSizes + BlockDecoded bookkeeping (TransformBlock.cpp:2444-2454)
    const int miRows = f.MiRows, miCols = f.MiCols;
    uint8_t* lftx = &m_lftx[(size_t)plane * miRows * miCols];
    for (int i = 0; i < stepY; i++) {
        for (int j = 0; j < stepX; j++) {
            for (int xx = 0; xx < subX + 1; xx++)
                for (int yy = 0; yy < subY + 1; yy++) {
                    const int rr = row + (i << subY) + yy, cc = col + (j << subX) + xx;
                    if (rr < miRows && cc < miCols) lftx[(size_t)rr * miCols + cc] = (uint8_t)txSz;
                }
            b.m_decoded.setFlag(plane, (subBlockMiRow >> subY) + i, (subBlockMiCol >> subX) + j);
        }
    }
}

static int countUnits(int unitSize, int frameSize) { return std::max((frameSize + (unitSize >> 1)) / unitSize, 1); }

void FrameEmitter::finish()
{
    FrameHeader& f = *m_frame;
    Av1bFrameHdr& h = m_hdr;
    const int miRows = f.MiRows, miCols = f.MiCols;
    // ---- deblocking inputs (LoopFilter.cpp:40-58,301-359)
    const LoopFilterParams& lf = f.m_loopFilter;
    for (int i = 0; i < 4; i++) h.lf.level[i] = lf.loop_filter_level[i];
    if (m_seq->NumPlanes == 1 || !(lf.loop_filter_level[0] || lf.loop_filter_level[1])) h.lf.level[2] = h.lf.level[3] = 0;
    h.lf.sharpness = lf.loop_filter_sharpness;
    h.lf.delta_enabled = lf.loop_filter_delta_enabled;
    h.lf.delta_lf_multi = f.m_deltaLf.delta_lf_multi;
    for (int i = 0; i < 8; i++) h.lf.ref_deltas[i] = lf.loop_filter_ref_deltas[i];
    for (int i = 0; i < 2; i++) h.lf.mode_deltas[i] = lf.loop_filter_mode_deltas[i];
    m_lfmi.resize((size_t)miRows * miCols);
    const uint8_t* txY = &m_lftx[0];
    const uint8_t* txU = &m_lftx[(size_t)miRows * miCols];
    const uint8_t* txV = &m_lftx[(size_t)2 * miRows * miCols];
    for (int r = 0; r < miRows; r++) {
        for (int c = 0; c < miCols; c++) {
            const ModeInfoBlock& info = f.getModeInfo(r, c);
            Av1bLfMi& m = m_lfmi[(size_t)r * miCols + c];
            const size_t k = (size_t)r * miCols + c;
            m.mi_size = (uint8_t)info.MiSize;
            const int mode = info.YMode;
            const int modeType = mode >= NEARESTMV && mode != GLOBALMV && mode != GLOBAL_GLOBALMV;
            const int ref = std::max(0, (int)info.RefFrames[0]);
            m.flags = (uint8_t)((info.Skip ? 1 : 0) | (modeType << 1) | ((ref & 7) << 2));
            m.tx = (uint16_t)(txY[k] | (txU[k] << 5) | (txV[k] << 10));
            for (int i = 0; i < 4; i++) m.delta_lf[i] = info.DeltaLFs[i];
        }
    }
    // ---- CDEF inputs (Cdef.cpp:41-101)
    const CdefParams& cd = f.m_cdef;
    const int r8n = miRows >> 1, c8n = miCols >> 1;
    m_cdef8.assign((size_t)r8n * c8n, 0xFF);
    bool anyCdef = false;
    const bool cdefOff = f.CodedLossless || f.allow_intrabc || !m_seq->enable_cdef;
    if (!cdefOff) {
        for (int r8 = 0; r8 < r8n; r8++) {
            for (int c8 = 0; c8 < c8n; c8++) {
                const int r = r8 * 2, c = c8 * 2;
                const int idx = cd.cdef_idx[r & ~15][c & ~15];
                if (idx == -1) continue;
                const bool skip = f.getModeInfo(r, c).Skip && f.getModeInfo(r + 1, c).Skip && f.getModeInfo(r, c + 1).Skip
                    && f.getModeInfo(r + 1, c + 1).Skip;
                if (skip) continue;
                m_cdef8[(size_t)r8 * c8n + c8] = (uint8_t)idx;
                anyCdef = true;
            }
        }
    }
    h.cdef.enabled = anyCdef;
    h.cdef.damping = cd.CdefDamping;
    for (int i = 0; i < 8; i++) {
        h.cdef.y_pri[i] = cd.cdef_y_pri_strength[i];
        h.cdef.y_sec[i] = cd.cdef_y_sec_strength[i];
        h.cdef.uv_pri[i] = cd.cdef_uv_pri_strength[i];
        h.cdef.uv_sec[i] = cd.cdef_uv_sec_strength[i];
    }
    // ---- loop-restoration inputs (LoopRestoration.cpp:49-134,191-219)
    const LoopRestorationpParams& lr = f.m_loopRestoration;
    h.lr.uses_lr = lr.UsesLr;
    if (lr.UsesLr) {
        for (int p = 0; p < m_seq->NumPlanes && p < 3; p++) {
            h.lr.frame_type[p] = (uint8_t)lr.FrameRestorationType[p];
            if (lr.FrameRestorationType[p] == RESTORE_NONE) continue;
            const int sub = p ? 1 : 0;
            const int us = lr.LoopRestorationSize[p];
            const int rows = countUnits(us, (f.FrameHeight + sub) >> sub), cols = countUnits(us, (f.UpscaledWidth + sub) >> sub);
            h.lr.unit_size[p] = (uint16_t)us;
            h.lr.unit_rows[p] = (uint16_t)rows;
            h.lr.unit_cols[p] = (uint16_t)cols;
            h.lr.unit_first[p] = (uint32_t)m_lru.size();
            for (int r = 0; r < rows; r++) {
                for (int c = 0; c < cols; c++) {
                    Av1bLrUnit u;
                    memset(&u, 0, sizeof(u));
                    u.type = (uint8_t)lr.LrType[p][r][c];
                    if (u.type == RESTORE_WIENER) {
                        for (int pass = 0; pass < 2; pass++)
                            for (int k = 0; k < 3; k++) u.wiener[pass][k] = lr.LrWiener[p][r][c][pass][k];
                    } else if (u.type == RESTORE_SGRPROJ) {
                        u.sgr_set = lr.LrSgrSet[p][r][c];
                        u.sgr_xqd[0] = lr.LrSgrXqd[p][r][c][0];
                        u.sgr_xqd[1] = lr.LrSgrXqd[p][r][c][1];
                    }
                    m_lru.push_back(u);
                }
            }
        }
    }
    layout();
}

static size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

void FrameEmitter::layout()
{
    Av1bFrameHdr& h = m_hdr;
    size_t off = align16(sizeof(Av1bFrameHdr));
    auto place = [&](uint32_t& o, size_t bytes) {
        o = (uint32_t)off;
        off = align16(off + bytes);
    };
    h.n_sb = (uint32_t)m_sbs.size();
    place(h.off_sb, m_sbs.size() * sizeof(Av1bSb));
    // [ordered ops of the dependent pass][inverse-transform-only ops of plain inter blocks]
    h.n_ops = (uint32_t)m_ops.size();
    place(h.off_ops, (m_ops.size() + m_itxOnly.size()) * sizeof(Av1bOp));
    for (auto& i : m_itx)
        if (i & 0x80000000u) i = (i & 0x7FFFFFFFu) + (uint32_t)m_ops.size();
    {
        // stable counting sort of the inverse-transform list by size class (max(w,h) = 4 | 8 | 16 | >= 32):
        // the device packs 32 / max(w,h) transform blocks into one warp
        auto opAt = [&](uint32_t i) -> const Av1bOp& { return i < m_ops.size() ? m_ops[i] : m_itxOnly[i - m_ops.size()]; };
        auto cls = [&](uint32_t i) {
            const Av1bOp& o = opAt(i);
            const int md = std::max(hk_tx_w[o.tx_size], hk_tx_h[o.tx_size]);
            return md <= 4 ? 0 : (md <= 8 ? 1 : (md <= 16 ? 2 : 3));
        };
        uint32_t cnt[5] = { 0, 0, 0, 0, 0 };
        for (uint32_t i : m_itx) cnt[cls(i) + 1]++;
        for (int k = 1; k < 5; k++) cnt[k] += cnt[k - 1];
        for (int k = 0; k < 4; k++) h.itx_class_end[k] = cnt[k + 1];
        m_perm.resize(m_itx.size());
        for (uint32_t i : m_itx) m_perm[cnt[cls(i)]++] = i;
        m_itx.swap(m_perm);
    }
    h.n_itx = (uint32_t)m_itx.size();
    place(h.off_itx, m_itx.size() * sizeof(uint32_t));
    h.n_iblk = (uint32_t)m_iblk.size();
    place(h.off_iblk, m_iblk.size() * sizeof(Av1bInterBlk));
    h.n_ipu = (uint32_t)m_ipu.size();
    place(h.off_ipu, m_ipu.size() * sizeof(Av1bIpu));
    h.n_aux = (uint32_t)m_aux.size();
    place(h.off_aux, m_aux.size() * sizeof(Av1bBlkAux));
    h.n_coef = (uint32_t)m_coef.size();
    place(h.off_coef, m_coef.size() * sizeof(int16_t));
    h.n_res = m_nRes;
    h.n_pal = (uint32_t)m_pal.size();
    place(h.off_pal, m_pal.size());
    place(h.off_lfmi, m_lfmi.size() * sizeof(Av1bLfMi));
    place(h.off_cdef8, m_cdef8.size());
    h.n_lru = (uint32_t)m_lru.size();
    place(h.off_lru, m_lru.size() * sizeof(Av1bLrUnit));
    h.total_bytes = (uint32_t)off;
    m_total = off;
}

void FrameEmitter::write(uint8_t* dst) const
{
    const Av1bFrameHdr& h = m_hdr;
    memcpy(dst, &h, sizeof(h));
    auto put = [&](uint32_t o, const void* p, size_t bytes) {
        if (bytes) memcpy(dst + o, p, bytes);
    };
    put(h.off_sb, m_sbs.data(), m_sbs.size() * sizeof(Av1bSb));
    put(h.off_ops, m_ops.data(), m_ops.size() * sizeof(Av1bOp));
    put(h.off_ops + (uint32_t)(m_ops.size() * sizeof(Av1bOp)), m_itxOnly.data(), m_itxOnly.size() * sizeof(Av1bOp));
    put(h.off_itx, m_itx.data(), m_itx.size() * sizeof(uint32_t));
    put(h.off_iblk, m_iblk.data(), m_iblk.size() * sizeof(Av1bInterBlk));
    put(h.off_ipu, m_ipu.data(), m_ipu.size() * sizeof(Av1bIpu));
    put(h.off_aux, m_aux.data(), m_aux.size() * sizeof(Av1bBlkAux));
    put(h.off_coef, m_coef.data(), m_coef.size() * sizeof(int16_t));
    put(h.off_pal, m_pal.data(), m_pal.size());
    put(h.off_lfmi, m_lfmi.data(), m_lfmi.size() * sizeof(Av1bLfMi));
    put(h.off_cdef8, m_cdef8.data(), m_cdef8.size());
    put(h.off_lru, m_lru.data(), m_lru.size() * sizeof(Av1bLrUnit));
}

}  // namespace av1b200

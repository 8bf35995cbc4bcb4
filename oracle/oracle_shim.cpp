// oracle_shim.cpp -- TEST INFRASTRUCTURE ONLY.  A C-ABI shim over the UNMODIFIED reference decoder
// (compiled from /root/reference by oracle/Makefile into oracle/_ref/libav1ref.a).  It is the
// checker for the parity tests and the CPU-baseline arm of bench.py; nothing in the product
// (av1dec_b200/) links, loads or calls it.
//
// Parity is PINNED: the library built here reproduces all 172 MD5s of bits/bits.md5
// (tests/test_oracle.py), and the stage-level entry points call the reference's own C++
// functions:
//   oracle_decode_ivf        YamiAv1::Decoder::decode/getOutput        decoder/Av1Decoder.cpp:49,203
//   oracle_decode_stages     decodeFrame + decode_frame_wrapup, frame captured after each stage
//                                                                      decoder/Av1Decoder.cpp:128-192
//   oracle_postfilter        LoopFilter::filter / Cdef::filter / LoopRestoration::filter on a
//                            hand-filled FrameHeader                   LoopFilter.cpp:40 Cdef.cpp:41 LoopRestoration.cpp:191
//   oracle_inverse_transform TransformBlock::inverseTransform          decoder/TransformBlock.cpp:2173
//   oracle_predict_intra     Block::IntraPredict::predict_intra / predict_chroma_from_luma on the
//                            ops of a synthetic command buffer         decoder/IntraPredict.cpp:563-667
//   oracle_predict_inter     Block::InterPredict::predict_inter on the prediction units of a
//                            synthetic command buffer (translational) decoder/InterPredict.cpp:962-1049
#include <algorithm>
#include <deque>
#include <functional>
#include <limits>
#include <list>
#include <memory>
#include <numeric>
#include <string>
#include <vector>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define private public
#define protected public
// Block's nested helper classes (IntraPredict, InterPredict, ...) are declared under the class's
// DEFAULT access, which no keyword redefinition reaches: Block.h alone is read with `class` spelt
// `struct` (its own includes come first, so nothing else is affected; same mangled names).
#include "../aom/enums.h"
#include "BitReader.h"
#include "BlockTree.h"
#include "EntropyDecoder.h"
#include "Tile.h"
#define class struct
#include "Block.h"
#undef class
#include "Av1Decoder.h"
#include "BitReader.h"
#include "Block.h"
#include "Cdef.h"
#include "EntropyDecoder.h"
#include "IntraPredict.h"
#include "InterPredict.h"
#include "LoopFilter.h"
#include "LoopRestoration.h"
#include "Parser.h"
#include "Partition.h"
#include "SuperBlock.h"
#include "Tile.h"
#include "TransformBlock.h"
#include "VideoFrame.h"
#undef private
#undef protected

#include "../include/av1b200_format.h"

using namespace YamiAv1;
using namespace Yami;

namespace {

uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

struct Sink {
    uint8_t* out;
    size_t cap, pos;
    bool overflow;
    void planes(const YuvFrame& f, int w, int h)
    {
        for (int p = 0; p < 3; p++) {
            const int pw = p ? (w >> 1) : w, ph = p ? (h >> 1) : h;
            for (int y = 0; y < ph; y++) {
                if (out && pos + pw <= cap) memcpy(out + pos, f.data[p] + (size_t)y * f.strides[p], pw);
                else if (out) overflow = true;
                pos += pw;
            }
        }
    }
};

// The reference's decodeFrame()/decode_frame_wrapup() sequence with a capture hook after each
// stage (the functions are private and not hookable, so the ~25 lines of glue are replayed here;
// every pixel is produced by the reference's own Tile::decode / filter classes).
bool decodeFrameStages(Decoder& d, TileGroup& tiles, int stage, Sink& sink)
{
    FrameHeader& h = *d.m_frame;
    std::shared_ptr<YuvFrame> frame = YuvFrame::create(h.FrameWidth, h.FrameHeight);
    for (auto& t : tiles)
        if (!t->decode(frame, d.m_store)) return false;
    d.frame_end_update_cdf(tiles);
    const int aw = h.MiCols * 4, ah = h.MiRows * 4;
    if (stage == 0) sink.planes(*frame, aw, ah);
    LoopFilter lf(d.m_frame);
    lf.filter(frame);
    if (stage == 1) sink.planes(*frame, aw, ah);
    Cdef cdef(d.m_frame);
    std::shared_ptr<YuvFrame> cdefFrame = cdef.filter(frame);
    if (stage == 2) sink.planes(*cdefFrame, h.FrameWidth, h.FrameHeight);
    LoopRestoration lr(d.m_frame, cdefFrame, frame);
    std::shared_ptr<YuvFrame> out = lr.filter();
    if (stage == 3) sink.planes(*out, h.FrameWidth, h.FrameHeight);
    h.motionVectorStorage();
    if (h.show_frame) d.m_output.push_back(out);
    d.updateFrameStore(h, out);
    d.m_parser->finishFrame();
    return true;
}

bool decodeUnitStages(Decoder& d, uint8_t* data, size_t size, int stage, Sink& sink, int& frames)
{
    BitReader reader(data, size);
    while (reader.getRemainingBitsCount() > 0) {
        obu_header hdr;
        if (!hdr.parse(reader)) return false;
        const uint64_t sz = hdr.obu_size;
        BitReader br(data + (reader.getPos() >> 3), sz);
        bool ok = true;
        if (hdr.obu_type == OBU_SEQUENCE_HEADER) ok = d.m_parser->parseSequenceHeader(br);
        else if (hdr.obu_type == OBU_TD) ok = d.m_parser->parseTemporalDelimiter(br);
        else if (hdr.obu_type == OBU_FRAME_HEADER) {
            d.m_frame = d.m_parser->parseFrameHeader(br);
            ok = bool(d.m_frame);
            if (ok && d.m_frame->show_existing_frame) d.showExistingFrame();
        } else if (hdr.obu_type == OBU_FRAME) {
            TileGroup group;
            d.m_frame = d.m_parser->parseFrame(br, group);
            ok = d.m_frame && decodeFrameStages(d, group, stage, sink);
            if (ok) frames++;
        } else if (hdr.obu_type == OBU_TILE_GROUP) {
            TileGroup group;
            ok = d.m_frame && d.m_parser->parseTileGroup(br, d.m_frame, group);
            if (ok) {
                d.m_tiles.insert(d.m_tiles.end(), group.begin(), group.end());
                if (d.m_tiles.size() == d.m_parser->m_frame->NumTiles) {
                    ok = decodeFrameStages(d, d.m_tiles, stage, sink);
                    d.m_tiles.clear();
                    if (ok) frames++;
                }
            }
        }
        if (!ok) return false;
        reader.skip(sz << 3);
    }
    return true;
}

}  // namespace

extern "C" {

// Whole-stream decode through the unmodified reference Decoder; output layout = the CLI's .yuv.
int oracle_decode_ivf(const uint8_t* ivf, size_t len, uint8_t* out, size_t cap, size_t* out_bytes, int* n_frames,
    uint64_t* luma_pixels)
{
    if (!ivf || len < 32 || memcmp(ivf, "DKIF", 4) != 0) return -1;
    size_t pos = ivf[6] | (ivf[7] << 8);
    Sink sink{ out, cap, 0, false };
    int frames = 0;
    uint64_t pixels = 0;
    Decoder dec;
    while (pos + 12 <= len) {
        const uint32_t sz = rd32(ivf + pos);
        pos += 12;
        if (pos + sz > len) break;
        dec.decode(const_cast<uint8_t*>(ivf + pos), sz);
        pos += sz;
        std::shared_ptr<YuvFrame> f;
        while ((f = dec.getOutput())) {
            frames++;
            pixels += (uint64_t)f->width * f->height;
            sink.planes(*f, f->width, f->height);
        }
    }
    if (out_bytes) *out_bytes = sink.pos;
    if (n_frames) *n_frames = frames;
    if (luma_pixels) *luma_pixels = pixels;
    return sink.overflow ? -2 : 0;
}

// Every decoded frame (shown or not) captured after `stage`:
//   0 = reconstruction (before filters), 1 = deblocked      -> MI-aligned area (MiCols*4 x MiRows*4)
//   2 = CDEF output,                      3 = final (LR)     -> visible area
int oracle_decode_stages(const uint8_t* ivf, size_t len, int stage, uint8_t* out, size_t cap, size_t* out_bytes, int* n_frames)
{
    if (!ivf || len < 32 || memcmp(ivf, "DKIF", 4) != 0) return -1;
    size_t pos = ivf[6] | (ivf[7] << 8);
    Sink sink{ out, cap, 0, false };
    int frames = 0;
    Decoder dec;
    while (pos + 12 <= len) {
        const uint32_t sz = rd32(ivf + pos);
        pos += 12;
        if (pos + sz > len) break;
        if (!decodeUnitStages(dec, const_cast<uint8_t*>(ivf + pos), sz, stage, sink, frames)) return -1;
        pos += sz;
        while (dec.getOutput()) {}
    }
    if (out_bytes) *out_bytes = sink.pos;
    if (n_frames) *n_frames = frames;
    return sink.overflow ? -2 : 0;
}

// The reference's in-loop filters on synthetic input.  The frame parameters come in the same
// flat layout the engine consumes (an Av1bFrameHdr with lf/cdef/lr params, Av1bLfMi[] and
// Av1bLrUnit[] sections); cdef_idx64 is CdefParams::cdef_idx at 64x64 granularity
// ((mi_rows+15)/16 x (mi_cols+15)/16, -1 = off).  `stages`: bit0 deblock, bit1 CDEF, bit2 LR.
// in/out planes cover the MI-aligned area.  Returns 0.
int oracle_postfilter(const uint8_t* cmd, const int8_t* cdef_idx64, int sb128, uint32_t stages, const uint8_t* const in[3],
    const int in_stride[3], uint8_t* const out[3], const int out_stride[3])
{
    const Av1bFrameHdr& hd = *(const Av1bFrameHdr*)cmd;
    auto seq = std::make_shared<SequenceHeader>();
    seq->BitDepth = 8;
    seq->subsampling_x = seq->subsampling_y = 1;
    seq->NumPlanes = 3;
    seq->mono_chrome = false;
    seq->use_128x128_superblock = sb128 != 0;
    seq->enable_cdef = true;
    seq->enable_restoration = true;
    ConstSequencePtr cseq = seq;
    auto fh = std::make_shared<FrameHeader>(cseq);
    FrameHeader& f = *fh;
    f.FrameWidth = hd.frame_w;
    f.FrameHeight = hd.frame_h;
    f.UpscaledWidth = hd.frame_w;
    f.compute_image_size();
    f.initGeometry();
    memset(&f.m_segmentation, 0, sizeof(f.m_segmentation));
    memset(&f.m_deltaLf, 0, sizeof(f.m_deltaLf));
    f.m_deltaLf.delta_lf_multi = hd.lf.delta_lf_multi;
    LoopFilterParams& lf = f.m_loopFilter;
    memset(&lf, 0, sizeof(lf));
    for (int i = 0; i < 4; i++) lf.loop_filter_level[i] = hd.lf.level[i];
    lf.loop_filter_sharpness = hd.lf.sharpness;
    lf.loop_filter_delta_enabled = hd.lf.delta_enabled;
    for (int i = 0; i < 8; i++) lf.loop_filter_ref_deltas[i] = hd.lf.ref_deltas[i];
    for (int i = 0; i < 2; i++) lf.loop_filter_mode_deltas[i] = hd.lf.mode_deltas[i];
    const Av1bLfMi* mis = (const Av1bLfMi*)(cmd + hd.off_lfmi);
    for (int r = 0; r < f.MiRows; r++)
        for (int c = 0; c < f.MiCols; c++) {
            const Av1bLfMi& m = mis[(size_t)r * f.MiCols + c];
            ModeInfoBlock& info = f.m_modeInfo[r][c];
            info.MiSize = (BLOCK_SIZE)m.mi_size;
            info.Skip = m.flags & 1;
            const int ref = (m.flags >> 2) & 7;
            info.RefFrames[0] = ref;
            info.RefFrames[1] = NONE_FRAME;
            info.YMode = ref == 0 ? DC_PRED : (((m.flags >> 1) & 1) ? NEWMV : GLOBALMV);
            for (int p = 0; p < 3; p++) info.LoopfilterTxSizes[p] = (TX_SIZE)((m.tx >> (5 * p)) & 31);
            for (int i = 0; i < 4; i++) info.DeltaLFs[i] = m.delta_lf[i];
        }
    CdefParams& cd = f.m_cdef;
    cd.CdefDamping = hd.cdef.damping;
    for (int i = 0; i < 8; i++) {
        cd.cdef_y_pri_strength[i] = hd.cdef.y_pri[i];
        cd.cdef_y_sec_strength[i] = hd.cdef.y_sec[i];
        cd.cdef_uv_pri_strength[i] = hd.cdef.uv_pri[i];
        cd.cdef_uv_sec_strength[i] = hd.cdef.uv_sec[i];
    }
    cd.cdef_idx.assign(f.MiRows, std::vector<int>(f.MiCols, -1));
    const int c64 = (f.MiCols + 15) / 16;
    for (int r = 0; r < f.MiRows; r += 16)
        for (int c = 0; c < f.MiCols; c += 16) cd.cdef_idx[r][c] = cdef_idx64 ? cdef_idx64[(r / 16) * c64 + c / 16] : -1;
    LoopRestorationpParams& lr = f.m_loopRestoration;
    lr.UsesLr = hd.lr.uses_lr;
    lr.LrType.resize(3);
    lr.LrWiener.resize(3);
    lr.LrSgrSet.resize(3);
    lr.LrSgrXqd.resize(3);
    const Av1bLrUnit* units = (const Av1bLrUnit*)(cmd + hd.off_lru);
    for (int p = 0; p < 3; p++) {
        lr.FrameRestorationType[p] = (RestorationType)hd.lr.frame_type[p];
        lr.LoopRestorationSize[p] = hd.lr.unit_size[p];
        if (!hd.lr.uses_lr || !hd.lr.frame_type[p]) continue;
        const int rows = hd.lr.unit_rows[p], cols = hd.lr.unit_cols[p];
        lr.LrType[p].assign(rows, std::vector<RestorationType>(cols));
        lr.LrWiener[p].assign(rows, std::vector<std::vector<std::vector<int8_t>>>(cols, std::vector<std::vector<int8_t>>(2, std::vector<int8_t>(3))));
        lr.LrSgrSet[p].assign(rows, std::vector<uint8_t>(cols));
        lr.LrSgrXqd[p].assign(rows, std::vector<std::vector<int8_t>>(cols, std::vector<int8_t>(2)));
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) {
                const Av1bLrUnit& u = units[hd.lr.unit_first[p] + r * cols + c];
                lr.LrType[p][r][c] = (RestorationType)u.type;
                for (int pass = 0; pass < 2; pass++)
                    for (int k = 0; k < 3; k++) lr.LrWiener[p][r][c][pass][k] = u.wiener[pass][k];
                lr.LrSgrSet[p][r][c] = u.sgr_set;
                lr.LrSgrXqd[p][r][c][0] = u.sgr_xqd[0];
                lr.LrSgrXqd[p][r][c][1] = u.sgr_xqd[1];
            }
    }
    const int aw = f.MiCols * 4, ah = f.MiRows * 4;
    std::shared_ptr<YuvFrame> frame = YuvFrame::create(f.FrameWidth, f.FrameHeight);
    for (int p = 0; p < 3; p++) {
        const int pw = p ? aw / 2 : aw, ph = p ? ah / 2 : ah;
        for (int y = 0; y < ph; y++) memcpy(frame->data[p] + (size_t)y * frame->strides[p], in[p] + (size_t)y * in_stride[p], pw);
    }
    ConstFramePtr cf = fh;
    std::shared_ptr<YuvFrame> result = frame;
    if (stages & 1) {
        LoopFilter filter(cf);
        filter.filter(frame);
    }
    std::shared_ptr<YuvFrame> cdefFrame = frame;
    bool visibleOnly = false;
    if (stages & 2) {
        Cdef cdef(cf);
        cdefFrame = cdef.filter(frame);
        result = cdefFrame;
        visibleOnly = true;
    }
    if (stages & 4) {
        LoopRestoration rest(cf, cdefFrame, frame);
        result = rest.filter();
        visibleOnly = visibleOnly || lr.UsesLr;
    }
    // Only the visible area of a CDEF/LR output is defined by the reference (they are
    // visible-region copies); a deblock-only result is defined on the MI-aligned area.
    const int ow = visibleOnly ? (int)f.FrameWidth : aw, oh = visibleOnly ? (int)f.FrameHeight : ah;
    for (int p = 0; p < 3; p++) {
        const int pw = p ? ow / 2 : ow, ph = p ? oh / 2 : oh;
        for (int y = 0; y < ph; y++) memcpy(out[p] + (size_t)y * out_stride[p], result->data[p] + (size_t)y * result->strides[p], pw);
    }
    return 0;
}

// Block::IntraPredict on every AV1B_OP_INTRA op of a (synthetic) command buffer, in list order, the
// way TransformBlock::decode() drives it (TransformBlock.cpp:2392-2426): prediction, chroma-from-
// luma, then the block is written into the frame (no residual).  The availability flags are the
// op's; the block-level state predict_intra reads (filter-intra, angle deltas, CfL alphas and
// MaxLumaW/H, the smooth-neighbour test behind get_filter_type) is set from the op.  `planes` is the
// frame (MI-aligned area), updated in place.  Returns the number of ops run, < 0 on error.
int oracle_predict_intra(const uint8_t* cmd, uint8_t* const planes[3], const int strides[3])
{
    const Av1bFrameHdr& hd = *(const Av1bFrameHdr*)cmd;
    auto seq = std::make_shared<SequenceHeader>();
    seq->BitDepth = 8;
    seq->subsampling_x = seq->subsampling_y = 1;
    seq->NumPlanes = 3;
    seq->mono_chrome = false;
    seq->use_128x128_superblock = hd.sb_log2 == 7;
    seq->enable_intra_edge_filter = hd.enable_intra_edge_filter != 0;
    ConstSequencePtr cseq = seq;
    auto fh = std::make_shared<FrameHeader>(cseq);
    FrameHeader& f = *fh;
    f.FrameWidth = hd.frame_w;
    f.FrameHeight = hd.frame_h;
    f.UpscaledWidth = hd.frame_w;
    f.compute_image_size();
    f.initGeometry();
    if ((int)f.MiCols != hd.mi_cols || (int)f.MiRows != hd.mi_rows) return -1;
    f.TileCols = f.TileRows = 1;
    f.MiColStarts.assign({ 0, (int)f.MiCols });
    f.MiRowStarts.assign({ 0, (int)f.MiRows });
    memset(&f.m_quant, 0, sizeof(f.m_quant));
    Tile tile(cseq, fh, 0);
    static const uint8_t no_bits[32] = { 0 };
    tile.m_entropy.reset(new EntropyDecoder(no_bits, sizeof(no_bits), true, tile.m_cdfs));
    const int aw = f.MiCols * 4, ah = f.MiRows * 4;
    std::shared_ptr<YuvFrame> frame = YuvFrame::create(f.FrameWidth, f.FrameHeight);
    for (int p = 0; p < 3; p++) {
        const int pw = p ? aw / 2 : aw, ph = p ? ah / 2 : ah;
        for (int y = 0; y < ph; y++) memcpy(frame->data[p] + (size_t)y * frame->strides[p], planes[p] + (size_t)y * strides[p], pw);
    }
    static const uint8_t wlog2[19] = { 2, 3, 4, 5, 6, 2, 3, 3, 4, 4, 5, 5, 6, 2, 4, 3, 5, 4, 6 };
    static const uint8_t hlog2[19] = { 2, 3, 4, 5, 6, 3, 2, 4, 3, 5, 4, 6, 5, 4, 2, 5, 3, 6, 4 };
    const Av1bOp* ops = (const Av1bOp*)(cmd + hd.off_ops);
    int n = 0;
    for (uint32_t k = 0; k < hd.n_ops; k++) {
        const Av1bOp& op = ops[k];
        if (op.kind != AV1B_OP_INTRA) continue;
        const int plane = op.plane, sub = plane ? 1 : 0;
        const int mi_row = (op.y << sub) >> 2, mi_col = (op.x << sub) >> 2;
        Block b(tile, mi_row, mi_col, BLOCK_4X4);
        b.use_filter_intra = (op.flags & AV1B_OPF_FILTER_INTRA) != 0;
        b.filter_intra_mode = (FILTER_INTRA_MODE)op.fi_mode;
        b.AngleDeltaY = b.AngleDeltaUV = op.angle_delta;
        b.CflAlphaU = b.CflAlphaV = op.cfl_alpha;
        b.MaxLumaW = op.max_luma_w;
        b.MaxLumaH = op.max_luma_h;
        // get_filter_type(): a smooth neighbour above (or, on the first MI row, to the left)
        const bool smooth = (op.flags & AV1B_OPF_EDGE_SMOOTH) != 0;
        b.AvailU = b.AvailUChroma = b.AvailL = b.AvailLChroma = false;
        if (mi_row > 0 || (plane && mi_row > 1)) {
            int r = mi_row - 1, c = mi_col;
            if (plane) {
                if (!(mi_col & 1)) c++;
                if (mi_row & 1) r--;
            }
            if (r >= 0 && c < (int)f.MiCols) {
                ModeInfoBlock& info = f.m_modeInfo[r][c];
                info.YMode = smooth ? SMOOTH_PRED : DC_PRED;
                info.UVMode = smooth ? UV_SMOOTH_PRED : UV_DC_PRED;
                info.RefFrames[0] = INTRA_FRAME;
                b.AvailU = b.AvailUChroma = true;
            }
        }
        if (!b.AvailU && mi_col > 0) {
            int r = mi_row, c = mi_col - 1;
            if (plane) {
                if (mi_col & 1) c--;
                if (!(mi_row & 1)) r++;
            }
            if (c >= 0 && r < (int)f.MiRows) {
                ModeInfoBlock& info = f.m_modeInfo[r][c];
                info.YMode = smooth ? SMOOTH_PRED : DC_PRED;
                info.UVMode = smooth ? UV_SMOOTH_PRED : UV_DC_PRED;
                info.RefFrames[0] = INTRA_FRAME;
                b.AvailL = b.AvailLChroma = true;
            }
        }
        const int lw = wlog2[op.tx_size], lh = hlog2[op.tx_size];
        std::vector<std::vector<uint8_t>> pred;
        Block::IntraPredict ip(b, frame, plane, op.x, op.y, lw, lh, pred);
        ip.predict_intra((op.flags & AV1B_OPF_HAVE_LEFT) != 0, (op.flags & AV1B_OPF_HAVE_ABOVE) != 0, (op.flags & AV1B_OPF_HAVE_ABOVE_RIGHT) != 0,
            (op.flags & AV1B_OPF_HAVE_BELOW_LEFT) != 0, op.mode);
        if (op.flags & AV1B_OPF_CFL) ip.predict_chroma_from_luma((TX_SIZE)op.tx_size);
        for (int i = 0; i < (1 << lh); i++)
            for (int j = 0; j < (1 << lw); j++) frame->setPixel(plane, op.x + j, op.y + i, pred[i][j]);
        n++;
    }
    for (int p = 0; p < 3; p++) {
        const int pw = p ? aw / 2 : aw, ph = p ? ah / 2 : ah;
        for (int y = 0; y < ph; y++) memcpy(planes[p] + (size_t)y * strides[p], frame->data[p] + (size_t)y * frame->strides[p], pw);
    }
    return n;
}

// Block::InterPredict::predict_inter for every AV1B_IPU_PRED unit of a (synthetic) command buffer:
// translational prediction, single reference or compound average / distance, from the reference
// pictures `refs[slot][plane]` (MI-aligned, one stride set for all).  ref_frame r reads store slot
// hd.ref_slot[r].  `out` receives the predicted picture (only samples covered by a unit change).
// Returns the number of units run, < 0 on error.
int oracle_predict_inter(const uint8_t* cmd, int n_refs, const uint8_t* const* refs, const int ref_strides[3], uint8_t* const out[3],
    const int out_strides[3])
{
    const Av1bFrameHdr& hd = *(const Av1bFrameHdr*)cmd;
    auto seq = std::make_shared<SequenceHeader>();
    seq->BitDepth = 8;
    seq->subsampling_x = seq->subsampling_y = 1;
    seq->NumPlanes = 3;
    seq->mono_chrome = false;
    seq->use_128x128_superblock = hd.sb_log2 == 7;
    ConstSequencePtr cseq = seq;
    auto fh = std::make_shared<FrameHeader>(cseq);
    FrameHeader& f = *fh;
    f.FrameWidth = hd.frame_w;
    f.FrameHeight = hd.frame_h;
    f.UpscaledWidth = hd.frame_w;
    f.compute_image_size();
    f.initGeometry();
    if ((int)f.MiCols != hd.mi_cols || (int)f.MiRows != hd.mi_rows) return -1;
    f.TileCols = f.TileRows = 1;
    f.MiColStarts.assign({ 0, (int)f.MiCols });
    f.MiRowStarts.assign({ 0, (int)f.MiRows });
    memset(&f.m_quant, 0, sizeof(f.m_quant));
    f.force_integer_mv = false;
    const int aw = f.MiCols * 4, ah = f.MiRows * 4;
    FrameStore store(NUM_REF_FRAMES);
    for (int sl = 0; sl < n_refs && sl < NUM_REF_FRAMES; sl++) {
        store[sl] = YuvFrame::create(f.FrameWidth, f.FrameHeight);
        for (int p = 0; p < 3; p++) {
            const int pw = p ? aw / 2 : aw, ph = p ? ah / 2 : ah;
            for (int y = 0; y < ph; y++) memcpy(store[sl]->data[p] + (size_t)y * store[sl]->strides[p], refs[sl * 3 + p] + (size_t)y * ref_strides[p], pw);
        }
        RefFrame& r = f.m_refInfo.m_refs[sl];
        r.RefValid = true;
        r.RefFrameWidth = r.RefUpscaledWidth = hd.frame_w;
        r.RefFrameHeight = hd.frame_h;
    }
    for (int r = LAST_FRAME; r <= ALTREF_FRAME; r++) {
        f.ref_frame_idx[r - LAST_FRAME] = hd.ref_slot[r] >= 0 ? hd.ref_slot[r] : 0;
        f.GmType[r] = IDENTITY;
    }
    Tile tile(cseq, fh, 0);
    static const uint8_t no_bits[32] = { 0 };
    tile.m_entropy.reset(new EntropyDecoder(no_bits, sizeof(no_bits), true, tile.m_cdfs));
    std::shared_ptr<YuvFrame> frame = YuvFrame::create(f.FrameWidth, f.FrameHeight);
    for (int p = 0; p < 3; p++) {
        const int pw = p ? aw / 2 : aw, ph = p ? ah / 2 : ah;
        for (int y = 0; y < ph; y++) memcpy(frame->data[p] + (size_t)y * frame->strides[p], out[p] + (size_t)y * out_strides[p], pw);
    }
    const Av1bIpu* units = (const Av1bIpu*)(cmd + hd.off_ipu);
    std::vector<std::vector<uint8_t>> mask;
    int n = 0;
    for (uint32_t k = 0; k < hd.n_ipu; k++) {
        const Av1bIpu& u = units[k];
        if (u.kind != AV1B_IPU_PRED || u.warp[0] || u.warp[1]) continue;
        const int plane = u.plane, sub = plane ? 1 : 0;
        const bool compound = (u.flags & AV1B_IPUF_COMPOUND) != 0;
        if (compound && u.comp_type != AV1B_COMP_AVERAGE) continue;
        const int mi_row = (u.y << sub) >> 2, mi_col = (u.x << sub) >> 2;
        Block b(tile, mi_row, mi_col, BLOCK_8X8);
        b.is_inter = true;
        b.use_intrabc = false;
        b.YMode = compound ? NEW_NEWMV : NEWMV;
        b.motion_mode = SIMPLE_TRANSLATION;
        b.compound_type = COMPOUND_AVERAGE;
        b.interintra = false;
        b.RefFrame[0] = u.ref_frame[0];
        b.RefFrame[1] = compound ? u.ref_frame[1] : NONE_FRAME;
        ModeInfoBlock& info = f.m_modeInfo[mi_row][mi_col];
        info.RefFrames[0] = b.RefFrame[0];
        info.RefFrames[1] = b.RefFrame[1];
        for (int l = 0; l < 2; l++) {
            info.Mvs[l].mv[0] = u.mv[l][0];
            info.Mvs[l].mv[1] = u.mv[l][1];
            info.InterpFilters[l] = (InterpFilter)u.filt[l];
        }
        Block::InterPredict ip(b, plane, *frame, store, mask);
        ip.predict_inter(u.x, u.y, u.w, u.h, mi_row, mi_col);
        n++;
    }
    for (int p = 0; p < 3; p++) {
        const int pw = p ? aw / 2 : aw, ph = p ? ah / 2 : ah;
        for (int y = 0; y < ph; y++) memcpy(out[p] + (size_t)y * out_strides[p], frame->data[p] + (size_t)y * frame->strides[p], pw);
    }
    return n;
}

// TransformBlock::inverseTransform() on `n` blocks.  For block i: tx_size[i], tx_type[i],
// lossless[i]; coefficients are int32 row-major min(w,32) x min(h,32) at coef + coef_off[i];
// the residual (w x h int32, BEFORE the flip mirroring of TransformBlock::decode) goes to
// res + res_off[i].
int oracle_inverse_transform(int n, const uint8_t* tx_size, const uint8_t* tx_type, const uint8_t* lossless,
    const int32_t* coef, const uint32_t* coef_off, int32_t* res, const uint32_t* res_off)
{
    auto seq = std::make_shared<SequenceHeader>();
    seq->BitDepth = 8;
    seq->subsampling_x = seq->subsampling_y = 1;
    seq->NumPlanes = 3;
    seq->use_128x128_superblock = false;
    ConstSequencePtr cseq = seq;
    auto fh = std::make_shared<FrameHeader>(cseq);
    FrameHeader& f = *fh;
    f.FrameWidth = 64;
    f.FrameHeight = 64;
    f.UpscaledWidth = 64;
    f.compute_image_size();
    f.initGeometry();
    f.TileCols = f.TileRows = 1;
    f.NumTiles = 1;
    f.MiColStarts = { 0, f.MiCols };
    f.MiRowStarts = { 0, f.MiRows };
    f.m_cdfs.reset(new Cdfs);
    static uint8_t dummy[64] = { 0 };
    Tile tile(cseq, fh, 0);
    tile.m_cdfs = *f.m_cdfs;
    tile.m_entropy.reset(new EntropyDecoder(dummy, sizeof(dummy), true, tile.m_cdfs));
    Block block(tile, 0, 0, BLOCK_64X64);
    block.is_inter = true;
    for (int i = 0; i < n; i++) {
        block.Lossless = lossless[i] != 0;
        std::unique_ptr<TransformBlock> tb(new TransformBlock(block, 0, 0, 0, 0, 0, (TX_SIZE)tx_size[i], false));
        tb->PlaneTxType = (TX_TYPE)tx_type[i];
        memset(tb->Dequant, 0, sizeof(tb->Dequant));
        const int tw = tb->tw, th = tb->th;
        for (int r = 0; r < th; r++)
            for (int c = 0; c < tw; c++) tb->Dequant[r][c] = coef[coef_off[i] + r * tw + c];
        tb->inverseTransform();
        for (int r = 0; r < tb->h; r++)
            for (int c = 0; c < tb->w; c++) res[res_off[i] + r * tb->w + c] = tb->Residual[r][c];
    }
    return 0;
}

}  // extern "C"

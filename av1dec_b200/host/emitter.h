// emitter.h -- turns one parsed frame (the reference front end's Tile/SuperBlock/Partition/
// Block/TransformBlock tree plus its FrameHeader) into the flat command buffer of
// include/av1b200_format.h.  This is the host half of the drop-in boundary: it walks the tree
// in the order Tile::decode() would (decoder/Tile.cpp:172) but serialises instead of touching
// pixels.
#pragma once
#include "../../include/av1b200_format.h"
#include <stddef.h>
#include <stdint.h>
#include <memory>
#include <vector>

namespace YamiAv1 {
class Tile;
class Block;
class TransformBlock;
struct FrameHeader;
struct SequenceHeader;
class Partition;
class SuperBlock;
}

namespace av1b200 {

class FrameEmitter {
public:
    void begin(YamiAv1::FrameHeader& frame, const YamiAv1::SequenceHeader& seq);
    void emitTile(YamiAv1::Tile& tile);
    void emitSb(YamiAv1::Tile& tile, YamiAv1::SuperBlock& sb); // one superblock of `tile` (emitTile = all of them, in order)
    void finish();
    size_t bytes() const { return m_total; }
    void write(uint8_t* dst) const;
    const Av1bFrameHdr& header() const { return m_hdr; }

private:
    void walk(YamiAv1::Partition& p);
    void emitBlock(YamiAv1::Block& b);
    void emitInter(YamiAv1::Block& b);
    void emitTb(YamiAv1::Block& b, YamiAv1::TransformBlock& t);
    void emitObmc(YamiAv1::Block& b, int plane, int w, int h);
    uint32_t auxFor(YamiAv1::Block& b);
    bool edgeSmooth(const YamiAv1::Block& b, int plane) const;
    void distanceWeights(int candRow, int candCol, int& fwd, int& bck) const;
    void layout();
    void scheduleSb(uint32_t first, size_t firstItx, int sbx, int sby);

    YamiAv1::FrameHeader* m_frame = nullptr;
    const YamiAv1::SequenceHeader* m_seq = nullptr;
    Av1bFrameHdr m_hdr;
    std::vector<Av1bSb> m_sbs;
    std::vector<Av1bOp> m_ops;
    std::vector<Av1bOp> m_itxOnly; // coded TBs of plain inter blocks (inverse transform only)
    std::vector<uint32_t> m_itx;
    std::vector<Av1bInterBlk> m_iblk;
    std::vector<Av1bIpu> m_ipu;
    std::vector<Av1bBlkAux> m_aux;
    std::vector<int16_t> m_coef;
    std::vector<uint8_t> m_pal;
    std::vector<Av1bLfMi> m_lfmi;
    std::vector<uint8_t> m_lftx; // [3][mi_rows*mi_cols] LoopfilterTxSizes
    std::vector<uint8_t> m_cdef8;
    std::vector<Av1bLrUnit> m_lru;
    std::vector<uint32_t> m_levels, m_count, m_perm;
    std::vector<uint16_t> m_sbDepth; // levels of each superblock scheduled so far (seeds of its neighbours' late ops)
    std::vector<Av1bOp> m_sorted;
    uint32_t m_nRes = 0;
    size_t m_total = 0;
    // per-block state
    uint32_t m_blockAux = 0xFFFFFFFFu;
    int m_maxLumaW = 0, m_maxLumaH = 0;
    bool m_gmReady = false;
};

}  // namespace av1b200

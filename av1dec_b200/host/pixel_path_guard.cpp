// pixel_path_guard.cpp -- the product links only the FRONT END of the reference (parser, entropy
// decoder, block syntax).  The reference's CPU pixel files (IntraPredict.cpp, LoopFilter.cpp,
// Cdef.cpp, LoopRestoration.cpp, VideoFrame.cpp, Av1Decoder.cpp) are NOT compiled in.  The few
// symbols of them that the mixed parse/decode translation units still reference are defined
// here as traps: if anything ever tried to reconstruct a pixel on the host, the process
// aborts instead of silently falling back to the CPU.
//
// The mixed translation units also DEFINE the reference's decode() tree walk (Tile / SuperBlock /
// Partition / Block ::decode, Tile.cpp:172, SuperBlock.cpp:46, Partition.cpp:207, Block.cpp:1600),
// kept alive by the BlockTree vtables.  The Makefile weakens those four symbols in the front-end
// objects; the strong definitions below replace them, and with them gone nothing references
// TransformBlock::decode / inverseTransform, Block::compute_prediction, InterPredict::predict_inter,
// blockWarp, Palette::predict_palette ... so the linker (--gc-sections) drops that code.
// tests/test_host.py checks with nm that none of it is left in the product.
#include "ref_access.h"
#include "IntraPredict.h"

#include <cstdio>
#include <cstdlib>

namespace {
[[noreturn]] void trap(const char* what)
{
    fprintf(stderr, "av1b200: FATAL: host pixel path reached (%s). Reconstruction runs on the GPU only.\n", what);
    abort();
}
}  // namespace

namespace YamiAv1 {

Block::IntraPredict::IntraPredict(const Block& block, const std::shared_ptr<YuvFrame>& yuv, int p, int startX, int startY,
    int log2w, int log2h, std::vector<std::vector<uint8_t>>& pred)
    : m_block(block)
    , m_tile(block.m_tile)
    , m_frame(block.m_frame)
    , m_sequence(block.m_sequence)
    , plane(p)
    , x(startX)
    , y(startY)
    , log2W(log2w)
    , log2H(log2h)
    , w(1 << log2w)
    , h(1 << log2h)
    , m_yuv(yuv)
    , m_pred(pred)
{
    trap("IntraPredict");
}

bool Tile::decode(std::shared_ptr<Yami::YuvFrame>&, const FrameStore&) { trap("Tile::decode"); }
bool SuperBlock::decode(std::shared_ptr<Yami::YuvFrame>&, const FrameStore&) { trap("SuperBlock::decode"); }
bool Partition::decode(std::shared_ptr<Yami::YuvFrame>&, const FrameStore&) { trap("Partition::decode"); }
bool Block::decode(std::shared_ptr<YuvFrame>&, const FrameStore&) { trap("Block::decode"); }

void Block::IntraPredict::predict_intra(int, int, bool, bool, int) { trap("predict_intra"); }
void Block::IntraPredict::predict_chroma_from_luma(TX_SIZE) { trap("predict_chroma_from_luma"); }

}  // namespace YamiAv1

// ref_access.h -- includes the reference decoder's own headers (from the reference tree, never
// copied) with member access opened up, so the command emitter can read the parse-time block
// tree exactly where Tile::decode() would have read it.
//
// Only the FRONT END of the reference (parser, entropy decoder, block syntax) is linked into
// the product; see av1dec_b200/build.py for the file list and pixel_path_guard.cpp for the
// trap that replaces its CPU pixel path.
#pragma once

// Standard headers first: the access hack must not leak into the C++ library.
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
#include <string.h>

#define private public
#define protected public
#include "BitReader.h"
#include "Parser.h"
#include "Tile.h"
#include "Block.h"
#include "Partition.h"
#include "SuperBlock.h"
#include "TransformBlock.h"
#include "VideoFrame.h"
#undef private
#undef protected

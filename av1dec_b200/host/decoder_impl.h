// decoder_impl.h -- internal glue shared by decoder.cpp, yami_adapter.cpp and capi.cpp.
#pragma once
#include "../../include/av1b200.h"
#include "emitter.h"
#include <memory>
#include <string>

namespace Yami {
struct YuvFrame;
}

namespace YamiAv1 {
class Decoder;
}

namespace av1b200 {

// Called once per frame with the finished command buffer (data == NULL: show_existing_frame,
// bytes = slot index).  Lets the benchmark record command buffers for the HBM-resident replay.
typedef void (*CmdSink)(void* user, const uint8_t* data, size_t bytes, uint32_t refresh_mask, int show);

struct DecoderOptions {
    int device = 0;
    uint32_t stages = AV1B_STAGE_ALL; // tests stop after a stage to compare against oracle dumps
    CmdSink sink = nullptr;
    void* sink_user = nullptr;
};

DecoderOptions& decoderOptions(YamiAv1::Decoder& d);
const char* decoderError(YamiAv1::Decoder& d);
av1b_ctx* decoderCtx(YamiAv1::Decoder& d);
bool decoderFormat(YamiAv1::Decoder& d, int& w, int& h);
void decoderFlush(YamiAv1::Decoder& d);
// Next output frame if it is already complete, or if more than keepInFlight frames are queued
// (then it blocks); null otherwise.  Lets a whole-stream loop parse ahead of the device.
std::shared_ptr<Yami::YuvFrame> decoderPollOutput(YamiAv1::Decoder& d, size_t keepInFlight);
// Run command emission + device submission on a second thread, overlapped with the parse of the
// next frame (used by the whole-stream entry point; getOutput() then waits for that thread).
void decoderSetAsync(YamiAv1::Decoder& d, bool on);
void decoderDrain(YamiAv1::Decoder& d);

}  // namespace av1b200

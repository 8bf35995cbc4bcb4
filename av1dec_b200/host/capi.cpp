// capi.cpp -- C ABI over YamiAv1::Decoder (include/av1b200_decoder.h).
#include "../../include/av1b200_decoder.h"
#include "Av1Decoder.h"
#include "decoder_impl.h"
#include "VideoFrame.h"

#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>

struct av1b_decoder {
    YamiAv1::Decoder dec;
    std::shared_ptr<Yami::YuvFrame> last;
};

extern "C" {

av1b_decoder* av1b_decoder_create(int device)
{
    av1b_decoder* d = new av1b_decoder;
    av1b200::decoderOptions(d->dec).device = device;
    return d;
}

void av1b_decoder_destroy(av1b_decoder* d)
{
    if (!d) return;
    d->last.reset();
    delete d;
}

void av1b_decoder_set_stages(av1b_decoder* d, uint32_t stages) { av1b200::decoderOptions(d->dec).stages = stages; }

void av1b_decoder_set_cmd_sink(av1b_decoder* d, av1b_cmd_sink sink, void* user)
{
    av1b200::decoderOptions(d->dec).sink = sink;
    av1b200::decoderOptions(d->dec).sink_user = user;
}

int av1b_decoder_decode(av1b_decoder* d, const uint8_t* data, size_t size)
{
    return d->dec.decode(const_cast<uint8_t*>(data), size) ? 0 : -1;
}

int av1b_decoder_get_output(av1b_decoder* d, int* width, int* height, const uint8_t* planes[3], int strides[3])
{
    d->last = d->dec.getOutput();
    if (!d->last) return 0;
    *width = d->last->width;
    *height = d->last->height;
    for (int p = 0; p < 3; p++) {
        planes[p] = d->last->data[p];
        strides[p] = d->last->strides[p];
    }
    return 1;
}

const char* av1b_decoder_error(av1b_decoder* d) { return av1b200::decoderError(d->dec); }

static uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

int av1b_decode_ivf(const uint8_t* ivf, size_t len, int device, uint32_t stages, uint8_t* out_yuv, size_t out_cap,
    size_t* out_bytes, int* n_frames, uint64_t* luma_pixels)
{
    if (!ivf || len < 32 || memcmp(ivf, "DKIF", 4) != 0) return -1;
    const size_t hdr = ivf[6] | (ivf[7] << 8);
    size_t pos = hdr, out = 0;
    int frames = 0, rc = 0;
    uint64_t pixels = 0;
    YamiAv1::Decoder dec;
    av1b200::decoderOptions(dec).device = device;
    av1b200::decoderOptions(dec).stages = stages;
    av1b200::decoderSetAsync(dec, getenv("AV1B200_SYNC_EMIT") == nullptr);
    // keepInFlight frames may still be on the device while the next temporal unit is parsed
    auto drain = [&](size_t keepInFlight) {
        std::shared_ptr<Yami::YuvFrame> f;
        while ((f = keepInFlight ? av1b200::decoderPollOutput(dec, keepInFlight) : dec.getOutput())) {
            frames++;
            pixels += (uint64_t)f->width * f->height;
            for (int p = 0; p < 3; p++) {
                const int w = p ? (f->width >> 1) : f->width, h = p ? (f->height >> 1) : f->height;
                if (out_yuv && out + (size_t)w * h <= out_cap) {
                    for (int y = 0; y < h; y++) memcpy(out_yuv + out + (size_t)y * w, f->data[p] + (size_t)y * f->strides[p], w);
                } else if (out_yuv) {
                    rc = -2;
                }
                out += (size_t)w * h;
            }
        }
    };
    while (pos + 12 <= len) {
        const uint32_t sz = rd32(ivf + pos);
        pos += 12;
        if (pos + sz > len) break;
        if (!dec.decode(const_cast<uint8_t*>(ivf + pos), sz)) {
            rc = -1;
            break;
        }
        pos += sz;
        drain(4);
    }
    if (rc != -1) {
        av1b200::decoderDrain(dec);
        drain(0);
    }
    if (out_bytes) *out_bytes = out;
    if (n_frames) *n_frames = frames;
    if (luma_pixels) *luma_pixels = pixels;
    return rc;
}

}  // extern "C"

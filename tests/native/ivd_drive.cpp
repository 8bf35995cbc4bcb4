// ivd_drive.cpp -- TEST PROGRAM: drives the Yami IVideoDecoder interface the reference declares
// (interface/VideoDecoderInterface.h:31-68, factory interface/VideoDecoderHost.h:26-41) exactly as
// a libyami client would: dlopen the decoder library, dlsym createVideoDecoder, then
//   start -> decode(temporal unit)* -> getOutput* -> decode(EOS) -> flush -> stop -> release.
// Every shown frame's visible I420 area is appended to <out.yuv>; the caller compares its MD5 with
// bits/bits.md5.  Modes:
//   plain   one straight pass
//   flush   decode the first unit, flush() (cached frames must be dropped), reset(), full pass
//   twice   full pass, stop(), start(), full pass again into the same file (2x the frames)
// usage: ivd_drive <libav1b200dec.so> <in.ivf> <out.yuv> [plain|flush|twice]
#include <VideoDecoderHost.h>
#include "VideoFrame.h"

#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace YamiMediaCodec;

static std::vector<std::vector<uint8_t>> read_ivf(const char* path)
{
    std::vector<std::vector<uint8_t>> units;
    FILE* f = fopen(path, "rb");
    if (!f) return units;
    uint8_t hdr[32];
    if (fread(hdr, 1, 32, f) != 32 || memcmp(hdr, "DKIF", 4)) {
        fclose(f);
        return units;
    }
    uint8_t fh[12];
    while (fread(fh, 1, 12, f) == 12) {
        const uint32_t n = fh[0] | (fh[1] << 8) | (fh[2] << 16) | ((uint32_t)fh[3] << 24);
        std::vector<uint8_t> u(n);
        if (n && fread(u.data(), 1, n, f) != n) break;
        units.push_back(std::move(u));
    }
    fclose(f);
    return units;
}

static int drain(IVideoDecoder* dec, FILE* out, int64_t want_pts)
{
    int n = 0;
    for (;;) {
        SharedPtr<VideoFrame> vf = dec->getOutput();
        if (!vf) break;
        if (vf->fourcc != YAMI_FOURCC_I420 || !vf->surface) {
            fprintf(stderr, "ivd_drive: unexpected frame format\n");
            exit(3);
        }
        (void)want_pts;
        const Yami::YuvFrame* y = (const Yami::YuvFrame*)vf->surface;
        const int w = (int)vf->crop.width, h = (int)vf->crop.height;
        for (int p = 0; p < 3; p++) {
            const int pw = p ? w >> 1 : w, ph = p ? h >> 1 : h;
            for (int r = 0; r < ph; r++) fwrite(y->data[p] + (size_t)r * y->strides[p], 1, pw, out);
        }
        n++;
    }
    return n;
}

static int full_pass(IVideoDecoder* dec, const std::vector<std::vector<uint8_t>>& units, FILE* out)
{
    int frames = 0;
    int64_t pts = 0;
    for (const auto& u : units) {
        VideoDecodeBuffer b;
        memset(&b, 0, sizeof(b));
        b.data = const_cast<uint8_t*>(u.data());
        b.size = u.size();
        b.timeStamp = pts++;
        const YamiStatus st = dec->decode(&b);
        if (st != YAMI_SUCCESS) {
            fprintf(stderr, "ivd_drive: decode() returned %d\n", (int)st);
            exit(4);
        }
        frames += drain(dec, out, b.timeStamp);
    }
    VideoDecodeBuffer eos;
    memset(&eos, 0, sizeof(eos));
    if (dec->decode(&eos) != YAMI_SUCCESS) exit(5);
    frames += drain(dec, out, -1);
    return frames;
}

int main(int argc, char** argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: ivd_drive <decoder.so> <in.ivf> <out.yuv> [plain|flush|twice]\n");
        return 2;
    }
    const std::string mode = argc > 4 ? argv[4] : "plain";
    void* lib = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        fprintf(stderr, "ivd_drive: dlopen: %s\n", dlerror());
        return 2;
    }
    YamiCreateVideoDecoderFuncPtr create = (YamiCreateVideoDecoderFuncPtr)dlsym(lib, "createVideoDecoder");
    YamiReleaseVideoDecoderFuncPtr release = (YamiReleaseVideoDecoderFuncPtr)dlsym(lib, "releaseVideoDecoder");
    if (!create || !release) {
        fprintf(stderr, "ivd_drive: factory symbols missing\n");
        return 2;
    }
    if (create("video/h264")) {
        fprintf(stderr, "ivd_drive: a decoder was created for a foreign MIME type\n");
        return 6;
    }
    IVideoDecoder* dec = create(YAMI_MIME_AV1);
    if (!dec) return 6;
    const auto units = read_ivf(argv[2]);
    if (units.empty()) {
        fprintf(stderr, "ivd_drive: cannot read %s\n", argv[2]);
        return 2;
    }
    FILE* out = fopen(argv[3], "wb");
    if (!out) return 2;
    dec->setNativeDisplay(nullptr);
    dec->setAllocator(nullptr);
    if (dec->start(nullptr) != YAMI_SUCCESS) return 7;
    int frames = 0;
    if (mode == "flush") {
        // seek-like use: decode one unit, do NOT fetch it, flush, start over
        VideoDecodeBuffer b;
        memset(&b, 0, sizeof(b));
        b.data = const_cast<uint8_t*>(units[0].data());
        b.size = units[0].size();
        if (dec->decode(&b) != YAMI_SUCCESS) return 4;
        dec->flush();
        if (dec->getOutput()) {
            fprintf(stderr, "ivd_drive: flush() left a cached frame\n");
            return 8;
        }
        if (dec->reset(nullptr) != YAMI_SUCCESS) return 7;
    }
    frames += full_pass(dec, units, out);
    const VideoFormatInfo* fi = dec->getFormatInfo();
    if (!fi || !fi->valid || fi->width <= 0 || fi->height <= 0) {
        fprintf(stderr, "ivd_drive: getFormatInfo() not valid after decoding\n");
        return 9;
    }
    printf("format %dx%d\n", (int)fi->width, (int)fi->height);
    if (mode == "twice") {
        dec->stop();
        if (dec->start(nullptr) != YAMI_SUCCESS) return 7;
        frames += full_pass(dec, units, out);
    }
    dec->flush();
    dec->stop();
    release(dec);
    fclose(out);
    printf("frames %d\n", frames);
    return 0;
}

// capi.cpp -- C ABI over YamiAv1::Decoder (include/av1b200_decoder.h).
#include "../../include/av1b200_decoder.h"
#include "Av1Decoder.h"
#include "decoder_impl.h"
#include "VideoFrame.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

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

// ---- whole-stream decode --------------------------------------------------------------------
namespace {

struct Tu {
    size_t off, size;
    bool rap; // random access point: sequence header + shown key frame first -> all state resets
};

// One OBU of a temporal unit: type and payload.
static bool next_obu(const uint8_t* p, size_t n, size_t& pos, int& type, const uint8_t*& payload, size_t& psize)
{
    if (pos >= n) return false;
    const uint8_t h = p[pos++];
    type = (h >> 3) & 15;
    if (h & 4) pos++; // extension byte
    if (h & 2) {
        uint64_t v = 0;
        for (int i = 0; i < 8; i++) {
            if (pos >= n) return false;
            const uint8_t b = p[pos++];
            v |= (uint64_t)(b & 0x7F) << (7 * i);
            if (!(b & 0x80)) break;
        }
        psize = (size_t)v;
    } else {
        psize = pos <= n ? n - pos : 0;
    }
    if (pos + psize > n) return false;
    payload = p + pos;
    pos += psize;
    return true;
}

// Spec 7.5 / 5.9.2: a temporal unit that carries a sequence header and whose first frame is a
// shown KEY_FRAME refreshes every reference slot and loads default CDFs -- nothing decoded
// before it is ever read again, so the stream can be cut there.
static bool tu_is_rap(const uint8_t* p, size_t n)
{
    size_t pos = 0;
    int type;
    const uint8_t* pl;
    size_t ps;
    bool seq = false, reduced = false;
    while (next_obu(p, n, pos, type, pl, ps)) {
        if (type == 1 && ps) { // OBU_SEQUENCE_HEADER: seq_profile(3) still_picture(1) reduced_still_picture_header(1)
            seq = true;
            reduced = (pl[0] >> 3) & 1;
        } else if (type == 3 || type == 6) { // OBU_FRAME_HEADER / OBU_FRAME
            if (!seq || !ps) return false;
            if (reduced) return true;
            // show_existing_frame(1) frame_type(2) show_frame(1)
            return (pl[0] >> 4) == 0x1;
        }
    }
    return false;
}

struct Sink {
    uint8_t* dst;  // caller's buffer (single-segment decode) or null
    size_t cap;
    std::vector<uint8_t>* vec; // private buffer (parallel segments)
    size_t bytes = 0;
    int frames = 0;
    uint64_t pixels = 0;
    bool overflow = false;
    void put(const Yami::YuvFrame& f)
    {
        frames++;
        pixels += (uint64_t)f.width * f.height;
        for (int p = 0; p < 3; p++) {
            const int w = p ? (f.width >> 1) : f.width, h = p ? (f.height >> 1) : f.height;
            const size_t n = (size_t)w * h;
            uint8_t* o = nullptr;
            if (vec) {
                vec->resize(bytes + n);
                o = vec->data() + bytes;
            } else if (dst && bytes + n <= cap) {
                o = dst + bytes;
            } else if (dst) {
                overflow = true;
            }
            if (o)
                for (int y = 0; y < h; y++) memcpy(o + (size_t)y * w, f.data[p] + (size_t)y * f.strides[p], w);
            bytes += n;
        }
    }
};

// Temporal units [first, last) through one decoder, outputs in order into `sink`.
static bool decode_range(YamiAv1::Decoder& dec, const uint8_t* ivf, const std::vector<Tu>& tus, size_t first, size_t last, Sink& sink)
{
    auto drain = [&](size_t keepInFlight) {
        // keepInFlight frames may still be on the device while the next temporal unit is parsed
        std::shared_ptr<Yami::YuvFrame> f;
        while ((f = keepInFlight ? av1b200::decoderPollOutput(dec, keepInFlight) : dec.getOutput())) sink.put(*f);
    };
    for (size_t i = first; i < last; i++) {
        if (!dec.decode(const_cast<uint8_t*>(ivf + tus[i].off), tus[i].size)) return false;
        drain(4);
    }
    av1b200::decoderDrain(dec);
    drain(0);
    return true;
}

}  // namespace

static bool scan_ivf(const uint8_t* ivf, size_t len, std::vector<Tu>& tus, std::vector<size_t>& seg)
{
    if (!ivf || len < 32 || memcmp(ivf, "DKIF", 4) != 0) return false;
    const size_t hdr = ivf[6] | (ivf[7] << 8);
    for (size_t pos = hdr; pos + 12 <= len;) {
        const uint32_t sz = rd32(ivf + pos);
        pos += 12;
        if (pos + sz > len) break;
        Tu t{ pos, sz, tu_is_rap(ivf + pos, sz) };
        if (tus.empty() || t.rap) seg.push_back(tus.size()); // segment = index of its first temporal unit
        tus.push_back(t);
        pos += sz;
    }
    return true;
}

int av1b_ivf_segments(const uint8_t* ivf, size_t len, uint32_t* seg_first, int cap)
{
    std::vector<Tu> tus;
    std::vector<size_t> seg;
    if (!scan_ivf(ivf, len, tus, seg)) return -1;
    for (size_t i = 0; i < seg.size() && seg_first && (int)i < cap; i++) seg_first[i] = (uint32_t)seg[i];
    return (int)seg.size();
}

// Streams that are cut into closed segments by random access points (every temporal unit of an
// all-intra stream, every GOP of a closed-GOP stream) decode segment-parallel: the entropy decode
// of one frame is serial, so this is the only host-side parallelism inside a stream.  Each worker
// owns a decoder (front end + device context); outputs are stitched back in stream order.
// AV1B200_GOP_THREADS caps the workers (default min(8, host threads); 1 = serial).
int av1b_decode_ivf(const uint8_t* ivf, size_t len, int device, uint32_t stages, uint8_t* out_yuv, size_t out_cap,
    size_t* out_bytes, int* n_frames, uint64_t* luma_pixels)
{
    std::vector<Tu> tus;
    std::vector<size_t> seg;
    if (!scan_ivf(ivf, len, tus, seg)) return -1;
    // The command emitter can run on a worker thread one frame behind the parser: lower latency
    // for a lone stream, but ~35 % more CPU in total (the tree of a whole frame is built and crosses
    // cores; inline emission streams superblock by superblock, decoder.cpp).  With more than two
    // decodes running in the process it is a throughput service: emit inline.  (The count is per
    // process: a rule relative to the machine's cores picked the expensive mode in every rank of
    // a multi-GPU job that saturated the box as a whole.)
    static std::atomic<int> active{ 0 };
    struct Busy {
        std::atomic<int>& n;
        explicit Busy(std::atomic<int>& a) : n(a) { n++; }
        ~Busy() { n--; }
    };
    auto setup = [&](YamiAv1::Decoder& dec) {
        av1b200::decoderOptions(dec).device = device;
        av1b200::decoderOptions(dec).stages = stages;
        const char* e = getenv("AV1B200_SYNC_EMIT");
        const bool async = e ? atoi(e) == 0 : active.load() <= 2;
        av1b200::decoderSetAsync(dec, async);
    };
    unsigned workers = std::min<unsigned>(8, std::max(1u, std::thread::hardware_concurrency()));
    if (const char* e = getenv("AV1B200_GOP_THREADS")) workers = (unsigned)std::max(1, atoi(e));
    workers = (unsigned)std::min<size_t>(workers, seg.size());
    if (strcmp(av1b_backend(), "cuda-sm_100a") != 0) workers = 1; // the test-only emulation is single-threaded
    int rc = 0;
    size_t out = 0;
    int frames = 0;
    uint64_t pixels = 0;
    if (workers <= 1) {
        Busy busy(active);
        YamiAv1::Decoder dec;
        setup(dec);
        Sink sink{ out_yuv, out_cap, nullptr };
        if (!tus.empty() && !decode_range(dec, ivf, tus, 0, tus.size(), sink)) rc = -1;
        else if (sink.overflow) rc = -2;
        out = sink.bytes, frames = sink.frames, pixels = sink.pixels;
    } else {
        seg.push_back(tus.size());
        const size_t nseg = seg.size() - 1;
        std::vector<std::vector<uint8_t>> bufs(nseg);
        std::vector<Sink> sinks(nseg, Sink{ nullptr, 0, nullptr });
        std::atomic<size_t> next{ 0 };
        std::atomic<bool> failed{ false };
        auto work = [&]() {
            Busy busy(active);
            YamiAv1::Decoder dec;
            setup(dec);
            for (size_t s; !failed && (s = next++) < nseg;) {
                sinks[s].vec = out_yuv ? &bufs[s] : nullptr;
                if (!decode_range(dec, ivf, tus, seg[s], seg[s + 1], sinks[s])) failed = true;
            }
        };
        std::vector<std::thread> pool;
        for (unsigned w = 1; w < workers; w++) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
        if (failed) rc = -1;
        for (size_t s = 0; s < nseg && rc != -1; s++) {
            if (out_yuv && out + sinks[s].bytes <= out_cap) memcpy(out_yuv + out, bufs[s].data(), sinks[s].bytes);
            else if (out_yuv) rc = -2;
            out += sinks[s].bytes;
            frames += sinks[s].frames;
            pixels += sinks[s].pixels;
        }
    }
    if (out_bytes) *out_bytes = out;
    if (n_frames) *n_frames = frames;
    if (luma_pixels) *luma_pixels = pixels;
    return rc;
}

}  // extern "C"

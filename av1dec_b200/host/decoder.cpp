// decoder.cpp -- YamiAv1::Decoder on top of the B200 engine.
//
// Frame lifecycle of the reference (decoder/Av1Decoder.cpp):
//     decode() OBU loop :49-109  -> kept on the host, drives the reference's own Parser
//     decodeFrame() :128-156     -> emit commands, av1b_frame_submit() (recon + filters on GPU)
//     showExistingFrame() :158   -> av1b_show_existing()
//     updateFrameStore() :111    -> refresh mask handed to the engine (device-resident refs)
//     getOutput() :203           -> async D2H into pooled pinned I420 frames
// There is no host pixel path: if the engine reports an error, decode() returns false.
#include "ref_access.h"
#include "Av1Decoder.h"
#include "decoder_impl.h"

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>

using namespace Yami;

namespace YamiAv1 {

namespace {

// Pinned host frame handed to the caller; same field layout as the reference's YuvFrame.
struct HostFrame : public YuvFrame {
    void* buf = nullptr;
    size_t cap = 0;
};

// Returned frames come back on whatever thread drops the last reference (the caller's), while the
// emit worker takes frames out: every access to free_list holds mu.
struct HostPool {
    std::mutex mu;
    std::vector<HostFrame*> free_list;
    ~HostPool()
    {
        for (HostFrame* f : free_list) {
            av1b_host_free(f->buf);
            delete f;
        }
    }
};

std::shared_ptr<YuvFrame> acquireHostFrame(const std::shared_ptr<HostPool>& pool, int w, int h)
{
    const int sy = (w + 63) & ~63, sc = ((w >> 1) + 63) & ~63;
    const size_t need = (size_t)sy * h + 2 * (size_t)sc * (h >> 1);
    HostFrame* f = nullptr;
    {
        std::lock_guard<std::mutex> lk(pool->mu);
        for (size_t i = 0; i < pool->free_list.size(); i++) {
            if (pool->free_list[i]->cap >= need) {
                f = pool->free_list[i];
                pool->free_list.erase(pool->free_list.begin() + i);
                break;
            }
        }
    }
    if (!f) {
        f = new HostFrame;
        f->buf = av1b_host_alloc(need);
        f->cap = need;
        if (!f->buf) {
            delete f;
            return nullptr;
        }
    }
    uint8_t* p = (uint8_t*)f->buf;
    f->pts = 0;
    f->width = w;
    f->height = h;
    f->data[0] = p;
    f->data[1] = p + (size_t)sy * h;
    f->data[2] = f->data[1] + (size_t)sc * (h >> 1);
    f->strides[0] = sy;
    f->strides[1] = f->strides[2] = sc;
    f->widths[0] = w;
    f->heights[0] = h;
    f->widths[1] = f->widths[2] = w / 2;
    f->heights[1] = f->heights[2] = h / 2;
    std::shared_ptr<HostPool> keep = pool;
    return std::shared_ptr<YuvFrame>(static_cast<YuvFrame*>(f), [keep](YuvFrame* y) {
        std::lock_guard<std::mutex> lk(keep->mu);
        keep->free_list.push_back(static_cast<HostFrame*>(y));
    });
}

}  // namespace

struct Decoder::Impl {
    std::unique_ptr<Parser> parser;
    FramePtr frame;
    TileGroup tiles;
    av1b_ctx* ctx = nullptr;
    int ctx_w = 0, ctx_h = 0; // size the device context was created for
    av1b200::FrameEmitter emitter;
    std::shared_ptr<HostPool> pool = std::make_shared<HostPool>();
    struct Pending {
        std::shared_ptr<YuvFrame> frame;
        uint64_t fence;
    };
    std::deque<Pending> output;   // guarded by mu when the emit thread runs
    int frame_w[AV1B_MAX_FRAME_IDS], frame_h[AV1B_MAX_FRAME_IDS]; // visible size of each device frame id
    av1b200::DecoderOptions opt;
    std::string error;
    // AV1B200_TIMING=1: cumulative host-side phase timers, printed when the decoder is destroyed
    bool timing = getenv("AV1B200_TIMING") != nullptr;
    double t_dbg[4] = { 0, 0, 0, 0 };
    double t_parse = 0, t_emit = 0, t_submit = 0, t_wait = 0;
    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

    bool fail(const char* what)
    {
        error = what;
        if (ctx) {
            error += ": ";
            error += av1b_last_error(ctx);
        }
        fprintf(stderr, "av1b200: %s\n", error.c_str());
        return false;
    }

    bool ensureCtx(const SequenceHeader& s)
    {
        if (ctx) return true;
        const int mw = s.max_frame_width_minus_1 + 1, mh = s.max_frame_height_minus_1 + 1;
        if (s.BitDepth != 8 || !s.subsampling_x || !s.subsampling_y || s.mono_chrome) return fail("only 8-bit 4:2:0 streams are supported (same envelope as the reference)");
        if (av1b_ctx_create(&ctx, opt.device, mw, mh, nullptr) != AV1B_OK) {
            ctx = nullptr; // av1b_ctx_create leaves *out null on failure; never keep a half-built context
            return fail("av1b_ctx_create");
        }
        ctx_w = mw;
        ctx_h = mh;
        return true;
    }

    bool queueOutput(int fid)
    {
        const int w = frame_w[fid], h = frame_h[fid];
        std::shared_ptr<YuvFrame> hf = acquireHostFrame(pool, w, h);
        if (!hf) return fail("pinned output allocation");
        uint8_t* dst[3] = { hf->data[0], hf->data[1], hf->data[2] };
        int st[3] = { hf->strides[0], hf->strides[1], hf->strides[2] };
        if (av1b_frame_download(ctx, fid, dst, st, w, h) != AV1B_OK) return fail("av1b_frame_download");
        uint64_t fence = 0;
        if (av1b_fence_record(ctx, &fence) != AV1B_OK) return fail("av1b_fence_record");
        {
            std::lock_guard<std::mutex> lk(mu);
            output.push_back(Pending{ hf, fence });
        }
        return true;
    }

    // ---- device side of one frame: serialise the parsed tree and hand it to the engine
    // Streaming emission (sync mode): a superblock is serialised right after it is parsed, while
    // its block tree (37 KB per transform block) is still in the cache, and freed at once -- the
    // tree of a whole frame (tens of MB even at CIF) is never built.  `streamed` = the emitter
    // already holds this frame's superblocks when submitFrame runs.
    bool streaming = false, streamed = false;

    bool beginFrame(const FramePtr& fr, const std::shared_ptr<const SequenceHeader>& seq)
    {
        FrameHeader& h = *fr;
        if (!ensureCtx(*seq)) return false;
        // outside the envelope (the reference asserts: Av1Decoder.cpp:194-200, InterPredict.cpp:395):
        // refuse loudly instead of decoding as if unscaled
        if (h.use_superres) return fail("super-resolution is not supported (same envelope as the reference)");
        if (!h.FrameIsIntra)
            for (int rf = LAST_FRAME; rf <= ALTREF_FRAME; rf++)
                if (h.is_scaled(rf)) return fail("scaled reference frames are not supported (same envelope as the reference)");
        if (h.MiCols * 4 > ((ctx_w + 7) & ~7) + 128 || (int)h.FrameWidth > ctx_w || (int)h.FrameHeight > ctx_h) {
            // a later, larger sequence header: rebuild the device context (references are lost, as
            // they would be useless at a new size anyway -- the new sequence starts with a key frame)
            if (av1b_sync(ctx) != AV1B_OK) return fail("av1b_sync");
            av1b_ctx_destroy(ctx);
            ctx = nullptr;
            if (!ensureCtx(*seq)) return false;
            if ((int)h.FrameWidth > ctx_w || (int)h.FrameHeight > ctx_h) return fail("frame larger than the sequence's maximum size");
        }
        emitter.begin(h, *seq);
        return true;
    }

    bool submitFrame(const FramePtr& fr, TileGroup& ts, const std::shared_ptr<const SequenceHeader>& seq)
    {
        FrameHeader& h = *fr;
        const bool had = streamed;
        streamed = false;
        if (!had && !beginFrame(fr, seq)) return false;
        const double t0 = now();
        for (auto& t : ts) {
            if (had) break;
            const double ta = now();
            emitter.emitTile(*t);
            const double tb = now();
            // the block tree is no longer needed (Tile::decode pops as it goes).  Freed here: handing it to
            // a background thread was measured (-8 % end to end: the frees contend for the callers' malloc arenas)
            t->m_sbs.clear();
            t_dbg[0] += tb - ta;
            t_dbg[1] += now() - tb;
        }
        const double tc = now();
        emitter.finish();
        const double t1 = now();
        t_dbg[2] += t1 - tc;
        t_emit += t1 - t0;
        const size_t bytes = emitter.bytes();
        void* slot = nullptr;
        if (av1b_cmd_acquire(ctx, bytes, &slot) != AV1B_OK) return fail("av1b_cmd_acquire");
        emitter.write((uint8_t*)slot);
        if (opt.sink) opt.sink(opt.sink_user, (const uint8_t*)slot, bytes, h.refresh_frame_flags, h.show_frame);
        int fid = -1;
        if (av1b_frame_submit(ctx, bytes, opt.stages, h.refresh_frame_flags, &fid) != AV1B_OK) return fail("av1b_frame_submit");
        frame_w[fid] = h.FrameWidth;
        frame_h[fid] = h.FrameHeight;
        if (h.show_frame && !queueOutput(fid)) return false;
        t_submit += now() - t1;
        return true;
    }

    bool submitShowExisting(int slotIdx, uint32_t refresh)
    {
        if (!ctx) return fail("show_existing_frame before any frame");
        int fid = -1;
        if (av1b_show_existing(ctx, slotIdx, refresh, &fid) != AV1B_OK) return fail("av1b_show_existing");
        if (opt.sink) opt.sink(opt.sink_user, nullptr, (size_t)slotIdx, refresh, 1);
        return queueOutput(fid);
    }

    // ---- optional emit thread: frame N is serialised + submitted while frame N+1 is parsed.
    // Parsing N+1 needs only host state of N (CDFs, motion vectors, reference bookkeeping), which
    // the parse thread finishes itself; the emit thread reads N's block tree and mode-info only.
    struct Job {
        FramePtr frame;
        TileGroup tiles;
        std::shared_ptr<const SequenceHeader> seq;
        int showSlot = -1;
        uint32_t refresh = 0;
    };
    bool async = false;
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job> jobs;
    bool busy = false, stopping = false, failed = false;

    void workerLoop()
    {
        for (;;) {
            Job job;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stopping || !jobs.empty(); });
                if (jobs.empty()) return;
                job = std::move(jobs.front());
                jobs.pop_front();
                busy = true;
            }
            const bool ok = job.showSlot >= 0 ? submitShowExisting(job.showSlot, job.refresh) : submitFrame(job.frame, job.tiles, job.seq);
            job = Job();
            {
                std::lock_guard<std::mutex> lk(mu);
                busy = false;
                if (!ok) failed = true;
            }
            cv.notify_all();
        }
    }

    bool post(Job&& job)
    {
        std::unique_lock<std::mutex> lk(mu);
        if (!worker.joinable()) worker = std::thread([this] { workerLoop(); });
        cv.wait(lk, [&] { return jobs.size() < 2; }); // bounded look-ahead
        if (failed) return false;
        jobs.push_back(std::move(job));
        lk.unlock();
        cv.notify_all();
        return true;
    }

    void drainWorker()
    {
        if (!async) return;
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return jobs.empty() && !busy; });
    }

    void stopWorker()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            stopping = true;
        }
        cv.notify_all();
        if (worker.joinable()) worker.join();
    }

    // Tile::parse (Tile.cpp:122-160) with the emitter called after every superblock; the
    // superblock is dropped instead of being kept in Tile::m_sbs.
    bool streamTile(Tile& t, const uint8_t* data, uint32_t size)
    {
        t.m_cdfs = *t.m_frame->m_cdfs;
        t.m_entropy.reset(new EntropyDecoder(data, size, t.m_frame->disable_cdf_update, t.m_cdfs));
        t.clear_above_context();
        for (int i = 0; i < FRAME_LF_COUNT; i++) t.DeltaLF[i] = 0;
        t.m_frame->m_loopRestoration.resetRefs(t.m_sequence->NumPlanes);
        const BLOCK_SIZE sbSize = t.m_sequence->use_128x128_superblock ? BLOCK_128X128 : BLOCK_64X64;
        const int sbSize4 = Num_4x4_Blocks_Wide[sbSize];
        for (int r = t.MiRowStart; r < t.MiRowEnd; r += sbSize4) {
            t.clear_left_context();
            for (int c = t.MiColStart; c < t.MiColEnd; c += sbSize4) {
                t.ReadDeltas = t.m_frame->m_deltaQ.delta_q_present;
                const double ta = now();
                SuperBlock sb(t, r, c, sbSize);
                sb.parse();
                const double tb = now();
                emitter.emitSb(t, sb);
                t_parse += tb - ta;
                t_dbg[0] += now() - tb;
            }
        }
        return true;
    }

    // Parser::parseTileGroup (Parser.cpp:474-524) over streamTile
    bool streamTileGroup(BitReader& br, const FramePtr& fr, TileGroup& group)
    {
        bool tile_start_and_end_present_flag = false;
        if (fr->NumTiles > 1 && !br.readT(tile_start_and_end_present_flag)) return false;
        if (tile_start_and_end_present_flag) return fail("tile_start_and_end_present_flag is not supported (the reference asserts)");
        parser->skipTrailingBits(br);
        const uint8_t* data = br.getCurrent();
        uint32_t size = (uint32_t)(br.getRemainingBitsCount() / 8);
        const int tg_end = (int)fr->NumTiles - 1;
        for (int TileNum = 0; TileNum <= tg_end; TileNum++) {
            uint32_t tileSize;
            if (TileNum == tg_end) tileSize = size;
            else {
                uint32_t tile_size_minus_1;
                BitReader reader(data, size);
                if (!reader.readLe(tile_size_minus_1, fr->TileSizeBytes)) return fail("read tile_size_minus_1 failed");
                data += fr->TileSizeBytes;
                size -= fr->TileSizeBytes;
                tileSize = tile_size_minus_1 + 1;
            }
            if (tileSize > size) return fail("tile size exceeds the remaining data");
            std::shared_ptr<Tile> tile(new Tile(parser->m_sequence, fr, TileNum));
            if (!streamTile(*tile, data, tileSize)) return false;
            data += tileSize;
            size -= tileSize;
            group.push_back(tile);
        }
        return true;
    }

    bool decodeFrame(TileGroup& ts)
    {
        FrameHeader& h = *frame;
        std::shared_ptr<const SequenceHeader> seq = parser->m_sequence;
        bool ok;
        if (async) {
            Job job;
            job.frame = frame;
            job.tiles = ts;
            job.seq = seq;
            ok = post(std::move(job));
        } else {
            ok = submitFrame(frame, ts, seq);
        }
        // host-only state the next parse depends on (Av1Decoder.cpp:139,190; Parser.cpp:1784)
        for (auto& t : ts) t->frame_end_update_cdf();
        h.motionVectorStorage();
        parser->finishFrame();
        return ok;
    }

    bool showExisting()
    {
        FrameHeader& h = *frame;
        bool ok;
        if (async) {
            Job job;
            job.showSlot = h.frame_to_show_map_idx;
            job.refresh = h.refresh_frame_flags;
            ok = post(std::move(job));
        } else {
            ok = submitShowExisting(h.frame_to_show_map_idx, h.refresh_frame_flags);
        }
        h.referenceFrameLoading();
        h.motionVectorStorage();
        parser->finishFrame();
        return ok;
    }
};

Decoder::Decoder()
    : m_impl(new Impl)
{
    m_impl->parser.reset(new Parser);
    const char* dev = getenv("AV1B200_DEVICE");
    if (dev) m_impl->opt.device = atoi(dev);
    const char* se = getenv("AV1B200_STREAM_EMIT");
    m_impl->streaming = se ? atoi(se) != 0 : true;
}

Decoder::~Decoder()
{
    m_impl->stopWorker();
    if (m_impl->timing)
        fprintf(stderr, "av1b200 timing: parse %.3fs emit %.3fs (walk %.3f, tree free %.3f, finish %.3f) submit %.3fs wait %.3fs\n", m_impl->t_parse, m_impl->t_emit,
            m_impl->t_dbg[0], m_impl->t_dbg[1], m_impl->t_dbg[2], m_impl->t_submit, m_impl->t_wait);
    if (m_impl->ctx) {
        av1b_sync(m_impl->ctx);
        m_impl->output.clear();
        av1b_ctx_destroy(m_impl->ctx);
    }
}

bool Decoder::decode(uint8_t* data, size_t size)
{
    Impl& d = *m_impl;
    Parser& parser = *d.parser;
    BitReader reader(data, size);
    while (reader.getRemainingBitsCount() > 0) {
        obu_header hdr;
        if (!hdr.parse(reader)) return false;
        const uint64_t payload = hdr.obu_size;
        BitReader br(data + (reader.getPos() >> 3), payload);
        bool ok = true;
        switch (hdr.obu_type) {
        case OBU_SEQUENCE_HEADER: ok = parser.parseSequenceHeader(br); break;
        case OBU_TD: ok = parser.parseTemporalDelimiter(br); break;
        case OBU_FRAME_HEADER:
            if (d.tiles.empty()) d.streamed = false; // (not a redundant copy between the tile groups of one frame)
            d.frame = parser.parseFrameHeader(br);
            ok = bool(d.frame);
            if (ok && d.frame->show_existing_frame) ok = d.showExisting();
            break;
        case OBU_TILE_GROUP: {
            if (!d.frame) return false;
            TileGroup group;
            if (d.streaming && !d.async) {
                if (d.tiles.empty() && !d.streamed) {
                    ok = d.beginFrame(d.frame, parser.m_sequence);
                    d.streamed = ok;
                }
                ok = ok && d.streamTileGroup(br, d.frame, group);
                if (!ok) d.streamed = false; // the emitter holds a partial frame: the next one starts afresh
            } else ok = parser.parseTileGroup(br, d.frame, group);
            if (ok) {
                d.tiles.insert(d.tiles.end(), group.begin(), group.end());
                if (d.tiles.size() == parser.m_frame->NumTiles) {
                    ok = d.decodeFrame(d.tiles);
                    d.tiles.clear();
                }
            }
            break;
        }
        case OBU_FRAME: {
            TileGroup group;
            const double tp = Impl::now();
            if (d.streaming && !d.async) {
                // Parser::parseFrame (Parser.cpp:1761-1772) with the tile group streamed
                d.frame = parser.parseFrameHeader(br);
                d.t_parse += Impl::now() - tp;
                if (d.frame) {
                    parser.skipTrailingBits(br);
                    d.streamed = d.beginFrame(d.frame, parser.m_sequence);
                    if (!d.streamed || !d.streamTileGroup(br, d.frame, group)) {
                        d.streamed = false;
                        return false;
                    }
                }
            } else {
                d.frame = parser.parseFrame(br, group);
                d.t_parse += Impl::now() - tp;
            }
            ok = d.frame ? d.decodeFrame(group) : false;
            d.tiles.clear();
            break;
        }
        case OBU_METADATA: ok = parser.parseMetadata(br); break;
        case OBU_PADDING: ok = parser.parsePadding(br); break;
        default: ok = parser.praseReserved(br); break;
        }
        if (!ok) return false;
        reader.skip(payload << 3);
    }
    return true;
}

std::shared_ptr<YuvFrame> Decoder::getOutput()
{
    Impl& d = *m_impl;
    d.drainWorker(); // reference semantics: every frame of the units decoded so far is available
    Impl::Pending p;
    {
        std::lock_guard<std::mutex> lk(d.mu);
        if (d.output.empty()) return nullptr;
        p = d.output.front();
        d.output.pop_front();
    }
    const double tw = Impl::now();
    const int wrc = av1b_fence_wait(d.ctx, p.fence);
    d.t_wait += Impl::now() - tw;
    if (wrc != AV1B_OK) {
        d.fail("av1b_fence_wait");
        return nullptr;
    }
    return p.frame;
}

}  // namespace YamiAv1

// ---- helpers for the C API / adapter ---------------------------------------------------------
namespace av1b200 {
static YamiAv1::Decoder::Impl* implOf(YamiAv1::Decoder& d) { return d.impl(); }
DecoderOptions& decoderOptions(YamiAv1::Decoder& d) { return implOf(d)->opt; }
const char* decoderError(YamiAv1::Decoder& d) { return implOf(d)->error.c_str(); }
av1b_ctx* decoderCtx(YamiAv1::Decoder& d) { return implOf(d)->ctx; }
bool decoderFormat(YamiAv1::Decoder& d, int& w, int& h)
{
    auto* i = implOf(d);
    if (!i->parser->m_sequence) return false;
    w = i->parser->m_sequence->max_frame_width_minus_1 + 1;
    h = i->parser->m_sequence->max_frame_height_minus_1 + 1;
    return true;
}
std::shared_ptr<Yami::YuvFrame> decoderPollOutput(YamiAv1::Decoder& d, size_t keepInFlight)
{
    auto* i = implOf(d);
    YamiAv1::Decoder::Impl::Pending p;
    {
        std::lock_guard<std::mutex> lk(i->mu);
        if (i->output.empty()) return nullptr;
        p = i->output.front();
        if (i->output.size() <= keepInFlight && !av1b_fence_done(i->ctx, p.fence)) return nullptr;
        i->output.pop_front();
    }
    if (av1b_fence_wait(i->ctx, p.fence) != AV1B_OK) return nullptr;
    return p.frame;
}
void decoderSetAsync(YamiAv1::Decoder& d, bool on) { implOf(d)->async = on; }
void decoderDrain(YamiAv1::Decoder& d) { implOf(d)->drainWorker(); }
void decoderFlush(YamiAv1::Decoder& d)
{
    auto* i = implOf(d);
    i->drainWorker();
    if (i->ctx) av1b_sync(i->ctx);
    std::lock_guard<std::mutex> lk(i->mu);
    i->output.clear();
}
}  // namespace av1b200

// yami_adapter.cpp -- the Yami-style IVideoDecoder the reference DECLARES but never implements
// (interface/VideoDecoderInterface.h:31-68, interface/VideoDecoderHost.h:32-34), implemented over
// the B200-backed YamiAv1::Decoder.  start/decode/getOutput as the north-star asks:
//   start()      -> YAMI_SUCCESS (the engine is created lazily from the sequence header)
//   decode(buf)  -> one temporal unit; YAMI_SUCCESS or YAMI_DECODE_INVALID_DATA
//   getOutput()  -> SharedPtr<VideoFrame>; surface = pointer to a host Yami::YuvFrame (I420),
//                   crop = visible area, fourcc = YAMI_FOURCC_I420; the frame is released when
//                   the VideoFrame is destroyed
#include <VideoDecoderHost.h>
#include "Av1Decoder.h"
#include "decoder_impl.h"
#include "VideoFrame.h"

#include <cstring>

namespace {

using namespace YamiMediaCodec;

class Av1B200VideoDecoder : public IVideoDecoder {
public:
    Av1B200VideoDecoder() { resetFormat(); }
    YamiStatus start(VideoConfigBuffer*) override
    {
        if (!m_dec) m_dec.reset(new YamiAv1::Decoder);
        return YAMI_SUCCESS;
    }
    YamiStatus reset(VideoConfigBuffer* b) override
    {
        stop();
        return start(b);
    }
    void stop() override
    {
        m_dec.reset();
        resetFormat();
    }
    void flush() override
    {
        if (m_dec) av1b200::decoderFlush(*m_dec);
    }
    YamiStatus decode(VideoDecodeBuffer* buffer) override
    {
        if (!m_dec) return YAMI_FAIL;
        if (!buffer || !buffer->data || !buffer->size) return YAMI_SUCCESS; // EOS marker
        m_pts = buffer->timeStamp;
        return m_dec->decode(buffer->data, buffer->size) ? YAMI_SUCCESS : YAMI_DECODE_INVALID_DATA;
    }
    SharedPtr<VideoFrame> getOutput() override
    {
        SharedPtr<VideoFrame> out;
        if (!m_dec) return out;
        std::shared_ptr<Yami::YuvFrame> yuv = m_dec->getOutput();
        if (!yuv) return out;
        // the holder keeps the pinned frame alive for as long as the VideoFrame exists
        auto* holder = new std::shared_ptr<Yami::YuvFrame>(yuv);
        VideoFrame* vf = new VideoFrame;
        memset(vf, 0, sizeof(*vf));
        vf->surface = (intptr_t)yuv.get();
        vf->timeStamp = m_pts;
        vf->crop.x = 0;
        vf->crop.y = 0;
        vf->crop.width = yuv->width;
        vf->crop.height = yuv->height;
        vf->fourcc = YAMI_FOURCC_I420;
        vf->user_data = (intptr_t)holder;
        out.reset(vf, [](VideoFrame* f) {
            delete (std::shared_ptr<Yami::YuvFrame>*)f->user_data;
            delete f;
        });
        return out;
    }
    const VideoFormatInfo* getFormatInfo() override
    {
        int w = 0, h = 0;
        if (m_dec && av1b200::decoderFormat(*m_dec, w, h)) {
            m_format.valid = true;
            m_format.width = m_format.surfaceWidth = w;
            m_format.height = m_format.surfaceHeight = h;
        }
        return &m_format;
    }
    void setNativeDisplay(NativeDisplay*) override {}
    void setAllocator(SurfaceAllocator*) override {}
    void releaseLock(bool) override {}

private:
    void resetFormat()
    {
        memset(&m_format, 0, sizeof(m_format));
        m_format.mimeType = const_cast<char*>(YAMI_MIME_AV1);
        m_format.fourcc = YAMI_FOURCC_I420;
    }
    std::unique_ptr<YamiAv1::Decoder> m_dec;
    VideoFormatInfo m_format;
    int64_t m_pts = 0;
};

}  // namespace

extern "C" {

YamiMediaCodec::IVideoDecoder* createVideoDecoder(const char* mimeType)
{
    if (!mimeType || strcmp(mimeType, YAMI_MIME_AV1) != 0) return nullptr;
    return new Av1B200VideoDecoder;
}

void releaseVideoDecoder(YamiMediaCodec::IVideoDecoder* p) { delete p; }

}

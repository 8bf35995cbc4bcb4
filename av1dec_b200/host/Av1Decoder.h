// Av1Decoder.h -- drop-in replacement for the reference's decoder/Av1Decoder.h.
//
// Same class name, namespace and public surface as the reference
// (decoder/Av1Decoder.h:47-51: Decoder(), ~Decoder(), decode(), getOutput()), so the reference's
// own CLI (tests/Av1Dec.cpp, DecodeInput.*, DecodeOutput.*) compiles and links UNCHANGED
// against this library.  The private part is different: frames live in B200 HBM, and the
// reconstruction / in-loop filters run as sm_100a kernels behind include/av1b200.h.
#pragma once

#include <memory>
#include <stddef.h>
#include <stdint.h>

namespace Yami {
struct YuvFrame;
}

namespace YamiAv1 {

class Decoder {
public:
    // One temporal unit (a sequence of OBUs).  false on a parse or device failure.
    bool decode(uint8_t* data, size_t size);
    // Next shown frame in output order (host I420, valid until the pointer is dropped), or null.
    std::shared_ptr<Yami::YuvFrame> getOutput();
    Decoder();
    ~Decoder();

    struct Impl;
    Impl* impl() const { return m_impl.get(); } // engine-side extras (av1b200_decoder.h), not part of the reference API

private:
    Decoder(const Decoder&) = delete;
    Decoder& operator=(const Decoder&) = delete;
    std::unique_ptr<Impl> m_impl;
};

}  // namespace YamiAv1

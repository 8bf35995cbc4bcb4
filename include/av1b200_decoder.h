/*
 * av1b200_decoder.h -- C ABI of the whole-stream decoder (host front end + B200 engine).
 *
 * A thin C surface over the drop-in C++ class YamiAv1::Decoder (av1dec_b200/host/Av1Decoder.h,
 * replacing the reference's decoder/Av1Decoder.h:47-51) for callers without a C++ ABI: the
 * Python tests / benchmark bind it with ctypes, and it is what a cgo/JNI/N-API host would bind.
 *
 *   reference call (tests/Av1Dec.cpp:199-221)        this ABI
 *   YamiAv1::Decoder decoder;                        av1b_decoder_create()
 *   decoder.decode(buf.data, buf.size)               av1b_decoder_decode()
 *   decoder.getOutput()                              av1b_decoder_get_output()
 *   DecodeInputVPX + the loop in Decode::run()       av1b_decode_ivf()   (one call per stream)
 */
#ifndef AV1B200_DECODER_H_
#define AV1B200_DECODER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct av1b_decoder av1b_decoder;

/* Per-frame command-buffer tap (benchmark: record command buffers for HBM-resident replay).
 * data == NULL means show_existing_frame of store slot `bytes`. */
typedef void (*av1b_cmd_sink)(void* user, const uint8_t* data, size_t bytes, uint32_t refresh_mask, int show);

av1b_decoder* av1b_decoder_create(int device);
void av1b_decoder_destroy(av1b_decoder* d);
/* Restrict the device stages run per frame (AV1B_STAGE_* mask, default all). Test hook. */
void av1b_decoder_set_stages(av1b_decoder* d, uint32_t stages);
void av1b_decoder_set_cmd_sink(av1b_decoder* d, av1b_cmd_sink sink, void* user);
/* Decode one temporal unit. 0 on success, -1 on failure (av1b_decoder_error() has the text). */
int av1b_decoder_decode(av1b_decoder* d, const uint8_t* data, size_t size);
/* Pop the next shown frame. Returns 1 and fills the out-params (pointers stay valid until the
 * next call on this decoder), 0 when no frame is pending, -1 on error. */
int av1b_decoder_get_output(av1b_decoder* d, int* width, int* height, const uint8_t* planes[3], int strides[3]);
const char* av1b_decoder_error(av1b_decoder* d);

/* Decode a complete IVF byte stream.  When out_yuv != NULL the shown frames are written back to
 * back as visible-area I420 (the layout the reference CLI writes and bits.md5 covers,
 * tests/DecodeOutput.cpp:48-69); *out_bytes receives the size.  Returns 0, or -1 on failure,
 * or -2 when out_cap is too small (out_bytes then holds the required size so far). */
int av1b_decode_ivf(const uint8_t* ivf, size_t len, int device, uint32_t stages, uint8_t* out_yuv, size_t out_cap,
    size_t* out_bytes, int* n_frames, uint64_t* luma_pixels);

/* Closed segments of an IVF stream: a segment starts at every temporal unit that carries a
 * sequence header followed by a shown key frame (a random access point: all reference slots and
 * CDFs reset there).  Writes the temporal-unit index of each segment start into seg_first (up to
 * cap entries) and returns the number of segments, 0 for a stream with no temporal unit, -1 for a
 * buffer that is not IVF.  av1b_decode_ivf decodes the segments in parallel (one front end +
 * device context per worker, AV1B200_GOP_THREADS workers, default min(8, host threads)) and
 * stitches the output back in stream order. */
int av1b_ivf_segments(const uint8_t* ivf, size_t len, uint32_t* seg_first, int cap);

#ifdef __cplusplus
}
#endif
#endif /* AV1B200_DECODER_H_ */

/*
 * av1b200.h -- C ABI of the B200 (sm_100a) AV1 reconstruction + in-loop-filter engine.
 *
 * This is the drop-in boundary for the reference's pixel path.  The reference (oddstone/av1dec)
 * runs, per frame,
 *     Tile::decode()                 decoder/Tile.cpp:172        block reconstruction
 *     Decoder::decode_frame_wrapup() decoder/Av1Decoder.cpp:171  deblock -> CDEF -> loop restoration
 *     Decoder::updateFrameStore()    decoder/Av1Decoder.cpp:111  reference refresh
 *     Decoder::getOutput()           decoder/Av1Decoder.cpp:203  shown-frame hand-off
 * on host memory.  Here the host front end (the reference's own parser / entropy decoder)
 * serialises what those calls read into one command buffer per frame (av1b200_format.h)
 * and the functions below replace them; frames and reference frames stay in HBM.
 *
 * Plain pointers and sizes only; no CUDA, torch or C++ types cross this boundary.
 * Every function returns 0 on success or a negative AV1B_E* code; av1b_last_error() gives text.
 * A context is single-threaded (like the reference's Decoder); contexts are independent and
 * many may share one GPU.
 */
#ifndef AV1B200_H_
#define AV1B200_H_

#include <stddef.h>
#include <stdint.h>
#include "av1b200_format.h"

#ifdef __cplusplus
extern "C" {
#endif

#define AV1B_OK 0
#define AV1B_EINVAL (-1)
#define AV1B_ECUDA (-2)
#define AV1B_ENOMEM (-3)
#define AV1B_ESTATE (-4)

/* stage bits for av1b_frame_submit() */
#define AV1B_STAGE_ITX 1u     /* TransformBlock::inverseTransform (TransformBlock.cpp:2173)   */
#define AV1B_STAGE_INTER 2u   /* InterPredict::predict_inter (InterPredict.cpp:962)           */
#define AV1B_STAGE_WAVE 4u    /* intra predict + residual add (TransformBlock.cpp:2376)       */
#define AV1B_STAGE_DEBLOCK 8u /* LoopFilter::filter (LoopFilter.cpp:40)                       */
#define AV1B_STAGE_CDEF 16u   /* Cdef::filter (Cdef.cpp:41)                                   */
#define AV1B_STAGE_LR 32u     /* LoopRestoration::filter (LoopRestoration.cpp:191)            */
#define AV1B_STAGE_RECON (AV1B_STAGE_ITX | AV1B_STAGE_INTER | AV1B_STAGE_WAVE)
#define AV1B_STAGE_POST (AV1B_STAGE_DEBLOCK | AV1B_STAGE_CDEF | AV1B_STAGE_LR)
#define AV1B_STAGE_ALL (AV1B_STAGE_RECON | AV1B_STAGE_POST)

typedef struct av1b_ctx av1b_ctx;

/* Library identity: "cuda-sm_100a" for the product, "emu" for the test-only host emulation. */
const char* av1b_backend(void);

/* Create an engine for frames up to max_w x max_h luma samples on CUDA device `device`.
 * `stream` is a cudaStream_t to enqueue on, or NULL to let the engine create its own.
 * Replaces the YuvFrame allocations of Decoder::decodeFrame (Av1Decoder.cpp:131). */
int av1b_ctx_create(av1b_ctx** out, int device, int max_w, int max_h, void* stream);
/* Destroying a context that owns its stream returns it (frame pool, pinned ring and all) to a
 * process-wide cache that av1b_ctx_create() draws from; av1b_pool_purge() really frees. */
void av1b_ctx_destroy(av1b_ctx* ctx);
void av1b_pool_purge(void);
const char* av1b_last_error(av1b_ctx* ctx);

/* Pinned command ring.  Returns a host pointer with room for `bytes`; blocks until the slot's
 * previous frame has been consumed by the device. */
int av1b_cmd_acquire(av1b_ctx* ctx, size_t bytes, void** host_ptr);

/* Reconstruct + filter the frame whose command buffer was written into the pointer returned
 * by the last av1b_cmd_acquire().  Asynchronous.  `stages` is a mask of AV1B_STAGE_*.
 * On return *frame_id names the device-resident output frame (valid until released by the
 * reference store).  refresh_mask: FrameHeader::refresh_frame_flags (Av1Decoder.cpp:115). */
int av1b_frame_submit(av1b_ctx* ctx, size_t bytes, uint32_t stages, uint32_t refresh_mask, int* frame_id);

/* Same, but the command buffer already lives in device memory (`dev_cmd`), `hdr` is a host
 * copy of its header.  Used to time the kernels with inputs resident in HBM. */
int av1b_frame_submit_resident(av1b_ctx* ctx, const void* dev_cmd, const Av1bFrameHdr* hdr, uint32_t stages,
    uint32_t refresh_mask, int* frame_id);

/* show_existing_frame (Decoder::showExistingFrame, Av1Decoder.cpp:158): returns the frame in
 * store slot `slot` and applies refresh_mask. */
int av1b_show_existing(av1b_ctx* ctx, int slot, uint32_t refresh_mask, int* frame_id);

/* Frame ids returned by submit / show_existing are pool indices below this bound. */
#define AV1B_MAX_FRAME_IDS 96

/* Asynchronous copy of the visible w x h (and chroma) area of a device frame to host planes
 * (pinned memory from av1b_host_alloc gives a true async copy).  Decoder::getOutput(). */
int av1b_frame_download(av1b_ctx* ctx, int frame_id, uint8_t* const dst[3], const int dst_stride[3], int w, int h);
/* Zero-copy output for a GPU consumer (the reference's VideoFrame.surface hook,
 * interface/VideoCommonDefs.h:257-283): device pointers and pitches of the three planes of a frame.
 * The context stream is made to wait for the frame; work the caller enqueues on that stream (the
 * one passed to av1b_ctx_create) afterwards sees finished samples.  The planes stay valid until the
 * frame leaves the reference store and the pool reuses it: hold it with av1b_frame_retain /
 * av1b_frame_release for longer. */
int av1b_frame_device_view(av1b_ctx* ctx, int frame_id, const uint8_t* planes[3], int pitches[3]);
int av1b_frame_retain(av1b_ctx* ctx, int frame_id);
int av1b_frame_release(av1b_ctx* ctx, int frame_id);
/* The visible area as NV12 (luma plane, then interleaved U/V rows) into device memory the caller
 * owns: dst_y / dst_uv with their pitches, on the context stream.  For consumers (encoders,
 * display surfaces) that take semi-planar input. */
int av1b_frame_to_nv12(av1b_ctx* ctx, int frame_id, uint8_t* dst_y, int pitch_y, uint8_t* dst_uv, int pitch_uv, int w, int h);
/* Block until everything enqueued so far (kernels and copies) has finished. */
int av1b_sync(av1b_ctx* ctx);
/* Frames are reconstructed on internal streams ("lanes", AV1B200_LANES, default 4) so that frames
 * that do not depend on each other overlap; downloads and fences run on the context stream and
 * wait for the frames they need.  av1b_join makes the context stream wait (on the device, without
 * blocking the host) for every frame submitted so far, and orders the next submitted frame behind
 * whatever the caller enqueues on that stream afterwards: call it before recording your own event
 * on a stream passed to av1b_ctx_create. */
int av1b_join(av1b_ctx* ctx);
/* Number of lanes of this context from now on (1..24; waits for the context to go idle first).
 * 1 makes every frame run alone on the device -- what a per-kernel timing wants. */
int av1b_set_lanes(av1b_ctx* ctx, int n);
/* CUDA-graph capture of a resident replay: with a stream of the caller's passed to av1b_ctx_create,
 * call av1b_set_capture(ctx, 1) (synchronises, so that nothing recorded earlier has to be waited
 * for), begin the capture on that stream, submit the frames (av1b_frame_submit_resident /
 * av1b_show_existing), av1b_join, end the capture, av1b_set_capture(ctx, 0).  While it is on the
 * engine issues no call that is illegal during capture (no event query, no allocation: run the
 * same sequence once before so that every buffer exists).  The lanes fork from and join back into
 * the captured stream through events, so the graph keeps the frame-level concurrency. */
int av1b_set_capture(av1b_ctx* ctx, int on);
/* Mark a fence after the work enqueued so far / wait for it: lets a caller overlap parsing of
 * the next frame with this frame's device work. */
int av1b_fence_record(av1b_ctx* ctx, uint64_t* fence);
int av1b_fence_wait(av1b_ctx* ctx, uint64_t fence);
int av1b_fence_done(av1b_ctx* ctx, uint64_t fence); /* 1 if the fence has been reached (non-blocking) */

void* av1b_host_alloc(size_t bytes); /* pinned host memory */
void av1b_host_free(void* p);
/* raw device memory helpers for the resident-command benchmark path */
void* av1b_dev_alloc(size_t bytes);
void av1b_dev_free(void* p);
int av1b_dev_upload(av1b_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes);
/* Synchronous copy of device memory (context stream order) to the host: NV12 / device-view tests. */
int av1b_dev_download(av1b_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes);

/* Test / stage-level entry points --------------------------------------------------------- */
/* Load host planes (MI-aligned area: w x h luma samples given) into the frame that the next
 * av1b_frame_submit() will treat as "current" -- lets a test or benchmark run the post-filter
 * stages (or a residual-only pass) on arbitrary input without a bitstream. */
int av1b_debug_set_input(av1b_ctx* ctx, const uint8_t* const src[3], const int src_stride[3], int w, int h);
/* Load host planes into reference store slot `slot`. */
int av1b_debug_set_ref(av1b_ctx* ctx, int slot, const uint8_t* const src[3], const int src_stride[3], int w, int h);
/* Copy the int16 residual arena of the last submitted frame to host (n values). */
int av1b_debug_get_residual(av1b_ctx* ctx, int16_t* dst, size_t n);
/* Number of kernel launches issued by this context so far. */
uint64_t av1b_launch_count(av1b_ctx* ctx);
/* Process-wide allocation counters: contexts created, contexts recycled from the pool, device
 * allocations, pinned host allocations.  A steady-state service should stop moving [0], [2], [3]. */
void av1b_debug_counters(uint64_t out[4]);
/* Per-stage device timing.  With profiling on, the engine brackets every launch group with CUDA
 * events on its stream; av1b_get_stage_times() synchronises and returns accumulated milliseconds
 * and call counts per stage (index: 0 itx, 1 inter, 2 wavefront, 3 deblock, 4 cdef, 5 lr). */
#define AV1B_N_STAGES 6
int av1b_set_profiling(av1b_ctx* ctx, int on);
int av1b_get_stage_times(av1b_ctx* ctx, double ms[AV1B_N_STAGES], uint64_t calls[AV1B_N_STAGES], int reset);
/* Device-to-device: make a copy of reference slot `slot` the "current" frame of the next submit
 * (benchmark: re-run the in-place post-filter chain on pristine, HBM-resident input). */
int av1b_debug_input_from_slot(av1b_ctx* ctx, int slot);
/* Wavefront trace (profiling aid).  av1b_debug_wave_trace(n) makes every later wavefront launch of
 * the process record, per superblock, eight 64-bit words into a device buffer of n entries
 * (n = 0 turns it off): superblock index, SM id, and %globaltimer nanoseconds at ticket taken,
 * dependencies satisfied, first level started, last level finished, neighbours signalled, tile
 * flushed.  av1b_debug_wave_trace_read copies the first n entries to the host (device-synchronous). */
int av1b_debug_wave_trace(size_t n_entries);
int av1b_debug_wave_trace_read(uint64_t* out, size_t n_entries);
/* sizeof() of the command-format structs (0 FrameHdr, 1 Op, 2 Sb, 3 Ipu, 4 InterBlk, 5 BlkAux,
 * 6 LfMi, 7 LrUnit): lets a foreign-language binding verify its mirror of av1b200_format.h. */
size_t av1b_struct_size(int which);

#ifdef __cplusplus
}
#endif
#endif /* AV1B200_H_ */

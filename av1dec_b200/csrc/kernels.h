// kernels.h -- internal interface between the engine (engine.cu) and the kernel files.
#pragma once
#include "dev.h"
#include "../../include/av1b200_format.h"

// Everything the reconstruction kernels need, passed by value as a kernel argument.
struct ReconCtx {
    const uint8_t* cmd;   // device copy of the frame command buffer
    FrameView cur;        // frame being reconstructed
    FrameView ref[8];     // reference frames by frame-store slot
    int16_t* res;         // compact residual arena (stage-level ITX test mode; used when rp[0] == null)
    int16_t* rp[3];       // residual planes in frame layout (int16), zero where no coded TB
    int rpitch[3];        // elements per row
    uint8_t* mask;        // luma-resolution scratch plane for diff-weighted compound masks
    int mask_pitch;
    const uint8_t* wedge; // wedge mask table [9][2][16][32*32]
    int* sync;            // [0] = SB ticket counter, [1 + r] = finished SBs of SB row r
    unsigned long long* trace; // profiling aid (av1b_debug_wave_trace), normally null
    unsigned trace_cap;
};

// In-loop filter context.
struct PostCtx {
    const uint8_t* cmd;
    FrameView src;  // reconstructed frame (never modified by the filters)
    FrameView deb;  // deblocked frame (== src when the deblocking stage does not run)
    FrameView cdef; // CDEF output
    FrameView lr;   // loop-restoration output
};

void launch_itx(const ReconCtx& c, const Av1bFrameHdr& h, av1b_stream_t st);
void launch_inter(const ReconCtx& c, const Av1bFrameHdr& h, av1b_stream_t st);
void launch_wave(const ReconCtx& c, const Av1bFrameHdr& h, av1b_stream_t st);

void launch_deblock(const PostCtx& c, const Av1bFrameHdr& h, av1b_stream_t st);
void launch_cdef(const PostCtx& c, const Av1bFrameHdr& h, av1b_stream_t st);
int launch_lr(const PostCtx& c, const Av1bFrameHdr& h, av1b_stream_t st); // returns the number of kernels launched
// planar 4:2:0 -> NV12 (device to device), visible w x h
void launch_to_nv12(const FrameView& src, uint8_t* dst_y, int pitch_y, uint8_t* dst_uv, int pitch_uv, int w, int h, av1b_stream_t st);

// Builds the wedge mask table (host memory, 9*2*16*1024 bytes).
void build_wedge_table(uint8_t* out);
#define AV1B_WEDGE_TABLE_BYTES (9 * 2 * 16 * 1024)

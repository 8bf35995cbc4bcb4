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
    int* sync;            // [0] = SB ticket counter, [1] = CTAs that left the kernel, [2 + r] = finished SBs of SB row r;
                          // all zero between launches (the last CTA out resets them)
    unsigned long long* trace; // profiling aid (av1b_debug_wave_trace), normally null
    unsigned trace_cap;
};

// The frame-header fields the filter kernels read, BY VALUE: a kernel argument lives in the constant
// bank, so a CTA starts on its tile without a dependent global load of the header first.
struct PostHdr {
    uint16_t frame_w, frame_h, mi_cols, mi_rows;
    uint32_t off_lfmi, off_cdef8, off_lru;
    Av1bLoopFilterParams lf;
    Av1bCdefParams cdef;
    Av1bLrParams lr;
};
inline PostHdr make_post_hdr(const Av1bFrameHdr& h)
{
    PostHdr p;
    p.frame_w = h.frame_w, p.frame_h = h.frame_h, p.mi_cols = h.mi_cols, p.mi_rows = h.mi_rows;
    p.off_lfmi = h.off_lfmi, p.off_cdef8 = h.off_cdef8, p.off_lru = h.off_lru;
    p.lf = h.lf, p.cdef = h.cdef, p.lr = h.lr;
    return p;
}

// In-loop filter context.
struct PostCtx {
    // TMA descriptors of the CDEF input planes (the whole padded plane as a 2-D byte tensor, box =
    // the CDEF tile with its halo); tma_ok == 0: not available, the kernel stages with loads
    Av1bTensorMap cdef_in[3];
    int tma_ok;
    int tma_x0, tma_y0; // tensor coordinates of plane sample (0, 0)
    PostHdr h;
    const uint8_t* cmd;
    FrameView src;  // reconstructed frame (never modified by the filters)
    FrameView deb;  // deblocked frame (== src when the deblocking stage does not run)
    FrameView cdef; // CDEF output
    FrameView lr;   // loop-restoration output
};

int launch_itx(const ReconCtx& c, const Av1bFrameHdr& h, av1b_stream_t st); // returns the number of kernels launched
void launch_inter(const ReconCtx& c, const Av1bFrameHdr& h, av1b_stream_t st);
void launch_wave(const ReconCtx& c, const Av1bFrameHdr& h, av1b_stream_t st);

void launch_deblock(const PostCtx& c, const Av1bFrameHdr& h, av1b_stream_t st);
void launch_cdef(const PostCtx& c, const Av1bFrameHdr& h, av1b_stream_t st);
// CDEF tile boxes the descriptors are encoded with: {bytes per row, rows} for luma / chroma
void cdef_tile_box(int plane, int* box_w, int* box_h);
int launch_lr(const PostCtx& c, const Av1bFrameHdr& h, av1b_stream_t st); // returns the number of kernels launched
// planar 4:2:0 -> NV12 (device to device), visible w x h
void launch_to_nv12(const FrameView& src, uint8_t* dst_y, int pitch_y, uint8_t* dst_uv, int pitch_uv, int w, int h, av1b_stream_t st);

// Builds the wedge mask table (host memory, 9*2*16*1024 bytes).
void build_wedge_table(uint8_t* out);
#define AV1B_WEDGE_TABLE_BYTES (9 * 2 * 16 * 1024)

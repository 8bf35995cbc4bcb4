// postfilter.cu -- output conversion (the in-loop filters live in deblock.cu, cdef.cu, lr.cu).
#include "dev.h"
#include "kernels.h"

// Planar 4:2:0 to NV12: luma rows copied, chroma rows interleaved U0 V0 U1 V1 ...  One thread per
// 4 luma samples / 2 chroma pairs; pure streaming (3 * w * h bytes moved).
__global__ void __launch_bounds__(256) nv12_kernel(FrameView src, uint8_t* dst_y, int pitch_y, uint8_t* dst_uv, int pitch_uv, int w, int h)
{
    const int cw = w >> 1, ch = h >> 1;
    const int lw4 = (w + 3) >> 2, cw2 = (cw + 1) >> 1;
    const int n_luma = lw4 * h, n_chroma = cw2 * ch;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_luma + n_chroma; t += gridDim.x * blockDim.x) {
        if (t < n_luma) {
            const int y = t / lw4, x = (t - y * lw4) * 4;
            const uint8_t* s = src.pl[0].p + (size_t)y * src.pl[0].stride + x;
            uint8_t* d = dst_y + (size_t)y * pitch_y + x;
            for (int k = 0; k < 4 && x + k < w; k++) d[k] = s[k];
        } else {
            const int u = t - n_luma;
            const int y = u / cw2, x = (u - y * cw2) * 2;
            const uint8_t* su = src.pl[1].p + (size_t)y * src.pl[1].stride + x;
            const uint8_t* sv = src.pl[2].p + (size_t)y * src.pl[2].stride + x;
            uint8_t* d = dst_uv + (size_t)y * pitch_uv + 2 * x;
            for (int k = 0; k < 2 && x + k < cw; k++) {
                d[2 * k] = su[k];
                d[2 * k + 1] = sv[k];
            }
        }
    }
}

void launch_to_nv12(const FrameView& src, uint8_t* dst_y, int pitch_y, uint8_t* dst_uv, int pitch_uv, int w, int h, av1b_stream_t st)
{
    const long long n = (long long)((w + 3) >> 2) * h + (long long)(((w >> 1) + 1) >> 1) * (h >> 1);
    if (n <= 0) return;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    AV1B_LAUNCH(nv12_kernel, (grid), (256), st, src, dst_y, pitch_y, dst_uv, pitch_uv, w, h);
}

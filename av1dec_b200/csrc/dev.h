// dev.h -- one source, two builds.
//
//  * nvcc (sm_100a): the product.  Kernels run on the GPU.
//  * g++ -DAV1B_EMU: a TEST-ONLY host emulation used to debug bit-exactness in a container
//    without a GPU.  Every kernel is written block-size agnostic (strided loops over
//    threadIdx/blockDim, __syncthreads between dependent phases), so running each CTA as a
//    single sequential "thread" is a valid schedule.  The emulation library is built under
//    tests/emu/ and is never loaded by the product path (av1dec_b200/__init__.py refuses it).
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef AV1B_EMU
// ------------------------------------------------------------------ host emulation
#include <algorithm>
#include <cstdlib>
#include <cstring>
struct EmuDim3 {
    unsigned x, y, z;
    EmuDim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef EmuDim3 dim3;
struct uint4 {
    unsigned x, y, z, w;
};
struct uint2 {
    unsigned x, y;
};
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{ x, y }; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{ x, y, z, w }; }
extern EmuDim3 threadIdx, blockIdx, blockDim, gridDim;
#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__
#define AV1B_UNROLL
#define AV1B_UNROLL4
#define AV1B_NOUNROLL
static inline void __syncthreads() {}
static inline void __syncwarp() {}
static inline void __threadfence() {}
static inline void __threadfence_block() {}
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
static inline int av1b_ld_acquire(const int* p) { return *p; }
static inline void av1b_st_release(int* p, int v) { *p = v; }
static inline void av1b_nanosleep(unsigned) {}
static inline int av1b_ld_relaxed(const int* p) { return *p; }
static inline unsigned long long av1b_gtime() { return 0; }
static inline unsigned av1b_smid() { return 0; }
using std::max;
using std::min;
// single-lane "warp": shuffles and votes see only lane 0
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
#define AV1B_NOINLINE
#define AV1B_ASSUME_SHARED(p) ((void)0)
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s)
{
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline uint32_t __vmaxu2(uint32_t a, uint32_t b)
{
    uint32_t lo = std::max(a & 0xFFFFu, b & 0xFFFFu), hi = std::max(a >> 16, b >> 16);
    return lo | (hi << 16);
}
static inline uint32_t __vminu2(uint32_t a, uint32_t b)
{
    uint32_t lo = std::min(a & 0xFFFFu, b & 0xFFFFu), hi = std::min(a >> 16, b >> 16);
    return lo | (hi << 16);
}
// ---- packed 16x2 / 8x4 SIMD intrinsics used by the CDEF kernel (single SASS instructions on
// sm_100a: VABSDIFF4, VIADD.16x2, VIMNMX[3].x16x2, VIADDMNMX.S16x2[.RELU])
static inline uint32_t emu_map2s(uint32_t a, uint32_t b, uint32_t c, int (*f)(int, int, int))
{
    const int lo = f((int16_t)(a & 0xFFFF), (int16_t)(b & 0xFFFF), (int16_t)(c & 0xFFFF));
    const int hi = f((int16_t)(a >> 16), (int16_t)(b >> 16), (int16_t)(c >> 16));
    return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16);
}
static inline uint32_t emu_map2u(uint32_t a, uint32_t b, uint32_t c, int (*f)(int, int, int))
{
    const int lo = f((int)(a & 0xFFFF), (int)(b & 0xFFFF), (int)(c & 0xFFFF));
    const int hi = f((int)(a >> 16), (int)(b >> 16), (int)(c >> 16));
    return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16);
}
static inline uint32_t __vabsdiffu4(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const int x = (a >> (8 * i)) & 0xFF, y = (b >> (8 * i)) & 0xFF;
        r |= (uint32_t)(x > y ? x - y : y - x) << (8 * i);
    }
    return r;
}
static inline uint32_t __vadd2(uint32_t a, uint32_t b) { return emu_map2s(a, b, 0, [](int x, int y, int) { return x + y; }); }
static inline uint32_t __vmaxs2(uint32_t a, uint32_t b) { return emu_map2s(a, b, 0, [](int x, int y, int) { return x > y ? x : y; }); }
static inline uint32_t __vmins2(uint32_t a, uint32_t b) { return emu_map2s(a, b, 0, [](int x, int y, int) { return x < y ? x : y; }); }
static inline uint32_t __viaddmin_s16x2(uint32_t a, uint32_t b, uint32_t c)
{
    return emu_map2s(a, b, c, [](int x, int y, int z) { const int s = (int16_t)(x + y); return s < z ? s : z; });
}
static inline uint32_t __viaddmax_s16x2(uint32_t a, uint32_t b, uint32_t c)
{
    return emu_map2s(a, b, c, [](int x, int y, int z) { const int s = (int16_t)(x + y); return s > z ? s : z; });
}
static inline uint32_t __viaddmin_s16x2_relu(uint32_t a, uint32_t b, uint32_t c)
{
    return emu_map2s(a, b, c, [](int x, int y, int z) { const int s = (int16_t)(x + y); const int m = s < z ? s : z; return m > 0 ? m : 0; });
}
static inline uint32_t __vimax3_u16x2(uint32_t a, uint32_t b, uint32_t c)
{
    return emu_map2u(a, b, c, [](int x, int y, int z) { return std::max(x, std::max(y, z)); });
}
static inline uint32_t __vimin3_u16x2(uint32_t a, uint32_t b, uint32_t c)
{
    return emu_map2u(a, b, c, [](int x, int y, int z) { return std::min(x, std::min(y, z)); });
}
static inline int av1b_dp4a_ss(uint32_t a, uint32_t b, int c)
{
    for (int i = 0; i < 4; i++) c += (int)(int8_t)((a >> (8 * i)) & 0xFF) * (int)(int8_t)((b >> (8 * i)) & 0xFF);
    return c;
}
// bulk asynchronous copies: the emulation copies at once, the barrier is always complete
static inline void av1b_mbar_init(unsigned long long*, unsigned) {}
static inline void av1b_mbar_expect_tx(unsigned long long*, unsigned) {}
static inline void av1b_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
static inline void av1b_mbar_wait(unsigned long long*, unsigned) {}
struct Av1bTensorMap;
static inline void av1b_tma_load_2d(void*, const Av1bTensorMap*, int, int, unsigned long long*) {}
struct alignas(64) Av1bTensorMap { unsigned long long opaque[16]; }; // CUtensorMap stand-in (never used by the emulation)
typedef void* av1b_stream_t;
template <class F> static inline void emu_launch(dim3 grid, F f)
{
    gridDim = grid;
    blockDim = EmuDim3(1, 1, 1);
    threadIdx = EmuDim3(0, 0, 0);
    for (unsigned z = 0; z < grid.z; z++)
        for (unsigned y = 0; y < grid.y; y++)
            for (unsigned x = 0; x < grid.x; x++) {
                blockIdx = EmuDim3(x, y, z);
                f();
            }
}
#define AV1B_LAUNCH(kern, grid, block, stream, ...) emu_launch(dim3 grid, [&] { kern(__VA_ARGS__); })
#define AV1T_CONST static const
#else
// ------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>
typedef cudaStream_t av1b_stream_t;
#define AV1B_UNROLL _Pragma("unroll")
#define AV1B_UNROLL4 _Pragma("unroll 4")
#define AV1B_NOUNROLL _Pragma("unroll 1")
#define AV1B_NOINLINE __noinline__
// address-space hint: lets the compiler turn generic accesses through p into LDS / STS
#define AV1B_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#define AV1B_LAUNCH(kern, grid, block, stream, ...) kern<<<dim3 grid, dim3 block, 0, stream>>>(__VA_ARGS__)
#define AV1T_CONST static __device__ const
static __device__ __forceinline__ int av1b_ld_acquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __forceinline__ void av1b_st_release(int* p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
static __device__ __forceinline__ void av1b_nanosleep(unsigned ns) { __nanosleep(ns); }
// ---- bulk asynchronous global -> shared copies (the TMA unit moves whole rows; SASS UBLKCP), completion
// counted in bytes on a shared-memory mbarrier.  Source, destination and size are multiples of 16.
static __device__ __forceinline__ void av1b_mbar_init(unsigned long long* bar, unsigned arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
static __device__ __forceinline__ void av1b_mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void av1b_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
// A 2-D tile (the box the descriptor was encoded with) at element coordinates (x, y) of the tensor:
// ONE request to the TMA unit (SASS UTMALDG), rows land packed in shared memory.
struct alignas(64) Av1bTensorMap { unsigned long long opaque[16]; }; // same size / alignment as CUtensorMap
static __device__ __forceinline__ void av1b_tma_load_2d(void* dst, const Av1bTensorMap* map, int x, int y, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)),
                 "l"(map), "r"(x), "r"(y), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
static __device__ __forceinline__ void av1b_mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done)
                     : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}
// polling load: served by L2 every time, no L1 invalidation (the acquire follows once the value is there)
static __device__ __forceinline__ int av1b_ld_relaxed(const int* p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __forceinline__ unsigned long long av1b_gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
static __device__ __forceinline__ unsigned av1b_smid()
{
    unsigned v;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
    return v;
}
#endif

// ------------------------------------------------------------------ shared helpers
#define AV1B_DEV static __device__ __forceinline__
#define AV1B_DEV_M __device__ __forceinline__ // member functions

// Barrier for a CTA of nt threads: a single-warp CTA (the latency-critical intra wavefront runs
// one warp per superblock) only needs warp-level ordering.
AV1B_DEV void block_sync(int nt)
{
    if (nt > 32) __syncthreads();
    else __syncwarp();
}

// Integer dot products (IDP.4A / IDP.2A).  a = four unsigned bytes (dp4a) or two signed 16-bit
// halves (dp2a); b = signed bytes; c = accumulator.
#ifdef AV1B_EMU
static inline int av1b_dp4a_us(uint32_t a, uint32_t b, int c)
{
    for (int i = 0; i < 4; i++) c += (int)((a >> (8 * i)) & 0xFF) * (int)(int8_t)((b >> (8 * i)) & 0xFF);
    return c;
}
static inline int av1b_dp2a_lo(uint32_t a, uint32_t b, int c)
{
    return c + (int)(int16_t)(a & 0xFFFF) * (int)(int8_t)(b & 0xFF) + (int)(int16_t)(a >> 16) * (int)(int8_t)((b >> 8) & 0xFF);
}
static inline int av1b_dp2a_hi(uint32_t a, uint32_t b, int c)
{
    return c + (int)(int16_t)(a & 0xFFFF) * (int)(int8_t)((b >> 16) & 0xFF) + (int)(int16_t)(a >> 16) * (int)(int8_t)((b >> 24) & 0xFF);
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
    sh &= 31;
    return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
}
static inline uint32_t av1b_dp4a_uu(uint32_t a, uint32_t b, uint32_t c)
{
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xFF) * ((b >> (8 * i)) & 0xFF);
    return c;
}
// d = (c << 16) | (sat_u8(a) << 8) | sat_u8(b)   (PTX cvt.pack.sat.u8.s32.b32)
static inline uint32_t av1b_pack_sat_u8(int a, int b, uint32_t c)
{
    const uint32_t ta = a < 0 ? 0 : (a > 255 ? 255 : a), tb = b < 0 ? 0 : (b > 255 ? 255 : b);
    return (c << 16) | (ta << 8) | tb;
}
#else
static __device__ __forceinline__ int av1b_dp4a_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
static __device__ __forceinline__ int av1b_dp4a_ss(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
static __device__ __forceinline__ uint32_t av1b_dp4a_uu(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// d = (c << 16) | (sat_u8(a) << 8) | sat_u8(b)   (I2IP.U8.S32.SAT)
static __device__ __forceinline__ uint32_t av1b_pack_sat_u8(int a, int b, uint32_t c)
{
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
static __device__ __forceinline__ int av1b_dp2a_lo(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
static __device__ __forceinline__ int av1b_dp2a_hi(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
#endif

// Four values, each saturated to 8 bits, packed little-endian (v0 in the low byte).
AV1B_DEV uint32_t pack_u8x4(int v0, int v1, int v2, int v3) { return av1b_pack_sat_u8(v1, v0, av1b_pack_sat_u8(v3, v2, 0u)); }

AV1B_DEV int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
AV1B_DEV int clip_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
AV1B_DEV int round2(int x, int n) { return n == 0 ? x : ((x + (1 << (n - 1))) >> n); }
// Four packed samples plus four int16 residuals, each clipped to 8 bits: bytes widened to 16x2,
// add + min(255) + max(0) per pair in one instruction (VIADDMNMX.S16x2.RELU), bytes gathered back.
// (Residuals of a conformant stream are 8 + BitDepth bits at most: the 16-bit add cannot wrap.)
AV1B_DEV uint32_t add_res4(uint32_t px, uint2 r)
{
    const uint32_t p01 = __byte_perm(px, 0u, 0x4140), p23 = __byte_perm(px, 0u, 0x4342);
    const uint32_t s01 = __viaddmin_s16x2_relu(p01, r.x, 0x00FF00FFu), s23 = __viaddmin_s16x2_relu(p23, r.y, 0x00FF00FFu);
    return __byte_perm(s01, s23, 0x6420);
}
AV1B_DEV int round2s(int x, int n) { return x >= 0 ? round2(x, n) : -round2(-x, n); }
AV1B_DEV int iabs(int v) { return v < 0 ? -v : v; }
AV1B_DEV int floor_log2(unsigned x)
{
    int s = -1;
    while (x) {
        x >>= 1;
        s++;
    }
    return s;
}

// Geometry of one device-resident plane.
struct PlaneView {
    uint8_t* p;  // address of sample (0,0)
    int stride;  // bytes between rows
};
struct FrameView {
    PlaneView pl[3];
};

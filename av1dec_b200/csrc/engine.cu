// engine.cu -- host side of the C ABI (include/av1b200.h): device frame pool, reference store,
// pinned command ring, stream plumbing and the per-frame launch sequence
//     H2D(command buffer) -> itx -> inter -> wavefront -> deblock(V,H) -> CDEF -> LR [-> D2H]
// Replaces Decoder::decodeFrame / decode_frame_wrapup / updateFrameStore of the reference
// (decoder/Av1Decoder.cpp:111-192).  No pixel is ever computed on the host here: when the
// CUDA runtime is unavailable every entry point fails with AV1B_ECUDA.
#include "kernels.h"
#include "../../include/av1b200.h"
#include "av1_tables_host.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

// allocation counters (av1b_debug_counters): a steady-state decode service should stop moving them
static unsigned long long* g_wave_trace = nullptr; // av1b_debug_wave_trace
static size_t g_wave_trace_cap = 0;
static std::atomic<uint64_t> g_n_ctx_new{ 0 }, g_n_ctx_reused{ 0 }, g_n_dev_alloc{ 0 }, g_n_pinned_alloc{ 0 };

// ------------------------------------------------------------------------------------------
// thin runtime layer (CUDA, or libc for the test-only emulation build)
// ------------------------------------------------------------------------------------------
#ifdef AV1B_EMU
EmuDim3 threadIdx, blockIdx, blockDim, gridDim;
typedef int rt_event_t;
static int rt_set_device(int) { return 0; }
static int rt_malloc(void** p, size_t n) { g_n_dev_alloc++; *p = calloc(1, n ? n : 1); return *p ? 0 : 1; }
static void rt_free(void* p) { free(p); }
static int rt_host_alloc(void** p, size_t n) { g_n_pinned_alloc++; *p = malloc(n ? n : 1); return *p ? 0 : 1; }
static void rt_host_free(void* p) { free(p); }
static int rt_h2d(void* d, const void* s, size_t n, av1b_stream_t) { memcpy(d, s, n); return 0; }
static int rt_d2h(void* d, const void* s, size_t n, av1b_stream_t) { memcpy(d, s, n); return 0; }
static int rt_copy2d(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, av1b_stream_t, int)
{
    for (size_t i = 0; i < h; i++) memcpy((uint8_t*)d + i * dp, (const uint8_t*)s + i * sp, w);
    return 0;
}
static int rt_memset(void* d, int v, size_t n, av1b_stream_t) { memset(d, v, n); return 0; }
static int rt_stream_create(av1b_stream_t* s) { *s = nullptr; return 0; }
static void rt_stream_destroy(av1b_stream_t) {}
static int rt_stream_sync(av1b_stream_t) { return 0; }
static int rt_event_create(rt_event_t* e) { *e = 0; return 0; }
static int rt_event_create_host(rt_event_t* e) { *e = 0; return 0; }
static void rt_event_destroy(rt_event_t) {}
static int rt_event_record(rt_event_t, av1b_stream_t) { return 0; }
static int rt_event_sync(rt_event_t) { return 0; }
static int rt_stream_wait(av1b_stream_t, rt_event_t) { return 0; }
static int rt_event_done(rt_event_t) { return 1; }
static const char* rt_error() { return "emu"; }
static int rt_check() { return 0; }
static int rt_tevent_create(rt_event_t* e) { *e = 0; return 0; }
static float rt_event_ms(rt_event_t, rt_event_t) { return 0.f; }
static int rt_d2d(void* d, const void* s, size_t n, av1b_stream_t) { memcpy(d, s, n); return 0; }
static int rt_device_sync() { return 0; }
static int rt_d2h_sync(void* d, const void* s, size_t n) { memcpy(d, s, n); return 0; }
const char* av1b_backend(void) { return "emu"; }
#else
typedef cudaEvent_t rt_event_t;
static int rt_set_device(int d) { return cudaSetDevice(d) != cudaSuccess; }
// Device memory comes from the device's stream-ordered pool (cudaMallocAsync on a stream of its own,
// release threshold = never): once the pool is warm an allocation in the middle of a decode is a
// pointer bump, where cudaMalloc / cudaFree synchronise the WHOLE device and stall every other
// decoder of the process for as long as its kernels take to drain.  The allocation stream carries
// nothing else, so synchronising it right away makes the block usable on any stream.  Callers free
// only what no stream uses any more (they synchronise first), as cudaFree required.
static cudaStream_t rt_alloc_stream()
{
    static std::mutex mu;
    static cudaStream_t streams[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    std::lock_guard<std::mutex> lk(mu);
    if (!streams[dev]) {
        cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    return streams[dev];
}
static int rt_malloc(void** p, size_t n)
{
    g_n_dev_alloc++;
    cudaStream_t s = rt_alloc_stream();
    if (cudaMallocAsync(p, n ? n : 1, s) != cudaSuccess) {
        cudaGetLastError();
        return cudaMalloc(p, n ? n : 1) != cudaSuccess; // pool exhausted / unsupported: the blocking path
    }
    return cudaStreamSynchronize(s) != cudaSuccess;
}
static void rt_free(void* p)
{
    if (p && cudaFreeAsync(p, rt_alloc_stream()) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(p);
    }
}
static int rt_host_alloc(void** p, size_t n) { g_n_pinned_alloc++; return cudaHostAlloc(p, n ? n : 1, cudaHostAllocDefault) != cudaSuccess; }
static void rt_host_free(void* p) { if (p) cudaFreeHost(p); }
static int rt_h2d(void* d, const void* s, size_t n, av1b_stream_t st) { return cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st) != cudaSuccess; }
static int rt_d2h(void* d, const void* s, size_t n, av1b_stream_t st) { return cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, st) != cudaSuccess; }
static int rt_copy2d(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, av1b_stream_t st, int to_host)
{
    return cudaMemcpy2DAsync(d, dp, s, sp, w, h, to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice, st) != cudaSuccess;
}
static int rt_memset(void* d, int v, size_t n, av1b_stream_t st) { return cudaMemsetAsync(d, v, n, st) != cudaSuccess; }
static int rt_stream_create(av1b_stream_t* s) { return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking) != cudaSuccess; }
static void rt_stream_destroy(av1b_stream_t s) { cudaStreamDestroy(s); }
static int rt_stream_sync(av1b_stream_t s) { return cudaStreamSynchronize(s) != cudaSuccess; }
static int rt_event_create(rt_event_t* e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess; }
// Events a HOST thread waits on (command-slot reuse, output fences): the waiter sleeps instead of
// spinning -- a decode service runs more host threads than cores (callers, emit workers, segment
// workers), and a core spent polling the device is taken from the parser.  AV1B200_SPIN_WAIT=1
// restores the spinning wait (lowest wake-up latency for a lone stream).
static int rt_event_create_host(rt_event_t* e)
{
    static const bool spin = [] {
        const char* v = getenv("AV1B200_SPIN_WAIT");
        return v && atoi(v) != 0;
    }();
    return cudaEventCreateWithFlags(e, cudaEventDisableTiming | (spin ? 0 : cudaEventBlockingSync)) != cudaSuccess;
}
static void rt_event_destroy(rt_event_t e)
{
    if (e) cudaEventDestroy(e);
}
static int rt_event_record(rt_event_t e, av1b_stream_t s) { return cudaEventRecord(e, s) != cudaSuccess; }
static int rt_event_sync(rt_event_t e) { return cudaEventSynchronize(e) != cudaSuccess; }
static int rt_stream_wait(av1b_stream_t s, rt_event_t e) { return cudaStreamWaitEvent(s, e, 0) != cudaSuccess; }
static int rt_event_done(rt_event_t e) { return cudaEventQuery(e) == cudaSuccess; }
static const char* rt_error() { return cudaGetErrorString(cudaGetLastError()); }
static int rt_check() { return cudaGetLastError() != cudaSuccess; }
static int rt_tevent_create(rt_event_t* e) { return cudaEventCreate(e) != cudaSuccess; }
static float rt_event_ms(rt_event_t a, rt_event_t b)
{
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}
static int rt_d2d(void* d, const void* s, size_t n, av1b_stream_t st) { return cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st) != cudaSuccess; }
static int rt_device_sync() { return cudaDeviceSynchronize() != cudaSuccess; }
static int rt_d2h_sync(void* d, const void* s, size_t n) { return cudaMemcpy(d, s, n, cudaMemcpyDeviceToHost) != cudaSuccess; }
const char* av1b_backend(void) { return "cuda-sm_100a"; }
#endif

// ------------------------------------------------------------------------------------------
// TMA descriptors.  cuTensorMapEncodeTiled is a driver entry point: fetched at run time so the
// library does not link libcuda.
// ------------------------------------------------------------------------------------------
#ifndef AV1B_EMU
#include <cuda.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tma_encoder()
{
    static std::once_flag once;
    static EncodeTiledFn fn = nullptr;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (getenv("AV1B200_NO_TMA") == nullptr && cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess
            && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else cudaGetLastError();
    });
    return fn;
}
// The padded plane (rows x pitch bytes starting at `base`) as a 2-D byte tensor with a box_w x box_h box.
static bool tma_encode_plane(Av1bTensorMap* out, uint8_t* base, size_t pitch, size_t rows, int box_w, int box_h)
{
    static_assert(sizeof(Av1bTensorMap) == sizeof(CUtensorMap) && alignof(Av1bTensorMap) >= alignof(CUtensorMap), "descriptor stand-in");
    EncodeTiledFn fn = tma_encoder();
    if (!fn) return false;
    const cuuint64_t dims[2] = { (cuuint64_t)pitch, (cuuint64_t)rows };
    const cuuint64_t strides[1] = { (cuuint64_t)pitch };
    const cuuint32_t box[2] = { (cuuint32_t)box_w, (cuuint32_t)box_h };
    const cuuint32_t estr[2] = { 1, 1 };
    return fn((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
        == CUDA_SUCCESS;
}
#else
static bool tma_encode_plane(Av1bTensorMap*, uint8_t*, size_t, size_t, int, int) { return false; }
#endif

// ------------------------------------------------------------------------------------------
// wedge mask table (spec 7.11.3.11; reference initialise_wedge_mask_table, InterPredict.cpp:835)
// ------------------------------------------------------------------------------------------
void build_wedge_table(uint8_t* out)
{
    enum { HORZ, VERT, OB27, OB63, OB117, OB153, NDIR };
    static uint8_t master[NDIR][64][64];
    for (int j = 0; j < 64; j++) {
        int shift = 16;
        for (int i = 0; i < 64; i += 2) {
            master[OB63][i][j] = hk_wedge_master_even[std::min(63, std::max(0, j - shift))];
            shift -= 1;
            master[OB63][i + 1][j] = hk_wedge_master_odd[std::min(63, std::max(0, j - shift))];
            master[VERT][i][j] = hk_wedge_master_vert[j];
            master[VERT][i + 1][j] = hk_wedge_master_vert[j];
        }
    }
    for (int i = 0; i < 64; i++)
        for (int j = 0; j < 64; j++) {
            int m = master[OB63][i][j];
            master[OB27][j][i] = (uint8_t)m;
            master[OB117][i][63 - j] = (uint8_t)(64 - m);
            master[OB153][63 - j][i] = (uint8_t)(64 - m);
            master[HORZ][j][i] = master[VERT][i][j];
        }
    static const int sizes[9] = { 3, 4, 5, 6, 7, 8, 9, 18, 19 }; // BLOCK_SIZE values that have wedges
    memset(out, 0, AV1B_WEDGE_TABLE_BYTES);
    for (int s = 0; s < 9; s++) {
        const int bs = sizes[s];
        const int w = hk_block_w[bs], h = hk_block_h[bs];
        const int shape = h > w ? 0 : (h < w ? 1 : 2);
        for (int wedge = 0; wedge < 16; wedge++) {
            const int dir = hk_wedge_codebook[shape][wedge][0];
            const int xoff = 32 - ((hk_wedge_codebook[shape][wedge][1] * w) >> 3);
            const int yoff = 32 - ((hk_wedge_codebook[shape][wedge][2] * h) >> 3);
            int sum = 0;
            for (int i = 0; i < w; i++) sum += master[dir][yoff][xoff + i];
            for (int i = 1; i < h; i++) sum += master[dir][yoff + i][xoff];
            const int avg = (sum + (w + h - 1) / 2) / (w + h - 1);
            const int flip = avg < 32;
            uint8_t* t0 = out + ((size_t)((s * 2 + flip) * 16 + wedge)) * 1024;
            uint8_t* t1 = out + ((size_t)((s * 2 + !flip) * 16 + wedge)) * 1024;
            for (int i = 0; i < h; i++)
                for (int j = 0; j < w; j++) {
                    const int m = master[dir][yoff + i][xoff + j];
                    t0[i * 32 + j] = (uint8_t)m;
                    t1[i * 32 + j] = (uint8_t)(64 - m);
                }
        }
    }
}

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
namespace {
enum { N_SLOTS = 8, N_FENCES = 64, PAD_X = 128, PAD_Y = 16, POOL_MAX = AV1B_MAX_FRAME_IDS, MAX_LANES = 24, MAIN_LANE = MAX_LANES };

// Frames are reconstructed on LANES: internal streams a context deals its frames to round-robin,
// so that frames with no dependency between them (intra-only frames, frames of different
// temporal layers) overlap on the GPU instead of queueing behind each other -- the superblock
// wavefront of a small frame fills a handful of SMs only.  Ordering comes from events: a frame
// waits for the `ready` event of every reference it reads, and a buffer taken from the pool waits
// for the lanes that last wrote or read it.  MAIN_LANE stands for the context's own stream
// (downloads, debug uploads).
struct DevFrame {
    uint8_t* base = nullptr;
    bool owned = true;        // own cudaMalloc (false: carved out of the context's slab)
    FrameView v;
    int refcnt = 0;
    rt_event_t ready = rt_event_t(); // recorded after the last kernel of the submit that wrote (or used) the frame
    rt_event_t copied = rt_event_t(); // recorded on the context stream after the last download / copy out of it
    int writer = MAIN_LANE;   // lane of that submit
    uint32_t readers = 0;     // lanes that read it since (bit MAIN_LANE = the context stream)
    bool settled = true;      // nothing recorded since the last full av1b_sync: no event to wait for
    Av1bTensorMap tm_cdef[3];  // TMA descriptors of the padded planes (CDEF tile boxes)
    bool tm_ok = false;
};

// Per-lane scratch: nothing here is shared between frames in flight on different lanes.
struct Lane {
    av1b_stream_t stream = nullptr;
    rt_event_t mark = rt_event_t(); // scratch event for cross-lane ordering
    int16_t* res_planes = nullptr; // frame-layout residual planes (luma aw x ah, chroma aw/2 x ah/2 each)
    uint8_t* mask_plane = nullptr; // luma-resolution compound-mask scratch (aw x ah)
    int* sync = nullptr;
    size_t sync_cap = 0;
};

struct CmdSlot {
    uint8_t* host = nullptr;
    uint8_t* dev = nullptr;
    size_t cap = 0;
    rt_event_t done = rt_event_t();
    bool pending = false;
};
}  // namespace

struct av1b_ctx {
    int device = 0;
    av1b_stream_t stream = nullptr;
    bool own_stream = false;
    int max_w = 0, max_h = 0, aw = 0, ah = 0;
    int stride_y = 0, stride_c = 0;
    size_t frame_bytes = 0;
    std::vector<DevFrame> frames;
    int ref_slot[8];
    CmdSlot slots[N_SLOTS];
    int cur_slot = -1;
    int16_t* res = nullptr; // compact residual arena (stage-level ITX test mode, lane 0 only)
    size_t res_cap = 0;
    Lane lanes[MAX_LANES];
    int n_lanes = 1;
    int lanes_made = 0; // lanes whose stream and event exist
    uint8_t* slab = nullptr; // one allocation holding the lanes' working set of frame buffers
    uint64_t frame_seq = 0;
    bool joined = true;     // no lane work outstanding relative to the context stream
    bool capturing = false; // the caller is capturing the context stream into a CUDA graph: no event queries
    rt_event_t main_mark = rt_event_t();
    uint8_t* wedge = nullptr;
    int pending_input = -1;
    rt_event_t fences[N_FENCES] = {};
    std::atomic<uint64_t> fence_next{ 1 }; // read by the caller's thread while the emit worker records
    uint64_t launches = 0;
    std::string err;
    // optional per-stage device timing (CUDA events around each launch group)
    bool profiling = false;
    struct Span {
        int stage;
        rt_event_t a, b;
    };
    std::vector<Span> spans;            // recorded, not yet resolved
    std::vector<rt_event_t> event_pool; // reusable timing events
    double stage_ms[AV1B_N_STAGES] = { 0 };
    uint64_t stage_calls[AV1B_N_STAGES] = { 0 };
};

namespace {
struct StageTimer {
    av1b_ctx* c;
    int stage;
    rt_event_t a, b;
    bool on;
    static rt_event_t get(av1b_ctx* c)
    {
        if (!c->event_pool.empty()) {
            rt_event_t e = c->event_pool.back();
            c->event_pool.pop_back();
            return e;
        }
        rt_event_t e;
        rt_tevent_create(&e);
        return e;
    }
    av1b_stream_t stream;
    StageTimer(av1b_ctx* ctx, int st, bool active, av1b_stream_t str)
        : c(ctx), stage(st), on(active && ctx->profiling), stream(str)
    {
        if (!on) return;
        a = get(c);
        b = get(c);
        rt_event_record(a, stream);
    }
    ~StageTimer()
    {
        if (!on) return;
        rt_event_record(b, stream);
        c->spans.push_back(av1b_ctx::Span{ stage, a, b });
    }
};
}  // namespace

static int fail(av1b_ctx* c, int code, const char* what)
{
    if (c) {
        c->err = what;
        if (code == AV1B_ECUDA) {
            c->err += ": ";
            c->err += rt_error();
        }
    }
    return code;
}

// Working set of the lanes: 4 buffers per frame in flight (reconstruction, deblocked, CDEF, LR) + 8 references + 1 pending output.
static size_t pool_soft_cap(const av1b_ctx* c) { return std::min<size_t>(POOL_MAX, (size_t)4 * c->n_lanes + 9); }

// Register one frame buffer (planes laid out inside `base`) with the pool.
static int frame_add(av1b_ctx* c, uint8_t* base, bool owned, int lane)
{
    DevFrame f;
    f.base = base;
    f.owned = owned;
    const size_t luma_rows = (size_t)c->ah + 2 * PAD_Y, chroma_rows = (size_t)c->ah / 2 + 2 * PAD_Y;
    uint8_t* y = f.base;
    uint8_t* u = y + luma_rows * c->stride_y;
    uint8_t* v = u + chroma_rows * c->stride_c;
    f.v.pl[0].p = y + (size_t)PAD_Y * c->stride_y + PAD_X;
    f.v.pl[0].stride = c->stride_y;
    f.v.pl[1].p = u + (size_t)PAD_Y * c->stride_c + PAD_X;
    f.v.pl[1].stride = c->stride_c;
    f.v.pl[2].p = v + (size_t)PAD_Y * c->stride_c + PAD_X;
    f.v.pl[2].stride = c->stride_c;
    {
        int bw, bh;
        f.tm_ok = true;
        uint8_t* bases[3] = { y, u, v };
        for (int p = 0; p < 3; p++) {
            cdef_tile_box(p, &bw, &bh);
            f.tm_ok = f.tm_ok && tma_encode_plane(&f.tm_cdef[p], bases[p], p ? c->stride_c : c->stride_y, p ? chroma_rows : luma_rows, bw, bh);
        }
    }
    if (rt_event_create(&f.ready)) return 1;
    if (rt_event_create(&f.copied)) {
        rt_event_destroy(f.ready);
        return 1;
    }
    f.writer = lane;
    c->frames.push_back(f);
    return 0;
}

static av1b_stream_t lane_stream(av1b_ctx* c, int lane) { return lane == MAIN_LANE ? c->stream : c->lanes[lane].stream; }

// A free frame buffer for a writer on `lane`.  Preference: a buffer only this lane touched (no
// wait at all); a buffer whose last writer has finished and that no other lane read (the waits
// frame_claim adds are on completed events); a new buffer while the pool may grow; any free one.
static int frame_alloc(av1b_ctx* c, int lane = MAIN_LANE)
{
    const uint32_t mine = (1u << lane) | (1u << MAIN_LANE);
    int idle = -1, any = -1;
    for (size_t i = 0; i < c->frames.size(); i++) {
        const DevFrame& f = c->frames[i];
        if (f.refcnt) continue;
        if (f.writer == lane && !(f.readers & ~(1u << lane))) return (int)i;
        if (idle < 0 && !(f.readers & ~mine) && (f.settled || (!c->capturing && rt_event_done(f.ready)))) idle = (int)i;
        if (any < 0) any = (int)i;
    }
    if (idle >= 0) return idle;
    if (c->capturing && any >= 0) return any; // no allocation while a graph is being captured
    // growing the pool costs a cudaMalloc (device-wide synchronisation): past the working set of
    // the lanes (3 buffers per frame in flight + 8 references + 1 pending output) reuse instead
    const size_t soft_cap = pool_soft_cap(c);
    if (c->frames.size() >= POOL_MAX || (any >= 0 && (c->n_lanes == 1 || c->frames.size() >= soft_cap))) return any;
    void* p = nullptr;
    if (rt_malloc(&p, c->frame_bytes)) return -1;
    if (frame_add(c, (uint8_t*)p, true, lane)) {
        rt_free(p);
        return -1;
    }
    return (int)c->frames.size() - 1;
}

// `lane` is about to WRITE frame f: order it after the lanes that wrote or read the buffer.
static int frame_claim(av1b_ctx* c, int f, int lane)
{
    DevFrame& fr = c->frames[f];
    av1b_stream_t st = lane_stream(c, lane);
    // exact: the submit that last wrote it, the last copy out of it on the context stream
    if (!fr.settled && fr.writer != lane && rt_stream_wait(st, fr.ready)) return 1;
    if ((fr.readers & (1u << MAIN_LANE)) && lane != MAIN_LANE && rt_stream_wait(st, fr.copied)) return 1;
    // conservative: everything queued so far on the other lanes that read it as a reference
    const uint32_t others = fr.readers & ~((1u << lane) | (1u << MAIN_LANE));
    for (int m = 0; m < MAX_LANES; m++) {
        if (!(others & (1u << m))) continue;
        if (rt_event_record(c->lanes[m].mark, c->lanes[m].stream) || rt_stream_wait(st, c->lanes[m].mark)) return 1;
    }
    fr.readers = 0;
    fr.writer = lane;
    fr.settled = false; // the claimer records `ready` when its submit ends
    return 0;
}

// `lane` is about to READ frame f.
static int frame_read(av1b_ctx* c, int f, int lane)
{
    DevFrame& fr = c->frames[f];
    if (!fr.settled && fr.writer != lane && rt_stream_wait(lane_stream(c, lane), fr.ready)) return 1;
    fr.readers |= 1u << lane;
    return 0;
}

// Make the context stream wait for everything queued on the lanes.  `rejoin`: also order the
// lanes' NEXT frame behind whatever the caller enqueues on the context stream from here on.
static int join_lanes(av1b_ctx* c, bool rejoin)
{
    if (c->joined) return 0;
    for (int m = 0; m < c->n_lanes; m++)
        if (rt_event_record(c->lanes[m].mark, c->lanes[m].stream) || rt_stream_wait(c->stream, c->lanes[m].mark)) return 1;
    c->joined = rejoin;
    return 0;
}

extern "C" {

// ---- process-wide caches -------------------------------------------------------------------
// Creating a context costs a stream, ~70 events, several cudaMalloc / cudaHostAlloc calls and a
// table upload; cudaFree synchronises the whole device.  A multi-stream decode service opens and
// closes one decoder per stream, so contexts (with their frame pools and pinned rings), the wedge
// table and pinned output buffers are recycled instead of being freed.
static std::mutex g_mu;
static std::vector<av1b_ctx*> g_ctx_pool;

static std::map<int, uint8_t*> g_wedge;           // per device
static std::map<void*, size_t> g_pinned_size;     // every live pinned block -> its bucket size
static std::map<size_t, std::vector<void*>> g_pinned_free;
static std::map<std::pair<int, size_t>, std::vector<void*>> g_dev_free; // (device, bucket) -> free device blocks

static size_t bucket_of(size_t bytes)
{
    size_t bucket = 64 << 10;
    while (bucket < bytes) bucket <<= 1;
    return bucket;
}

// Device blocks for the command ring, recycled process-wide in power-of-two buckets like the
// pinned ones: contexts are handed from stream to stream, and a ring slot that had to grow for a
// bigger frame would otherwise cost a cudaMalloc + cudaHostAlloc in the middle of a decode.
static void* dev_bucket_alloc(int device, size_t bucket)
{
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto& fl = g_dev_free[std::make_pair(device, bucket)];
        if (!fl.empty()) {
            void* p = fl.back();
            fl.pop_back();
            return p;
        }
    }
    void* p = nullptr;
    return rt_malloc(&p, bucket) ? nullptr : p;
}
static void dev_bucket_free(int device, size_t bucket, void* p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_mu);
    g_dev_free[std::make_pair(device, bucket)].push_back(p);
}

void* av1b_host_alloc(size_t bytes);
void av1b_host_free(void* p);

static void ctx_free(av1b_ctx* c)
{
    rt_set_device(c->device);
    if (c->stream) rt_stream_sync(c->stream);
    for (int m = 0; m < c->lanes_made; m++) rt_stream_sync(c->lanes[m].stream);
    for (auto& f : c->frames) {
        if (f.owned) rt_free(f.base);
        rt_event_destroy(f.ready);
        rt_event_destroy(f.copied);
    }
    rt_free(c->slab);
    for (int m = 0; m < c->lanes_made; m++) {
        Lane& L = c->lanes[m];
        rt_free(L.res_planes);
        rt_free(L.mask_plane);
        rt_free(L.sync);
        rt_event_destroy(L.mark);
        rt_stream_destroy(L.stream);
    }
    rt_event_destroy(c->main_mark);
    for (int i = 0; i < N_SLOTS; i++) {
        av1b_host_free(c->slots[i].host);
        dev_bucket_free(c->device, c->slots[i].cap, c->slots[i].dev);
        rt_event_destroy(c->slots[i].done);
    }
    for (int i = 0; i < N_FENCES; i++) rt_event_destroy(c->fences[i]);
    for (auto e : c->event_pool) rt_event_destroy(e);
    rt_free(c->res);
    if (c->own_stream && c->stream) rt_stream_destroy(c->stream);
    delete c;
}

void av1b_pool_purge(void);
static std::string g_last_create_error;

static int ctx_init(av1b_ctx* c, int device, int max_w, int max_h, int aw, int ah, void* stream)
{
    c->device = device;
    for (int i = 0; i < 8; i++) c->ref_slot[i] = -1;
    if (rt_set_device(device)) return fail(c, AV1B_ECUDA, "cudaSetDevice");
    if (stream) c->stream = (av1b_stream_t)stream;
    else {
        if (rt_stream_create(&c->stream)) return fail(c, AV1B_ECUDA, "cudaStreamCreate");
        c->own_stream = true;
    }
    c->max_w = max_w;
    c->max_h = max_h;
    c->aw = aw;
    c->ah = ah;
    c->stride_y = c->aw + 2 * PAD_X;
    c->stride_c = c->aw / 2 + 2 * PAD_X;
    c->frame_bytes = (size_t)(c->ah + 2 * PAD_Y) * c->stride_y + 2 * (size_t)(c->ah / 2 + 2 * PAD_Y) * c->stride_c;
    {
        // AV1B200_LANES: frames in flight per context (default 4, 1 = strictly serial)
        const char* e = getenv("AV1B200_LANES");
        int n = e ? atoi(e) : 4;
        c->n_lanes = n < 1 ? 1 : (n > MAX_LANES ? MAX_LANES : n);
    }
    for (int m = 0; m < c->n_lanes; m++) {
        if (rt_stream_create(&c->lanes[m].stream) || rt_event_create(&c->lanes[m].mark)) return fail(c, AV1B_ECUDA, "lane stream");
        c->lanes_made = m + 1;
    }
    if (rt_event_create(&c->main_mark)) return fail(c, AV1B_ECUDA, "cudaEventCreate");
    {
        // the working set comes from ONE allocation made here: no cudaMalloc (a device-wide
        // synchronisation) in the middle of a decode; only a deeper pipeline grows the pool later
        const size_t n = pool_soft_cap(c);
        void* p = nullptr;
        if (rt_malloc(&p, c->frame_bytes * n)) return fail(c, AV1B_ENOMEM, "frame slab");
        c->slab = (uint8_t*)p;
        for (size_t i = 0; i < n; i++)
            if (frame_add(c, c->slab + i * c->frame_bytes, false, MAIN_LANE)) return fail(c, AV1B_ECUDA, "cudaEventCreate");
    }
    for (int i = 0; i < N_SLOTS; i++)
        if (rt_event_create_host(&c->slots[i].done)) return fail(c, AV1B_ECUDA, "cudaEventCreate");
    for (int i = 0; i < N_FENCES; i++)
        if (rt_event_create_host(&c->fences[i])) return fail(c, AV1B_ECUDA, "cudaEventCreate");
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_wedge.find(device);
        if (it == g_wedge.end()) {
            std::vector<uint8_t> wt(AV1B_WEDGE_TABLE_BYTES);
            build_wedge_table(wt.data());
            void* p = nullptr;
            if (rt_malloc(&p, AV1B_WEDGE_TABLE_BYTES)) return fail(c, AV1B_ENOMEM, "wedge table alloc");
            if (rt_h2d(p, wt.data(), AV1B_WEDGE_TABLE_BYTES, c->stream) || rt_stream_sync(c->stream))
                return fail(c, AV1B_ECUDA, "wedge table upload");
            it = g_wedge.emplace(device, (uint8_t*)p).first;
        }
        c->wedge = it->second;
    }
    return AV1B_OK;
}

int av1b_ctx_create(av1b_ctx** out, int device, int max_w, int max_h, void* stream)
{
    if (!out || max_w <= 0 || max_h <= 0 || max_w > 16384 || max_h > 16384) return AV1B_EINVAL;
#ifndef AV1B_EMU
    // Many short streams with event waits between them: with the default 8 hardware work queues
    // a wait at the head of a queue stalls unrelated streams behind it.  Only effective when the
    // CUDA context does not exist yet; hosts that initialise CUDA first set it themselves.
    static const int once = setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    (void)once;
#endif
    const int aw = (max_w + 127) & ~127, ah = (max_h + 127) & ~127;
    if (!stream) {
        std::lock_guard<std::mutex> lk(g_mu);
        for (size_t i = 0; i < g_ctx_pool.size(); i++) {
            av1b_ctx* c = g_ctx_pool[i];
            if (c->device == device && c->aw >= aw && c->ah >= ah && (size_t)c->aw * c->ah <= 2 * (size_t)aw * ah) {
                g_ctx_pool.erase(g_ctx_pool.begin() + i);
                g_n_ctx_reused++;
                *out = c;
                return AV1B_OK;
            }
        }
    }
    // A context is either fully built or not handed out at all: on failure the partial one is
    // freed (its handles are value-initialised, so ctx_free only destroys what exists) and *out
    // stays null.  Out of device memory: idle pooled contexts hold frame slabs -- purge them and
    // try once more before giving up.
    *out = nullptr;
    for (int attempt = 0; attempt < 2; attempt++) {
        av1b_ctx* c = new av1b_ctx;
        g_n_ctx_new++;
        const int rc = ctx_init(c, device, max_w, max_h, aw, ah, stream);
        if (rc == AV1B_OK) {
            *out = c;
            return AV1B_OK;
        }
        g_last_create_error = c->err;
        ctx_free(c);
        if (rc != AV1B_ENOMEM || attempt) return rc;
        av1b_pool_purge();
    }
    return AV1B_ENOMEM;
}

void av1b_ctx_destroy(av1b_ctx* c)
{
    if (!c) return;
    rt_set_device(c->device);
    for (int m = 0; m < c->n_lanes; m++) rt_stream_sync(c->lanes[m].stream);
    rt_stream_sync(c->stream);
    if (c->own_stream && c->wedge) {
        // recycle: reset the decode state, keep every allocation (everything is idle: no hazards)
        for (auto& f : c->frames) {
            f.refcnt = 0;
            f.readers = 0;
            f.settled = true;
        }
        c->joined = true;
        c->capturing = false;
        for (int i = 0; i < 8; i++) c->ref_slot[i] = -1;
        for (int i = 0; i < N_SLOTS; i++) c->slots[i].pending = false;
        c->cur_slot = -1; // the next stream starts on the slots that already own a buffer
        c->frame_seq = 0;
        c->pending_input = -1;
        c->profiling = false;
        for (auto& sp : c->spans) {
            c->event_pool.push_back(sp.a);
            c->event_pool.push_back(sp.b);
        }
        c->spans.clear();
        for (int i = 0; i < AV1B_N_STAGES; i++) {
            c->stage_ms[i] = 0;
            c->stage_calls[i] = 0;
        }
        c->err.clear();
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_ctx_pool.size() < 512) {
            g_ctx_pool.push_back(c);
            return;
        }
    }
    ctx_free(c);
}

void av1b_pool_purge(void)
{
    std::vector<av1b_ctx*> pool;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        pool.swap(g_ctx_pool);
    }
    for (av1b_ctx* c : pool) ctx_free(c);
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_pinned_free)
        for (void* p : kv.second) {
            g_pinned_size.erase(p);
            rt_host_free(p);
        }
    g_pinned_free.clear();
    for (auto& kv : g_dev_free) {
        rt_set_device(kv.first.first);
        for (void* p : kv.second) rt_free(p);
    }
    g_dev_free.clear();
}

const char* av1b_last_error(av1b_ctx* c) { return c ? c->err.c_str() : "null context"; }

int av1b_cmd_acquire(av1b_ctx* c, size_t bytes, void** host_ptr)
{
    if (!c || !host_ptr) return AV1B_EINVAL;
    rt_set_device(c->device);
    const int s = (c->cur_slot + 1) % N_SLOTS;
    CmdSlot& sl = c->slots[s];
    if (sl.pending) {
        if (rt_event_sync(sl.done)) return fail(c, AV1B_ECUDA, "slot wait");
        sl.pending = false;
    }
    if (sl.cap < bytes) {
        // the slot is idle (its event was waited for above): trade both halves for a bigger bucket
        av1b_host_free(sl.host);
        dev_bucket_free(c->device, sl.cap, sl.dev);
        sl.host = sl.dev = nullptr;
        sl.cap = 0;
        const size_t cap = bucket_of(bytes);
        void* h = av1b_host_alloc(cap);
        if (!h) return fail(c, AV1B_ENOMEM, "pinned command slot");
        void* d = dev_bucket_alloc(c->device, cap);
        if (!d) {
            av1b_host_free(h);
            return fail(c, AV1B_ENOMEM, "device command slot");
        }
        sl.host = (uint8_t*)h;
        sl.dev = (uint8_t*)d;
        sl.cap = cap;
    }
    c->cur_slot = s;
    *host_ptr = sl.host;
    return AV1B_OK;
}

static int submit_impl(av1b_ctx* c, int lane, const uint8_t* dev_cmd, const Av1bFrameHdr* hdr, uint32_t stages, uint32_t refresh_mask,
    int* frame_id)
{
    const Av1bFrameHdr& h = *hdr;
    if (h.magic != AV1B_MAGIC || h.version != AV1B_FORMAT_VERSION) return fail(c, AV1B_EINVAL, "bad command buffer magic/version");
    if (h.mi_cols * 4 > c->aw || h.mi_rows * 4 > c->ah || (h.sb_cols << h.sb_log2) > c->aw || (h.sb_rows << h.sb_log2) > c->ah)
        return fail(c, AV1B_EINVAL, "frame larger than the context");
    Lane& L = c->lanes[lane];
    av1b_stream_t st = L.stream;
    // scratch.  A lone ITX stage (stage-level test) writes the compact arena; any fuller submit
    // writes residuals into frame-layout int16 planes that the inter and dependent passes read.
    const bool arena_mode = (stages & AV1B_STAGE_RECON) == AV1B_STAGE_ITX;
    const size_t plane_elems = (size_t)c->aw * c->ah * 3 / 2;
    if (!arena_mode && h.n_itx) {
        // first coded residual of the stream: every lane gets its planes now (no cudaMalloc later,
        // whichever lane a frame lands on)
        for (int m = 0; m < c->n_lanes; m++) {
            void* p = nullptr;
            if (c->lanes[m].res_planes) continue;
            if (rt_malloc(&p, plane_elems * sizeof(int16_t))) return fail(c, AV1B_ENOMEM, "residual planes");
            c->lanes[m].res_planes = (int16_t*)p;
        }
    }
    if (arena_mode && h.n_res > c->res_cap) {
        for (int m = 0; m < c->n_lanes; m++) rt_stream_sync(c->lanes[m].stream);
        rt_stream_sync(c->stream);
        rt_free(c->res);
        c->res = nullptr;
        c->res_cap = 0;
        void* p = nullptr;
        size_t cap = (size_t)h.n_res + h.n_res / 2 + 4096;
        if (rt_malloc(&p, cap * sizeof(int16_t))) return fail(c, AV1B_ENOMEM, "residual arena");
        c->res = (int16_t*)p;
        c->res_cap = cap;
    }
    if (arena_mode) {
        // the arena is shared: order this lane behind everything queued on the others
        for (int m = 0; m < c->n_lanes; m++)
            if (m != lane && (rt_event_record(c->lanes[m].mark, c->lanes[m].stream) || rt_stream_wait(st, c->lanes[m].mark)))
                return fail(c, AV1B_ECUDA, "stream wait");
    }
    const size_t sync_need = 2 + (size_t)std::max<uint32_t>(h.n_sb, h.sb_rows); // ticket, exit counter, progress per superblock
    for (int m = 0; m < c->n_lanes; m++) { // all lanes together, like the residual planes
        Lane& Lm = c->lanes[m];
        if (sync_need <= Lm.sync_cap) continue;
        rt_stream_sync(Lm.stream);
        rt_free(Lm.sync);
        Lm.sync = nullptr;
        Lm.sync_cap = 0;
        void* p = nullptr;
        if (rt_malloc(&p, (sync_need + 64) * sizeof(int))) return fail(c, AV1B_ENOMEM, "sync buffer");
        // zeroed ONCE: every wavefront launch leaves its counters at zero again (the last CTA out
        // resets them), so a frame does not pay a memset node for them
        if (rt_memset(p, 0, (sync_need + 64) * sizeof(int), Lm.stream) || rt_stream_sync(Lm.stream)) return fail(c, AV1B_ECUDA, "memset");
        Lm.sync = (int*)p;
        Lm.sync_cap = sync_need + 64;
    }
    // frames: the one being written, then the references it reads
    int cur = c->pending_input;
    const bool preloaded = cur >= 0; // debug input: already holds samples written on the context stream
    c->pending_input = -1;
    if (cur < 0) cur = frame_alloc(c, lane);
    if (cur < 0) return fail(c, AV1B_ENOMEM, "frame pool exhausted");
    c->frames[cur].refcnt++;
    if (preloaded) {
        if (frame_read(c, cur, lane)) return fail(c, AV1B_ECUDA, "stream wait");
        c->frames[cur].readers = 0;
        c->frames[cur].writer = lane;
        c->frames[cur].settled = false;
    } else if (frame_claim(c, cur, lane)) return fail(c, AV1B_ECUDA, "stream wait");
    ReconCtx rc;
    memset(&rc, 0, sizeof(rc));
    rc.cmd = dev_cmd;
    rc.cur = c->frames[cur].v;
    for (int i = 0; i < 8; i++)
        if (c->ref_slot[i] >= 0) {
            rc.ref[i] = c->frames[c->ref_slot[i]].v;
            if (h.n_iblk && frame_read(c, c->ref_slot[i], lane)) return fail(c, AV1B_ECUDA, "stream wait");
        }
    rc.res = c->res;
    if (!arena_mode && h.n_itx) {
        rc.rp[0] = L.res_planes;
        rc.rp[1] = L.res_planes + (size_t)c->aw * c->ah;
        rc.rp[2] = rc.rp[1] + (size_t)(c->aw / 2) * (c->ah / 2);
        rc.rpitch[0] = c->aw;
        rc.rpitch[1] = rc.rpitch[2] = c->aw / 2;
        if (stages & AV1B_STAGE_ITX) {
            // zero only the area this frame can touch (SB-aligned rows of the luma plane + chroma)
            const size_t rows = (size_t)h.sb_rows << h.sb_log2;
            const size_t separate = rows * c->aw + 2 * (rows / 2) * (c->aw / 2);
            const size_t joined = (size_t)(rc.rp[2] - rc.rp[0]) + (rows / 2) * (c->aw / 2); // the planes are contiguous
            if (joined <= separate + separate / 4) {
                // one node instead of three: a small frame's cost is launches, not bytes
                if (rt_memset(rc.rp[0], 0, joined * sizeof(int16_t), st)) return fail(c, AV1B_ECUDA, "memset");
            } else if (rt_memset(rc.rp[0], 0, rows * c->aw * sizeof(int16_t), st)
                || rt_memset(rc.rp[1], 0, (rows / 2) * (c->aw / 2) * sizeof(int16_t), st)
                || rt_memset(rc.rp[2], 0, (rows / 2) * (c->aw / 2) * sizeof(int16_t), st))
                return fail(c, AV1B_ECUDA, "memset");
        }
    }
    if (h.n_iblk && (stages & AV1B_STAGE_INTER)) {
        {
            for (int m = 0; m < c->n_lanes; m++) {
                void* p = nullptr;
                if (c->lanes[m].mask_plane) continue;
                if (rt_malloc(&p, (size_t)c->aw * c->ah)) return fail(c, AV1B_ENOMEM, "mask plane");
                c->lanes[m].mask_plane = (uint8_t*)p;
            }
        }
        rc.mask = L.mask_plane;
        rc.mask_pitch = c->aw;
    }
    rc.wedge = c->wedge;
    rc.sync = L.sync;
    rc.trace = g_wave_trace;
    rc.trace_cap = (unsigned)g_wave_trace_cap;
    if (stages & AV1B_STAGE_ITX) {
        StageTimer t(c, 0, h.n_itx != 0, st);
        c->launches += launch_itx(rc, h, st);
    }
    if (stages & AV1B_STAGE_INTER) {
        StageTimer t(c, 1, h.n_iblk != 0, st);
        launch_inter(rc, h, st);
        c->launches += h.n_iblk ? 1 : 0;
    }
    if (stages & AV1B_STAGE_WAVE) {
        StageTimer t(c, 2, h.n_ops != 0, st);
        launch_wave(rc, h, st);
        c->launches += h.n_ops ? 1 : 0;
    }
    PostCtx pc;
    memset(&pc, 0, sizeof(pc));
    pc.cmd = dev_cmd;
    pc.h = make_post_hdr(h);
    pc.src = c->frames[cur].v;
    int final_frame = cur, deb = -1, cdef = -1, lr = -1;
    pc.deb = pc.src;
    if ((stages & AV1B_STAGE_DEBLOCK) && (h.lf.level[0] || h.lf.level[1])) {
        // out of place: both edge passes read the reconstruction and write the deblocked frame
        deb = frame_alloc(c, lane);
        if (deb < 0) return fail(c, AV1B_ENOMEM, "frame pool exhausted");
        c->frames[deb].refcnt++;
        if (frame_claim(c, deb, lane)) return fail(c, AV1B_ECUDA, "stream wait");
        pc.deb = c->frames[deb].v;
        StageTimer t(c, 3, true, st);
        launch_deblock(pc, h, st);
        c->launches += 1;
        final_frame = deb;
    }
    pc.cdef = pc.deb;
    if ((stages & AV1B_STAGE_CDEF) && h.cdef.enabled) {
        cdef = frame_alloc(c, lane);
        if (cdef < 0) return fail(c, AV1B_ENOMEM, "frame pool exhausted");
        c->frames[cdef].refcnt++;
        if (frame_claim(c, cdef, lane)) return fail(c, AV1B_ECUDA, "stream wait");
        pc.cdef = c->frames[cdef].v;
        {
            const DevFrame& in = c->frames[deb >= 0 ? deb : cur]; // what the kernel reads as pc.deb
            pc.tma_ok = in.tm_ok;
            pc.tma_x0 = PAD_X;
            pc.tma_y0 = PAD_Y;
            for (int p = 0; p < 3; p++) pc.cdef_in[p] = in.tm_cdef[p];
        }
        StageTimer t(c, 4, true, st);
        launch_cdef(pc, h, st);
        c->launches += 1;
        final_frame = cdef;
    }
    if ((stages & AV1B_STAGE_LR) && h.lr.uses_lr) {
        lr = frame_alloc(c, lane);
        if (lr < 0) return fail(c, AV1B_ENOMEM, "frame pool exhausted");
        c->frames[lr].refcnt++;
        if (frame_claim(c, lr, lane)) return fail(c, AV1B_ECUDA, "stream wait");
        pc.lr = c->frames[lr].v;
        StageTimer t(c, 5, true, st);
        c->launches += launch_lr(pc, h, st);
        final_frame = lr;
    }
    if (rt_check()) return fail(c, AV1B_ECUDA, "kernel launch");
    // one completion point for every buffer this submit wrote or used as an intermediate
    {
        const int used[4] = { cur, deb, cdef, lr };
        for (int k = 0; k < 4; k++)
            if (used[k] >= 0 && rt_event_record(c->frames[used[k]].ready, st)) return fail(c, AV1B_ECUDA, "event record");
    }
    c->joined = false;
    // reference refresh (Decoder::updateFrameStore)
    for (int i = 0; i < 8; i++) {
        if (refresh_mask & (1u << i)) {
            if (c->ref_slot[i] >= 0) c->frames[c->ref_slot[i]].refcnt--;
            c->ref_slot[i] = final_frame;
            c->frames[final_frame].refcnt++;
        }
    }
    c->frames[cur].refcnt--;
    if (deb >= 0) c->frames[deb].refcnt--;
    if (cdef >= 0) c->frames[cdef].refcnt--;
    if (lr >= 0) c->frames[lr].refcnt--;
    if (frame_id) *frame_id = final_frame;
    return AV1B_OK;
}

// Lane of the next frame.  The first frame after a join also orders every lane behind the work
// already queued on the context stream (the caller's own stream, when one was supplied).
static int next_lane(av1b_ctx* c)
{
    if (c->joined) {
        if (rt_event_record(c->main_mark, c->stream)) return -1;
        for (int m = 0; m < c->n_lanes; m++)
            if (rt_stream_wait(c->lanes[m].stream, c->main_mark)) return -1;
    }
    return (int)(c->frame_seq++ % (uint64_t)c->n_lanes);
}

int av1b_frame_submit(av1b_ctx* c, size_t bytes, uint32_t stages, uint32_t refresh_mask, int* frame_id)
{
    if (!c || c->cur_slot < 0) return AV1B_ESTATE;
    rt_set_device(c->device);
    CmdSlot& sl = c->slots[c->cur_slot];
    if (bytes > sl.cap || bytes < sizeof(Av1bFrameHdr)) return fail(c, AV1B_EINVAL, "command size");
    const int lane = next_lane(c);
    if (lane < 0) return fail(c, AV1B_ECUDA, "lane");
    if (rt_h2d(sl.dev, sl.host, bytes, c->lanes[lane].stream)) return fail(c, AV1B_ECUDA, "command upload");
    int r = submit_impl(c, lane, sl.dev, (const Av1bFrameHdr*)sl.host, stages, refresh_mask, frame_id);
    if (r) return r;
    if (rt_event_record(sl.done, c->lanes[lane].stream)) return fail(c, AV1B_ECUDA, "event record");
    sl.pending = true;
    return AV1B_OK;
}

int av1b_frame_submit_resident(av1b_ctx* c, const void* dev_cmd, const Av1bFrameHdr* hdr, uint32_t stages,
    uint32_t refresh_mask, int* frame_id)
{
    if (!c || !dev_cmd || !hdr) return AV1B_EINVAL;
    rt_set_device(c->device);
    const int lane = next_lane(c);
    if (lane < 0) return fail(c, AV1B_ECUDA, "lane");
    return submit_impl(c, lane, (const uint8_t*)dev_cmd, hdr, stages, refresh_mask, frame_id);
}

int av1b_show_existing(av1b_ctx* c, int slot, uint32_t refresh_mask, int* frame_id)
{
    if (!c || slot < 0 || slot > 7) return AV1B_EINVAL;
    const int f = c->ref_slot[slot];
    if (f < 0) return fail(c, AV1B_ESTATE, "show_existing_frame of an empty slot");
    for (int i = 0; i < 8; i++) {
        if (refresh_mask & (1u << i)) {
            c->frames[f].refcnt++;
            if (c->ref_slot[i] >= 0) c->frames[c->ref_slot[i]].refcnt--;
            c->ref_slot[i] = f;
        }
    }
    if (frame_id) *frame_id = f;
    return AV1B_OK;
}

int av1b_frame_download(av1b_ctx* c, int frame_id, uint8_t* const dst[3], const int dst_stride[3], int w, int h)
{
    if (!c || frame_id < 0 || frame_id >= (int)c->frames.size()) return AV1B_EINVAL;
    rt_set_device(c->device);
    if (frame_read(c, frame_id, MAIN_LANE)) return fail(c, AV1B_ECUDA, "stream wait");
    const FrameView& v = c->frames[frame_id].v;
    for (int p = 0; p < 3; p++) {
        const int pw = p ? (w >> 1) : w, ph = p ? (h >> 1) : h;
        if (!dst[p] || pw <= 0 || ph <= 0) continue;
        if (rt_copy2d(dst[p], dst_stride[p], v.pl[p].p, v.pl[p].stride, pw, ph, c->stream, 1)) return fail(c, AV1B_ECUDA, "download");
    }
    if (rt_event_record(c->frames[frame_id].copied, c->stream)) return fail(c, AV1B_ECUDA, "event record");
    return AV1B_OK;
}

int av1b_frame_device_view(av1b_ctx* c, int frame_id, const uint8_t* planes[3], int pitches[3])
{
    if (!c || !planes || !pitches || frame_id < 0 || frame_id >= (int)c->frames.size()) return AV1B_EINVAL;
    rt_set_device(c->device);
    if (frame_read(c, frame_id, MAIN_LANE)) return fail(c, AV1B_ECUDA, "stream wait");
    // the caller reads on the context stream: whoever reuses the buffer waits for that stream's tail
    if (rt_event_record(c->frames[frame_id].copied, c->stream)) return fail(c, AV1B_ECUDA, "event record");
    for (int p = 0; p < 3; p++) {
        planes[p] = c->frames[frame_id].v.pl[p].p;
        pitches[p] = c->frames[frame_id].v.pl[p].stride;
    }
    return AV1B_OK;
}

int av1b_frame_retain(av1b_ctx* c, int frame_id)
{
    if (!c || frame_id < 0 || frame_id >= (int)c->frames.size()) return AV1B_EINVAL;
    c->frames[frame_id].refcnt++;
    return AV1B_OK;
}

int av1b_frame_release(av1b_ctx* c, int frame_id)
{
    if (!c || frame_id < 0 || frame_id >= (int)c->frames.size() || c->frames[frame_id].refcnt <= 0) return AV1B_EINVAL;
    rt_set_device(c->device);
    // the holder read it on the context stream up to now
    c->frames[frame_id].readers |= 1u << MAIN_LANE;
    if (rt_event_record(c->frames[frame_id].copied, c->stream)) return fail(c, AV1B_ECUDA, "event record");
    c->frames[frame_id].refcnt--;
    return AV1B_OK;
}

int av1b_frame_to_nv12(av1b_ctx* c, int frame_id, uint8_t* dst_y, int pitch_y, uint8_t* dst_uv, int pitch_uv, int w, int h)
{
    if (!c || !dst_y || !dst_uv || frame_id < 0 || frame_id >= (int)c->frames.size() || w <= 0 || h <= 0 || w > c->aw || h > c->ah
        || pitch_y < w || pitch_uv < ((w >> 1) << 1))
        return AV1B_EINVAL;
    rt_set_device(c->device);
    if (frame_read(c, frame_id, MAIN_LANE)) return fail(c, AV1B_ECUDA, "stream wait");
    launch_to_nv12(c->frames[frame_id].v, dst_y, pitch_y, dst_uv, pitch_uv, w, h, c->stream);
    c->launches++;
    if (rt_check() || rt_event_record(c->frames[frame_id].copied, c->stream)) return fail(c, AV1B_ECUDA, "nv12");
    return AV1B_OK;
}

int av1b_sync(av1b_ctx* c)
{
    if (!c) return AV1B_EINVAL;
    rt_set_device(c->device);
    for (int m = 0; m < c->n_lanes; m++)
        if (rt_stream_sync(c->lanes[m].stream)) return fail(c, AV1B_ECUDA, "sync");
    if (rt_stream_sync(c->stream)) return fail(c, AV1B_ECUDA, "sync");
    c->joined = true;
    // everything is finished: no recorded event has to be waited for any more
    for (auto& f : c->frames) {
        f.readers = 0;
        f.settled = true;
    }
    return AV1B_OK;
}

int av1b_set_capture(av1b_ctx* c, int on)
{
    if (!c) return AV1B_EINVAL;
    c->capturing = on != 0;
    // on: start from a settled context, nothing recorded outside the capture has to be waited for.
    // off: events recorded during the capture belong to the graph and must not be waited for by
    // later work -- nothing ran yet, so everything counts as settled again.
    if (on) return av1b_sync(c);
    for (auto& f : c->frames) {
        f.readers = 0;
        f.settled = true;
    }
    c->joined = true;
    return AV1B_OK;
}

int av1b_set_lanes(av1b_ctx* c, int n)
{
    if (!c || n < 1 || n > MAX_LANES) return AV1B_EINVAL;
    rt_set_device(c->device);
    if (av1b_sync(c)) return AV1B_ECUDA;
    for (int m = c->lanes_made; m < n; m++) {
        if (rt_stream_create(&c->lanes[m].stream) || rt_event_create(&c->lanes[m].mark)) return fail(c, AV1B_ECUDA, "lane stream");
        c->lanes_made = m + 1;
    }
    c->n_lanes = n;
    // the working set of the new lane count, allocated now rather than in the middle of a decode
    while (c->frames.size() < pool_soft_cap(c)) {
        void* p = nullptr;
        if (rt_malloc(&p, c->frame_bytes)) return fail(c, AV1B_ENOMEM, "frame pool");
        if (frame_add(c, (uint8_t*)p, true, MAIN_LANE)) {
            rt_free(p);
            return fail(c, AV1B_ECUDA, "cudaEventCreate");
        }
    }
    return AV1B_OK;
}

int av1b_join(av1b_ctx* c)
{
    if (!c) return AV1B_EINVAL;
    rt_set_device(c->device);
    if (join_lanes(c, true)) return fail(c, AV1B_ECUDA, "join");
    return AV1B_OK;
}

int av1b_fence_record(av1b_ctx* c, uint64_t* fence)
{
    if (!c || !fence) return AV1B_EINVAL;
    rt_set_device(c->device);
    const uint64_t id = c->fence_next++;
    if (join_lanes(c, false)) return fail(c, AV1B_ECUDA, "join");
    if (rt_event_record(c->fences[id % N_FENCES], c->stream)) return fail(c, AV1B_ECUDA, "fence record");
    *fence = id;
    return AV1B_OK;
}

int av1b_fence_done(av1b_ctx* c, uint64_t fence)
{
    if (!c) return 1;
    if (fence == 0) return 1;
    rt_set_device(c->device);
    // a slot that has been re-recorded since holds a LATER point of the same stream: its
    // completion implies the older fence's, so query it rather than assuming "long retired"
    return rt_event_done(c->fences[fence % N_FENCES]);
}

int av1b_fence_wait(av1b_ctx* c, uint64_t fence)
{
    if (!c) return AV1B_EINVAL;
    if (fence == 0) return AV1B_OK;
    rt_set_device(c->device);
    // (a re-recorded slot is a later point of the same stream: waiting on it covers this fence)
    if (rt_event_sync(c->fences[fence % N_FENCES])) return fail(c, AV1B_ECUDA, "fence wait");
    return AV1B_OK;
}

void* av1b_host_alloc(size_t bytes)
{
    const size_t bucket = bucket_of(bytes);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto& fl = g_pinned_free[bucket];
        if (!fl.empty()) {
            void* p = fl.back();
            fl.pop_back();
            return p;
        }
    }
    void* p = nullptr;
    if (rt_host_alloc(&p, bucket)) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    g_pinned_size[p] = bucket;
    return p;
}
void av1b_host_free(void* p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_pinned_size.find(p);
    if (it == g_pinned_size.end()) return;
    g_pinned_free[it->second].push_back(p);
}
void* av1b_dev_alloc(size_t bytes)
{
    void* p = nullptr;
    return rt_malloc(&p, bytes) ? nullptr : p;
}
void av1b_dev_free(void* p) { rt_free(p); }
int av1b_dev_upload(av1b_ctx* c, void* dev_dst, const void* host_src, size_t bytes)
{
    if (!c) return AV1B_EINVAL;
    rt_set_device(c->device);
    if (rt_h2d(dev_dst, host_src, bytes, c->stream) || rt_stream_sync(c->stream)) return fail(c, AV1B_ECUDA, "upload");
    return AV1B_OK;
}

int av1b_dev_download(av1b_ctx* c, void* host_dst, const void* dev_src, size_t bytes)
{
    if (!c) return AV1B_EINVAL;
    rt_set_device(c->device);
    if (rt_d2h(host_dst, dev_src, bytes, c->stream) || rt_stream_sync(c->stream)) return fail(c, AV1B_ECUDA, "download");
    return AV1B_OK;
}

static int upload_planes(av1b_ctx* c, int f, const uint8_t* const src[3], const int src_stride[3], int w, int h)
{
    if (frame_claim(c, f, MAIN_LANE)) return fail(c, AV1B_ECUDA, "stream wait");
    const FrameView& v = c->frames[f].v;
    for (int p = 0; p < 3; p++) {
        const int pw = p ? (w >> 1) : w, ph = p ? (h >> 1) : h;
        if (rt_copy2d(v.pl[p].p, v.pl[p].stride, src[p], src_stride[p], pw, ph, c->stream, 0)) return fail(c, AV1B_ECUDA, "upload planes");
    }
    if (rt_event_record(c->frames[f].ready, c->stream) || rt_stream_sync(c->stream)) return fail(c, AV1B_ECUDA, "upload planes");
    return AV1B_OK;
}

int av1b_debug_set_input(av1b_ctx* c, const uint8_t* const src[3], const int src_stride[3], int w, int h)
{
    if (!c || w > c->aw || h > c->ah) return AV1B_EINVAL;
    rt_set_device(c->device);
    int f = c->pending_input >= 0 ? c->pending_input : frame_alloc(c);
    if (f < 0) return fail(c, AV1B_ENOMEM, "frame pool exhausted");
    c->pending_input = f;
    return upload_planes(c, f, src, src_stride, w, h);
}

int av1b_debug_set_ref(av1b_ctx* c, int slot, const uint8_t* const src[3], const int src_stride[3], int w, int h)
{
    if (!c || slot < 0 || slot > 7 || w > c->aw || h > c->ah) return AV1B_EINVAL;
    rt_set_device(c->device);
    int f = frame_alloc(c);
    if (f < 0) return fail(c, AV1B_ENOMEM, "frame pool exhausted");
    if (c->ref_slot[slot] >= 0) c->frames[c->ref_slot[slot]].refcnt--;
    c->ref_slot[slot] = f;
    c->frames[f].refcnt++;
    return upload_planes(c, f, src, src_stride, w, h);
}

int av1b_debug_get_residual(av1b_ctx* c, int16_t* dst, size_t n)
{
    if (!c || n > c->res_cap) return AV1B_EINVAL;
    rt_set_device(c->device);
    if (join_lanes(c, false)) return fail(c, AV1B_ECUDA, "join");
    if (rt_d2h(dst, c->res, n * sizeof(int16_t), c->stream) || rt_stream_sync(c->stream)) return fail(c, AV1B_ECUDA, "residual download");
    return AV1B_OK;
}

uint64_t av1b_launch_count(av1b_ctx* c) { return c ? c->launches : 0; }

int av1b_debug_wave_trace(size_t n)
{
    if (g_wave_trace) {
        rt_device_sync();
        rt_free(g_wave_trace);
        g_wave_trace = nullptr;
        g_wave_trace_cap = 0;
    }
    if (!n) return AV1B_OK;
    void* p = nullptr;
    if (rt_malloc(&p, n * 64)) return AV1B_ENOMEM;
    if (rt_memset(p, 0, n * 64, nullptr) || rt_device_sync()) {
        rt_free(p);
        return AV1B_ECUDA;
    }
    g_wave_trace = (unsigned long long*)p;
    g_wave_trace_cap = n;
    return AV1B_OK;
}

int av1b_debug_wave_trace_read(uint64_t* out, size_t n)
{
    if (!g_wave_trace || n > g_wave_trace_cap) return AV1B_EINVAL;
    rt_device_sync();
    return rt_d2h_sync(out, g_wave_trace, n * 64) ? AV1B_ECUDA : AV1B_OK;
}

void av1b_debug_counters(uint64_t out[4])
{
    out[0] = g_n_ctx_new, out[1] = g_n_ctx_reused, out[2] = g_n_dev_alloc, out[3] = g_n_pinned_alloc;
}

int av1b_set_profiling(av1b_ctx* c, int on)
{
    if (!c) return AV1B_EINVAL;
    c->profiling = on != 0;
    return AV1B_OK;
}

int av1b_get_stage_times(av1b_ctx* c, double ms[AV1B_N_STAGES], uint64_t calls[AV1B_N_STAGES], int reset)
{
    if (!c) return AV1B_EINVAL;
    rt_set_device(c->device);
    for (int m = 0; m < c->n_lanes; m++)
        if (rt_stream_sync(c->lanes[m].stream)) return fail(c, AV1B_ECUDA, "sync");
    if (rt_stream_sync(c->stream)) return fail(c, AV1B_ECUDA, "sync");
    for (auto& sp : c->spans) {
        c->stage_ms[sp.stage] += rt_event_ms(sp.a, sp.b);
        c->stage_calls[sp.stage]++;
        c->event_pool.push_back(sp.a);
        c->event_pool.push_back(sp.b);
    }
    c->spans.clear();
    for (int i = 0; i < AV1B_N_STAGES; i++) {
        if (ms) ms[i] = c->stage_ms[i];
        if (calls) calls[i] = c->stage_calls[i];
        if (reset) {
            c->stage_ms[i] = 0;
            c->stage_calls[i] = 0;
        }
    }
    return AV1B_OK;
}

int av1b_debug_input_from_slot(av1b_ctx* c, int slot)
{
    if (!c || slot < 0 || slot > 7 || c->ref_slot[slot] < 0) return AV1B_EINVAL;
    rt_set_device(c->device);
    int f = c->pending_input >= 0 ? c->pending_input : frame_alloc(c);
    if (f < 0) return fail(c, AV1B_ENOMEM, "frame pool exhausted");
    c->pending_input = f;
    if (frame_read(c, c->ref_slot[slot], MAIN_LANE) || frame_claim(c, f, MAIN_LANE)) return fail(c, AV1B_ECUDA, "stream wait");
    if (rt_d2d(c->frames[f].base, c->frames[c->ref_slot[slot]].base, c->frame_bytes, c->stream)
        || rt_event_record(c->frames[f].ready, c->stream) || rt_event_record(c->frames[c->ref_slot[slot]].copied, c->stream))
        return fail(c, AV1B_ECUDA, "d2d copy");
    return AV1B_OK;
}

size_t av1b_struct_size(int which)
{
    switch (which) {
    case 0: return sizeof(Av1bFrameHdr);
    case 1: return sizeof(Av1bOp);
    case 2: return sizeof(Av1bSb);
    case 3: return sizeof(Av1bIpu);
    case 4: return sizeof(Av1bInterBlk);
    case 5: return sizeof(Av1bBlkAux);
    case 6: return sizeof(Av1bLfMi);
    case 7: return sizeof(Av1bLrUnit);
    default: return 0;
    }
}

}  // extern "C"

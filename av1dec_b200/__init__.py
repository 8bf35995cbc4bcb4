"""av1dec_b200 -- B200-native AV1 reconstruction + in-loop-filter engine behind the decoder API
of oddstone/av1dec.

Python is plumbing only: this module binds the two C-ABI shared libraries with ctypes

    lib/libav1b200.so      include/av1b200.h          the sm_100a engine (kernels + frame store)
    lib/libav1b200dec.so   include/av1b200_decoder.h  host front end + drop-in YamiAv1::Decoder

and mirrors the reference's decoder interface (decoder/Av1Decoder.h:47-51: ``decode`` /
``getOutput``) as :class:`Decoder`.  There is no CPU fallback: if the CUDA libraries are missing
or no GPU is present, loading / decoding raises.
"""
import os as _os

# Work from many decoder streams is ordered by events; with CUDA's default 8 hardware queues a
# wait at the head of a queue stalls unrelated streams.  Must be set before CUDA initialises.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")

STAGE_ITX, STAGE_INTER, STAGE_WAVE, STAGE_DEBLOCK, STAGE_CDEF, STAGE_LR = 1, 2, 4, 8, 16, 32
STAGE_RECON = STAGE_ITX | STAGE_INTER | STAGE_WAVE
STAGE_POST = STAGE_DEBLOCK | STAGE_CDEF | STAGE_LR
STAGE_ALL = STAGE_RECON | STAGE_POST

CMD_SINK = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t, C.c_uint32, C.c_int)

ENGINE_SYMBOLS = [
    "av1b_backend", "av1b_ctx_create", "av1b_ctx_destroy", "av1b_last_error", "av1b_cmd_acquire",
    "av1b_frame_submit", "av1b_frame_submit_resident", "av1b_show_existing", "av1b_frame_download",
    "av1b_dev_download", "av1b_frame_device_view", "av1b_frame_retain", "av1b_frame_release", "av1b_frame_to_nv12", "av1b_sync", "av1b_fence_record", "av1b_fence_wait", "av1b_fence_done", "av1b_pool_purge", "av1b_host_alloc", "av1b_host_free",
    "av1b_dev_alloc", "av1b_dev_free", "av1b_dev_upload", "av1b_debug_set_input", "av1b_debug_set_ref",
    "av1b_debug_get_residual", "av1b_debug_counters", "av1b_join", "av1b_set_lanes", "av1b_set_capture", "av1b_launch_count", "av1b_set_profiling", "av1b_get_stage_times",
    "av1b_debug_input_from_slot", "av1b_struct_size", "av1b_debug_wave_trace", "av1b_debug_wave_trace_read",
]
STAGE_NAMES = ["itx", "inter", "wave", "deblock", "cdef", "lr"]
DECODER_SYMBOLS = [
    "av1b_decoder_create", "av1b_decoder_destroy", "av1b_decoder_set_stages", "av1b_decoder_set_cmd_sink",
    "av1b_decoder_decode", "av1b_decoder_get_output", "av1b_decoder_error", "av1b_decode_ivf", "av1b_ivf_segments",
    "createVideoDecoder", "releaseVideoDecoder",
]


class EngineError(RuntimeError):
    pass


def _bind_engine(lib):
    u8pp = C.POINTER(C.POINTER(C.c_uint8))
    lib.av1b_backend.restype = C.c_char_p
    lib.av1b_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.av1b_ctx_destroy.argtypes = [C.c_void_p]
    lib.av1b_ctx_destroy.restype = None
    lib.av1b_last_error.argtypes = [C.c_void_p]
    lib.av1b_last_error.restype = C.c_char_p
    lib.av1b_cmd_acquire.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.av1b_frame_submit.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.POINTER(C.c_int)]
    lib.av1b_frame_submit_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_int)]
    lib.av1b_show_existing.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_int)]
    lib.av1b_frame_download.argtypes = [C.c_void_p, C.c_int, u8pp, C.POINTER(C.c_int), C.c_int, C.c_int]
    lib.av1b_sync.argtypes = [C.c_void_p]
    lib.av1b_fence_record.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    lib.av1b_fence_wait.argtypes = [C.c_void_p, C.c_uint64]
    lib.av1b_host_alloc.argtypes = [C.c_size_t]
    lib.av1b_host_alloc.restype = C.c_void_p
    lib.av1b_host_free.argtypes = [C.c_void_p]
    lib.av1b_host_free.restype = None
    lib.av1b_dev_alloc.argtypes = [C.c_size_t]
    lib.av1b_dev_alloc.restype = C.c_void_p
    lib.av1b_dev_free.argtypes = [C.c_void_p]
    lib.av1b_dev_free.restype = None
    lib.av1b_dev_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.av1b_debug_set_input.argtypes = [C.c_void_p, u8pp, C.POINTER(C.c_int), C.c_int, C.c_int]
    lib.av1b_debug_set_ref.argtypes = [C.c_void_p, C.c_int, u8pp, C.POINTER(C.c_int), C.c_int, C.c_int]
    lib.av1b_debug_get_residual.argtypes = [C.c_void_p, C.POINTER(C.c_int16), C.c_size_t]
    lib.av1b_launch_count.argtypes = [C.c_void_p]
    lib.av1b_launch_count.restype = C.c_uint64
    lib.av1b_join.argtypes = [C.c_void_p]
    lib.av1b_dev_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.av1b_frame_device_view.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    lib.av1b_frame_retain.argtypes = [C.c_void_p, C.c_int]
    lib.av1b_frame_release.argtypes = [C.c_void_p, C.c_int]
    lib.av1b_frame_to_nv12.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.av1b_debug_counters.argtypes = [C.POINTER(C.c_uint64)]
    lib.av1b_debug_counters.restype = None
    lib.av1b_set_lanes.argtypes = [C.c_void_p, C.c_int]
    lib.av1b_set_capture.argtypes = [C.c_void_p, C.c_int]
    lib.av1b_set_profiling.argtypes = [C.c_void_p, C.c_int]
    lib.av1b_get_stage_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]
    lib.av1b_debug_input_from_slot.argtypes = [C.c_void_p, C.c_int]
    lib.av1b_struct_size.argtypes = [C.c_int]
    lib.av1b_struct_size.restype = C.c_size_t
    return lib


def _bind_decoder(lib):
    lib.av1b_decoder_create.argtypes = [C.c_int]
    lib.av1b_decoder_create.restype = C.c_void_p
    lib.av1b_decoder_destroy.argtypes = [C.c_void_p]
    lib.av1b_decoder_destroy.restype = None
    lib.av1b_decoder_set_stages.argtypes = [C.c_void_p, C.c_uint32]
    lib.av1b_decoder_set_stages.restype = None
    lib.av1b_decoder_set_cmd_sink.argtypes = [C.c_void_p, CMD_SINK, C.c_void_p]
    lib.av1b_decoder_set_cmd_sink.restype = None
    lib.av1b_decoder_decode.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.av1b_decoder_get_output.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                            C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_int)]
    lib.av1b_decoder_error.argtypes = [C.c_void_p]
    lib.av1b_decoder_error.restype = C.c_char_p
    lib.av1b_decode_ivf.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_uint32, C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    lib.av1b_ivf_segments.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.c_int]
    return lib


_engine = None
_decoder = None


def engine_path():
    return os.path.join(LIB_DIR, "libav1b200.so")


def decoder_path():
    return os.path.join(LIB_DIR, "libav1b200dec.so")


def load_engine(path=None):
    """Load the CUDA engine library.  Raises if it is missing or is not the sm_100a build."""
    global _engine
    if _engine is not None and path is None:
        return _engine
    p = path or engine_path()
    if not os.path.exists(p):
        raise EngineError(f"{p} not found: run `make` (nvcc, sm_100a); there is no CPU fallback")
    lib = _bind_engine(C.CDLL(p))
    if path is None:
        if lib.av1b_backend() != b"cuda-sm_100a":
            raise EngineError("refusing to use a non-CUDA engine build as the product")
        _engine = lib
    return lib


def load_decoder(path=None):
    """Load the host decoder library (front end + drop-in Decoder)."""
    global _decoder
    if _decoder is not None and path is None:
        return _decoder
    p = path or decoder_path()
    if path is None:
        load_engine()
    if not os.path.exists(p):
        raise EngineError(f"{p} not found: run `make`")
    lib = _bind_decoder(C.CDLL(p))
    if path is None:
        _decoder = lib
    return lib


def decode_ivf(data, device=0, stages=STAGE_ALL, want_yuv=True, lib=None):
    """Decode a whole IVF byte string.  Returns (yuv_bytes_or_None, n_frames, shown_luma_pixels).

    The yuv layout is the reference CLI's output file (tests/DecodeOutput.cpp:48-69)."""
    lib = lib or load_decoder()
    out_bytes, n_frames, pixels = C.c_size_t(0), C.c_int(0), C.c_uint64(0)
    if len(data) < 32 or data[:4] != b"DKIF":
        raise EngineError("not an IVF stream")
    # size the output from the IVF header (width, height, frame count) so one decode suffices
    w, h = int.from_bytes(data[12:14], "little"), int.from_bytes(data[14:16], "little")
    n = int.from_bytes(data[24:28], "little")
    cap = min(max(w * h * 3 // 2 * max(n, 1) + (64 << 10), 1 << 16), 1 << 31)
    while True:
        buf = C.create_string_buffer(cap) if want_yuv else None
        rc = lib.av1b_decode_ivf(data, len(data), device, stages, buf, cap if want_yuv else 0,
                                 C.byref(out_bytes), C.byref(n_frames), C.byref(pixels))
        if rc == -2 and want_yuv:
            cap = max(out_bytes.value, cap * 2)
            continue
        if rc != 0:
            raise EngineError(f"av1b_decode_ivf failed rc={rc}")
        break
    yuv = buf.raw[:out_bytes.value] if want_yuv else None
    return yuv, n_frames.value, pixels.value


def alloc_counters(lib=None):
    """(contexts created, contexts recycled, device allocations, pinned allocations) so far."""
    lib = lib or load_engine()
    out = (C.c_uint64 * 4)()
    lib.av1b_debug_counters(out)
    return tuple(int(v) for v in out)


def ivf_segments(data, lib=None):
    """Temporal-unit index of the first unit of every closed segment (random access point) of an
    IVF byte string; av1b_decode_ivf decodes these segments in parallel."""
    lib = lib or load_decoder()
    n = lib.av1b_ivf_segments(data, len(data), None, 0)
    if n < 0:
        raise EngineError("not an IVF stream")
    arr = (C.c_uint32 * max(n, 1))()
    lib.av1b_ivf_segments(data, len(data), arr, n)
    return list(arr[:n])


class Decoder:
    """Mirror of the reference's ``YamiAv1::Decoder`` (decoder/Av1Decoder.h:47-51)."""

    def __init__(self, device=0, stages=STAGE_ALL, lib=None):
        self._lib = lib or load_decoder()
        self._h = self._lib.av1b_decoder_create(device)
        self._sink = None
        if stages != STAGE_ALL:
            self._lib.av1b_decoder_set_stages(self._h, stages)

    def set_cmd_sink(self, fn):
        """fn(data_bytes_or_None, nbytes_or_slot, refresh_mask, show) per frame."""
        def tramp(_user, ptr, n, refresh, show):
            fn(C.string_at(ptr, n) if ptr else None, n, refresh, show)
        self._sink = CMD_SINK(tramp)
        self._lib.av1b_decoder_set_cmd_sink(self._h, self._sink, None)

    def decode(self, data):
        """One temporal unit.  Returns True/False like the reference."""
        return self._lib.av1b_decoder_decode(self._h, data, len(data)) == 0

    def get_output(self):
        """Next shown frame as (width, height, [Y, U, V] bytes) or None."""
        w, h = C.c_int(0), C.c_int(0)
        planes = (C.POINTER(C.c_uint8) * 3)()
        strides = (C.c_int * 3)()
        rc = self._lib.av1b_decoder_get_output(self._h, C.byref(w), C.byref(h), planes, strides)
        if rc != 1:
            return None
        out = []
        for p in range(3):
            pw, ph = (w.value >> 1, h.value >> 1) if p else (w.value, h.value)
            rows = [C.string_at(C.addressof(planes[p].contents) + y * strides[p], pw) for y in range(ph)]
            out.append(b"".join(rows))
        return w.value, h.value, out

    def error(self):
        return self._lib.av1b_decoder_error(self._h).decode()

    def close(self):
        if self._h:
            self._lib.av1b_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def iter_ivf(data):
    """Yield the temporal units of an IVF byte string."""
    hdr = int.from_bytes(data[6:8], "little")
    pos = hdr
    while pos + 12 <= len(data):
        sz = int.from_bytes(data[pos:pos + 4], "little")
        pos += 12
        yield data[pos:pos + sz]
        pos += sz

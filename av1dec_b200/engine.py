"""Thin numpy/ctypes wrapper over the engine C ABI (include/av1b200.h) for tests and bench.py.

Everything here is plumbing: allocate a context, hand it command buffers, move planes in and out.
"""
import ctypes as C

import numpy as np

from . import STAGE_ALL, EngineError, load_engine


def _plane_args(planes):
    arr = [np.ascontiguousarray(p, dtype=np.uint8) for p in planes]
    ptrs = (C.POINTER(C.c_uint8) * 3)(*[a.ctypes.data_as(C.POINTER(C.c_uint8)) for a in arr])
    strides = (C.c_int * 3)(*[a.strides[0] for a in arr])
    return arr, ptrs, strides


class Engine:
    def __init__(self, max_w, max_h, device=0, stream=None, lib=None):
        self.lib = lib or load_engine()
        self.ctx = C.c_void_p()
        rc = self.lib.av1b_ctx_create(C.byref(self.ctx), device, max_w, max_h, stream)
        if rc != 0:
            msg = self.lib.av1b_last_error(self.ctx).decode() if self.ctx else "?"
            raise EngineError(f"av1b_ctx_create rc={rc}: {msg}")

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError(f"{what} rc={rc}: {self.lib.av1b_last_error(self.ctx).decode()}")

    def submit(self, cmd, stages=STAGE_ALL, refresh_mask=0):
        """Upload a command buffer (bytes) through the pinned ring and run the stages."""
        ptr = C.c_void_p()
        self._check(self.lib.av1b_cmd_acquire(self.ctx, len(cmd), C.byref(ptr)), "av1b_cmd_acquire")
        C.memmove(ptr, cmd, len(cmd))
        fid = C.c_int(-1)
        self._check(self.lib.av1b_frame_submit(self.ctx, len(cmd), stages, refresh_mask, C.byref(fid)), "av1b_frame_submit")
        return fid.value

    def submit_resident(self, dev_cmd, hdr_bytes, stages=STAGE_ALL, refresh_mask=0):
        fid = C.c_int(-1)
        self._check(self.lib.av1b_frame_submit_resident(self.ctx, dev_cmd, hdr_bytes, stages, refresh_mask, C.byref(fid)),
                    "av1b_frame_submit_resident")
        return fid.value

    def show_existing(self, slot, refresh_mask):
        fid = C.c_int(-1)
        self._check(self.lib.av1b_show_existing(self.ctx, slot, refresh_mask, C.byref(fid)), "av1b_show_existing")
        return fid.value

    def upload(self, data):
        """Copy bytes to a fresh device allocation; returns the device pointer (int)."""
        p = self.lib.av1b_dev_alloc(len(data))
        if not p:
            raise EngineError("av1b_dev_alloc failed")
        self._check(self.lib.av1b_dev_upload(self.ctx, p, data, len(data)), "av1b_dev_upload")
        return p

    def free(self, dev_ptr):
        self.lib.av1b_dev_free(dev_ptr)

    def set_input(self, planes, w, h):
        keep, ptrs, strides = _plane_args(planes)
        self._check(self.lib.av1b_debug_set_input(self.ctx, ptrs, strides, w, h), "av1b_debug_set_input")

    def set_ref(self, slot, planes, w, h):
        keep, ptrs, strides = _plane_args(planes)
        self._check(self.lib.av1b_debug_set_ref(self.ctx, slot, ptrs, strides, w, h), "av1b_debug_set_ref")

    def download(self, frame_id, w, h):
        out = [np.empty((h, w), np.uint8), np.empty((h >> 1, w >> 1), np.uint8), np.empty((h >> 1, w >> 1), np.uint8)]
        ptrs = (C.POINTER(C.c_uint8) * 3)(*[a.ctypes.data_as(C.POINTER(C.c_uint8)) for a in out])
        strides = (C.c_int * 3)(*[a.strides[0] for a in out])
        self._check(self.lib.av1b_frame_download(self.ctx, frame_id, ptrs, strides, w, h), "av1b_frame_download")
        self.sync()
        return out

    def device_view(self, frame_id):
        """([device pointers], [pitches]) of the frame's planes; valid on the context stream."""
        ptrs = (C.c_void_p * 3)()
        pitches = (C.c_int * 3)()
        self._check(self.lib.av1b_frame_device_view(self.ctx, frame_id, ptrs, pitches), "av1b_frame_device_view")
        return [int(p) for p in ptrs], [int(p) for p in pitches]

    def read_device(self, dev_ptr, nbytes):
        buf = np.empty(nbytes, np.uint8)
        self._check(self.lib.av1b_dev_download(self.ctx, buf.ctypes.data_as(C.c_void_p), C.c_void_p(dev_ptr), nbytes), "av1b_dev_download")
        return buf

    def to_nv12(self, frame_id, w, h):
        """The frame as NV12: (luma (h, w), interleaved chroma (h/2, 2*(w/2))), converted on the device."""
        cw = (w >> 1) * 2
        self.lib.av1b_dev_alloc.restype = C.c_void_p
        dy = self.lib.av1b_dev_alloc(C.c_size_t(w * h))
        duv = self.lib.av1b_dev_alloc(C.c_size_t(max(cw, 1) * max(h >> 1, 1)))
        try:
            self._check(self.lib.av1b_frame_to_nv12(self.ctx, frame_id, C.c_void_p(dy), w, C.c_void_p(duv), cw, w, h), "av1b_frame_to_nv12")
            y = self.read_device(dy, w * h).reshape(h, w)
            uv = self.read_device(duv, cw * (h >> 1)).reshape(h >> 1, cw)
        finally:
            self.lib.av1b_dev_free(C.c_void_p(dy))
            self.lib.av1b_dev_free(C.c_void_p(duv))
        return y, uv

    def residual(self, n):
        out = np.empty(n, np.int16)
        self._check(self.lib.av1b_debug_get_residual(self.ctx, out.ctypes.data_as(C.POINTER(C.c_int16)), n), "av1b_debug_get_residual")
        return out

    def sync(self):
        self._check(self.lib.av1b_sync(self.ctx), "av1b_sync")

    def set_lanes(self, n):
        """Frames in flight on the device at once (1 = strictly one after the other)."""
        self._check(self.lib.av1b_set_lanes(self.ctx, n), "av1b_set_lanes")

    def set_capture(self, on):
        """Bracket a CUDA-graph capture of submits on the context's (caller-supplied) stream."""
        self._check(self.lib.av1b_set_capture(self.ctx, 1 if on else 0), "av1b_set_capture")

    def join(self):
        """Make the context stream wait (device-side) for every frame submitted so far."""
        self._check(self.lib.av1b_join(self.ctx), "av1b_join")

    def set_profiling(self, on=True):
        self._check(self.lib.av1b_set_profiling(self.ctx, 1 if on else 0), "av1b_set_profiling")

    def stage_times(self, reset=True):
        """({stage: ms}, {stage: calls}) accumulated since the last reset (synchronises)."""
        from . import STAGE_NAMES
        ms = (C.c_double * 6)()
        calls = (C.c_uint64 * 6)()
        self._check(self.lib.av1b_get_stage_times(self.ctx, ms, calls, 1 if reset else 0), "av1b_get_stage_times")
        return ({n: ms[i] for i, n in enumerate(STAGE_NAMES)}, {n: int(calls[i]) for i, n in enumerate(STAGE_NAMES)})

    def input_from_slot(self, slot):
        self._check(self.lib.av1b_debug_input_from_slot(self.ctx, slot), "av1b_debug_input_from_slot")

    def launches(self):
        return int(self.lib.av1b_launch_count(self.ctx))

    def close(self):
        if self.ctx:
            self.lib.av1b_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

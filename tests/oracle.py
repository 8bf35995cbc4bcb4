"""ctypes binding of oracle/_ref/liboracle.so -- the UNMODIFIED reference decoder behind a C shim
(oracle/oracle_shim.cpp).  TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "liboracle.so")
CLI = os.path.join(ROOT, "oracle", "_ref", "av1dec")

_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB} missing: run `make oracle` where /root/reference is mounted")
        l = C.CDLL(LIB)
        u8pp = C.POINTER(C.POINTER(C.c_uint8))
        l.oracle_decode_ivf.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                        C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
        l.oracle_decode_stages.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t,
                                           C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        l.oracle_postfilter.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_uint32, u8pp, C.POINTER(C.c_int), u8pp,
                                        C.POINTER(C.c_int)]
        l.oracle_inverse_transform.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p]
        l.oracle_predict_intra.argtypes = [C.c_char_p, u8pp, C.POINTER(C.c_int)]
        l.oracle_predict_inter.argtypes = [C.c_char_p, C.c_int, u8pp, C.POINTER(C.c_int), u8pp, C.POINTER(C.c_int)]
        _lib = l
    return _lib


def decode_ivf(data):
    """(yuv bytes, n_frames, luma_pixels) through the reference Decoder."""
    n, frames, pix = C.c_size_t(0), C.c_int(0), C.c_uint64(0)
    lib().oracle_decode_ivf(data, len(data), None, 0, C.byref(n), C.byref(frames), C.byref(pix))
    buf = C.create_string_buffer(max(n.value, 1))
    rc = lib().oracle_decode_ivf(data, len(data), buf, n.value, C.byref(n), C.byref(frames), C.byref(pix))
    assert rc == 0
    return buf.raw[:n.value], frames.value, pix.value


def decode_stages(data, stage):
    """All frames captured after `stage` (0 recon, 1 deblock, 2 cdef, 3 final) as bytes."""
    n, frames = C.c_size_t(0), C.c_int(0)
    lib().oracle_decode_stages(data, len(data), stage, None, 0, C.byref(n), C.byref(frames))
    buf = C.create_string_buffer(max(n.value, 1))
    rc = lib().oracle_decode_stages(data, len(data), stage, buf, n.value, C.byref(n), C.byref(frames))
    assert rc == 0
    return buf.raw[:n.value], frames.value


def postfilter(synth, stages):
    """Run the reference's LoopFilter / Cdef / LoopRestoration on a synth.SynthFrame.
    stages: bit0 deblock, bit1 CDEF, bit2 LR.  Returns [Y, U, V] (MI-aligned shape; only the
    visible area is defined once CDEF or LR ran)."""
    ins = [np.ascontiguousarray(p) for p in synth.planes]
    outs = [np.zeros_like(p) for p in ins]
    mk = lambda arrs: (C.POINTER(C.c_uint8) * 3)(*[a.ctypes.data_as(C.POINTER(C.c_uint8)) for a in arrs])
    st = lambda arrs: (C.c_int * 3)(*[a.strides[0] for a in arrs])
    idx = np.ascontiguousarray(synth.cdef_idx64, dtype=np.int8)
    rc = lib().oracle_postfilter(synth.cmd, idx.ctypes.data, int(synth.sb128), stages, mk(ins), st(ins), mk(outs), st(outs))
    assert rc == 0
    return outs


def inverse_transform(batch):
    """Reference TransformBlock::inverseTransform on a synth.make_itx_batch() batch.
    Returns a list of int32 residual arrays (h x w), BEFORE the flip mirroring."""
    from av1dec_b200.synth import TX_H, TX_W
    n = len(batch)
    ts = np.array([b[0] for b in batch], np.uint8)
    tt = np.array([b[1] for b in batch], np.uint8)
    ll = np.array([b[2] for b in batch], np.uint8)
    coef_off, res_off, co, ro = [], [], 0, 0
    for b in batch:
        coef_off.append(co)
        res_off.append(ro)
        co += len(b[3])
        ro += TX_W[b[0]] * TX_H[b[0]]
    coef = np.concatenate([b[3].astype(np.int32) for b in batch])
    res = np.zeros(ro, np.int32)
    coef_off = np.array(coef_off, np.uint32)
    res_off = np.array(res_off, np.uint32)
    rc = lib().oracle_inverse_transform(n, ts.ctypes.data, tt.ctypes.data, ll.ctypes.data, coef.ctypes.data,
                                        coef_off.ctypes.data, res.ctypes.data, res_off.ctypes.data)
    assert rc == 0
    return [res[res_off[i]:res_off[i] + TX_W[b[0]] * TX_H[b[0]]].reshape(TX_H[b[0]], TX_W[b[0]]) for i, b in enumerate(batch)]


def predict_intra(cmd, planes):
    """Block::IntraPredict::predict_intra (+ chroma-from-luma) for every intra op of a synthetic
    command buffer, in list order, over the picture `planes` ([Y, U, V], MI-aligned).  Returns the
    resulting planes."""
    outs = [np.ascontiguousarray(p).copy() for p in planes]
    ptrs = (C.POINTER(C.c_uint8) * 3)(*[a.ctypes.data_as(C.POINTER(C.c_uint8)) for a in outs])
    strides = (C.c_int * 3)(*[a.strides[0] for a in outs])
    n = lib().oracle_predict_intra(cmd, ptrs, strides)
    assert n >= 0
    return outs, n


def predict_inter(cmd, refs, shape_like):
    """Block::InterPredict::predict_inter for every translational prediction unit of a synthetic
    command buffer from the reference pictures `refs` (list of [Y, U, V] per store slot).  Returns
    ([Y, U, V] predicted picture, units run); samples no unit covers stay 0."""
    rr = [[np.ascontiguousarray(p) for p in planes] for planes in refs]
    flat = [a for planes in rr for a in planes]
    rptr = (C.POINTER(C.c_uint8) * len(flat))(*[a.ctypes.data_as(C.POINTER(C.c_uint8)) for a in flat])
    rstr = (C.c_int * 3)(*[a.strides[0] for a in rr[0]])
    outs = [np.zeros_like(p) for p in shape_like]
    optr = (C.POINTER(C.c_uint8) * 3)(*[a.ctypes.data_as(C.POINTER(C.c_uint8)) for a in outs])
    ostr = (C.c_int * 3)(*[a.strides[0] for a in outs])
    n = lib().oracle_predict_inter(cmd, len(rr), rptr, rstr, optr, ostr)
    assert n >= 0
    return outs, n

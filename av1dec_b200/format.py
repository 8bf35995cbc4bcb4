"""ctypes mirror of include/av1b200_format.h (the per-frame command buffer) plus a small builder.

Used by the tests and the benchmark to drive the engine's stage-level entry points with
synthetic frames (BASELINE.json config 4) -- the same flat sections the C++ emitter writes.
Sizes are cross-checked against the C structs by tests/test_host.py (av1b_struct_size).
"""
import ctypes as C

MAGIC = 0x42315641
VERSION = 4

OP_INTER_RES, OP_INTRA, OP_PALETTE, OP_INTERINTRA, OP_INTRABC = range(5)
OPF_HAVE_LEFT, OPF_HAVE_ABOVE, OPF_HAVE_ABOVE_RIGHT, OPF_HAVE_BELOW_LEFT = 1, 2, 4, 8
OPF_EDGE_SMOOTH, OPF_HAS_RESID, OPF_CFL, OPF_FILTER_INTRA = 0x10, 0x20, 0x40, 0x80


class Op(C.Structure):
    _fields_ = [("x", C.c_uint16), ("y", C.c_uint16), ("plane", C.c_uint8), ("kind", C.c_uint8),
                ("tx_size", C.c_uint8), ("tx_type", C.c_uint8), ("mode", C.c_uint8), ("angle_delta", C.c_int8),
                ("flags", C.c_uint8), ("fi_mode", C.c_uint8), ("cfl_alpha", C.c_int8), ("nz_rows", C.c_uint8),
                ("nz_cols", C.c_uint8), ("lossless", C.c_uint8), ("coef_off", C.c_uint32), ("res_off", C.c_uint32),
                ("aux", C.c_uint32), ("max_luma_w", C.c_uint16), ("max_luma_h", C.c_uint16)]


class Sb(C.Structure):
    _fields_ = [("first_op", C.c_uint32), ("n_ops", C.c_uint32), ("wait_l1", C.c_uint8), ("wait_l2", C.c_uint8), ("wait_a1", C.c_uint8), ("wait_a2", C.c_uint8),
                ("pub_r1", C.c_uint8), ("pub_b1", C.c_uint8), ("pad", C.c_uint8 * 2)]


class Ipu(C.Structure):
    _fields_ = [("x", C.c_uint16), ("y", C.c_uint16), ("w", C.c_uint8), ("h", C.c_uint8), ("plane", C.c_uint8),
                ("kind", C.c_uint8), ("mv", (C.c_int16 * 2) * 2), ("ref_slot", C.c_int8 * 2),
                ("ref_frame", C.c_uint8 * 2), ("filt", C.c_uint8 * 2), ("warp", C.c_uint8 * 2), ("flags", C.c_uint8),
                ("comp_type", C.c_uint8), ("fwd_w", C.c_uint8), ("bck_w", C.c_uint8), ("aux", C.c_uint32)]


class InterBlk(C.Structure):
    _fields_ = [("first_ipu", C.c_uint32), ("n_ipu", C.c_uint16), ("flags", C.c_uint16), ("x", C.c_uint16),
                ("y", C.c_uint16), ("cx", C.c_uint16), ("cy", C.c_uint16), ("bw", C.c_uint8), ("bh", C.c_uint8),
                ("cw", C.c_uint8), ("ch", C.c_uint8), ("pad", C.c_uint32)]


class BlkAux(C.Structure):
    _fields_ = [("warp_params", C.c_int32 * 6), ("warp_abgd", C.c_int16 * 4), ("mi_size", C.c_uint8),
                ("interintra_mode", C.c_uint8), ("wedge_interintra", C.c_uint8), ("wedge_index", C.c_uint8),
                ("wedge_sign", C.c_uint8), ("mask_type", C.c_uint8), ("pal_size_y", C.c_uint8),
                ("pal_size_uv", C.c_uint8), ("pal_colors", (C.c_uint8 * 8) * 3), ("pal_map_off", C.c_uint32 * 2),
                ("pal_map_stride", C.c_uint16 * 2), ("base_x", C.c_uint16 * 2), ("base_y", C.c_uint16 * 2),
                ("pad", C.c_uint8 * 8)]


class LfMi(C.Structure):
    _fields_ = [("mi_size", C.c_uint8), ("flags", C.c_uint8), ("tx", C.c_uint16), ("delta_lf", C.c_int8 * 4)]


class LrUnit(C.Structure):
    _fields_ = [("type", C.c_uint8), ("sgr_set", C.c_uint8), ("sgr_xqd", C.c_int8 * 2),
                ("wiener", (C.c_int8 * 3) * 2), ("pad", C.c_uint8 * 2)]


class LoopFilterParams(C.Structure):
    _fields_ = [("level", C.c_uint8 * 4), ("sharpness", C.c_uint8), ("delta_enabled", C.c_uint8),
                ("delta_lf_multi", C.c_uint8), ("pad", C.c_uint8), ("ref_deltas", C.c_int8 * 8),
                ("mode_deltas", C.c_int8 * 2), ("pad2", C.c_uint8 * 6)]


class CdefParams(C.Structure):
    _fields_ = [("enabled", C.c_uint8), ("damping", C.c_uint8), ("pad", C.c_uint8 * 2), ("y_pri", C.c_uint8 * 8),
                ("y_sec", C.c_uint8 * 8), ("uv_pri", C.c_uint8 * 8), ("uv_sec", C.c_uint8 * 8)]


class LrParams(C.Structure):
    _fields_ = [("uses_lr", C.c_uint8), ("frame_type", C.c_uint8 * 3), ("unit_size", C.c_uint16 * 3),
                ("unit_rows", C.c_uint16 * 3), ("unit_cols", C.c_uint16 * 3), ("pad", C.c_uint16),
                ("unit_first", C.c_uint32 * 3)]


class FrameHdr(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("version", C.c_uint32), ("total_bytes", C.c_uint32),
                ("frame_w", C.c_uint16), ("frame_h", C.c_uint16), ("mi_cols", C.c_uint16), ("mi_rows", C.c_uint16),
                ("sb_cols", C.c_uint16), ("sb_rows", C.c_uint16), ("sb_log2", C.c_uint8),
                ("enable_intra_edge_filter", C.c_uint8), ("frame_is_intra", C.c_uint8), ("allow_intrabc", C.c_uint8),
                ("ref_slot", C.c_int8 * 8), ("ref_w", C.c_uint16 * 8), ("ref_h", C.c_uint16 * 8),
                ("gm_params", (C.c_int32 * 6) * 8), ("gm_abgd", (C.c_int16 * 4) * 8), ("gm_warp_ok", C.c_uint8 * 8),
                ("off_sb", C.c_uint32), ("n_sb", C.c_uint32), ("off_ops", C.c_uint32), ("n_ops", C.c_uint32),
                ("off_itx", C.c_uint32), ("n_itx", C.c_uint32), ("itx_class_end", C.c_uint32 * 4), ("off_iblk", C.c_uint32), ("n_iblk", C.c_uint32),
                ("off_ipu", C.c_uint32), ("n_ipu", C.c_uint32), ("off_aux", C.c_uint32), ("n_aux", C.c_uint32),
                ("off_coef", C.c_uint32), ("n_coef", C.c_uint32), ("n_res", C.c_uint32), ("off_pal", C.c_uint32),
                ("n_pal", C.c_uint32), ("off_lfmi", C.c_uint32), ("off_cdef8", C.c_uint32), ("off_lru", C.c_uint32),
                ("n_lru", C.c_uint32), ("lf", LoopFilterParams), ("cdef", CdefParams), ("lr", LrParams)]


STRUCTS = {0: FrameHdr, 1: Op, 2: Sb, 3: Ipu, 4: InterBlk, 5: BlkAux, 6: LfMi, 7: LrUnit}


def _align16(v):
    return (v + 15) & ~15


def build(hdr, sections):
    """Assemble a command buffer.  `sections` maps a header offset-field name (e.g. 'off_ops') to
    bytes; fields not given stay 0.  Returns bytes with hdr.total_bytes filled in."""
    off = _align16(C.sizeof(FrameHdr))
    placed = []
    for name, blob in sections.items():
        setattr(hdr, name, off)
        placed.append((off, blob))
        off = _align16(off + len(blob))
    hdr.magic, hdr.version, hdr.total_bytes = MAGIC, VERSION, off
    buf = bytearray(off)
    buf[:C.sizeof(FrameHdr)] = bytes(hdr)
    for o, blob in placed:
        buf[o:o + len(blob)] = blob
    return bytes(buf)


TX_MAXDIM = [4, 8, 16, 32, 64, 8, 8, 16, 16, 32, 32, 64, 64, 16, 16, 32, 32, 64, 64]


def sort_itx_list(hdr, itx, tx_sizes):
    """Sort an inverse-transform work list (op indices) by size class and fill hdr.itx_class_end.
    tx_sizes[i] is the TX_SIZE of op itx[i].  Returns the sorted list (numpy uint32)."""
    import numpy as np
    itx = np.asarray(itx, np.uint32)
    md = np.array(TX_MAXDIM)[np.asarray(tx_sizes, np.int64)] if len(itx) else np.zeros(0, np.int64)
    cls = np.where(md <= 4, 0, np.where(md <= 8, 1, np.where(md <= 16, 2, 3)))
    order = np.argsort(cls, kind="stable")
    ends = np.cumsum(np.bincount(cls, minlength=4)) if len(itx) else np.zeros(4, np.int64)
    for k in range(4):
        hdr.itx_class_end[k] = int(ends[k])
    return itx[order]

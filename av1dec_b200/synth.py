"""Synthetic frames and frame parameters for the per-kernel configurations (BASELINE.json
config 4): random partitions / transform sizes / skip / reference classes,
deblock levels, CDEF presets and loop-restoration units, plus pixel planes in the distributions

    'U'  uniform [0,255] per sample
    'B'  blocky-smooth: per-8x8 DC ~ clip(N(128,40)) + per-sample U[-3,3]

PRNG: splitmix64, seed 20261018 (+ frame index), as BASELINE.md fixes it.  The output is a
command buffer in the engine's own format (av1dec_b200.format) -- the same bytes feed the GPU
stage entry points and, through oracle/oracle_shim.cpp, the reference's C++ filter classes.
"""
import ctypes as C

import numpy as np

from . import format as F

SEED = 20261018
_MASK = (1 << 64) - 1


class SplitMix64:
    """Vectorised splitmix64."""

    def __init__(self, seed):
        self.state = np.uint64(seed & _MASK)

    def u64(self, n):
        with np.errstate(over="ignore"):
            idx = np.arange(1, n + 1, dtype=np.uint64)
            z = self.state + idx * np.uint64(0x9E3779B97F4A7C15)
            self.state = z[-1] if n else self.state
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def randint(self, lo, hi, shape):
        """Uniform integers in [lo, hi]."""
        n = int(np.prod(shape))
        return (((self.u64(n) >> np.uint64(11)) % np.uint64(hi - lo + 1)).astype(np.int64) + lo).reshape(shape)

    def uniform(self, shape):
        n = int(np.prod(shape))
        return ((self.u64(n) >> np.uint64(11)).astype(np.float64) / float(1 << 53)).reshape(shape)

    def normal(self, shape):
        u1 = np.maximum(self.uniform(shape), 1e-12)
        u2 = self.uniform(shape)
        return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def _up(a, cell, rows, cols):
    return np.repeat(np.repeat(a, cell, axis=0), cell, axis=1)[:rows, :cols]


def make_planes(rng, w, h, dist="B"):
    """Planes covering the MI-aligned area of a w x h frame."""
    aw, ah = 8 * ((w + 7) // 8), 8 * ((h + 7) // 8)
    planes = []
    for sub in (0, 1, 1):
        pw, ph = aw >> sub, ah >> sub
        if dist == "U":
            p = rng.randint(0, 255, (ph, pw))
        else:
            bs = 8 >> sub
            dc = np.clip(128 + 40 * rng.normal(((ph + bs - 1) // bs, (pw + bs - 1) // bs)), 0, 255)
            p = np.clip(np.rint(_up(dc, bs, ph, pw)) + rng.randint(-3, 3, (ph, pw)), 0, 255)
        planes.append(p.astype(np.uint8))
    return planes


class SynthFrame:
    pass


def make_postfilter_frame(w, h, seed=SEED, dist="B", delta_lf=False, lr_unit=64, sb128=False, levels=None,
                          lr_types=(0, 1, 2)):
    """Random but structurally valid inputs for deblock + CDEF + loop restoration."""
    rng = SplitMix64(seed)
    mi_cols, mi_rows = 2 * ((w + 7) >> 3), 2 * ((h + 7) >> 3)
    s = SynthFrame()
    s.w, s.h, s.mi_cols, s.mi_rows, s.sb128 = w, h, mi_cols, mi_rows, sb128
    s.planes = make_planes(rng, w, h, dist)
    # ---- random partition: leaf level per MI (0: 64x64 ... 4: 4x4), optional 2:1 split of the leaf
    level = np.zeros((mi_rows, mi_cols), np.int64)
    props = []
    for l in range(5):
        cell = 16 >> l
        r, c = (mi_rows + cell - 1) // cell, (mi_cols + cell - 1) // cell
        split = _up(rng.uniform((r, c)) < (0.75, 0.6, 0.5, 0.35, 0.0)[l], cell, mi_rows, mi_cols)
        if l < 4:
            level = np.where((level == l) & split, l + 1, level)
        props.append({
            "shape": _up(rng.randint(0, 3, (r, c)) % 3 if l < 4 else np.zeros((r, c), np.int64), cell, mi_rows, mi_cols),
            "skip": _up((rng.uniform((r, c)) < 0.25).astype(np.int64), cell, mi_rows, mi_cols),
            "ref": _up(np.where(rng.uniform((r, c)) < 0.3, 0, rng.randint(1, 7, (r, c))), cell, mi_rows, mi_cols),
            "mode": _up(rng.randint(0, 1, (r, c)), cell, mi_rows, mi_cols),
        })
    pick = lambda key: np.choose(level, [p[key] for p in props])
    shape, skip, ref, mode = pick("shape"), pick("skip"), pick("ref"), pick("mode")
    sq_bs = np.array([12, 9, 6, 3, 0])[level]
    s.mi_size = (sq_bs - shape).astype(np.uint8)  # shape 1: S x S/2 (enum-1), 2: S/2 x S (enum-2)
    tx_y = np.choose(shape, [np.array([4, 3, 2, 1, 0])[level], np.array([12, 10, 8, 6, 0])[level],
                             np.array([11, 9, 7, 5, 0])[level]])
    tx_uv = np.choose(shape, [np.array([3, 2, 1, 0, 0])[level], np.array([10, 8, 6, 0, 0])[level],
                              np.array([9, 7, 5, 0, 0])[level]])
    s.tx = (tx_y | (tx_uv << 5) | (tx_uv << 10)).astype(np.uint16)
    s.flags = (skip | (mode << 1) | (ref << 2)).astype(np.uint8)
    s.delta_lf = np.zeros((mi_rows, mi_cols, 4), np.int8)
    if delta_lf:
        r, c = (mi_rows + 15) // 16, (mi_cols + 15) // 16
        for i in range(4):
            s.delta_lf[:, :, i] = _up(rng.randint(-8, 8, (r, c)), 16, mi_rows, mi_cols)
    lfmi = np.zeros((mi_rows, mi_cols), dtype=[("mi_size", "u1"), ("flags", "u1"), ("tx", "<u2"), ("dlf", "i1", 4)])
    lfmi["mi_size"], lfmi["flags"], lfmi["tx"], lfmi["dlf"] = s.mi_size, s.flags, s.tx, s.delta_lf
    # ---- frame-level parameters
    hdr = F.FrameHdr()
    hdr.frame_w, hdr.frame_h, hdr.mi_cols, hdr.mi_rows = w, h, mi_cols, mi_rows
    hdr.sb_log2 = 7 if sb128 else 6
    sb4 = 1 << (hdr.sb_log2 - 2)
    hdr.sb_cols, hdr.sb_rows = (mi_cols + sb4 - 1) // sb4, (mi_rows + sb4 - 1) // sb4
    lv = levels if levels is not None else [int(v) for v in rng.randint(0, 63, (4,))]
    if levels is None and lv[0] == 0 and lv[1] == 0:
        lv[0] = 17
    for i in range(4):
        hdr.lf.level[i] = lv[i]
    hdr.lf.sharpness = int(rng.randint(0, 7, (1,))[0])
    hdr.lf.delta_enabled = 1
    for i, v in enumerate((1, 0, 0, 0, -1, 0, -1, -1)):  # default ref deltas (Parser.cpp:1920-1930)
        hdr.lf.ref_deltas[i] = v
    hdr.lf.delta_lf_multi = 1 if delta_lf else 0
    hdr.cdef.enabled = 1
    hdr.cdef.damping = int(rng.randint(3, 6, (1,))[0])
    for i in range(8):
        hdr.cdef.y_pri[i] = int(rng.randint(0, 15, (1,))[0])
        hdr.cdef.uv_pri[i] = int(rng.randint(0, 15, (1,))[0])
        hdr.cdef.y_sec[i] = (0, 1, 2, 4)[int(rng.randint(0, 3, (1,))[0])]
        hdr.cdef.uv_sec[i] = (0, 1, 2, 4)[int(rng.randint(0, 3, (1,))[0])]
    r64, c64 = (mi_rows + 15) // 16, (mi_cols + 15) // 16
    idx64 = rng.randint(0, 7, (r64, c64))
    idx64 = np.where(rng.uniform((r64, c64)) < 0.05, -1, idx64).astype(np.int8)
    s.cdef_idx64 = idx64
    # cdef8: preset per 8x8, 0xFF where idx == -1 or all four MIs skip (Cdef.cpp:76-82)
    sk = skip.astype(bool)
    allskip = sk[0::2, 0::2] & sk[1::2, 0::2] & sk[0::2, 1::2] & sk[1::2, 1::2]
    idx8 = _up(idx64.astype(np.int64), 8, mi_rows // 2, mi_cols // 2)
    s.cdef8 = np.where((idx8 < 0) | allskip, 0xFF, idx8).astype(np.uint8)
    # ---- loop restoration units (LoopRestoration.cpp:35-38 count_units_in_frame)
    units = []
    hdr.lr.uses_lr = 1
    for p in range(3):
        sub = 1 if p else 0
        us = lr_unit if p == 0 else max(32, lr_unit >> 1)
        rows = max((((h + sub) >> sub) + (us >> 1)) // us, 1)
        cols = max((((w + sub) >> sub) + (us >> 1)) // us, 1)
        hdr.lr.frame_type[p] = 3  # RESTORE_SWITCHABLE: per-unit type
        hdr.lr.unit_size[p], hdr.lr.unit_rows[p], hdr.lr.unit_cols[p] = us, rows, cols
        hdr.lr.unit_first[p] = len(units)
        types = np.array(lr_types)[rng.randint(0, len(lr_types) - 1, (rows * cols,))]
        c0 = rng.randint(-5, 10, (rows * cols, 2)) if p == 0 else np.zeros((rows * cols, 2), np.int64)
        c1 = rng.randint(-23, 8, (rows * cols, 2))
        c2 = rng.randint(-17, 46, (rows * cols, 2))
        sets = rng.randint(0, 15, (rows * cols,))
        x0 = rng.randint(-96, 31, (rows * cols,))
        x1 = rng.randint(-32, 95, (rows * cols,))
        for k in range(rows * cols):
            u = F.LrUnit()
            u.type = int(types[k])
            u.sgr_set = int(sets[k])
            # r == 0 rules of read_lr_unit (Parser.cpp:2198-2206)
            st = int(sets[k])
            q0, q1 = int(x0[k]), int(x1[k])
            if st >= 10 and st <= 13:  # r0 == 0
                q0 = 0
            if st >= 14:  # r1 == 0
                q1 = int(np.clip(128 - q0, -32, 95))
            u.sgr_xqd[0], u.sgr_xqd[1] = q0, q1
            for ps in range(2):
                u.wiener[ps][0], u.wiener[ps][1], u.wiener[ps][2] = int(c0[k, ps]), int(c1[k, ps]), int(c2[k, ps])
            units.append(bytes(u))
    hdr.n_lru = len(units)
    s.cmd = F.build(hdr, {"off_lfmi": lfmi.tobytes(), "off_cdef8": s.cdef8.tobytes(), "off_lru": b"".join(units)})
    s.hdr = F.FrameHdr.from_buffer_copy(s.cmd[:C.sizeof(F.FrameHdr)])
    return s


# legal (tx_size, tx_type) pairs: 64-point -> DCT only, 32-point -> DCT / IDTX, else all 16
TX_W = [4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64]
TX_H = [4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16]


def legal_tx_types(tx_size):
    m = max(TX_W[tx_size], TX_H[tx_size])
    if m == 64:
        return [0]
    if m == 32:
        return [0, 9]
    return list(range(16))


def make_itx_batch(n, seed=SEED, lossless_frac=0.05, sizes=None):
    """n random transform blocks: (tx_size, tx_type, lossless, coef(int16 list), nz_rows, nz_cols)."""
    rng = SplitMix64(seed)
    out = []
    szs = rng.randint(0, 18, (n,)) if sizes is None else np.array(sizes)[rng.randint(0, len(sizes) - 1, (n,))]
    tys = rng.randint(0, 15, (n,))
    ll = rng.uniform((n,))
    for i in range(n):
        ts = int(szs[i])
        lossless = ts == 0 and ll[i] < lossless_frac
        legal = legal_tx_types(ts)
        tt = 0 if lossless else legal[int(tys[i]) % len(legal)]
        tw, th = min(TX_W[ts], 32), min(TX_H[ts], 32)
        area = tw * th
        eob = int(min(area, max(1, rng.randint(1, max(1, area // 4), (1,))[0])))
        coef = np.zeros(area, np.int64)
        # low-frequency-biased positions, Laplacian-ish magnitudes, ~1% at the +-2^15 clip
        pos = np.unique((rng.uniform((eob,)) ** 2 * area).astype(np.int64))
        mag = (-np.log(np.maximum(rng.uniform((len(pos),)), 1e-9)) * (200 if not lossless else 20)).astype(np.int64)
        sign = np.where(rng.uniform((len(pos),)) < 0.5, -1, 1)
        val = np.clip(mag * sign, -32768, 32767)
        clipsel = rng.uniform((len(pos),)) < 0.01
        val = np.where(clipsel, np.where(sign < 0, -32768, 32767), val)
        if lossless:
            val = np.clip(val, -2000, 2000)
        rows, cols = pos // tw, pos % tw
        coef[rows * tw + cols] = val
        nz = np.nonzero(coef)[0]
        if len(nz) == 0:
            coef[0] = 64
            nz = np.array([0])
        out.append((ts, tt, int(lossless), coef.astype(np.int16), int((nz // tw).max()) + 1, int((nz % tw).max()) + 1))
    return out


def make_itx_extent_batch(seed=SEED):
    """Transform blocks whose non-zero coefficients fill a given leading box: every transform size
    x every legal type x a ladder of (nz_rows, nz_cols) extents -- 1x1 (DC only), a few low
    frequencies, half, full -- so that each zero-aware variant of the generated butterflies
    (itx_gen.h, selected by the extents) runs in the row pass and in the column pass."""
    rng = SplitMix64(seed)
    out = []
    ladder = (1, 2, 4, 5, 8, 11, 16, 23, 32)
    k = 0
    for ts in range(19):
        tw, th = min(TX_W[ts], 32), min(TX_H[ts], 32)
        for tt in legal_tx_types(ts):
            for e, nzr in enumerate(ladder):
                if nzr > th:
                    break
                nzc = min(tw, ladder[(e + k) % len(ladder)])
                k += 1
                vals = rng.randint(-1500, 1500, (nzr, nzc))
                vals[nzr - 1, nzc - 1] = 777  # the extents are exact
                coef = np.zeros((th, tw), np.int64)
                coef[:nzr, :nzc] = vals
                out.append((ts, tt, 0, coef.reshape(-1).astype(np.int16), nzr, nzc))
    return out


def make_itx_cmd(batch):
    """Command buffer holding only inverse-transform work (ops + itx list + coefficient arena)."""
    ops, coefs, itx = [], [], []
    coef_off = res_off = 0
    for i, (ts, tt, lossless, coef, nzr, nzc) in enumerate(batch):
        op = F.Op()
        op.tx_size, op.tx_type, op.lossless = ts, tt, lossless
        op.flags = F.OPF_HAS_RESID
        op.nz_rows, op.nz_cols = nzr, nzc
        op.coef_off, op.res_off = coef_off, res_off
        coef_off += len(coef)
        res_off += TX_W[ts] * TX_H[ts]
        ops.append(bytes(op))
        coefs.append(coef.tobytes())
        itx.append(i)
    hdr = F.FrameHdr()
    hdr.frame_w = hdr.frame_h = 64
    hdr.mi_cols = hdr.mi_rows = 16
    hdr.sb_cols = hdr.sb_rows = 1
    hdr.sb_log2 = 6
    hdr.n_ops, hdr.n_itx, hdr.n_coef, hdr.n_res = len(ops), len(itx), coef_off, res_off
    itx_sorted = F.sort_itx_list(hdr, itx, [b[0] for b in batch])
    cmd = F.build(hdr, {"off_ops": b"".join(ops), "off_itx": itx_sorted.tobytes(), "off_coef": b"".join(coefs)})
    return cmd, res_off


def make_itx_frame(w, h, seed=SEED, coded_frac=0.7):
    """A whole frame of transform blocks (luma + both chroma planes) following a random
    partition: every leaf block carries one luma TB of its own size (<= 64x64) and chroma TBs,
    `coded_frac` of them coded with random coefficients.  Returns (cmd_bytes, n_tb, n_res, algo_bytes)
    where algo_bytes = sum over coded TBs of 2*tw*th (coefficients read) + 2*w*h (residual written)."""
    rng = SplitMix64(seed)
    mi_cols, mi_rows = 2 * ((w + 7) >> 3), 2 * ((h + 7) >> 3)
    level = np.zeros((mi_rows, mi_cols), np.int64)
    for l in range(4):
        cell = 16 >> l
        r, c = (mi_rows + cell - 1) // cell, (mi_cols + cell - 1) // cell
        split = _up(rng.uniform((r, c)) < (0.8, 0.65, 0.5, 0.4)[l], cell, mi_rows, mi_cols)
        level = np.where((level == l) & split, l + 1, level)
    tbs = []  # (plane, x, y, tx_size)
    for l in range(5):
        cell = 16 >> l  # MI units
        size = 64 >> l
        ys, xs = np.nonzero(level[::cell, ::cell][: (mi_rows + cell - 1) // cell, : (mi_cols + cell - 1) // cell] == l)
        tx = [4, 3, 2, 1, 0][l]
        for yy, xx in zip(ys.tolist(), xs.tolist()):
            tbs.append((0, xx * size, yy * size, tx))
            if l < 4:
                ctx = [3, 2, 1, 0][l]
                tbs.append((1, xx * size // 2, yy * size // 2, ctx))
                tbs.append((2, xx * size // 2, yy * size // 2, ctx))
            elif (xx & 1) and (yy & 1):
                tbs.append((1, (xx - 1) * 2, (yy - 1) * 2, 0))
                tbs.append((2, (xx - 1) * 2, (yy - 1) * 2, 0))
    n = len(tbs)
    coded = rng.uniform((n,)) < coded_frac
    tys = rng.randint(0, 15, (n,))
    ops = np.zeros(n, dtype=[("x", "<u2"), ("y", "<u2"), ("plane", "u1"), ("kind", "u1"), ("tx_size", "u1"), ("tx_type", "u1"),
                             ("mode", "u1"), ("angle", "i1"), ("flags", "u1"), ("fi", "u1"), ("cfl", "i1"), ("nzr", "u1"),
                             ("nzc", "u1"), ("lossless", "u1"), ("coef_off", "<u4"), ("res_off", "<u4"), ("aux", "<u4"),
                             ("mlw", "<u2"), ("mlh", "<u2")])
    assert ops.dtype.itemsize == 32
    coefs, itx = [], []
    coef_off = res_off = 0
    algo = 0
    for i, (pl, x, y, tx) in enumerate(tbs):
        ops[i]["x"], ops[i]["y"], ops[i]["plane"], ops[i]["tx_size"] = x, y, pl, tx
        if not coded[i]:
            continue
        legal = legal_tx_types(tx)
        tw, th = min(TX_W[tx], 32), min(TX_H[tx], 32)
        area = tw * th
        k = max(1, area // 8)
        pos = np.unique((rng.uniform((k,)) ** 2 * area).astype(np.int64))
        coef = np.zeros(area, np.int16)
        coef[pos] = np.clip((rng.normal((len(pos),)) * 120).astype(np.int64), -2000, 2000)
        coef[0] = 64
        nz = np.nonzero(coef)[0]
        ops[i]["tx_type"] = legal[int(tys[i]) % len(legal)]
        ops[i]["flags"] = F.OPF_HAS_RESID
        ops[i]["nzr"], ops[i]["nzc"] = int((nz // tw).max()) + 1, int((nz % tw).max()) + 1
        ops[i]["coef_off"], ops[i]["res_off"] = coef_off, res_off
        coef_off += area
        res_off += TX_W[tx] * TX_H[tx]
        algo += 2 * area + 2 * TX_W[tx] * TX_H[tx]
        coefs.append(coef.tobytes())
        itx.append(i)
    hdr = F.FrameHdr()
    hdr.frame_w, hdr.frame_h, hdr.mi_cols, hdr.mi_rows = w, h, mi_cols, mi_rows
    hdr.sb_log2 = 6
    hdr.sb_cols, hdr.sb_rows = (mi_cols + 15) // 16, (mi_rows + 15) // 16
    hdr.n_ops, hdr.n_itx, hdr.n_coef, hdr.n_res = n, len(itx), coef_off, res_off
    itx_sorted = F.sort_itx_list(hdr, itx, [tbs[i][3] for i in itx])
    cmd = F.build(hdr, {"off_ops": ops.tobytes(), "off_itx": itx_sorted.tobytes(), "off_coef": b"".join(coefs)})
    return cmd, len(itx), res_off, algo


def make_inter_frame(w, h, seed=SEED, compound_frac=0.25, max_mv=512, fixed_mv=None, fast=True, same_ref=False):
    """A whole frame of translational inter blocks (8x8 .. 64x64, random partition): random
    1/8-pel motion vectors within +-max_mv/8 samples, random dual interpolation filters, references
    in store slots 0 / 1 (RefFrame 1 / 2), `compound_frac` of the blocks compound-average.
    `fixed_mv` = (row, col) in 1/8 luma pel gives every block that vector for both lists;
    `fast=False` leaves the AV1B_IBF_FAST / AV1B_IPUF_FAST flags off (general kernel);
    `same_ref` makes both lists of a compound block read store slot 0.
    Returns (cmd_bytes, n_blocks, algo_bytes); algo_bytes = (1 + refs) * samples (read + write)."""
    rng = SplitMix64(seed)
    mi_cols, mi_rows = 2 * ((w + 7) >> 3), 2 * ((h + 7) >> 3)
    level = np.zeros((mi_rows, mi_cols), np.int64)
    for l in range(3):
        cell = 16 >> l
        r, c = (mi_rows + cell - 1) // cell, (mi_cols + cell - 1) // cell
        split = _up(rng.uniform((r, c)) < (0.8, 0.6, 0.45)[l], cell, mi_rows, mi_cols)
        level = np.where((level == l) & split, l + 1, level)
    xs_all, ys_all, sz_all = [], [], []
    for l in range(4):
        cell, size = 16 >> l, 64 >> l
        ys, xs = np.nonzero(level[::cell, ::cell] == l)
        xs_all.append(xs * size)
        ys_all.append(ys * size)
        sz_all.append(np.full(len(xs), size))
    bx, by, bs = np.concatenate(xs_all), np.concatenate(ys_all), np.concatenate(sz_all)
    nb = len(bx)
    comp = rng.uniform((nb,)) < compound_frac
    mv = rng.randint(-max_mv, max_mv, (nb, 2, 2))
    if fixed_mv is not None:
        mv[:, :, 0], mv[:, :, 1] = fixed_mv[0], fixed_mv[1]
    filt = rng.randint(0, 2, (nb, 2))
    ref0 = rng.randint(1, 2, (nb,))
    if same_ref:
        ref0[:] = 1
    ipu_t = np.dtype([("x", "<u2"), ("y", "<u2"), ("w", "u1"), ("h", "u1"), ("plane", "u1"), ("kind", "u1"),
                      ("mv", "<i2", (2, 2)), ("ref_slot", "i1", 2), ("ref_frame", "u1", 2), ("filt", "u1", 2),
                      ("warp", "u1", 2), ("flags", "u1"), ("comp_type", "u1"), ("fwd_w", "u1"), ("bck_w", "u1"), ("aux", "<u4")])
    assert ipu_t.itemsize == 32
    ipu = np.zeros(nb * 3, ipu_t)
    for pl in range(3):
        sub = 1 if pl else 0
        v = ipu[pl::3]
        v["x"], v["y"], v["w"], v["h"], v["plane"] = bx >> sub, by >> sub, bs >> sub, bs >> sub, pl
        v["mv"] = mv
        v["ref_frame"][:, 0] = ref0
        v["ref_frame"][:, 1] = 3 - ref0
        v["ref_slot"][:, 0] = ref0 - 1
        v["ref_slot"][:, 1] = np.where(comp, 0 if same_ref else 2 - ref0, -1)
        v["filt"] = filt
        v["flags"] = comp.astype(np.uint8) | (0x08 if fast else 0)  # AV1B_IPUF_COMPOUND, AV1B_IPUF_FAST
        v["comp_type"] = 2                  # AV1B_COMP_AVERAGE
        v["aux"] = 0xFFFFFFFF
    blk_t = np.dtype([("first_ipu", "<u4"), ("n_ipu", "<u2"), ("flags", "<u2"), ("x", "<u2"), ("y", "<u2"), ("cx", "<u2"),
                      ("cy", "<u2"), ("bw", "u1"), ("bh", "u1"), ("cw", "u1"), ("ch", "u1"), ("pad", "<u4")])
    assert blk_t.itemsize == 24
    blk = np.zeros(nb, blk_t)
    blk["first_ipu"], blk["n_ipu"], blk["flags"] = np.arange(nb) * 3, 3, 1 | (4 if fast else 0)  # HAS_CHROMA | FAST
    blk["x"], blk["y"], blk["cx"], blk["cy"] = bx, by, bx >> 1, by >> 1
    blk["bw"] = blk["bh"] = bs
    blk["cw"] = blk["ch"] = bs >> 1
    hdr = F.FrameHdr()
    hdr.frame_w, hdr.frame_h, hdr.mi_cols, hdr.mi_rows = w, h, mi_cols, mi_rows
    hdr.sb_log2 = 6
    hdr.sb_cols, hdr.sb_rows = (mi_cols + 15) // 16, (mi_rows + 15) // 16
    for rf in (1, 2):
        hdr.ref_slot[rf] = rf - 1
        hdr.ref_w[rf], hdr.ref_h[rf] = w, h
    hdr.n_iblk, hdr.n_ipu = nb, nb * 3
    cmd = F.build(hdr, {"off_iblk": blk.tobytes(), "off_ipu": ipu.tobytes()})
    samples = (bs.astype(np.int64) ** 2 * 3 // 2)
    algo = int((samples * (2 + comp.astype(np.int64))).sum())
    return cmd, nb, algo


def _morton(x, y):
    z = 0
    for b in range(8):
        z |= ((x >> b) & 1) << (2 * b) | ((y >> b) & 1) << (2 * b + 1)
    return z


def make_intra_frame(w, h, seed=SEED, sb_log2=6, intra_frac=0.6, levelled=True, sizes=(8, 16, 32), mode_set=None, fi=True, cfl=True, rect=False, segments=True):
    """A frame for the superblock wavefront (prediction only): a random `intra_frac` of the blocks
    of a non-intra frame are intra-predicted over whatever the frame already holds (the caller
    supplies it as the input picture, standing for the inter prediction), so every edge carries
    real data.  Per superblock one block size (8 / 16 / 32 luma, half of it chroma), blocks in
    raster order inside the superblock, every mode (DC, directional with angle deltas and the edge
    filter, smooth, Paeth, filter-intra, chroma-from-luma).  The availability flags follow the
    decoding order exactly -- above-right only where the neighbour precedes the block, below-left
    only at the left column next to a finished superblock -- so no op reads samples a later op
    overwrites, and the dependency levels (level | run length << 16 in res_off, ops stably sorted
    by level inside each superblock) are computed the way the host emitter does (4x4-cell map).
    `levelled=False` keeps the ops in decoding order, one per step (res_off = 0): the reference
    the level analysis is checked against.  `sizes` / `mode_set` / `fi` / `cfl` narrow the mix
    (profiling aid: tools/wave_prof.py); `rect` splits half of the blocks into two rectangular
    transform blocks (side by side or stacked, flags per half); `segments=False` leaves the overlap
    hints of every superblock zero (the classic two-superblock-lag wavefront).  Returns the command buffer."""
    rng = SplitMix64(seed)
    sb = 1 << sb_log2
    mi_cols, mi_rows = 2 * ((w + 7) >> 3), 2 * ((h + 7) >> 3)
    fw, fh = mi_cols * 4, mi_rows * 4  # MI-aligned plane size (luma)
    sb_cols, sb_rows = (fw + sb - 1) // sb, (fh + sb - 1) // sb
    op_t = np.dtype([("x", "<u2"), ("y", "<u2"), ("plane", "u1"), ("kind", "u1"), ("tx_size", "u1"), ("tx_type", "u1"),
                     ("mode", "u1"), ("angle", "i1"), ("flags", "u1"), ("fi", "u1"), ("cfl", "i1"), ("nzr", "u1"),
                     ("nzc", "u1"), ("lossless", "u1"), ("coef_off", "<u4"), ("res_off", "<u4"), ("aux", "<u4"),
                     ("mlw", "<u2"), ("mlh", "<u2")])
    assert op_t.itemsize == 32
    all_ops, sbs = [], []
    import os
    split_area = int(os.environ.get("AV1B200_WAVE_SPLIT", "512"))
    seed_pct = [int(v) for v in os.environ.get("AV1B200_WAVE_SEED", "50,50,80").split(",")]  # like host/emitter.cpp
    depth = {}
    sq_tx = {4: 0, 8: 1, 16: 2, 32: 3, 64: 4}
    rect_tx = {(4, 8): 5, (8, 4): 6, (8, 16): 7, (16, 8): 8, (16, 32): 9, (32, 16): 10, (32, 64): 11, (64, 32): 12}
    size_pick = rng.randint(0, 2, (sb_rows, sb_cols))
    sizes = tuple(sizes) if len(sizes) == 3 else tuple(sizes[i % len(sizes)] for i in range(3))
    for r in range(sb_rows):
        for c in range(sb_cols):
            bs = sizes[int(size_pick[r, c])]
            x0, y0 = c * sb, r * sb
            nbx, nby = min(sb, fw - x0) // bs, min(sb, fh - y0) // bs
            n_blk = nbx * nby
            pick = rng.uniform((max(n_blk, 1),)) < intra_frac
            modes = rng.randint(0, 12, (max(n_blk, 1), 2))
            deltas = rng.randint(-3, 3, (max(n_blk, 1), 2))
            misc = rng.randint(0, 255, (max(n_blk, 1), 4))
            ops = []
            for by in range(nby):
                for bx in range(nbx):
                    k = by * nbx + bx
                    if not pick[k]:
                        continue
                    for pl in range(3):
                        sub = 1 if pl else 0
                        n = sb >> sub
                        s = bs >> sub
                        x, y = (x0 >> sub) + bx * s, (y0 >> sub) + by * s
                        pw, ph = fw >> sub, fh >> sub
                        fl = 0
                        if x > 0:
                            fl |= F.OPF_HAVE_LEFT
                        if y > 0:
                            fl |= F.OPF_HAVE_ABOVE
                        top_row = by == 0
                        # above-right: in the superblock row above, or an earlier block of this
                        # superblock in DECODING (z) order -- what AV1's availability rule gives
                        if y > 0 and x + s < pw and (top_row or (bx + 1 < nbx and _morton(bx + 1, by - 1) < _morton(bx, by))):
                            fl |= F.OPF_HAVE_ABOVE_RIGHT
                        if bx == 0 and c > 0 and by + 1 < nby and y + 2 * s <= ph:
                            fl |= F.OPF_HAVE_BELOW_LEFT
                        mode = int(modes[k, 1 if pl else 0])
                        if mode_set is not None:
                            mode = mode_set[mode % len(mode_set)]
                        o = np.zeros((), op_t)
                        o["x"], o["y"], o["plane"], o["kind"], o["tx_size"] = x, y, pl, F.OP_INTRA, sq_tx[s]
                        o["mode"] = mode
                        if 1 <= mode <= 8:
                            o["angle"] = int(deltas[k, 1 if pl else 0])
                            if misc[k, 0] & 1:
                                fl |= F.OPF_EDGE_SMOOTH
                        if fi and pl == 0 and s <= 32 and misc[k, 1] < 48:
                            o["mode"], o["angle"], o["fi"] = 0, 0, int(misc[k, 2]) % 5
                            fl |= F.OPF_FILTER_INTRA
                        if cfl and pl > 0 and misc[k, 3] < 64:
                            o["mode"], o["angle"] = 0, 0
                            o["cfl"] = int(misc[k, 2 if pl == 1 else 1]) % 31 - 15
                            o["mlw"], o["mlh"] = x0 + bx * bs + bs, y0 + by * bs + bs
                            fl = (fl | F.OPF_CFL) & ~F.OPF_EDGE_SMOOTH
                        o["flags"] = fl
                        split = int(misc[k, 0] >> 1) & 3 if (rect and s >= 8 and not (fl & F.OPF_FILTER_INTRA and s > 32)) else 0
                        if split == 1:  # two transform blocks side by side (s/2 x s)
                            a, b = o.copy(), o.copy()
                            a["tx_size"] = b["tx_size"] = rect_tx[(s // 2, s)]
                            fa = (fl & ~F.OPF_HAVE_ABOVE_RIGHT) | (F.OPF_HAVE_ABOVE_RIGHT if y > 0 else 0)
                            fb = (fl | F.OPF_HAVE_LEFT) & ~F.OPF_HAVE_BELOW_LEFT
                            a["flags"], b["flags"] = fa, fb
                            b["x"] = x + s // 2
                            ops.extend([a, b])
                        elif split == 2:  # two transform blocks stacked (s x s/2)
                            a, b = o.copy(), o.copy()
                            a["tx_size"] = b["tx_size"] = rect_tx[(s, s // 2)]
                            fa = (fl & ~F.OPF_HAVE_BELOW_LEFT) | (F.OPF_HAVE_BELOW_LEFT if (fl & F.OPF_HAVE_LEFT) else 0)
                            fb = (fl | F.OPF_HAVE_ABOVE) & ~F.OPF_HAVE_ABOVE_RIGHT
                            a["flags"], b["flags"] = fa, fb
                            b["y"] = y + s // 2
                            ops.extend([a, b])
                        else:
                            ops.append(o)
            # large blocks as row strips of 256 samples, one op each (like host/emitter.cpp)
            if split_area > 0:
                split_ops = []
                for o in ops:
                    area = TX_W[int(o["tx_size"])] * TX_H[int(o["tx_size"])]
                    if int(o["flags"]) & F.OPF_FILTER_INTRA or area < max(512, split_area):
                        split_ops.append(o)
                        continue
                    lg = 1
                    while lg < 3 and (area >> (8 + lg)) > 1:
                        lg += 1
                    for i in range(1 << lg):
                        so = o.copy()
                        so["fi"] = (lg | (i << 2)) << 3
                        split_ops.append(so)
                ops = split_ops
            # dependency levels: 4x4-cell map per plane, like host/emitter.cpp scheduleSb
            cell = [np.zeros((sb >> 2, sb >> 2), np.int64), np.zeros((sb >> 3, sb >> 3), np.int64), np.zeros((sb >> 3, sb >> 3), np.int64)]
            levels = []
            seed_l2 = depth.get((r, c - 1), 0) * seed_pct[0] // 100 if (segments and c > 0) else 0
            d_ar = depth.get((r - 1, c + 1), 0) if (segments and r > 0) else 0
            seed_a1, seed_a2 = d_ar * seed_pct[1] // 100, d_ar * seed_pct[2] // 100
            for o in ops:
                if levels and not (int(o["flags"]) & F.OPF_FILTER_INTRA) and (int(o["fi"]) >> 5):
                    levels.append(levels[-1])  # a further strip of the block before
                    continue
                pl = int(o["plane"])
                sub = 1 if pl else 0
                nc = (sb >> sub) >> 2
                tw, th = TX_W[int(o["tx_size"])], TX_H[int(o["tx_size"])]
                x, y = int(o["x"]) - (x0 >> sub), int(o["y"]) - (y0 >> sub)

                def rd(p, cxa, cxb, cya, cyb, ncp):
                    cxa, cya, cxb, cyb = max(cxa, 0), max(cya, 0), min(cxb, ncp - 1), min(cyb, ncp - 1)
                    if cxa > cxb or cya > cyb:
                        return 0
                    return int(cell[p][cya:cyb + 1, cxa:cxb + 1].max())
                cx0, cy0, cx1, cy1 = x >> 2, y >> 2, (x + tw - 1) >> 2, (y + th - 1) >> 2
                lvl = rd(pl, cx0, cx1, cy0, cy1, nc)
                ar = 2 * tw if int(o["flags"]) & F.OPF_HAVE_ABOVE_RIGHT else tw
                bl = 2 * th if int(o["flags"]) & F.OPF_HAVE_BELOW_LEFT else th
                if y > 0:
                    lvl = max(lvl, rd(pl, (x - 1) >> 2, (x + ar - 1) >> 2, (y - 1) >> 2, (y - 1) >> 2, nc))
                if x > 0:
                    lvl = max(lvl, rd(pl, (x - 1) >> 2, (x - 1) >> 2, (y - 1) >> 2, (y + bl - 1) >> 2, nc))
                if int(o["flags"]) & F.OPF_CFL:
                    lvl = max(lvl, rd(0, (2 * x) >> 2, (2 * (x + tw) - 1) >> 2, (2 * y) >> 2, (2 * (y + th) - 1) >> 2, nc * 2))
                n_pl, q_pl = sb >> sub, (sb >> 1) >> sub
                if x == 0 and y + bl > q_pl:
                    lvl = max(lvl, seed_l2)
                if y == 0 and x + ar > n_pl:
                    lvl = max(lvl, seed_a2 if x + ar - n_pl > q_pl else seed_a1)
                lvl += 1
                cell[pl][cy0:cy1 + 1, cx0:cx1 + 1] = lvl
                levels.append(lvl)
            depth[(r, c)] = max(levels) if levels else 0
            if not levelled:
                sbs.append((len(all_ops), len(ops), 0, 0, 0, 0, 0, 0, 0))
                all_ops.extend(ops)
                continue
            # overlap hints (av1b200_format.h, Av1bSb): the first level that reads each half of the
            # left superblock's right column / the above-right superblock's bottom row, and the level
            # after which each early-published half of the own border is final -- like
            # host/emitter.cpp scheduleSb
            NEVER = 0xFF
            wl1 = wl2 = wa1 = wa2 = NEVER
            pr1 = pb1 = 1
            for o, lvl in zip(ops, levels):
                sub = 1 if int(o["plane"]) else 0
                n, q = sb >> sub, (sb >> 1) >> sub
                tw, th = TX_W[int(o["tx_size"])], TX_H[int(o["tx_size"])]
                xr, yr = int(o["x"]) - (x0 >> sub), int(o["y"]) - (y0 >> sub)
                fl = int(o["flags"])
                lv8 = min(lvl, 0xFE)
                if xr == 0 and c > 0:  # reads the left superblock's right column, rows yr-1 .. yr+bl-1
                    bl = 2 * th if fl & F.OPF_HAVE_BELOW_LEFT else th
                    if yr - 1 < q:
                        wl1 = min(wl1, lv8)
                    if yr + bl > q:
                        wl2 = min(wl2, lv8)
                if yr == 0 and r > 0:  # reads the row above, columns xr-1 .. xr+ar-1
                    ar = 2 * tw if fl & F.OPF_HAVE_ABOVE_RIGHT else tw
                    if xr + ar > n:
                        wa1 = min(wa1, lv8)
                        if xr + ar - n > q:
                            wa2 = min(wa2, lv8)
                if xr + tw == n and yr < q:
                    pr1 = max(pr1, lvl)
                if yr + th == n and xr < q:
                    pb1 = max(pb1, lvl)
            if not segments:
                wl1 = wl2 = wa1 = wa2 = pr1 = pb1 = 0
            elif pr1 > 0xFE or pb1 > 0xFE:
                pr1 = pb1 = 0
            order = np.argsort(np.asarray(levels, np.int64), kind="stable") if ops else []
            sorted_ops = [ops[i] for i in order]
            lv = [levels[i] for i in order]
            rem = 0
            for k in range(len(sorted_ops) - 1, -1, -1):
                rem = rem + 1 if (k + 1 < len(sorted_ops) and lv[k + 1] == lv[k]) else 1
                sorted_ops[k]["res_off"] = (lv[k] & 0xFFFF) | (min(rem, 0xFFFF) << 16)
            sbs.append((len(all_ops), len(sorted_ops), wl1, wl2, wa1, wa2, pr1, pb1, 0))
            all_ops.extend(sorted_ops)
    hdr = F.FrameHdr()
    hdr.frame_w, hdr.frame_h, hdr.mi_cols, hdr.mi_rows = w, h, mi_cols, mi_rows
    hdr.sb_log2, hdr.sb_cols, hdr.sb_rows = sb_log2, sb_cols, sb_rows
    hdr.enable_intra_edge_filter, hdr.frame_is_intra = 1, 0
    hdr.n_sb, hdr.n_ops = len(sbs), len(all_ops)
    ops_blob = np.array(all_ops, op_t).tobytes() if all_ops else b""
    sb_blob = np.array(sbs, np.dtype([("first", "<u4"), ("n", "<u4"), ("wl1", "u1"), ("wl2", "u1"), ("wa1", "u1"), ("wa2", "u1"), ("pr1", "u1"), ("pb1", "u1"), ("pad", "<u2")])).tobytes()
    return F.build(hdr, {"off_sb": sb_blob, "off_ops": ops_blob})

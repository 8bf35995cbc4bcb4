"""Parity checks shared by the GPU tests (product CUDA library) and the CPU tests (the test-only
emulation build of the same kernel sources, tests/emu/).  Every check compares the engine with
the reference's own C++ code through oracle.py, bit-exactly."""
import ctypes as C
import hashlib
import os

import numpy as np

import av1dec_b200 as pkg
import oracle
from av1dec_b200 import synth
from av1dec_b200.engine import Engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_ENGINE = os.path.join(ROOT, "tests", "emu", "libav1b200_emu.so")
EMU_DECODER = os.path.join(ROOT, "tests", "emu", "libav1b200dec_emu.so")


def emu_available():
    return os.path.exists(EMU_ENGINE) and os.path.exists(EMU_DECODER)


def emu_engine():
    return pkg.load_engine(EMU_ENGINE)


def emu_decoder():
    emu_engine()
    return pkg.load_decoder(EMU_DECODER)


def flip_residual(res, tx_type):
    ud = tx_type in (4, 8, 14, 6)
    lr = tx_type in (5, 7, 15, 6)
    if ud:
        res = res[::-1, :]
    if lr:
        res = res[:, ::-1]
    return res


def check_itx(lib, n=400, seed=synth.SEED, sizes=None, extents=False):
    batch = synth.make_itx_extent_batch(seed=seed) if extents else synth.make_itx_batch(n, seed=seed, sizes=sizes)
    cmd, n_res = synth.make_itx_cmd(batch)
    eng = Engine(64, 64, lib=lib)
    eng.submit(cmd, stages=pkg.STAGE_ITX)
    got = eng.residual(n_res)
    eng.close()
    want = oracle.inverse_transform(batch)
    off = 0
    for i, b in enumerate(batch):
        w, h = synth.TX_W[b[0]], synth.TX_H[b[0]]
        ref = np.clip(flip_residual(want[i], b[1]), -32768, 32767).astype(np.int16)
        mine = got[off:off + w * h].reshape(h, w)
        assert np.array_equal(ref, mine), f"itx mismatch block {i}: tx_size={b[0]} tx_type={b[1]} lossless={b[2]}"
        off += w * h
    return len(batch)


POST_STAGE_BITS = {1: pkg.STAGE_DEBLOCK, 2: pkg.STAGE_CDEF, 4: pkg.STAGE_LR}


def run_postfilter(lib, s, stages):
    """stages: oracle bit mask (1 deblock, 2 cdef, 4 lr).  Returns engine planes (MI-aligned)."""
    aw, ah = s.mi_cols * 4, s.mi_rows * 4
    eng = Engine(aw, ah, lib=lib)
    eng.set_input(s.planes, aw, ah)
    mask = 0
    for bit, st in POST_STAGE_BITS.items():
        if stages & bit:
            mask |= st
    fid = eng.submit(s.cmd, stages=mask)
    out = eng.download(fid, aw, ah)
    eng.close()
    return out


def run_inter(lib, w, h, refs, **kw):
    """One synthetic frame of translational inter blocks (synth.make_inter_frame) predicted from
    `refs` (planes per store slot).  Returns the reconstructed planes."""
    from av1dec_b200 import STAGE_INTER
    cmd, _, _ = synth.make_inter_frame(w, h, **kw)
    eng = Engine(w, h, lib=lib)
    for slot, planes in enumerate(refs):
        eng.set_ref(slot, planes, w, h)
    fid = eng.submit(cmd, stages=STAGE_INTER)
    out = eng.download(fid, w, h)
    eng.close()
    return out


def check_inter_properties(lib, w, h):
    """Size-independent properties of motion compensation (no oracle needed):
    a zero vector copies the reference, an integer vector copies a shifted window, the average of
    two identical predictions is that prediction, and the fast kernel equals the general one on
    random sub-pel vectors (including windows that leave the frame)."""
    rng = synth.SplitMix64(synth.SEED + 5)
    refs = [synth.make_planes(rng, w, h, "B"), synth.make_planes(rng, w, h, "B")]
    # (1) zero motion, mixed single / compound from the same reference: identity
    out = run_inter(lib, w, h, refs, fixed_mv=(0, 0), same_ref=True)
    for p in range(3):
        assert np.array_equal(out[p], refs[0][p]), f"zero-motion plane {p}"
    # (2) integer motion (+16 rows, -24 columns luma): shifted copy inside the frame
    dy, dx = 16, -24
    out = run_inter(lib, w, h, refs, fixed_mv=(dy * 8, dx * 8), same_ref=True, compound_frac=0.0)
    for p in range(3):
        sy, sx = (dy, dx) if p == 0 else (dy // 2, dx // 2)
        hh, ww = out[p].shape
        assert np.array_equal(out[p][0:hh - sy, -sx:ww], refs[0][p][sy:hh, 0:ww + sx]), f"integer-motion plane {p}"
    # (3) fast kernel == general kernel, random sub-pel vectors up to 64 samples, 30 % compound
    a = run_inter(lib, w, h, refs, seed=synth.SEED + 9, compound_frac=0.3, fast=True)
    b = run_inter(lib, w, h, refs, seed=synth.SEED + 9, compound_frac=0.3, fast=False)
    for p in range(3):
        assert np.array_equal(a[p], b[p]), f"fast vs general kernel, plane {p}"


def check_inter_vs_oracle(lib, w, h, **kw):
    """Motion compensation against the REFERENCE's Block::InterPredict::predict_inter
    (oracle_predict_inter) on a synthetic frame of translational blocks: random 1/8-pel vectors
    (windows leaving the frame included), the four interpolation filters per direction, single and
    compound-average prediction, block sizes 8..64 -- at full frame size."""
    import oracle
    rng = synth.SplitMix64(synth.SEED + 5)
    refs = [synth.make_planes(rng, w, h, "B"), synth.make_planes(rng, w, h, "B")]
    cmd, nb, _ = synth.make_inter_frame(w, h, **kw)
    got = run_inter(lib, w, h, refs, **kw)
    want, n = oracle.predict_inter(cmd, refs, refs[0])
    assert n == 3 * nb
    for p in range(3):
        hh, ww = got[p].shape
        assert np.array_equal(got[p], want[p][:hh, :ww]), f"inter vs reference InterPredict: plane {p}, {int((got[p] != want[p][:hh, :ww]).sum())} samples differ"
    return n


def run_wave(lib, w, h, planes, cmd):
    """Superblock wavefront alone over the input picture `planes`.  Returns the planes."""
    from av1dec_b200 import STAGE_WAVE
    eng = Engine(w, h, lib=lib)
    eng.set_input(planes, w, h)
    fid = eng.submit(cmd, stages=STAGE_WAVE)
    out = eng.download(fid, w, h)
    eng.close()
    return out


def check_wave(lib, w, h, sb_log2, ref_lib=None):
    """Superblock wavefront on a synthetic frame (synth.make_intra_frame): the level-scheduled
    command buffer through `lib` must equal the same ops in plain decoding order through
    `ref_lib` (default: `lib` itself).  Under emulation the ops of a level run in reverse order, so
    this checks the level analysis; on the GPU it checks the cross-superblock synchronisation,
    the shared-memory tile and the warp-per-op execution at full frame size."""
    rng = synth.SplitMix64(synth.SEED + 11)
    planes = synth.make_planes(rng, w, h, "B")
    seq = run_wave(ref_lib or lib, w, h, planes, synth.make_intra_frame(w, h, sb_log2=sb_log2, levelled=False))
    lev = run_wave(lib, w, h, planes, synth.make_intra_frame(w, h, sb_log2=sb_log2, levelled=True))
    changed = 0
    for p in range(3):
        assert np.array_equal(seq[p], lev[p]), f"wavefront sb_log2={sb_log2} plane {p}"
        changed += int((lev[p] != planes[p]).sum())
    assert changed > w * h // 4  # the ops really did something
    return lev


def check_wave_vs_oracle(lib, w, h, sb_log2, seed=0, **kw):
    """Superblock wavefront against the REFERENCE's Block::IntraPredict (oracle_predict_intra) on a
    synthetic frame: every mode x size the generator draws, edge filter / upsampling, filter-intra,
    chroma-from-luma, frame-edge clamps -- at frame sizes whose coordinates run into the thousands."""
    import oracle
    rng = synth.SplitMix64(synth.SEED + 11 + seed)
    planes = synth.make_planes(rng, w, h, "B")
    cmd = synth.make_intra_frame(w, h, seed=synth.SEED + seed, sb_log2=sb_log2, levelled=True, **kw)
    got = run_wave(lib, w, h, planes, cmd)
    mw, mh = 2 * ((w + 7) >> 3) * 4, 2 * ((h + 7) >> 3) * 4
    full = [np.zeros((mh >> (1 if p else 0), mw >> (1 if p else 0)), np.uint8) for p in range(3)]
    for p in range(3):
        full[p][:planes[p].shape[0], :planes[p].shape[1]] = planes[p]
    want, n = oracle.predict_intra(cmd, full)
    assert n > 0
    for p in range(3):
        hh, ww = got[p].shape
        assert np.array_equal(got[p], want[p][:hh, :ww]), f"wavefront vs reference IntraPredict: plane {p}, {int((got[p] != want[p][:hh, :ww]).sum())} samples differ"
    return n


def check_output_paths(lib, w, h):
    """Zero-copy device view and NV12 conversion of a frame equal its planar download."""
    import ctypes as C
    from av1dec_b200 import format as F
    rng = synth.SplitMix64(synth.SEED + 21)
    planes = synth.make_planes(rng, w, h, "U")
    hdr = F.FrameHdr()
    hdr.frame_w, hdr.frame_h = w, h
    hdr.mi_cols, hdr.mi_rows = 2 * ((w + 7) >> 3), 2 * ((h + 7) >> 3)
    hdr.sb_log2, hdr.sb_cols, hdr.sb_rows = 6, (hdr.mi_cols + 15) // 16, (hdr.mi_rows + 15) // 16
    eng = Engine(w, h, lib=lib)
    eng.set_input(planes, w, h)
    fid = eng.submit(F.build(hdr, {}), stages=0)  # no stage: the input picture is the frame
    ptrs, pitches = eng.device_view(fid)
    for p in range(3):
        ph, pw = (h, w) if p == 0 else (h >> 1, w >> 1)
        got = eng.read_device(ptrs[p], pitches[p] * ph).reshape(ph, pitches[p])[:, :pw]
        assert np.array_equal(got, planes[p][:ph, :pw]), f"device view plane {p}"
    y, uv = eng.to_nv12(fid, w, h)
    assert np.array_equal(y, planes[0][:h, :w])
    assert np.array_equal(uv[:, 0::2], planes[1][:h >> 1, :w >> 1]) and np.array_equal(uv[:, 1::2], planes[2][:h >> 1, :w >> 1])
    eng.close()


def check_postfilter(lib, w, h, stages, **kw):
    s = synth.make_postfilter_frame(w, h, **kw)
    got = run_postfilter(lib, s, stages)
    want = oracle.postfilter(s, stages)
    visible = bool(stages & 6)
    for p in range(3):
        cw, ch = (w, h) if visible else (s.mi_cols * 4, s.mi_rows * 4)
        if p:
            cw, ch = cw >> 1, ch >> 1
        a, b = got[p][:ch, :cw], want[p][:ch, :cw]
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            raise AssertionError(f"postfilter stages={stages} plane {p}: {len(bad)} samples differ, first at (y,x)={tuple(bad[0])} "
                                 f"got {a[tuple(bad[0])]} want {b[tuple(bad[0])]}")
    return s


def stream_md5(dec_lib, path, device=0):
    data = open(path, "rb").read()
    yuv, frames, pixels = pkg.decode_ivf(data, device=device, lib=dec_lib)
    return hashlib.md5(yuv).hexdigest(), frames, pixels


def stage_frames(dec_lib, data, stages):
    """Shown frames (visible area, concatenated) from the engine with a restricted stage mask."""
    yuv, frames, _ = pkg.decode_ivf(data, stages=stages, lib=dec_lib)
    return yuv, frames


IVD_DRIVE = os.path.join(ROOT, "tests", "native", "ivd_drive")


def drive_ivideodecoder(decoder_so, name, mode, tmp_path):
    """Run the libyami-style client (tests/native/ivd_drive.cpp: dlopen + createVideoDecoder +
    start / decode / getOutput / flush / reset / stop through the IVideoDecoder vtable).
    Returns (md5 of the written I420, frames)."""
    import subprocess
    out = os.path.join(str(tmp_path), f"ivd_{mode}.yuv")
    r = subprocess.run([IVD_DRIVE, decoder_so, os.path.join(ROOT, "tests", "golden", "bits", name), out, mode],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, f"ivd_drive {mode}: rc={r.returncode} {r.stderr[-400:]}"
    frames = int([l for l in r.stdout.splitlines() if l.startswith("frames")][0].split()[1])
    return hashlib.md5(open(out, "rb").read()).hexdigest(), frames, open(out, "rb").read()

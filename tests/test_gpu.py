"""Parity tests proper: the CUDA path (through the C ABI) against the reference, bit-exact.
Run on the B200 box: python -m pytest tests -m gpu."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import av1dec_b200 as pkg
import checks
import oracle
from av1dec_b200 import synth
from conftest import BITS, ROOT, all_streams

pytestmark = pytest.mark.gpu
need_oracle = pytest.mark.skipif(not oracle.available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def dec():
    return pkg.load_decoder()


@pytest.fixture(scope="module")
def eng():
    return pkg.load_engine()


def test_backend_is_cuda(eng):
    assert eng.av1b_backend() == b"cuda-sm_100a"


def test_every_conformance_stream_md5_exact(dec, md5_table):
    """BASELINE.json config 2: the full bits/ set through the GPU reconstruction path."""
    bad, pixels = [], 0
    for f in all_streams():
        got, frames, px = checks.stream_md5(dec, os.path.join(BITS, f))
        pixels += px
        if got != md5_table[f]:
            bad.append(f)
    assert not bad, f"{len(bad)} streams differ: {bad[:10]}"
    assert pixels > 23_000_000


def test_dropin_cli_unchanged_reference_main(md5_table, tmp_path):
    """The reference's own tests/Av1Dec.cpp, compiled unchanged against this library."""
    cli = os.path.join(ROOT, "av1dec_b200", "bin", "av1dec")
    for name in ("av1-1-b8-06-mfmv.ivf", "av1-1-b8-02-allintra.ivf", "Halo_426x240_1frames_intrabc.ivf"):
        out = tmp_path / "o.yuv"
        subprocess.run([cli, "-i", os.path.join(BITS, name), str(out)], check=True, stdout=subprocess.DEVNULL, timeout=300)
        assert hashlib.md5(out.read_bytes()).hexdigest() == md5_table[name], name


@pytest.mark.parametrize("env", [{"AV1B200_WAVE_OVERLAP": "0"}, {"AV1B200_WAVE_SPLIT": "0", "AV1B200_WAVE_SEED": "0,0,0", "AV1B200_SPIN_WAIT": "1"},
                                 {"AV1B200_WAVE_SEED": "90,90,100", "AV1B200_WAVE_WARPS": "16"}])
def test_scheduling_knobs_read_at_start_up(md5_table, tmp_path, env):
    """The emitter's scheduling choices (overlap hints, level seeds, row strips) and the host wait
    mode are read once per process: each setting decodes intra-heavy and inter streams through the
    CLI in a process of its own and must give the same pictures."""
    cli = os.path.join(ROOT, "av1dec_b200", "bin", "av1dec")
    for name in ("av1-1-b8-02-allintra.ivf", "av1-1-b8-06-mfmv.ivf", "av1-1-b8-00-quantizer-20.ivf", "av1-1-b8-04-cdfupdate.ivf"):
        if name not in md5_table:
            continue
        out = tmp_path / "o.yuv"
        subprocess.run([cli, "-i", os.path.join(BITS, name), str(out)], check=True, stdout=subprocess.DEVNULL, timeout=300, env={**os.environ, **env})
        assert hashlib.md5(out.read_bytes()).hexdigest() == md5_table[name], (name, env)


@pytest.mark.skipif(not os.path.exists(checks.IVD_DRIVE), reason="tests/native/ivd_drive not built")
def test_ivideodecoder_vtable(md5_table, tmp_path):
    """interface/VideoDecoderInterface.h:31-68 through dlopen + createVideoDecoder on the product
    library: start / decode / getOutput, flush + reset, stop + start; MD5-checked."""
    for name, n in (("av1-1-b8-06-mfmv.ivf", 4), ("av1-1-b8-02-allintra.ivf", 39)):
        got, frames, one = checks.drive_ivideodecoder(pkg.decoder_path(), name, "plain", tmp_path)
        assert (got, frames) == (md5_table[name], n), name
    name = "av1-1-b8-06-mfmv.ivf"
    got, frames, one = checks.drive_ivideodecoder(pkg.decoder_path(), name, "plain", tmp_path)
    got, frames, _ = checks.drive_ivideodecoder(pkg.decoder_path(), name, "flush", tmp_path)
    assert (got, frames) == (md5_table[name], 4)
    _, frames, two = checks.drive_ivideodecoder(pkg.decoder_path(), name, "twice", tmp_path)
    assert frames == 8 and two == one + one


def test_decoder_class_decode_getoutput(dec, md5_table):
    name = "av1-1-b8-03-sizeup.ivf" if os.path.exists(os.path.join(BITS, "av1-1-b8-03-sizeup.ivf")) else "av1-1-b8-04-cdfupdate.ivf"
    data = open(os.path.join(BITS, name), "rb").read()
    d = pkg.Decoder()
    md5 = hashlib.md5()
    for unit in pkg.iter_ivf(data):
        assert d.decode(unit)
        while True:
            o = d.get_output()
            if o is None:
                break
            for p in o[2]:
                md5.update(p)
    d.close()
    assert md5.hexdigest() == md5_table[name]


@need_oracle
@pytest.mark.parametrize("name", ["foreman_qcif_i.ivf", "av1-1-b8-02-allintra.ivf"])
def test_stage_boundaries_vs_reference_dumps(dec, name):
    """Per-plane equality after reconstruction, deblock, CDEF and LR on intra-only streams (each
    frame is independent of the filtered references, so restricted stage masks stay comparable)."""
    data = open(os.path.join(BITS, name), "rb").read()
    masks = [pkg.STAGE_RECON, pkg.STAGE_RECON | pkg.STAGE_DEBLOCK, pkg.STAGE_RECON | pkg.STAGE_DEBLOCK | pkg.STAGE_CDEF, pkg.STAGE_ALL]
    for stage, mask in enumerate(masks):
        got, n = checks.stage_frames(dec, data, mask)
        want, n2 = oracle.decode_stages(data, stage)
        assert n == n2
        if stage >= 2:
            assert got == want, f"{name}: stage {stage} differs"
        else:
            # oracle dumps the MI-aligned area for stages 0/1; compare its visible sub-rectangle
            from av1dec_b200 import iter_ivf  # noqa: F401
            w, h = (176, 144) if name.startswith("foreman") else (352, 288)
            aw, ah = 8 * ((w + 7) // 8), 8 * ((h + 7) // 8)
            assert (aw, ah) == (w, h)
            assert got == want, f"{name}: stage {stage} differs"


@need_oracle
def test_inverse_transform_vs_reference(eng):
    assert checks.check_itx(eng, n=20000) == 20000
    assert checks.check_itx(eng, extents=True) > 800  # every zero-aware butterfly variant, both passes
    for ts in range(19):
        checks.check_itx(eng, n=64, seed=7000 + ts, sizes=[ts])


@need_oracle
@pytest.mark.parametrize("w,h,kw", [
    (200, 136, {}),
    (226, 226, {"delta_lf": True}),
    (64, 64, {"dist": "U"}),
    (18, 34, {"lr_unit": 256}),
    (352, 288, {"lr_unit": 128, "sb128": True}),
    (1920, 1080, {"dist": "U", "delta_lf": True, "lr_unit": 256}),
    (1920, 1080, {}),
])
def test_postfilter_stages_vs_reference(eng, w, h, kw):
    for stages in (1, 2, 4, 7):
        checks.check_postfilter(eng, w, h, stages, **kw)


@need_oracle
def test_postfilter_chain_4k_vs_reference(eng):
    """BASELINE.json config 4 at full size: 3840x2160 deblock + CDEF + LR, bit-exact."""
    checks.check_postfilter(eng, 3840, 2160, 7)


def test_postfilter_4k_is_deterministic_and_launches_kernels(eng):
    s = synth.make_postfilter_frame(3840, 2160, seed=synth.SEED + 1)
    a = checks.run_postfilter(eng, s, 7)
    b = checks.run_postfilter(eng, s, 7)
    for p in range(3):
        assert np.array_equal(a[p], b[p])
    # LR with every unit RESTORE_NONE and CDEF off must return the deblocked frame unchanged
    s2 = synth.make_postfilter_frame(640, 360, lr_types=(0,))
    only_deblock = checks.run_postfilter(eng, s2, 1)
    with_lr = checks.run_postfilter(eng, s2, 1 | 4)
    for p in range(3):
        h, w = (360, 640) if p == 0 else (180, 320)
        assert np.array_equal(only_deblock[p][:h, :w], with_lr[p][:h, :w])


@pytest.mark.parametrize("env", [{"AV1B200_LANES": "1", "AV1B200_GOP_THREADS": "1"},
                                 {"AV1B200_LANES": "16", "AV1B200_GOP_THREADS": "8"},
                                 {"AV1B200_LANES": "3", "AV1B200_WAVE_WARPS": "16"},
                                 {"AV1B200_LANES": "8", "AV1B200_WAVE_WARPS": "4", "AV1B200_WAVE_GRIDQ": "4"}])
def test_concurrency_knobs_do_not_change_pixels(dec, md5_table, env, monkeypatch):
    """Frame lanes, closed-segment workers and the wavefront build only reorder work: an
    all-intra stream (39 independent frames), an inter stream with compound / OBMC references and
    an intrabc frame must come out bit-identical under every setting."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    pkg.load_engine().av1b_pool_purge()  # pooled contexts keep the lane count they were created with
    try:
        for name in ("av1-1-b8-02-allintra.ivf", "av1-1-b8-06-mfmv.ivf", "av1-1-b8-04-cdfupdate.ivf",
                     "Halo_426x240_1frames_intrabc.ivf"):
            got, _, _ = checks.stream_md5(dec, os.path.join(BITS, name))
            assert got == md5_table[name], (name, env)
    finally:
        pkg.load_engine().av1b_pool_purge()  # contexts created under these settings must not be recycled


def test_segments_of_all_intra_stream(dec):
    data = open(os.path.join(BITS, "av1-1-b8-02-allintra.ivf"), "rb").read()
    assert pkg.ivf_segments(data) == list(range(39))


@pytest.mark.parametrize("w,h", [(640, 384), (3840, 2160)])
def test_inter_prediction_properties(eng, w, h):
    """Motion compensation on whole synthetic frames (4K = BASELINE's per-kernel size): identity,
    shifted copy, and the fast kernel against the general one."""
    checks.check_inter_properties(eng, w, h)


@pytest.mark.skipif(not checks.emu_available(), reason="tests/emu not built")
@pytest.mark.parametrize("w,h,sb_log2", [(1920, 1080, 6), (1920, 1080, 7), (3840, 2160, 6)])
def test_wavefront_full_frames_vs_sequential_emulation(eng, w, h, sb_log2):
    """The wavefront kernel at sizes no conformance stream reaches (up to 60x34 superblocks in
    flight): GPU level-scheduled execution against the emulation's sequential one."""
    checks.check_wave(eng, w, h, sb_log2, ref_lib=checks.emu_engine())


@pytest.mark.parametrize("warps", ["8", "16", "4"])
def test_wavefront_hand_offs_are_race_free(eng, warps, monkeypatch):
    """The overlapping wavefront (early hand-offs by one warp, look-ahead waits, progress words)
    run 8 times over a dense 1080p intra frame under each CTA shape: a lost ordering between a
    border store and the progress word, or a halo read too early, shows up as a differing run."""
    monkeypatch.setenv("AV1B200_WAVE_WARPS", warps)
    w, h = 1920, 1080
    rng = synth.SplitMix64(synth.SEED + 31)
    planes = synth.make_planes(rng, w, h, "B")
    cmd = synth.make_intra_frame(w, h, sb_log2=6, intra_frac=0.95, rect=True, sizes=(8, 16, 32, 64))
    first = checks.run_wave(eng, w, h, planes, cmd)
    for _ in range(7):
        again = checks.run_wave(eng, w, h, planes, cmd)
        for p in range(3):
            assert np.array_equal(first[p], again[p]), f"wavefront not deterministic ({warps} warps, plane {p})"


@pytest.mark.skipif(not oracle.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("w,h,kw", [(1920, 1080, dict(compound_frac=0.3)), (1920, 1080, dict(compound_frac=0.3, fast=False, max_mv=2048, seed=99)),
                                    (3840, 2160, dict(compound_frac=0.25, max_mv=1024)), (3840, 2160, dict(compound_frac=0.5, fast=False, seed=5))])
def test_inter_prediction_vs_reference(eng, w, h, kw):
    """The inter pass (fast translational kernel and the general predictor) against the REFERENCE's
    Block::InterPredict::predict_inter (oracle_predict_inter) at 1080p and 4K: random 1/8-pel
    vectors including windows that leave the frame, every interpolation filter pair, single and
    compound-average prediction, blocks of 8..64 samples."""
    checks.check_inter_vs_oracle(eng, w, h, **kw)


@pytest.mark.skipif(not oracle.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("w,h,sb_log2,kw", [(1920, 1080, 6, {}), (1920, 1080, 6, dict(segments=False, intra_frac=0.9)),
                                             (1920, 1080, 7, dict(rect=True, intra_frac=1.0)),
                                             (3840, 2160, 6, dict(rect=True, sizes=(8, 32, 64), intra_frac=0.9)),
                                             (3840, 2160, 7, dict(rect=True, sizes=(16, 32, 64)))])
def test_intra_prediction_vs_reference(eng, w, h, sb_log2, kw):
    """The wavefront kernel against the REFERENCE's Block::IntraPredict::predict_intra /
    predict_chroma_from_luma (oracle_predict_intra) at 1080p and 4K: every mode, square and
    rectangular transform sizes up to 64 (blocks of 512 samples and more run as row strips on
    several warps), edge filter / upsampling, filter-intra, CfL -- with coordinates in the thousands
    (16-bit packed op fields); overlapping superblocks (hints + seeded levels) and, with
    `segments=False`, the classic two-superblock-lag schedule."""
    checks.check_wave_vs_oracle(eng, w, h, sb_log2, **kw)


@pytest.mark.parametrize("w,h", [(178, 94), (3840, 2160)])
def test_device_view_and_nv12(eng, w, h):
    checks.check_output_paths(eng, w, h)


@pytest.mark.parametrize("name", ["av1-1-b8-06-mfmv.ivf", "av1-1-b8-02-allintra.ivf"])
def test_resident_replay_captured_as_cuda_graph(name, md5_table):
    """A stream's resident replay captured into a CUDA graph (av1b_set_capture) leaves the same
    last frame as the same submits made directly -- with eight lanes forking and joining inside
    the capture."""
    import ctypes as C
    import sys
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from av1dec_b200 import format as F
    from av1dec_b200.engine import Engine
    data = open(os.path.join(BITS, name), "rb").read()
    rs = bench.record_stream(pkg, None, name, data, md5_table[name], 0)
    side = torch.cuda.Stream()
    eng = Engine(rs.max_w, rs.max_h, device=0, stream=side.cuda_stream)
    eng.set_lanes(8)
    hdr_size = C.sizeof(F.FrameHdr)
    frames = [(None, None, r, n) if b is None else (eng.upload(b), b[:hdr_size], r, -1) for b, n, r, s in rs.host_frames]

    def once():
        last = -1
        for ptr, hdr, refresh, slot in frames:
            last = eng.show_existing(slot, refresh) if ptr is None else eng.submit_resident(ptr, hdr, pkg.STAGE_ALL, refresh)
        eng.join()
        return last
    for _ in range(5):  # every lane gets used, every buffer exists
        fid = once()
    eng.sync()
    want = eng.download(fid, rs.max_w, rs.max_h)
    eng.set_capture(True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
        fid_g = once()
    eng.set_capture(False)
    for _ in range(2):
        with torch.cuda.stream(side):
            g.replay()
    side.synchronize()
    got = eng.download(fid_g, rs.max_w, rs.max_h)
    for p in range(3):
        assert np.array_equal(want[p], got[p]), (name, p)
    eng.close()

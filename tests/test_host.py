"""CPU-side tests: the C-ABI libraries load and export every declared symbol, the command-format
mirror matches the C structs, the product carries no CPU pixel path, and the host logic (command
emitter + decoder lifecycle) is bit-exact when the SAME kernel sources run under the test-only
emulation build (tests/emu)."""
import ctypes as C
import hashlib
import os
import re
import subprocess

import pytest

import av1dec_b200 as pkg
import checks
import oracle
from av1dec_b200 import format as F
from conftest import BITS, ROOT

have_product = os.path.exists(pkg.engine_path()) and os.path.exists(pkg.decoder_path())
need_product = pytest.mark.skipif(not have_product, reason="product libraries not built")
need_emu = pytest.mark.skipif(not checks.emu_available(), reason="tests/emu not built")
need_oracle = pytest.mark.skipif(not oracle.available(), reason="oracle/_ref not built")


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(av1b_[a-z0-9_]+)\s*\(", src)) - {"av1b_cmd_sink"})


@need_product
def test_engine_exports_every_declared_symbol():
    lib = C.CDLL(pkg.engine_path())
    names = declared("av1b200.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    lib.av1b_backend.restype = C.c_char_p
    assert lib.av1b_backend() == b"cuda-sm_100a"


@need_product
def test_decoder_exports_every_declared_symbol():
    pkg.load_engine()
    lib = C.CDLL(pkg.decoder_path())
    for n in declared("av1b200_decoder.h") + ["createVideoDecoder", "releaseVideoDecoder"]:
        assert hasattr(lib, n), n


@need_product
def test_format_mirror_matches_c_structs():
    lib = pkg.load_engine()
    lib.av1b_struct_size.restype = C.c_size_t
    for k, st in F.STRUCTS.items():
        assert C.sizeof(st) == lib.av1b_struct_size(k), st.__name__


@need_product
def test_product_has_no_cpu_pixel_path():
    """Neither the reference's pixel translation units nor the pixel functions that live in its mixed
    parse / decode units are in the product: the decode() tree walk is replaced by abort() traps
    (host/pixel_path_guard.cpp) and what hung below it is dropped at link time (--gc-sections)."""
    out = subprocess.run(["nm", "-C", "-S", "--defined-only", pkg.decoder_path()], stdout=subprocess.PIPE, text=True).stdout
    for sym in ("LoopFilter::filter", "Cdef::filter", "LoopRestoration::filter", "YuvFrame::create",
                "IntraPredict::directionalIntraPredict", "IntraPredict::dcPredict",
                # pixel code of the mixed units
                "TransformBlock::inverseTransform", "TransformBlock::decode", "TransformBlock::reconstruct",
                "InterPredict::predict_inter", "InterPredict::blockWarp", "InterPredict::blockInterPrediction",
                "InterPredict::maskBlend", "InterPredict::overlappedMotionCompensation", "Block::compute_prediction",
                "Palette::predict_palette", "Tile::decode"):
        assert sym not in out, sym
    # what the vtables still name are the traps: a few bytes each, not the reference's bodies
    sizes = {}
    for line in out.splitlines():
        f = line.split(None, 3)
        if len(f) == 4 and "::decode(std::shared_ptr<Yami::YuvFrame>&" in f[3]:
            sizes[f[3].split("(")[0]] = int(f[1], 16)
    assert "YamiAv1::Block::decode" in sizes  # vtable slot of BlockTree::decode
    assert all(v <= 64 for v in sizes.values()), sizes
    assert "IntraPredict::predict_intra" not in out  # its only caller (TransformBlock::decode) is gone


@need_emu
@pytest.mark.parametrize("name", ["4x4.ivf", "64x64.ivf", "foreman_qcif_i.ivf", "Halo_426x240_1frames_intrabc.ivf",
                                  "av1-1-b8-06-mfmv.ivf", "av1-1-b8-04-cdfupdate.ivf", "av1-1-b8-00-quantizer-00.ivf",
                                  "av1-1-b8-00-quantizer-33.ivf", "av1-1-b8-01-size-66x66.ivf",
                                  "av1-1-b8-01-size-226x226.ivf", "av1-1-b8-01-size-16x18.ivf"])
def test_emitter_and_lifecycle_bit_exact_under_emulation(name, md5_table):
    got, frames, _ = checks.stream_md5(checks.emu_decoder(), os.path.join(BITS, name))
    assert got == md5_table[name]


@need_emu
def test_decoder_class_mirrors_reference_api(md5_table):
    """decode() per temporal unit + getOutput() loop, as tests/Av1Dec.cpp:205-221 drives it."""
    name = "av1-1-b8-01-size-34x34.ivf"
    data = open(os.path.join(BITS, name), "rb").read()
    dec = pkg.Decoder(lib=checks.emu_decoder())
    md5 = hashlib.md5()
    n = 0
    for unit in pkg.iter_ivf(data):
        assert dec.decode(unit)
        while True:
            out = dec.get_output()
            if out is None:
                break
            n += 1
            for p in out[2]:
                md5.update(p)
    assert dec.get_output() is None
    dec.close()
    assert n == 2 and md5.hexdigest() == md5_table[name]


@need_emu
@pytest.mark.skipif(not os.path.exists(checks.IVD_DRIVE), reason="tests/native/ivd_drive not built")
def test_ivideodecoder_vtable_under_emulation(md5_table, tmp_path):
    """The declared-only Yami interface (interface/VideoDecoderInterface.h:31-68), driven the way
    a libyami client would: start / decode / getOutput, flush + reset, stop + start."""
    name = "av1-1-b8-06-mfmv.ivf"
    got, frames, one = checks.drive_ivideodecoder(checks.EMU_DECODER, name, "plain", tmp_path)
    assert (got, frames) == (md5_table[name], 4)
    got, frames, _ = checks.drive_ivideodecoder(checks.EMU_DECODER, name, "flush", tmp_path)
    assert (got, frames) == (md5_table[name], 4)
    _, frames, two = checks.drive_ivideodecoder(checks.EMU_DECODER, name, "twice", tmp_path)
    assert frames == 8 and two == one + one


@need_emu
@pytest.mark.parametrize("w,h,sb_log2", [(328, 200, 6), (456, 264, 7)])
def test_wavefront_level_schedule_under_emulation(w, h, sb_log2):
    checks.check_wave(checks.emu_engine(), w, h, sb_log2)


@need_emu
@pytest.mark.skipif(not oracle.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("kw", [dict(compound_frac=0.3), dict(compound_frac=0.3, fast=False, seed=1234, max_mv=2048)])
def test_inter_prediction_vs_reference_under_emulation(kw):
    """Translational motion compensation against the reference's Block::InterPredict."""
    checks.check_inter_vs_oracle(checks.emu_engine(), 640, 384, **kw)


@need_emu
@pytest.mark.skipif(not oracle.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("w,h,sb_log2,kw", [(640, 360, 6, {}), (1280, 720, 7, dict(rect=True, sizes=(16, 32, 64))),
                                             (1920, 1080, 6, dict(rect=True, intra_frac=1.0))])
def test_intra_prediction_vs_reference_under_emulation(w, h, sb_log2, kw):
    """Every intra op of a synthetic frame against the reference's Block::IntraPredict."""
    checks.check_wave_vs_oracle(checks.emu_engine(), w, h, sb_log2, **kw)


@need_emu
def test_device_view_and_nv12_under_emulation():
    checks.check_output_paths(checks.emu_engine(), 178, 94)


@need_emu
def test_inter_prediction_properties_under_emulation():
    checks.check_inter_properties(checks.emu_engine(), 192, 128)


@need_emu
def test_ivf_segments_split_at_random_access_points():
    lib = checks.emu_decoder()
    rd = lambda n: open(os.path.join(BITS, n), "rb").read()
    # every temporal unit of the all-intra stream is sequence header + shown key frame
    assert pkg.ivf_segments(rd("av1-1-b8-02-allintra.ivf"), lib=lib) == list(range(39))
    # ordinary streams: one key frame up front, one segment
    assert pkg.ivf_segments(rd("av1-1-b8-00-quantizer-00.ivf"), lib=lib) == [0]
    assert pkg.ivf_segments(rd("av1-1-b8-06-mfmv.ivf"), lib=lib) == [0]
    with pytest.raises(pkg.EngineError):
        pkg.ivf_segments(b"garbage" * 10, lib=lib)


@need_emu
def test_decode_rejects_garbage():
    dec = pkg.Decoder(lib=checks.emu_decoder())
    assert dec.decode(b"\xff" * 40) is False
    assert dec.get_output() is None
    dec.close()
    with pytest.raises(pkg.EngineError):
        pkg.decode_ivf(b"not an ivf file at all, just bytes" * 2, lib=checks.emu_decoder())


@need_emu
@need_oracle
def test_itx_vs_reference_under_emulation():
    assert checks.check_itx(checks.emu_engine(), n=600) == 600
    assert checks.check_itx(checks.emu_engine(), extents=True) > 800  # every zero-aware butterfly variant
    # every transform size on its own, so a failure names the size
    for ts in range(19):
        checks.check_itx(checks.emu_engine(), n=40, seed=1000 + ts, sizes=[ts])


@need_emu
@need_oracle
@pytest.mark.parametrize("w,h,kw", [
    (200, 136, {}),
    (226, 226, {"delta_lf": True}),
    (64, 64, {"dist": "U"}),
    (352, 288, {"lr_unit": 128, "sb128": True}),
    (18, 34, {"lr_unit": 256}),
    (640, 360, {"dist": "U", "delta_lf": True, "lr_unit": 256}),
])
def test_postfilter_vs_reference_under_emulation(w, h, kw):
    for stages in (1, 2, 4, 7):
        checks.check_postfilter(checks.emu_engine(), w, h, stages, **kw)

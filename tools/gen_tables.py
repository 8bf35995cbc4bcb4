#!/usr/bin/env python3
"""Generate av1dec_b200/csrc/av1_tables.h.

The tables are the normative constants of the AV1 specification (interpolation
kernels, warp kernels, filter-intra taps, smooth weights, wedge generators, ...).
There is no network in the build container, so the numbers are *extracted* from
the reference tree (where they are transcribed from the spec) and re-emitted in
this project's own packed layout (int8/int16/uint8 device constant arrays).
Only numeric data crosses; no reference code is copied.

Sources (reference file:line of each array):
  Subpel_Filters            decoder/InterPredict.cpp:99
  Warped_Filters            decoder/InterPredict.cpp:407
  Wedge_Master_*            decoder/InterPredict.cpp:712-731
  Wedge_Codebook            decoder/InterPredict.cpp:753
  Ii_Weights_1d             decoder/InterPredict.cpp:557
  Obmc_Mask_*               decoder/InterPredict.cpp:632-646
  Quant_Dist_*              decoder/InterPredict.cpp:919-930
  Intra_Filter_Taps         decoder/IntraPredict.cpp:59
  Dr_Intra_Derivative       decoder/IntraPredict.cpp:175
  Sm_Weights_Tx_*           decoder/IntraPredict.cpp:510-518
  Intra_Edge_Kernel         decoder/IntraPredict.cpp:324
  Cdef_*                    decoder/Cdef.cpp:60,103,107,120,200
  Sgr_Params                decoder/Av1Common.h:206
  Transform_Row_Shift       decoder/TransformBlock.cpp:2168
  Cos128_Lookup             decoder/TransformBlock.cpp:1771 (cross-checked against round(4096*cos))

Usage: python tools/gen_tables.py [/root/reference] > av1dec_b200/csrc/av1_tables.h
       python tools/gen_tables.py --host [/root/reference] > av1dec_b200/csrc/av1_tables_host.h
"""
import math
import re
import sys

ARGS = [a for a in sys.argv[1:] if not a.startswith("--")]
HOST = "--host" in sys.argv  # emit the host-side subset (hk_ prefix) used by engine.cu
REF = ARGS[0] if ARGS else "/root/reference"

ENUMS = {
    "WEDGE_HORIZONTAL": 0, "WEDGE_VERTICAL": 1, "WEDGE_OBLIQUE27": 2,
    "WEDGE_OBLIQUE63": 3, "WEDGE_OBLIQUE117": 4, "WEDGE_OBLIQUE153": 5,
    "MAX_FRAME_DISTANCE": 31,
}


def strip_comments(s):
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    s = re.sub(r"//[^\n]*", "", s)
    return s


def extract(path, name):
    """Return the nested-list value of C array `name` defined in `path`."""
    src = strip_comments(open(f"{REF}/{path}").read())
    m = re.search(r"\b" + re.escape(name) + r"\s*(\[[^=;{]*\])+\s*=\s*\{", src)
    if not m:
        raise SystemExit(f"table {name} not found in {path}")
    i = m.end() - 1
    depth = 0
    j = i
    while True:
        c = src[j]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    body = src[i:j + 1]
    for k, v in ENUMS.items():
        body = re.sub(r"\b" + k + r"\b", str(v), body)
    body = body.replace("{", "[").replace("}", "]")
    body = re.sub(r",\s*\]", "]", body)
    return eval(body)  # numeric literals only


def flat(x):
    if isinstance(x, list):
        out = []
        for e in x:
            out.extend(flat(e))
        return out
    return [x]


def dims(x):
    d = []
    while isinstance(x, list):
        d.append(len(x))
        x = x[0]
    return d


def emit(ctype, name, val, per_line=16):
    d = dims(val)
    f = flat(val)
    n = 1
    for k in d:
        n *= k
    assert n == len(f), (name, d, len(f))
    dd = "".join(f"[{k}]" for k in d)
    if HOST:
        name = "h" + name
    out = [f"AV1T_CONST {ctype} {name}{dd} = {{"]
    for i in range(0, len(f), per_line):
        out.append("    " + ", ".join(str(v) for v in f[i:i + per_line]) + ",")
    out.append("};")
    return "\n".join(out)


def main():
    T = []
    sub = extract("decoder/InterPredict.cpp", "Subpel_Filters")
    assert dims(sub) == [6, 16, 8]
    T.append(emit("int16_t", "k_subpel_filters", sub))
    # the same taps halved (every tap is even) and packed as signed bytes, four per word
    packed = [[[sum(((t >> 1) & 0xFF) << (8 * i) for i, t in enumerate(f[k:k + 4])) for k in (0, 4)] for f in row] for row in sub]
    assert all(t % 2 == 0 and -128 <= (t >> 1) <= 127 for row in sub for f in row for t in f)
    T.append(emit("uint32_t", "k_subpel_packed", packed, per_line=8))
    warp = extract("decoder/InterPredict.cpp", "Warped_Filters")
    assert dims(warp) == [193, 8]
    T.append(emit("int16_t", "k_warped_filters", warp))
    for nm, out in (("Wedge_Master_Oblique_Odd", "k_wedge_master_odd"),
                    ("Wedge_Master_Oblique_Even", "k_wedge_master_even"),
                    ("Wedge_Master_Vertical", "k_wedge_master_vert")):
        T.append(emit("uint8_t", out, extract("decoder/InterPredict.cpp", nm)))
    T.append(emit("uint8_t", "k_wedge_codebook", extract("decoder/InterPredict.cpp", "Wedge_Codebook")))
    T.append(emit("uint8_t", "k_ii_weights_1d", extract("decoder/InterPredict.cpp", "Ii_Weights_1d")))
    # OBMC masks packed back to back: length 2 @0, 4 @2, 8 @6, 16 @14, 32 @30 (offset = len-2)
    obmc = []
    for n in (2, 4, 8, 16, 32):
        obmc += extract("decoder/InterPredict.cpp", f"Obmc_Mask_{n}")
    T.append(emit("uint8_t", "k_obmc_mask", obmc))
    T.append(emit("uint8_t", "k_quant_dist_weight", extract("decoder/InterPredict.cpp", "Quant_Dist_Weight")))
    T.append(emit("uint8_t", "k_quant_dist_lookup", extract("decoder/InterPredict.cpp", "Quant_Dist_Lookup")))
    fi = extract("decoder/IntraPredict.cpp", "Intra_Filter_Taps")
    assert dims(fi) == [5, 8, 7]
    T.append(emit("int8_t", "k_intra_filter_taps", fi))
    T.append(emit("int16_t", "k_dr_intra_derivative", extract("decoder/IntraPredict.cpp", "Dr_Intra_Derivative")))
    # smooth weights packed: 4 @0, 8 @4, 16 @12, 32 @28, 64 @60 (offset = n-4)
    sm = []
    for n in (4, 8, 16, 32, 64):
        sm += extract("decoder/IntraPredict.cpp", f"Sm_Weights_Tx_{n}x{n}")
    T.append(emit("uint8_t", "k_sm_weights", sm))
    T.append(emit("uint8_t", "k_intra_edge_kernel", extract("decoder/IntraPredict.cpp", "Intra_Edge_Kernel")))
    T.append(emit("uint8_t", "k_mode_to_angle", extract("decoder/IntraPredict.cpp", "Mode_To_Angle")))
    T.append(emit("uint8_t", "k_cdef_uv_dir", extract("decoder/Cdef.cpp", "Cdef_Uv_Dir")))
    T.append(emit("uint8_t", "k_cdef_pri_taps", extract("decoder/Cdef.cpp", "Cdef_Pri_Taps")))
    T.append(emit("uint8_t", "k_cdef_sec_taps", extract("decoder/Cdef.cpp", "Cdef_Sec_Taps")))
    T.append(emit("int8_t", "k_cdef_directions", extract("decoder/Cdef.cpp", "Cdef_Directions")))
    T.append(emit("int16_t", "k_cdef_div_table", extract("decoder/Cdef.cpp", "Div_Table")))
    T.append(emit("uint8_t", "k_sgr_params", extract("decoder/Av1Common.h", "Sgr_Params")))
    T.append(emit("uint8_t", "k_tx_row_shift", extract("decoder/TransformBlock.cpp", "Transform_Row_Shift")))
    cos = extract("decoder/TransformBlock.cpp", "Cos128_Lookup")
    mine = [int(math.floor(4096 * math.cos(i * math.pi / 128) + 0.5)) for i in range(65)]
    assert cos == mine, "Cos128 table is not round(4096*cos(i*pi/128))"
    T.append(emit("int16_t", "k_cos128", cos))
    for nm, out in (("Tx_Width", "k_tx_w"), ("Tx_Height", "k_tx_h"),
                    ("Tx_Width_Log2", "k_tx_wlog2"), ("Tx_Height_Log2", "k_tx_hlog2"),
                    ("Block_Width", "k_block_w"), ("Block_Height", "k_block_h")):
        T.append(emit("uint8_t", out, extract("decoder/Av1Common.h", nm)))
    # Tx_*_Log2 again as 19 three-bit fields of one word: (PACKED >> 3*tx_size) & 7 costs no memory
    # access on the wavefront's critical path
    for nm, out in (("Tx_Width_Log2", "AV1T_TX_WLOG2_PACKED"), ("Tx_Height_Log2", "AV1T_TX_HLOG2_PACKED")):
        v = extract("decoder/Av1Common.h", nm)
        assert len(v) == 19 and max(v) < 8
        T.append(f"#define {out} 0x{sum(x << (3 * i) for i, x in enumerate(v)):x}ULL")

    print("// GENERATED by tools/gen_tables.py -- AV1 specification constant tables. Do not edit.")
    print("// Data only; see the generator's docstring for provenance of each array.")
    print("#pragma once")
    print("#include <stdint.h>")
    if HOST:
        keep = ("hk_wedge_master", "hk_wedge_codebook", "hk_block_w", "hk_block_h", "hk_quant_dist", "hk_tx_")
        T = [t.replace("AV1T_CONST", "static const") for t in T if not t.startswith("#define") and t.split()[2].startswith(keep)]
    else:
        print("#ifndef AV1T_CONST")
        print("#define AV1T_CONST static const")
        print("#endif")
    print()
    print("\n\n".join(T))


if __name__ == "__main__":
    main()

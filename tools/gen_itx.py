#!/usr/bin/env python3
"""Generator of av1dec_b200/csrc/itx_gen.h: the AV1 inverse DCT (4..64 points) and ADST (8, 16
points) as ZERO-AWARE STRAIGHT-LINE code.

The AV1 inverse transforms are normative flow graphs (spec 7.13.2.3 / 7.13.2.7 / 7.13.2.8): a fixed
sequence of butterfly rotations B(a, b, angle) with Round2(., 12) and clamped Hadamard steps
H(a, b); bit-exactness means evaluating exactly that graph.  This script holds the graph as DATA
(`dct_graph`, `adst_graph`), evaluates it symbolically over hash-consed value nodes, and emits one
function per (transform, number of leading non-zero inputs K):

  * an input known to be zero removes every multiply / add it would feed: a rotation with one zero
    operand is a single multiply, a Hadamard with one is a clamp, with two it disappears;
  * identical sub-expressions are shared (a DC-only 64-point DCT is ONE multiply, three clamps and
    a broadcast -- not 64 columns of butterflies);
  * the bit-reversal / ADST permutations cost nothing (they only rename nodes);
  * the angle table folds into immediates.

The device code picks K from the op's nz_rows / nz_cols (the emitter records the extent of the
non-zero coefficients): real streams are dominated by blocks with a handful of low-frequency
coefficients.  The 64-point transforms never see more than 32 inputs (AV1 zeroes the rest).

Run:  python tools/gen_itx.py            (rewrites av1dec_b200/csrc/itx_gen.h)
"""
import math
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "av1dec_b200", "csrc", "itx_gen.h")


def load_cos_table():
    """cos(i * pi / 128) in Q12, i = 0..64 -- the normative table (spec 7.13.2.1), taken from
    av1_tables.h so that both stay one source."""
    src = open(os.path.join(ROOT, "av1dec_b200", "csrc", "av1_tables.h")).read()
    m = re.search(r"k_cos128\[65\]\s*=\s*\{([^}]*)\}", src)
    vals = [int(v) for v in m.group(1).replace("\n", " ").split(",") if v.strip()]
    assert len(vals) == 65 and vals[0] == 4096 and vals[64] == 0 and vals[32] == 2896
    for i, v in enumerate(vals):  # the table is what the formula says, to the unit
        assert abs(v - 4096 * math.cos(math.pi * i / 128)) <= 1.0, i
    return vals


COS = load_cos_table()


def cos128(angle):
    a = angle & 255
    if a <= 64:
        return COS[a]
    if a <= 128:
        return -COS[128 - a]
    if a <= 192:
        return -COS[a - 128]
    return COS[256 - a]


def sin128(angle):
    return cos128(angle - 64)


def brev(bits, x):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


# ------------------------------------------------------------------------------ symbolic values
class Graph:
    """Hash-consed expression nodes.  Node 0 is the constant zero."""

    def __init__(self):
        self.nodes = [("zero",)]
        self.index = {("zero",): 0}

    def mk(self, *key):
        if key not in self.index:
            self.index[key] = len(self.nodes)
            self.nodes.append(key)
        return self.index[key]

    def inp(self, i):
        return self.mk("in", i)

    def mulr(self, a, ca, b, cb):
        """Round2(a * ca + b * cb, 12)"""
        if a == 0 or ca == 0:
            a, ca = 0, 0
        if b == 0 or cb == 0:
            b, cb = 0, 0
        if a == 0 and b == 0:
            return 0
        if a == 0:
            return self.mk("mul1", b, cb)
        if b == 0:
            return self.mk("mul1", a, ca)
        return self.mk("mul2", a, ca, b, cb)

    def rot(self, T, a, b, angle, flip):
        """the spec's B(a, b, angle, flip)"""
        c, s = cos128(angle), sin128(angle)
        x = self.mulr(T[a], c, T[b], -s)
        y = self.mulr(T[a], s, T[b], c)
        if flip:
            T[b], T[a] = x, y
        else:
            T[a], T[b] = x, y

    def had(self, T, a, b, flip):
        """the spec's H(a, b, flip, r): clamped sum and difference"""
        if flip:
            a, b = b, a
        x, y = T[a], T[b]
        if x == 0 and y == 0:
            return
        if y == 0:
            T[a] = T[b] = self.mk("clip", x)
        elif x == 0:
            T[a], T[b] = self.mk("clip", y), self.mk("clipneg", y)
        else:
            T[a], T[b] = self.mk("add", x, y), self.mk("sub", x, y)

    def neg(self, a):
        return 0 if a == 0 else self.mk("neg", a)


# ------------------------------------------------------------------------------ the flow graphs
def dct_graph(g, T, n, skip_final=False):
    """Inverse DCT of 1 << n points (AV1 spec 7.13.2.3, steps in the order the spec lists them;
    a step applies when n is large enough for the indices it names)."""
    N = 1 << n
    T[:] = [T[brev(n, i)] for i in range(N)]
    steps = []
    B = lambda a, b, ang, fl=False: steps.append(("B", a, b, ang, fl))
    H = lambda a, b, fl=False: steps.append(("H", a, b, fl))
    if n == 6:
        for i in range(16):
            B(32 + i, 63 - i, 63 - 4 * brev(4, i))
    if n >= 5:
        for i in range(8):
            B(16 + i, 31 - i, 6 + (brev(3, 7 - i) << 3))
    if n == 6:
        for i in range(16):
            H(32 + 2 * i, 33 + 2 * i, bool(i & 1))
    if n >= 4:
        for i in range(4):
            B(8 + i, 15 - i, 12 + (brev(2, 3 - i) << 4))
    if n >= 5:
        for i in range(8):
            H(16 + 2 * i, 17 + 2 * i, bool(i & 1))
    if n == 6:
        for i in range(4):
            for j in range(2):
                B(62 - 4 * i - j, 33 + 4 * i + j, 60 - 16 * brev(2, i) + 64 * j, True)
    if n >= 3:
        for i in range(2):
            B(4 + i, 7 - i, 56 - 32 * i)
    if n >= 4:
        for i in range(4):
            H(8 + 2 * i, 9 + 2 * i, bool(i & 1))
    if n >= 5:
        for i in range(2):
            for j in range(2):
                B(30 - 4 * i - j, 17 + 4 * i + j, 24 + (j << 6) + ((1 - i) << 5), True)
    if n == 6:
        for i in range(8):
            for j in range(2):
                H(32 + 4 * i + j, 35 + 4 * i - j, bool(i & 1))
    for i in range(2):
        B(2 * i, 2 * i + 1, 32 + 16 * i, bool(1 - i))
    if n >= 3:
        for i in range(2):
            H(4 + 2 * i, 5 + 2 * i, bool(i))
    if n >= 4:
        for i in range(2):
            B(14 - i, 9 + i, 48 + 64 * i, True)
    if n >= 5:
        for i in range(4):
            for j in range(2):
                H(16 + 4 * i + j, 19 + 4 * i - j, bool(i & 1))
    if n == 6:
        for i in range(2):
            for j in range(4):
                B(61 - 8 * i - j, 34 + 8 * i + j, 56 - 32 * i + (j >> 1) * 64, True)
    for i in range(2):
        H(i, 3 - i)
    if n >= 3:
        B(6, 5, 32, True)
    if n >= 4:
        for i in range(2):
            for j in range(2):
                H(8 + 4 * i + j, 11 + 4 * i - j, bool(i))
    if n >= 5:
        for i in range(4):
            B(29 - i, 18 + i, 48 + (i >> 1) * 64, True)
    if n == 6:
        for i in range(4):
            for j in range(4):
                H(32 + 8 * i + j, 39 + 8 * i - j, bool(i & 1))
    if n >= 3:
        for i in range(4):
            H(i, 7 - i)
    if n >= 4:
        for i in range(2):
            B(13 - i, 10 + i, 32, True)
    if n >= 5:
        for i in range(2):
            for j in range(4):
                H(16 + 8 * i + j, 23 + 8 * i - j, bool(i))
    if n == 6:
        for i in range(8):
            B(59 - i, 36 + i, 48 if i < 4 else 112, True)
    if n >= 4:
        for i in range(8):
            H(i, 15 - i)
    if n >= 5:
        for i in range(4):
            B(27 - i, 20 + i, 32, True)
    if n == 6:
        for i in range(8):
            H(32 + i, 47 - i)
            H(48 + i, 63 - i, True)
    if n >= 5:
        for i in range(16):
            H(i, 31 - i)
    if n == 6:
        for i in range(8):
            B(55 - i, 40 + i, 32, True)
        for i in range(32):
            H(i, 63 - i)
    if skip_final:  # everything but the last Hadamard stage H(i, N-1-i): the two halves are still independent
        last, steps = steps[-(N // 2):], steps[:-(N // 2)]
        assert all(st == ("H", i, N - 1 - i, False) for i, st in enumerate(last))
    for st in steps:
        if st[0] == "B":
            g.rot(T, st[1], st[2], st[3], st[4])
        else:
            g.had(T, st[1], st[2], st[3])


def adst_graph(g, T, n):
    """Inverse ADST of 8 or 16 points (AV1 spec 7.13.2.7 / 7.13.2.8)."""
    N = 1 << n
    T[:] = [T[i - 1] if (i & 1) else T[N - i - 1] for i in range(N)]
    if n == 3:
        for i in range(4):
            g.rot(T, 2 * i, 2 * i + 1, 60 - 16 * i, True)
        for i in range(4):
            g.had(T, i, 4 + i, False)
        for i in range(2):
            g.rot(T, 4 + 3 * i, 5 + i, 48 - 32 * i, True)
        for i in range(2):
            for j in range(2):
                g.had(T, 4 * j + i, 2 + 4 * j + i, False)
        for i in range(2):
            g.rot(T, 2 + 4 * i, 3 + 4 * i, 32, True)
    else:
        for i in range(8):
            g.rot(T, 2 * i, 2 * i + 1, 62 - 8 * i, True)
        for i in range(8):
            g.had(T, i, 8 + i, False)
        for i in range(2):
            g.rot(T, 8 + 2 * i, 9 + 2 * i, 56 - 32 * i, True)
            g.rot(T, 13 + 2 * i, 12 + 2 * i, 8 + 32 * i, True)
        for i in range(4):
            for j in range(2):
                g.had(T, 8 * j + i, 4 + 8 * j + i, False)
        for i in range(2):
            for j in range(2):
                g.rot(T, 4 + 8 * j + 3 * i, 5 + 8 * j + i, 48 - 32 * i, True)
        for i in range(2):
            for j in range(4):
                g.had(T, 4 * j + i, 2 + 4 * j + i, False)
        for i in range(4):
            g.rot(T, 2 + 4 * i, 3 + 4 * i, 32, True)
    c = list(T)
    for i in range(N):
        a = (i >> 3) & 1
        b = ((i >> 2) & 1) ^ ((i >> 3) & 1)
        cc = ((i >> 1) & 1) ^ ((i >> 2) & 1)
        d = (i & 1) ^ ((i >> 1) & 1)
        idx = ((d << 3) | (cc << 2) | (b << 1) | a) >> (4 - n)
        T[i] = g.neg(c[idx]) if (i & 1) else c[idx]


# ------------------------------------------------------------------------------ emission
def emit(name, n, k, graph_fn, odd_half=False):
    """One function: T[0 .. k-1] are the inputs (the rest are zero), T[0 .. N-1] the outputs.
    odd_half: the ODD half of a DCT of 1 << n points on its own -- T[0 .. k-1] are the odd-indexed
    inputs x[1], x[3], ..., the outputs T[0 .. N/2-1] are positions N/2 .. N-1 of the flow graph
    before its last Hadamard stage (the even half is the DCT of half the size on the even-indexed
    inputs; the caller combines the two: out[i] = clip(E[i] + O[N/2-1-i]), out[N-1-i] = clip(E[i] - O[N/2-1-i]))."""
    N = 1 << n
    g = Graph()
    if odd_half:
        T = [g.inp(i >> 1) if (i & 1) and (i >> 1) < k else 0 for i in range(N)]
        graph_fn(g, T, n, skip_final=True)
        assert all(t == 0 for t in T[:N // 2])
        T = T[N // 2:]
        N = N // 2
    else:
        T = [g.inp(i) if i < k else 0 for i in range(N)]
        graph_fn(g, T, n)
    live, stack = set(), [t for t in T if t]
    while stack:
        v = stack.pop()
        if v in live or v == 0:
            continue
        live.add(v)
        node = g.nodes[v]
        if node[0] in ("mul1", "clip", "clipneg", "neg"):
            stack.append(node[1])
        elif node[0] == "mul2":
            stack += [node[1], node[3]]
        elif node[0] in ("add", "sub"):
            stack += [node[1], node[2]]
    lines = ["AV1B_DEV void %s(int* T, const int lo, const int hi)" % name, "{"]
    ops = {"mul": 0, "addsub": 0, "clip": 0}
    for v in sorted(live):
        node = g.nodes[v]
        kind = node[0]
        if kind == "in":
            e = "T[%d]" % node[1]
        elif kind == "mul1":
            e = "(v%d * %d + 2048) >> 12" % (node[1], node[2])
            ops["mul"] += 1
        elif kind == "mul2":
            e = "(v%d * %d + v%d * %d + 2048) >> 12" % (node[1], node[2], node[3], node[4])
            ops["mul"] += 2
        elif kind == "add":
            e = "clip3(lo, hi, v%d + v%d)" % (node[1], node[2])
            ops["addsub"] += 1
        elif kind == "sub":
            e = "clip3(lo, hi, v%d - v%d)" % (node[1], node[2])
            ops["addsub"] += 1
        elif kind == "clip":
            e = "clip3(lo, hi, v%d)" % node[1]
            ops["clip"] += 1
        elif kind == "clipneg":
            e = "clip3(lo, hi, -v%d)" % node[1]
            ops["clip"] += 1
        else:
            e = "-v%d" % node[1]
        lines.append("    const int v%d = %s;" % (v, e))
    for i in range(N):
        lines.append("    T[%d] = %s;" % (i, "v%d" % T[i] if T[i] else "0"))
    lines.append("}")
    lines[0] = "// %d multiplies, %d clamped add/sub, %d clamps\n" % (ops["mul"], ops["addsub"], ops["clip"]) + lines[0]
    return "\n".join(lines)


def dispatcher(prefix, n, ks):
    N = 1 << n
    out = ["// %s of %d points, nz = number of leading inputs that may be non-zero, r = clamp width of the Hadamard steps" % (prefix, N),
           "AV1B_DEV void %s%d(int* T, int r, int nz)" % (prefix, N), "{",
           "    const int hi = (1 << (r - 1)) - 1, lo = -hi - 1;"]
    for i, k in enumerate(ks):
        cond = "if (nz <= %d) " % k if i + 1 < len(ks) else ""
        out.append("    %s%s%s%d_k%d(T, lo, hi);" % ("else " if i else "", cond, prefix, N, k))
    out.append("}")
    return "\n".join(out)


def main():
    parts = ["// itx_gen.h -- GENERATED by tools/gen_itx.py; do not edit.",
             "// Zero-aware straight-line AV1 inverse DCT / ADST butterflies (see the generator for the design).",
             "#pragma once", '#include "dev.h"', "", "namespace itx {", ""]
    # three variants per transform: DC only, the low-frequency quarter, everything (more variants
    # cost instruction-cache and registers on dense blocks and buy little on sparse ones)
    plan = [("idct", dct_graph, {2: [1, 4], 3: [1, 4, 8], 4: [1, 4, 16], 5: [1, 8, 16, 32], 6: [1, 8, 32]}),
            ("iadst", adst_graph, {3: [1, 4, 8], 4: [1, 4, 16]})]
    for prefix, fn, sizes in plan:
        for n, ks in sizes.items():
            for k in ks:
                parts.append(emit("%s%d_k%d" % (prefix, 1 << n, k), n, k, fn))
                parts.append("")
            parts.append(dispatcher(prefix, n, ks))
            parts.append("")
    # the odd halves of the 32- and 64-point DCT (two lanes of different warps share one transform)
    for n, ks in ((5, [4, 16]), (6, [4, 16])):
        for k in ks:
            parts.append(emit("idct%d_odd_k%d" % (1 << n, k), n, k, dct_graph, odd_half=True))
            parts.append("")
        parts.append(dispatcher("idct%d_odd" % (1 << n), 0, ks).replace("of 1 points", "odd half").replace("idct%d_odd1" % (1 << n), "idct%d_odd" % (1 << n)))
        parts.append("")
    parts.append("}  // namespace itx")
    open(OUT, "w").write("\n".join(parts) + "\n")
    print("wrote", OUT, sum(p.count("\n") + 1 for p in parts), "lines")


if __name__ == "__main__":
    main()

"""Pins the oracle: the unmodified reference built by oracle/Makefile must reproduce every MD5 of
the reference's own golden file bits/bits.md5 (testscript/conformance.py:42-56) -- SURVEY.md 8c."""
import hashlib
import os

import pytest

import oracle
from conftest import BITS, all_streams

pytestmark = pytest.mark.skipif(not oracle.available(), reason="oracle/_ref not built (needs /root/reference)")


def test_oracle_reproduces_all_golden_md5(md5_table):
    streams = all_streams()
    assert len(streams) == 172
    bad = []
    for f in streams:
        yuv, frames, _ = oracle.decode_ivf(open(os.path.join(BITS, f), "rb").read())
        if hashlib.md5(yuv).hexdigest() != md5_table[f]:
            bad.append(f)
    assert not bad, f"oracle differs from bits.md5 on {bad}"


def test_oracle_stage_dump_final_equals_output():
    data = open(os.path.join(BITS, "av1-1-b8-02-allintra.ivf"), "rb").read()
    yuv, frames, _ = oracle.decode_ivf(data)
    final, n = oracle.decode_stages(data, 3)
    assert n == frames == 39 and final == yuv

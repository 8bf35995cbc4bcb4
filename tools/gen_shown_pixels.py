#!/usr/bin/env python3
"""Writes tests/golden/shown_pixels.json: shown luma pixels per conformance stream, counted by
decoding each stream with the reference (oracle/_ref/liboracle.so).  Needs /root/reference-built oracle."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402

bits = os.path.join(ROOT, "tests", "golden", "bits")
out = {}
for f in sorted(os.listdir(bits)):
    if f.endswith(".ivf"):
        _, frames, px = oracle.decode_ivf(open(os.path.join(bits, f), "rb").read())
        out[f] = px
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "shown_pixels.json"), "w"), indent=0, sort_keys=True)
print(len(out), "streams", sum(out.values()), "shown luma pixels")

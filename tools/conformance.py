#!/usr/bin/env python3
"""End-to-end conformance gate: decode every tests/golden/bits/*.ivf with a CLI binary and
compare the MD5 of the written .yuv with tests/golden/bits/bits.md5 (the reference's own gate,
testscript/conformance.py:115-141, restated).  Exit code = number of failures.

usage: tools/conformance.py [--bin av1dec_b200/bin/av1dec] [--pattern quantizer] [--jobs N]
"""
import argparse
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BITS = os.path.join(ROOT, "tests", "golden", "bits")


def load_md5():
    table = {}
    for line in open(os.path.join(BITS, "bits.md5")):
        parts = line.split()
        if len(parts) == 2:
            table[parts[1]] = parts[0]
    return table


def run_one(binary, path):
    with tempfile.NamedTemporaryFile(suffix=".yuv") as tmp:
        try:
            p = subprocess.run([binary, "-i", path, tmp.name], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
        except subprocess.TimeoutExpired:
            return None, "timeout"
        data = open(tmp.name, "rb").read()
        tail = p.stdout.decode(errors="replace")[-300:] if p.returncode else ""
        return hashlib.md5(data).hexdigest(), f"rc={p.returncode} {tail}" if p.returncode else ""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bin", default=os.path.join(ROOT, "av1dec_b200", "bin", "av1dec"))
    ap.add_argument("--pattern", default="")
    ap.add_argument("--jobs", type=int, default=4)
    a = ap.parse_args()
    md5 = load_md5()
    files = sorted(f for f in os.listdir(BITS) if f.endswith(".ivf") and a.pattern in f)
    fails = []
    with cf.ThreadPoolExecutor(a.jobs) as ex:
        futs = {ex.submit(run_one, a.bin, os.path.join(BITS, f)): f for f in files}
        for fut in cf.as_completed(futs):
            f = futs[fut]
            got, note = fut.result()
            if got != md5.get(f):
                fails.append(f)
                print(f"FAIL {f}: got {got} want {md5.get(f)} {note}", flush=True)
    print(f"conformance: {len(files) - len(fails)}/{len(files)} MD5-exact ({a.bin})")
    return len(fails)


if __name__ == "__main__":
    sys.exit(main())

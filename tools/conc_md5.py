"""All conformance streams decoded by 14 concurrent callers of av1b_decode_ivf (the inline / streaming
emission path a saturated service runs), twice over, every output MD5-checked against bits.md5.
Usage: python tools/conc_md5.py   (needs the GPU)"""
import os, sys, hashlib, concurrent.futures as cf
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import av1dec_b200 as pkg
pkg.load_engine(); lib = pkg.load_decoder()
bits = os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "tests", "golden", "bits")
table = {l.split()[1]: l.split()[0] for l in open(os.path.join(bits, "bits.md5")) if len(l.split()) == 2}
names = sorted(n for n in os.listdir(bits) if n.endswith(".ivf") and n in table)
datas = {n: open(os.path.join(bits, n), "rb").read() for n in names}
def one(n):
    yuv, frames, px = pkg.decode_ivf(datas[n], device=0, lib=lib)
    return n, hashlib.md5(yuv).hexdigest() == table[n]
bad = []
for rep in range(2):
    with cf.ThreadPoolExecutor(14) as ex:
        for n, ok in ex.map(one, sorted(names, key=lambda n: -len(datas[n])) * 2):
            if not ok: bad.append(n)
print("streams", len(names), "bad", bad[:10], "OK" if not bad else "FAIL")

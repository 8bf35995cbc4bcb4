"""Profiling aid: the inverse transform alone on the synthetic 4K frame of bench.py's itx leg.
Usage: python tools/itx_prof.py [reps]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import av1dec_b200 as pkg
from av1dec_b200 import format as F
from av1dec_b200 import synth
from av1dec_b200.engine import Engine

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
W, H = 3840, 2160
hdr_size = C.sizeof(F.FrameHdr)
cmd, n_tb, n_samples, algo = synth.make_itx_frame(W, H)[:4]
eng = Engine(W, H, device=0)
eng.set_lanes(1)
dev_cmd = eng.upload(cmd)
for _ in range(2):
    eng.submit_resident(dev_cmd, cmd[:hdr_size], pkg.STAGE_ITX, 0)
eng.sync()
eng.set_profiling(True)
for _ in range(reps):
    eng.submit_resident(dev_cmd, cmd[:hdr_size], pkg.STAGE_ITX, 0)
ms, calls = eng.stage_times()
eng.close()
print({"us": round(ms["itx"] / max(calls["itx"], 1) * 1e3, 1), "tbs": n_tb, "samples": n_samples})

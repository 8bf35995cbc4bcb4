"""Profiling aid: the deblock -> CDEF -> LR chain alone on synthetic frames (default 2 x 4K).
Usage: python tools/pf_prof.py [frames] [passes] [W H]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import av1dec_b200 as pkg
from av1dec_b200 import format as F
from av1dec_b200 import synth
from av1dec_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (3840, 2160)
hdr_size = C.sizeof(F.FrameHdr)
eng = Engine(W, H, device=0, stream=None)
eng.set_lanes(1)
frames = []
for i in range(n):
    sf = synth.make_postfilter_frame(W, H, seed=synth.SEED + i, dist="B", lr_unit=int(os.environ.get("PF_LR_UNIT", "64")),
                                     lr_types=tuple(int(t) for t in os.environ.get("PF_LR_TYPES", "0,1,2").split(",")))
    eng.set_ref(i, sf.planes, sf.mi_cols * 4, sf.mi_rows * 4)
    frames.append((eng.upload(sf.cmd), sf.cmd[:hdr_size]))
eng.set_profiling(True)
for k in range(n * passes):
    i = k % n
    eng.input_from_slot(i)
    eng.submit_resident(frames[i][0], frames[i][1], pkg.STAGE_POST, 0)
eng.sync()
ms, calls = eng.stage_times()
print({k: round(ms[k] / max(calls[k], 1) * 1e3, 1) for k in ("deblock", "cdef", "lr")})
eng.close()

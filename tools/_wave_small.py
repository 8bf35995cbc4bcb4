import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo")); sys.path.insert(0, os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "tests"))
import numpy as np
import av1dec_b200 as pkg
from av1dec_b200 import synth
from av1dec_b200.engine import Engine
lib = pkg.load_engine()
w, h, seg, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
kw = eval(sys.argv[5]) if len(sys.argv) > 5 else {}
rng = synth.SplitMix64(1)
planes = synth.make_planes(rng, w, h, "B")
cmd = synth.make_intra_frame(w, h, segments=bool(seg), **kw)
eng = Engine(w, h, lib=lib)
for i in range(reps):
    print("start", w, h, seg, i, flush=True)
    eng.set_input(planes, w, h)
    fid = eng.submit(cmd, stages=pkg.STAGE_WAVE)
    out = eng.download(fid, w, h)
    print("done", flush=True)
eng.close()

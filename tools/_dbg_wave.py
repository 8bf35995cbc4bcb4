import sys, os, ctypes as C, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
import av1dec_b200 as pkg
from av1dec_b200 import format as F
from av1dec_b200.engine import Engine
pkg.load_engine(); pkg.load_decoder()
name = sys.argv[1]
data = open(os.path.join(bench.BITS, name), 'rb').read()
frames = []
dec = pkg.Decoder(device=0)
dec.set_cmd_sink(lambda buf, n, refresh, show: frames.append((buf, n, refresh, show)))
mw = mh = 0
for unit in pkg.iter_ivf(data):
    dec.decode(unit)
    while True:
        o = dec.get_output()
        if o is None: break
        mw, mh = max(mw, o[0]), max(mh, o[1])
dec.close()
side = torch.cuda.Stream()
eng = Engine(mw, mh, device=0, stream=side.cuda_stream)
eng.set_lanes(1)
hdr_size = C.sizeof(F.FrameHdr)
fr = [(None, None, r, n) if b is None else (eng.upload(b), b[:hdr_size], r, -1) for b, n, r, s in frames]
def replay():
    for ptr, hdr, refresh, slot in fr:
        if ptr is None: eng.show_existing(slot, refresh)
        else: eng.submit_resident(ptr, hdr, pkg.STAGE_ALL, refresh)
for _ in range(3): replay()
eng.sync(); eng.set_profiling(True); eng.stage_times(reset=True)
for _ in range(5): replay()
ms, calls = eng.stage_times()
print(os.environ.get('AV1B_DBG_WAVE'), name, 'wave ms per replay', round(ms['wave'] / 5, 3))

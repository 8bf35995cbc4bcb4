"""Profiling aid: the superblock wavefront alone on synthetic 4K frames of a chosen mix.
Usage: python tools/wave_prof.py [config ...]   (configs: see CONFIGS; default = all)
Prints one line per config: us per frame, ops, superblocks."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import av1dec_b200 as pkg
from av1dec_b200 import format as F
from av1dec_b200 import synth
from av1dec_b200.engine import Engine

CONFIGS = {
    "mix": dict(),
    "empty": dict(intra_frac=0.0),
    "bs8": dict(sizes=(8,)),
    "bs16": dict(sizes=(16,)),
    "bs32": dict(sizes=(32,)),
    "bs8_dc": dict(sizes=(8,), mode_set=[0], fi=False, cfl=False),
    "bs8_dir": dict(sizes=(8,), mode_set=[1, 2, 3, 4, 5, 6, 7, 8], fi=False, cfl=False),
    "bs8_full": dict(sizes=(8,), intra_frac=1.0),
    "bs32_dc": dict(sizes=(32,), mode_set=[0], fi=False, cfl=False),
    "mix_sb128": dict(sb_log2=7),
    "mix_noseg": dict(segments=False),
    "bs8_dc_noseg": dict(sizes=(8,), mode_set=[0], fi=False, cfl=False, segments=False),
}

W, H = 3840, 2160
hdr_size = C.sizeof(F.FrameHdr)
names = [] if os.environ.get("WAVE_TRACE") else (sys.argv[1:] or list(CONFIGS))
reps = int(os.environ.get("WAVE_PROF_REPS", "6"))
for name in names:
    cmd = synth.make_intra_frame(W, H, **CONFIGS[name])
    hdr = F.FrameHdr.from_buffer_copy(cmd[:hdr_size])
    eng = Engine(W, H, device=0)
    eng.set_lanes(1)
    rng = synth.SplitMix64(synth.SEED + 78)
    eng.set_ref(0, synth.make_planes(rng, W, H, "B"), W, H)
    dev_cmd = eng.upload(cmd)

    def once():
        eng.input_from_slot(0)
        eng.submit_resident(dev_cmd, cmd[:hdr_size], pkg.STAGE_WAVE, 0)
    for _ in range(2):
        once()
    eng.sync()
    eng.set_profiling(True)
    for _ in range(reps):
        once()
    ms, calls = eng.stage_times()
    eng.close()
    print(name, {"us": round(ms["wave"] / max(calls["wave"], 1) * 1e3, 1), "ops": int(hdr.n_ops), "sbs": int(hdr.n_sb)}, flush=True)


def trace(name="mix"):
    """WAVE_TRACE=1: per-superblock timestamps of one launch and the critical path through them."""
    import numpy as np
    lib = pkg.load_engine()
    cmd = synth.make_intra_frame(W, H, **CONFIGS[name])
    hdr = F.FrameHdr.from_buffer_copy(cmd[:hdr_size])
    eng = Engine(W, H, device=0)
    eng.set_lanes(1)
    rng = synth.SplitMix64(synth.SEED + 78)
    eng.set_ref(0, synth.make_planes(rng, W, H, "B"), W, H)
    dev_cmd = eng.upload(cmd)
    n = int(hdr.n_sb)
    lib.av1b_debug_wave_trace.argtypes = [C.c_size_t]
    lib.av1b_debug_wave_trace_read.argtypes = [C.c_void_p, C.c_size_t]
    for it in range(3):
        if it == 2:
            assert lib.av1b_debug_wave_trace(n + 64) == 0
        eng.input_from_slot(0)
        eng.submit_resident(dev_cmd, cmd[:hdr_size], pkg.STAGE_WAVE, 0)
        eng.sync()
    buf = np.zeros((n + 64, 8), np.uint64)
    assert lib.av1b_debug_wave_trace_read(buf.ctypes.data, n + 64) == 0
    lvlog = buf[n:].reshape(32, 16).astype(np.int64)  # level log of the middle superblock (csrc/recon.cu)
    buf = buf[:n]
    if os.environ.get("WAVE_LEVEL_LOG"):
        base = lvlog[lvlog[:, 0] > 0, 0].min() if (lvlog[:, 0] > 0).any() else 0
        print("level log (ns from the first barrier exit): level, barrier exit, [op start-end per warp 0..6], warp 7 at the barrier")
        for lv in range(32):
            if lvlog[lv, 0] == 0:
                continue
            ops = " ".join(f"w{w}:{lvlog[lv, 1 + 2 * w] - base}-{lvlog[lv, 2 + 2 * w] - base}" for w in range(7) if lvlog[lv, 1 + 2 * w])
            print(f"  L{lv:2d} exit {lvlog[lv, 0] - base:7d}  {ops}  w7@bar {lvlog[lv, 15] - base}")
    lib.av1b_debug_wave_trace(0)
    eng.close()
    if os.environ.get("WAVE_TRACE_DUMP"):
        np.save(os.path.join(os.environ["WAVE_TRACE_DUMP"], f"wave_trace_{name}.npy"), buf)
        open(os.path.join(os.environ["WAVE_TRACE_DUMP"], f"wave_trace_{name}.cmd"), "wb").write(cmd)
    t = buf[:, 2:].astype(np.int64)
    t0 = t[:, 0].min()
    t -= t0
    cols, rows = int(hdr.sb_cols), int(hdr.sb_rows)
    ops = np.frombuffer(cmd, np.uint32, count=4 * n, offset=int(hdr.off_sb)).reshape(n, 4)[:, 1]
    dur = {"wait": t[:, 1] - t[:, 0], "halo": t[:, 2] - t[:, 1], "ops": t[:, 3] - t[:, 2], "signal": t[:, 4] - t[:, 3], "flush": t[:, 5] - t[:, 4]}
    print(name, "total us", t[:, 5].max() / 1e3, "SMs used", len(set(buf[:, 1].tolist())))
    for k, v in dur.items():
        print(f"  {k:7s} mean {v.mean() / 1e3:8.2f} us  p50 {np.median(v) / 1e3:8.2f}  max {v.max() / 1e3:8.2f}")
    print("  ops time per op (ns): ", (dur["ops"].sum() / max(ops.sum(), 1)))
    # critical path: from the superblock that finished last, step to whichever gate opened last
    sb = int(np.argmax(t[:, 4]))
    acc = {"ops": 0, "halo": 0, "signal": 0, "handoff": 0, "own_cta": 0}
    steps = 0
    while True:
        r, c = divmod(sb, cols)
        acc["ops"] += int(dur["ops"][sb])
        acc["halo"] += int(dur["halo"][sb])
        acc["signal"] += int(dur["signal"][sb])
        steps += 1
        deps = []
        if c > 0:
            deps.append(sb - 1)
        if r > 0:
            deps.append((r - 1) * cols + min(c + 1, cols - 1))
        if not deps:
            break
        gate = max(deps, key=lambda d: t[d, 4])
        ready = t[sb, 1]  # wait satisfied
        if t[sb, 0] > t[gate, 4]:  # the CTA took the ticket after the gate had opened: it was busy elsewhere
            acc["own_cta"] += int(ready - t[gate, 4])
        else:
            acc["handoff"] += int(ready - t[gate, 4])
        sb = gate
    print("  critical path:", steps, "superblocks;", {k: round(v / 1e3, 1) for k, v in acc.items()}, "us")


if os.environ.get("WAVE_TRACE"):
    for nm in (sys.argv[1:] or ["mix"]):
        trace(nm)

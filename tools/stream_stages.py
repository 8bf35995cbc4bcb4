#!/usr/bin/env python
"""Per-stage device time of single conformance streams (resident replay, one CUDA stream).

    python tools/stream_stages.py [--reps N] name.ivf [name.ivf ...]

Records each stream's command buffers once (MD5-gated, like bench.py), uploads them, then replays
the stream alone with per-stage CUDA-event profiling.  Prints one JSON line per stream: frames,
shown pixels, total ms per replay and the per-stage split.  Used to find which kernel bounds the
latency of one stream (the critical path of the multi-stream bench) and as the ncu target for the
wavefront kernel."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("names", nargs="+")
    a = ap.parse_args()
    import torch
    import av1dec_b200 as pkg
    from av1dec_b200 import format as F
    from av1dec_b200.engine import Engine
    pkg.load_engine()
    pkg.load_decoder()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    hdr_size = C.sizeof(F.FrameHdr)
    streams = {n: (n, d, w) for n, d, w in bench.load_streams()}
    side = torch.cuda.Stream(device=dev)
    for name in a.names:
        n, data, want = streams[name]
        rs = bench.record_stream(pkg, None, n, data, want, 0)
        eng = Engine(max(rs.max_w, 16), max(rs.max_h, 16), device=0, stream=side.cuda_stream)
        frames = []
        for buf, nb, refresh, show in rs.host_frames:
            frames.append((None, None, refresh, nb) if buf is None else (eng.upload(buf), buf[:hdr_size], refresh, -1))

        def replay():
            for ptr, hdr, refresh, slot in frames:
                if ptr is None:
                    eng.show_existing(slot, refresh)
                else:
                    eng.submit_resident(ptr, hdr, pkg.STAGE_ALL, refresh)
        for _ in range(3):
            replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(side)
        for _ in range(a.reps):
            replay()
        eng.join()
        e.record(side)
        e.synchronize()
        total = s.elapsed_time(e) / a.reps
        eng.set_lanes(1)  # the per-stage split is taken with one frame on the device at a time
        eng.set_profiling(True)
        eng.stage_times(reset=True)
        for _ in range(a.reps):
            replay()
        torch.cuda.synchronize()
        ms, calls = eng.stage_times()
        print(json.dumps({"stream": name, "frames": len(frames), "mpix": rs.pixels / 1e6, "w": rs.max_w, "h": rs.max_h,
                          "ms_per_replay": total, "mpix_per_s": rs.pixels / total / 1e3,
                          "stage_ms": {k: v / a.reps for k, v in ms.items()}, "calls": {k: v // a.reps for k, v in calls.items()}}))
        eng.set_profiling(False)
        eng.close()


if __name__ == "__main__":
    main()

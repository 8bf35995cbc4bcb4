"""End-to-end throughput of the whole conformance set through av1b_decode_ivf (host buffers in,
I420 out), N caller threads pulling streams from one queue.  Usage: python tools/e2e_set.py [threads] [passes]"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import av1dec_b200 as pkg

threads = int(sys.argv[1]) if len(sys.argv) > 1 else max(1, (os.cpu_count() or 4) * 3 // 4)
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pkg.load_engine()
lib = pkg.load_decoder()
bits = os.path.join(ROOT, "tests", "golden", "bits")
names = sorted(f for f in os.listdir(bits) if f.endswith(".ivf"))
datas = sorted((open(os.path.join(bits, n), "rb").read() for n in names), key=len, reverse=True)


def run(jobs):
    it = iter(jobs)
    lock = threading.Lock()
    px = [0]
    log = []
    t0 = time.perf_counter()

    def work():
        while True:
            with lock:
                d = next(it, None)
            if d is None:
                return
            t1 = time.perf_counter()
            _, _, p = pkg.decode_ivf(d, device=0, want_yuv=True, lib=lib)
            t2 = time.perf_counter()
            with lock:
                px[0] += p
                log.append((t2 - t1, t1 - t0, len(d)))
    ts = [threading.Thread(target=work) for _ in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    if os.environ.get("E2E_LOG"):
        log.sort(reverse=True)
        print("  longest jobs (s, start, bytes):", [(round(a, 3), round(b, 3), c) for a, b, c in log[:6]], "sum of job times", round(sum(a for a, _, _ in log), 2))
    return px[0], dt


run(datas)
for _ in range(3):
    run(datas * passes)
c0 = pkg.alloc_counters()
p, dt = run(datas * passes)
print("  alloc counters delta (ctx new, ctx reused, device allocs, pinned allocs):", tuple(b - a for a, b in zip(c0, pkg.alloc_counters())))
print("e2e", threads, "threads:", round(p / dt / 1e6, 1), "Mpix/s", round(dt / passes * 1e3, 1), "ms per set", flush=True)

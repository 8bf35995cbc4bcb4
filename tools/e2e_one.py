"""Single-stream end-to-end latency: av1b_decode_ivf on one conformance stream, three times.

    python tools/e2e_one.py av1-1-b8-02-allintra.ivf      (AV1B200_TIMING=1 prints the host phase timers)
"""
import sys, os, time
sys.path.insert(0,'/root/repo')
import av1dec_b200 as pkg
pkg.load_engine(); pkg.load_decoder()
name=sys.argv[1]
data=open(os.path.join('/root/repo/tests/golden/bits',name),'rb').read()
for i in range(3):
    t=time.perf_counter(); yuv,frames,px=pkg.decode_ivf(data, device=0); dt=time.perf_counter()-t
    print(name, 'frames',frames,'ms %.1f'%(dt*1e3), 'Mpix/s %.1f'%(px/dt/1e6), flush=True)

"""One CIF stream through the sequence encoder (p64b_enc_*, frames pipelined three deep): host time per encode() call.
A single stream is latency-bound on a GPU (396 macroblocks per launch); reported for completeness."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p64_b200 import y4m
from p64_b200.encoder import Encoder
it = y4m.IT_CIF
clip = y4m.synth_clip(it, 60, seed=4)
for kw in (dict(q=8), dict(rate=384000), dict(rate=2000000), dict(rate=384000, host_vlc=True)):
    enc = Encoder(it, 1, **kw)
    ts = []
    for fr in clip:
        t0 = time.perf_counter(); enc.encode(fr[None]); ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); enc.finish(); tf = time.perf_counter() - t0
    print(kw, "per-frame ms: first %.2f median %.3f max(after 5) %.3f finish %.2f  ovf %d bytes %d" % (ts[0] * 1e3, np.median(ts) * 1e3, max(ts[5:]) * 1e3, tf * 1e3, enc.overflows(0), len(enc.data(0))))
    enc.close()

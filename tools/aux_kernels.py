"""Exercises the auxiliary kernels at the bench's batch size (256 CIF streams) for ncu / timing: ingest_chroma_kernel
(420paldv, the most expensive conversion), plane_stats_kernel, mb_decode_kernel.  Prints CUDA-event timings."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p64_b200 import y4m  # noqa: E402
from p64_b200.encoder import DeviceContext, Parser, encode_clip, make_step  # noqa: E402

IT, S = y4m.IT_CIF, 256
clip = y4m.synth_clip(IT, 3, seed=3)
ctx = DeviceContext(IT, S)
ctx.set_input_chroma("420paldv")
raw = np.repeat(clip[1][None], S, 0)
t0 = time.perf_counter(); ctx.convert_frames(raw); t_conv = time.perf_counter() - t0
t0 = time.perf_counter(); ctx.convert_frames(raw); t_conv = min(t_conv, time.perf_counter() - t0)
ctx.close()

ctx = DeviceContext(IT, S)
for f in range(3):
    ctx.encode_frames(make_step(f == 0, 8, 1, 31), np.repeat(clip[f][None], S, 0))
t0 = time.perf_counter(); ctx.statistics(); t_stat = time.perf_counter() - t0
t0 = time.perf_counter(); ctx.statistics(); t_stat = min(t_stat, time.perf_counter() - t0)
ctx.close()

data = encode_clip(IT, clip, q=8, me_mode=1, search_limit=31)
p = Parser(data)
pics = []
while True:
    r = p.next_picture()
    if r is None:
        break
    pics.append(r)
p.close()
ctx = DeviceContext(IT, S)
t_dec = []
for mbs, lv, tr, rep in pics:
    m = np.repeat(mbs[None], S, 0); l = np.repeat(lv[None], S, 0)
    t0 = time.perf_counter(); ctx.decode_frames(m, l); t_dec.append(time.perf_counter() - t0)
ctx.close()
print(json.dumps({"convert_frames_wall_ms": t_conv * 1e3, "statistics_wall_ms": t_stat * 1e3, "decode_frames_wall_ms": [t * 1e3 for t in t_dec],
                  "note": "host wall clock incl. PCIe copies of 256 CIF frames; kernel times are in the ncu launch list"}))

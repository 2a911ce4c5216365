"""How content-dependent is the exhaustive search's exact early exit?  Times p64b_ctx_motion_estimation_dev on 256 CIF
pairs of four kinds and reports ms per launch, executed / algorithmic packed-SAD ops (device counter).
    python tools/me_content_sweep.py  > profiles/rNN_me_content_sweep.jsonl"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p64_b200 import y4m  # noqa: E402
from p64_b200.encoder import DeviceContext  # noqa: E402

IT, S = y4m.IT_CIF, 256
W, H = y4m.DIMS[IT]
ALG = 343473 * 64 * S
rng = np.random.default_rng(1)


def pairs(kind):
    ref = np.empty((S, H, W), np.uint8); cur = np.empty((S, H, W), np.uint8)
    for s in range(S):
        if kind == "bench clip (texture + noise 20, pan <= 2, temporal noise 3)":
            c = y4m.synth_clip(IT, 2, seed=1000 + s % 8, pan=((s % 5) - 2, (s % 3) - 1))
            ref[s], cur[s] = c[0, :W * H].reshape(H, W), c[1, :W * H].reshape(H, W)
        elif kind == "camera pan (9,-7) per frame, same texture (the previous vector is the hint)":
            c = y4m.synth_clip(IT, 2, seed=1000 + s % 8, pan=(9, -7))
            ref[s], cur[s] = c[0, :W * H].reshape(H, W), c[1, :W * H].reshape(H, W)
        elif kind == "static scene + temporal noise 3":
            c = y4m.synth_clip(IT, 2, seed=1000 + s % 8, pan=(0, 0))
            ref[s], cur[s] = c[0, :W * H].reshape(H, W), c[1, :W * H].reshape(H, W)
        elif kind == "shifted copy of uniform noise, noise 4 (BASELINE configs[3])":
            ref[s], cur[s] = y4m.random_pair(IT, 50 + s % 8, shift=((s % 29) - 14, (s % 23) - 11), noise=4)
        else:
            ref[s] = rng.integers(0, 256, (H, W)); cur[s] = rng.integers(0, 256, (H, W))
    return ref, cur


ctx = DeviceContext(IT, S)
stream = torch.cuda.Stream()
ctx.set_cuda_stream(stream.cuda_stream)
out = torch.zeros(S * 396 * 8, dtype=torch.int32, device="cuda")
for kind in ["bench clip (texture + noise 20, pan <= 2, temporal noise 3)", "camera pan (9,-7) per frame, same texture (the previous vector is the hint)",
             "static scene + temporal noise 3",
             "shifted copy of uniform noise, noise 4 (BASELINE configs[3])", "unrelated uniform noise (worst case: nothing can be skipped)"]:
    ref, cur = pairs(kind)
    r, c = torch.from_numpy(ref).cuda(), torch.from_numpy(cur).cuda()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for _ in range(3):
            ctx.motion_estimation_dev(r.data_ptr(), c.data_ptr(), S, 1, 31, out.data_ptr())
        ctx.me_executed(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20):
            ctx.motion_estimation_dev(r.data_ptr(), c.data_ptr(), S, 1, 31, out.data_ptr())
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    ex = ctx.me_executed() / 20
    print(json.dumps({"content": kind, "ms_per_launch": round(ms, 4), "executed_share_of_algorithmic": round(ex / ALG, 3),
                      "algorithmic_T_ops_s": round(ALG / ms / 1e9, 2), "executed_T_ops_s": round(ex / ms / 1e9, 2)}))
ctx.close()

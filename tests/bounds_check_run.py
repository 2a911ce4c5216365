"""TEST INFRASTRUCTURE.  Run by tests/test_bounds_check.py in a subprocess with P64B_LIB pointing at the -DP64B_BOUNDS_CHECK build of the
library: drives the motion-estimation, macroblock and decoder kernels over edge / corner macroblocks (every picture size, vectors
that reach the frame edges, every search mode and range), ragged stream / pair counts and rate control, then reads the device's
violation counter.  Also checks the debug build still produces the oracle's bytes (the checks must not change results)."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_encode_stream  # noqa: E402
from p64_b200 import _lib, y4m  # noqa: E402
from p64_b200.encoder import Decoder, DeviceContext, Encoder  # noqa: E402

L = _lib.lib()
v, line = C.c_uint32(), C.c_uint32()
assert L.p64b_debug_oob(0, C.byref(v), C.byref(line)) == 1, "not a -DP64B_BOUNDS_CHECK build"
rng = np.random.default_rng(5)
cases = 0
for it in (y4m.IT_QCIF, y4m.IT_CIF, y4m.IT_NTSC):
    w, h = y4m.DIMS[it]
    for S, pan, kw in [(1, (13, -14), dict(q=8, me_mode=1, search_limit=31)), (3, (-15, 15), dict(q=5, me_mode=1, search_limit=15)),
                       (2, (7, 3), dict(q=12, me_mode=1, search_limit=3)), (5, (-9, 11), dict(q=8, me_mode=0)),
                       (3, (4, -6), dict(rate=64000 if it == y4m.IT_QCIF else 256000, me_mode=1, search_limit=31)),
                       (1, (0, 0), dict(q=8, force_intra=True)), (2, (2, 1), dict(q=31, me_mode=0, host_vlc=True))]:
        clips = [y4m.synth_clip(it, 4, seed=int(rng.integers(1 << 30)), pan=pan, noise=int(rng.integers(0, 25))) for _ in range(S)]
        if S > 1:                    # unrelated noise in one stream: every candidate survives, long survivor lists
            clips[-1] = np.stack([rng.integers(0, 256, clips[0].shape[1]).astype(np.uint8) for _ in range(4)])
        enc = Encoder(it, S, **kw)
        for f in range(4):
            enc.encode(np.stack([c[f] for c in clips]))
        enc.finish()
        got = [enc.data(s) for s in range(S)]
        enc.close()
        want, recons, _ = oracle_encode_stream(it, clips[0], **{k: x for k, x in kw.items() if k != "host_vlc"})
        assert got[0] == want, (it, S, kw)
        dec = Decoder(got[0]); fr = dec.frames(); dec.close()
        assert np.array_equal(fr[-1], recons[-1])
        cases += 1
    # motion estimation alone on ragged pair counts, incl. flat pairs (all ties) and the SAD surface variant
    import torch
    ctx = DeviceContext(it, 1)
    nmb = (w // 16) * (h // 16)
    for n_pairs in (1, 5, 7):
        ref = rng.integers(0, 256, (n_pairs, h, w)).astype(np.uint8)
        cur = np.roll(ref, (int(rng.integers(-15, 16)), int(rng.integers(-15, 16))), axis=(1, 2)).copy()
        ref[0] = 128; cur[0] = 128
        r, c = torch.from_numpy(ref).cuda(), torch.from_numpy(cur).cuda()
        out = torch.zeros(n_pairs * nmb * 8, dtype=torch.int32, device="cuda")
        surf = torch.zeros(n_pairs * nmb * 961, dtype=torch.int32, device="cuda")
        ctx.set_cuda_stream(torch.cuda.current_stream().cuda_stream)
        for mode, limit in ((1, 31), (1, 15), (1, 1), (0, 15)):
            ctx.motion_estimation_dev(r.data_ptr(), c.data_ptr(), n_pairs, mode, limit, out.data_ptr())
        _lib.check(L.p64b_ctx_sad_surface_dev(ctx.h, C.c_void_p(r.data_ptr()), C.c_void_p(c.data_ptr()), n_pairs, C.c_void_p(out.data_ptr()), C.c_void_p(surf.data_ptr())))
        torch.cuda.synchronize()
        cases += 1
    ctx.close()
assert L.p64b_debug_oob(0, C.byref(v), C.byref(line)) == 1
print(json.dumps({"cases": cases, "violations": int(v.value), "last_line": int(line.value)}))
sys.exit(1 if v.value else 0)

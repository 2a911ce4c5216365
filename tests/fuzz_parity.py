"""TEST INFRASTRUCTURE.  Randomised parity soak (not collected by pytest: it runs for minutes).  Random picture size, content, quantiser or
bit rate, search mode / range, input chroma type, stream count; the device path (device-side VLC / rate control, batch of
streams) must give, stream by stream, the bytes of the CPU oracle + host bit writer, and the decoder must reproduce the
encoder's reconstruction.   python tests/fuzz_parity.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))      # the oracle is used here as the checker only
from helpers import oracle_encode_stream  # noqa: E402
from oracle import oracle as O  # noqa: E402
from p64_b200 import y4m  # noqa: E402
from p64_b200.encoder import Decoder, Encoder  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
CHROMAS = ["420jpeg", "420jpeg", "420jpeg", "420mpeg2", "420paldv", "422", "411", "444", "mono"]


def content(it, nf, chroma):
    w, h = y4m.DIMS[it]
    kind = rng.integers(0, 6)
    seed = int(rng.integers(0, 1 << 30))
    if kind <= 2:
        clip = y4m.synth_payloads(it, nf, seed, chroma) if chroma != "420jpeg" else \
            y4m.synth_clip(it, nf, seed, pan=(int(rng.integers(-7, 8)), int(rng.integers(-7, 8))), noise=int(rng.integers(0, 30)),
                           temporal_noise=int(rng.integers(0, 6)))
        return clip
    n = O.payload_bytes(w, h, chroma)
    r = np.random.default_rng(seed)
    if kind == 3:                                   # unrelated noise frames, flat frames in between
        return np.stack([r.integers(0, 256, n).astype(np.uint8) if f % 3 else np.full(n, int(r.integers(0, 256)), np.uint8) for f in range(nf)])
    if kind == 4:                                   # slowly changing low-amplitude noise: many type-2/4/5/7 decisions
        base = r.integers(100, 140, n)
        return np.stack([np.clip(base + r.integers(-2, 3, n), 0, 255).astype(np.uint8) for _ in range(nf)])
    base = (np.arange(n) % 256).astype(np.int64)    # ramps: ties everywhere
    return np.stack([((base + 3 * f) % 256).astype(np.uint8) for f in range(nf)])


t0 = time.time()
cases = fails = 0
while time.time() - t0 < budget:
    it = int(rng.integers(0, 3))
    nf = int(rng.integers(2, 7)) if rng.random() < 0.8 else int(rng.integers(8, 16))     # long clips replay the captured rate-control graphs
    S = int(rng.choice([1, 1, 2, 3, 5]))
    chroma = str(rng.choice(CHROMAS))
    rate = int(rng.choice([0, 0, 0, 48000, 64000, 128000, 384000, 2000000]))
    q = 0 if rate and rng.random() < 0.7 else int(rng.integers(1, 32))
    full = bool(rng.integers(0, 2))
    limit = int(rng.choice([31, 31, 15, 8, 3])) if full else 15
    intra = rng.random() < 0.1
    clips = [content(it, nf, chroma) for _ in range(S)]
    kw = dict(q=q, rate=rate, me_mode=int(full), search_limit=limit, force_intra=intra)
    if rng.random() < 0.4:                          # -f, -k, -a: they enter the buffer model and the temporal references
        fr = [(30000, 1001), (25, 1), (15, 1), (10, 1)][int(rng.integers(0, 4))]
        kw.update(frame_rate=fr, frame_skip=int(rng.integers(1, 4)), start_frame=int(rng.integers(0, 70)))
    devs = None if rng.random() < 0.6 else [0] * int(rng.integers(1, 5))         # the batch partitioned over device contexts (one GPU, listed k times)
    enc = Encoder(it, S, input_chroma=chroma, devices=devs, balance_links=bool(devs) and rng.random() < 0.3, **kw)
    for f in range(nf):
        enc.encode(np.stack([c[f] for c in clips]))
    enc.finish()
    got = [enc.data(s) for s in range(S)]
    ovf = [enc.overflows(s) for s in range(S)]
    enc.close()
    for s in range(S):
        want, recons, wovf = oracle_encode_stream(it, clips[s], input_chroma=None if chroma == "420jpeg" else chroma, **kw)
        ok = got[s] == want and ovf[s] == wovf
        if ok and s == 0:
            dec = Decoder(got[s]); fr = dec.frames(); dec.close()
            # a picture is written once per temporal-reference step to the next one (p64.c:1047-1054): -k repeats pictures
            ok = len(fr) == (nf - 1) * kw.get("frame_skip", 1) + 1 and np.array_equal(fr[-1], recons[-1])
        if not ok:
            fails += 1
            print("MISMATCH", dict(it=it, nf=nf, S=S, s=s, chroma=chroma, devs=devs, **kw), len(got[s]), len(want), ovf[s], wovf, flush=True)
    cases += 1
print(f"fuzz: {cases} cases, {fails} mismatches, {time.time() - t0:.0f} s")
sys.exit(1 if fails else 0)

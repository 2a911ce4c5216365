"""Generates tests/golden/ingest.json: md5s of what the reference's OWN Y4M reader (vidinput.c / y4m_input.c compiled into
oracle/_ref/libp64ref.so, called through oracle/ref_shim.c:ref_y4m_frames) hands the encoder for seeded frame payloads of
every chroma type it accepts (y4m_input.c:587-655).  Run where /root/reference is mounted:
`python tests/golden/make_ingest_golden.py`.  The tests regenerate the same payloads from the seeds."""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

TYPES = ["420jpeg", "420mpeg2", "420paldv", "422", "411", "444", "444alpha", "mono"]
DIMS = [(352, 288), (176, 144), (352, 240)]


def payloads(w, h, chroma, seed):
    """three frames: uniform noise, smooth ramp + noise, saturated 0/255 checker noise (filter overshoot, clamps)"""
    rng = np.random.default_rng(seed)
    n = O.payload_bytes(w, h, chroma)
    a = rng.integers(0, 256, n).astype(np.uint8)
    b = ((np.arange(n) // 3) % 256 + rng.integers(-3, 4, n)).clip(0, 255).astype(np.uint8)
    c = np.where(rng.integers(0, 2, n) > 0, 255, 0).astype(np.uint8)
    return [a, b, c]


def main():
    out = {}
    tmp = tempfile.mkdtemp()
    for (w, h) in DIMS:
        for k, chroma in enumerate(TYPES):
            seed = 5000 + 10 * k + w
            pay = payloads(w, h, chroma, seed)
            O.write_y4m_raw(f"{tmp}/a.y4m", w, h, pay, chroma, frame_params=b" Xparam" if k % 2 else b"")
            ref = O.ref_y4m_frames(f"{tmp}/a.y4m", w, h)
            assert len(ref) == len(pay)
            out[f"{w}x{h}_{chroma}"] = dict(w=w, h=h, chroma=chroma, seed=seed,
                                            payload_md5=hashlib.md5(b"".join(p.tobytes() for p in pay)).hexdigest(),
                                            frames_md5=[hashlib.md5(f.tobytes()).hexdigest() for f in ref])
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ingest.json"), "w"), indent=1)
    print(len(out), "cases")


if __name__ == "__main__":
    main()

"""Test-only helpers: the CPU oracle (oracle/) driven through the PRODUCT's host bit writer, restating the
reference's sequence / rate control loop (p64.c:524-786) in Python so whole streams can be checked without a GPU."""
import hashlib
import json
import os

import numpy as np

from oracle import oracle as O
from p64_b200 import y4m
from p64_b200.encoder import MB_DTYPE, BitWriter

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "streams.json")))


def golden_clip(name):
    g = GOLDEN[name]
    clip = y4m.synth_payloads(g["image_type"], g["n_frames"] + g["args"].get("start", 0), g["seed"], g["args"].get("chroma", "420jpeg"))
    assert hashlib.md5(clip.tobytes()).hexdigest() == g["clip_md5"], "synthetic clip generator drifted"
    return g, clip[g["args"].get("start", 0):]      # -a with Y4M input discards StartFrame frames first (p64.c:562-565)


def golden_kwargs(g):
    """reference CLI arguments of a golden case -> keyword arguments of p64_b200.encoder.Encoder"""
    a = g["args"]
    return dict(q=a.get("q", 0), rate=a.get("rate", 0), me_mode=1 if a.get("full_search") else 0,
                search_limit=a.get("search_limit") or 15, force_intra=bool(a.get("intra_only")),
                **({"input_chroma": a["chroma"]} if a.get("chroma") else {}),
                **({"start_frame": a["start"], "frame_skip": a["frame_skip"], "last_frame": a["last"]} if a.get("frame_skip") else {}),
                **({"frame_rate": (a["frame_rate"], 1)} if a.get("frame_rate") else {}))


def recs_to_mb(recs):
    mb = np.zeros(len(recs), MB_DTYPE)
    mb["mtype"], mb["cbp"], mb["mvx"], mb["mvy"], mb["quant"] = recs[:, 0], recs[:, 1], recs[:, 2], recs[:, 3], recs[:, 4]
    return mb


def levels_to_i8(levels):
    """int32 oracle levels -> the ABI's int8 storage (intra DC 1..254 wraps into the same byte)."""
    return levels.astype(np.uint8).view(np.int8)


def oracle_encode_stream(image_type, clip, *, q=0, rate=0, me_mode=0, search_limit=15, force_intra=False,
                         frame_rate=(30000, 1001), frame_skip=1, start_frame=0, input_chroma=None, last_frame=None):
    """Returns (.p64 bytes, per-frame recon list, overflow count).  input_chroma: `clip` holds unconverted Y4M payloads."""
    if input_chroma:
        w, h = y4m.DIMS[image_type]
        clip = np.stack([O.y4m_payload_to_encoder_frame(fr, w, h, input_chroma) for fr in clip])
    enc = O.Encoder(image_type)
    bw = BitWriter(image_type)
    ngob = enc.ngob
    qdfact, qoffs = 1, 1
    iq = q
    if rate:
        qdfact = rate // 320
        if not iq:
            iq = min(max(10000000 // rate, 1), 31)
    if not iq:
        iq = 8
    gquant, boff, ovfl = iq, 0, 0
    denom = ngob * 33 * frame_rate[0] // frame_rate[1]

    def c_int(v):                                   # the reference computes in C `int`: products wrap at 32 bits
        return (v + 2 ** 31) % 2 ** 32 - 2 ** 31

    def c_div(a, b):                                # C division truncates toward zero
        return abs(a) // abs(b) * (1 if (a >= 0) == (b > 0) else -1)

    def contents(g, m):                             # BufferContents(), p64.c:233-236
        return bw.tell() + boff - c_div(c_int((g * 33 + m) * rate * frame_skip), denom)

    recons = []
    cur = start_frame
    for f, fr in enumerate(clip):
        first = f == 0
        bw.picture_header(cur % 32)
        enc.begin_frame(fr, me_mode, search_limit)
        for g in range(ngob):
            if rate and not first:
                c = contents(g, 0)
                gquant = min(max(c_div(c_int(c), qdfact) + qoffs, 1), 31)
            bw.gob_header(g, gquant)
            for m in range(33):
                over = bool(rate) and contents(g, m) > rate // 4
                ovfl += over
                rec, lv = enc.encode_mb(g, m, gquant, force_intra, over)
                bw.mb(m, recs_to_mb(rec[None])[0], levels_to_i8(lv))
        enc.end_frame()
        recons.append(enc.recon())
        if rate:
            if first:
                boff = (rate // 4) // 2 - contents(ngob, 0)
            boff -= c_div(c_int(rate * frame_skip * frame_rate[1]), frame_rate[0])      # the product wraps before the division (p64.c:677)
        cur += frame_skip
    # p64.c:600-602: "limit file growth" -- CurrentFrame is clamped to LastFrame+1 (-b; unknown = the last frame coded)
    last = last_frame if last_frame is not None else cur - frame_skip
    bw.picture_header(min(cur, last + 1) % 32)
    bw.finish()
    return bw.data(), recons, ovfl

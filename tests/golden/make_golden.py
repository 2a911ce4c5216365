"""Generates tests/golden/streams.json: md5/size of the .p64 streams the UNMODIFIED reference encoder
(oracle/_ref/p64_ref = stock three-step search, p64_ref_fs = FastBME toggled at me.c:351-352) writes for
seeded synthetic clips, plus md5s of the frames its own decoder reconstructs from them (closed loop,
p64.c:971-1013).  Run here (where /root/reference is mounted): `python tests/golden/make_golden.py`.
The GPU box has no reference tree; tests compare against this committed file.
"""
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from p64_b200 import y4m  # noqa: E402

CASES = [
    # name, image_type, n_frames, seed, kwargs for ref_encode
    ("cif30_q8_tss", y4m.IT_CIF, 30, 1234, dict(q=8)),
    ("cif30_q8_full31", y4m.IT_CIF, 30, 1234, dict(q=8, full_search=True, search_limit=31)),
    ("cif8_q8_full15", y4m.IT_CIF, 8, 1234, dict(q=8, full_search=True)),
    ("cif6_q3_tss", y4m.IT_CIF, 6, 99, dict(q=3)),
    ("cif6_q31_full31", y4m.IT_CIF, 6, 99, dict(q=31, full_search=True, search_limit=31)),
    ("qcif12_q8_intra", y4m.IT_QCIF, 12, 4321, dict(q=8, intra_only=True)),
    ("qcif12_q8_tss", y4m.IT_QCIF, 12, 4321, dict(q=8)),
    ("qcif140_q10_tss", y4m.IT_QCIF, 140, 7, dict(q=10)),          # forced-intra refresh fires (LastIntra > 131)
    ("ntsc7_q8_tss", y4m.IT_NTSC, 7, 77, dict(q=8)),
    ("ntsc7_q8_full31", y4m.IT_NTSC, 7, 77, dict(q=8, full_search=True, search_limit=31)),
    ("cif30_r384000_tss", y4m.IT_CIF, 30, 1234, dict(rate=384000)),
    ("cif30_r384000_full31", y4m.IT_CIF, 30, 1234, dict(rate=384000, full_search=True, search_limit=31)),
    ("cif12_r128000_tss", y4m.IT_CIF, 12, 1234, dict(rate=128000)),  # buffer-overflow branch (p64.c:776-783)
    ("cif12_r64000_full31", y4m.IT_CIF, 12, 1234, dict(rate=64000, full_search=True, search_limit=31)),
    ("qcif20_r64000_tss", y4m.IT_QCIF, 20, 4321, dict(rate=64000)),
    # ingest: Y4M chroma types the reference's reader converts (y4m_input.c:195-545) before the encoder sees the frame
    ("cif5_q8_tss_420mpeg2", y4m.IT_CIF, 5, 31, dict(q=8, chroma="420mpeg2")),
    ("qcif6_q6_full31_420paldv", y4m.IT_QCIF, 6, 32, dict(q=6, full_search=True, search_limit=31, chroma="420paldv")),
    ("cif4_q8_tss_422", y4m.IT_CIF, 4, 33, dict(q=8, chroma="422")),
    ("qcif5_q8_tss_411", y4m.IT_QCIF, 5, 34, dict(q=8, chroma="411")),
    ("qcif4_q8_tss_444", y4m.IT_QCIF, 4, 35, dict(q=8, chroma="444")),
    ("ntsc3_q8_tss_444alpha", y4m.IT_NTSC, 3, 36, dict(q=8, chroma="444alpha")),
    ("qcif5_r64000_tss_mono", y4m.IT_QCIF, 5, 37, dict(rate=64000, chroma="mono")),
    # -a / -k / -b: StartFrame, FrameSkip and where LastFrame falls between two coded frames (temporal references, the
    # clamp of the trailing picture header p64.c:600-602, the per-frame deduction of the buffer model)
    ("qcif6_q8_tss_a3_k2_b14", y4m.IT_QCIF, 6, 38, dict(q=8, start=3, frame_skip=2, last=14)),
    ("qcif5_r64000_a30_k3_b42", y4m.IT_QCIF, 5, 39, dict(rate=64000, start=30, frame_skip=3, last=42)),
    # -r 2000000 -k 3 -f 15 on CIF: (CurrentGOB*33+CurrentMDU)*Rate*FrameSkip exceeds 2^31 and WRAPS in the reference's C int
    # arithmetic (p64.c:233-236) -- the buffer model then sees a negative deduction and overflows; part of the behaviour
    ("cif4_r2000000_full31_a50_k3_f15", y4m.IT_CIF, 4, 77, dict(rate=2000000, full_search=True, search_limit=31, start=50,
                                                                frame_skip=3, last=59, frame_rate=15)),
    # -r 1100000 -k 2 at the default 30000/1001: Rate*FrameSkip*FrameRateDiv of the per-frame deduction (p64.c:677) exceeds 2^31
    # and wraps BEFORE the division by FrameRate (305 overflow macroblocks follow); -r 1000000 does not wrap
    ("qcif5_r1100000_k2", y4m.IT_QCIF, 5, 40, dict(rate=1100000, start=0, frame_skip=2, last=8)),
    ("qcif5_r1000000_k2", y4m.IT_QCIF, 5, 40, dict(rate=1000000, start=0, frame_skip=2, last=8)),
]


def main():
    if not O.have_ref():
        raise SystemExit("oracle/_ref missing: run `make -C oracle` where /root/reference is mounted")
    out = {}
    tmp = tempfile.mkdtemp()
    for name, it, nf, seed, kw in CASES:
        kw = dict(kw)
        chroma = kw.pop("chroma", "420jpeg")
        start, skip, last = kw.pop("start", 0), kw.pop("frame_skip", 1), kw.pop("last", None)
        fr = kw.pop("frame_rate", None)
        clip = y4m.synth_payloads(it, nf + start, seed, chroma)
        y4m.write_y4m(f"{tmp}/c.y4m", it, clip, chroma=chroma)
        extra = ("-k", str(skip), "-b", str(last)) if last is not None else ()     # (a later -b overrides ref_encode's)
        extra += ("-f", str(fr)) if fr else ()
        log = O.ref_encode(f"{tmp}/c.y4m", f"{tmp}/o.p64", it, nf, start=start, extra=extra, **kw)
        if chroma != "420jpeg":
            kw["chroma"] = chroma
        if last is not None:
            kw.update(start=start, frame_skip=skip, last=last)
        if fr:
            kw["frame_rate"] = fr
        data = open(f"{tmp}/o.p64", "rb").read()
        O.ref_decode(f"{tmp}/o.p64", f"{tmp}/dec")
        _, _, dec = y4m.read_y4m(f"{tmp}/dec.y4m")
        out[name] = dict(image_type=it, n_frames=nf, seed=seed, args=kw, clip_md5=hashlib.md5(clip.tobytes()).hexdigest(),
                         size=len(data), md5=hashlib.md5(data).hexdigest(), overflows=log.count("Buffer Overflow!"),
                         decoded_frames=len(dec), last_recon_md5=hashlib.md5(dec[-1].tobytes()).hexdigest(),
                         recon_md5=[hashlib.md5(f.tobytes()).hexdigest() for f in dec[:: max(1, len(dec) // 6)]])
        print(name, len(data), out[name]["md5"], "ovf", out[name]["overflows"], "dec", len(dec))
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "streams.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

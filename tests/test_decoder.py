"""Decoder: the host bit-stream parser (p64b_parser_*) and the device's inverse half
(mb_decode_kernel through p64b_ctx_decode_frames / p64b_dec_*).  CPU: parsing the reference encoder's golden streams
recovers exactly the records and levels the encoder's hot path (oracle) produced.  GPU: decoding those streams gives the
frames the REFERENCE DECODER wrote (md5s in tests/golden/streams.json); foreign-stream features the encoder never emits
(skipped macroblocks, missing GOBs, MQUANT types, temporal-reference gaps) against a NumPy/oracle model."""
import hashlib
import subprocess

import numpy as np
import pytest

from helpers import GOLDEN, golden_clip, golden_kwargs, levels_to_i8, oracle_encode_stream, recs_to_mb
from oracle import oracle as O
from p64_b200 import y4m

PLAIN = [n for n in GOLDEN if not GOLDEN[n]["args"].get("chroma")]


def _oracle_stream_with_records(name):
    """the oracle encoder's stream for a golden case + per-frame (records, levels) as the encoder produced them"""
    from p64_b200.encoder import BitWriter
    g, clip = golden_clip(name)
    kw = golden_kwargs(g)
    assert not kw.get("rate")
    enc, bw = O.Encoder(g["image_type"]), BitWriter(g["image_type"])
    frames = []
    for f, fr in enumerate(clip):
        bw.picture_header(f % 32)
        recs, lv = enc.encode_frame(fr, kw["q"], kw["me_mode"], kw["search_limit"], force_intra=kw["force_intra"])
        mbs, lv8 = recs_to_mb(recs), levels_to_i8(lv)
        for gob in range(enc.ngob):
            bw.gob_header(gob, kw["q"])
            for m in range(33):
                bw.mb(m, mbs[gob * 33 + m], lv8[gob * 33 + m])
        frames.append((mbs, lv8, enc.recon().copy()))
    bw.picture_header(len(clip) % 32)
    bw.finish()
    data = bw.data()
    assert hashlib.md5(data).hexdigest() == g["md5"]
    return g, data, frames


@pytest.mark.parametrize("name", ["qcif12_q8_tss", "cif6_q3_tss", "ntsc7_q8_full31", "qcif12_q8_intra", "cif6_q31_full31"])
def test_parser_recovers_encoder_records_and_levels(name):
    from p64_b200.encoder import Parser
    g, data, frames = _oracle_stream_with_records(name)
    p = Parser(data)
    assert p.image_type == g["image_type"]
    tcoef = np.array([1, 1, 1, 1, 0, 1, 1, 0, 1, 1], bool)
    for f, (mbs, lv8, _) in enumerate(frames):
        got = p.next_picture()
        assert got is not None, f
        gm, gl, tr, rep = got
        assert (tr, rep) == (f % 32, 1)
        assert np.all(gm["reserved"] == 1)
        for k in ("mtype", "cbp", "mvx", "mvy", "quant"):
            assert np.array_equal(gm[k], mbs[k]), (k, f)
        # levels of blocks the stream carries (CBP bit set and a coefficient type); the rest is not transmitted
        for i in range(len(mbs)):
            for c in range(6):
                if tcoef[mbs["mtype"][i]] and (mbs["cbp"][i] >> (5 - c)) & 1:
                    assert np.array_equal(gl[i, c], lv8[i, c]), (f, i, c)
                else:
                    assert not gl[i, c].any()
    assert p.next_picture() is None
    p.close()


def test_parser_rejects_garbage_and_survives_truncation():
    from p64_b200._lib import P64Error
    from p64_b200.encoder import Parser
    with pytest.raises(P64Error):
        Parser(b"\x12\x34\x56\x78" * 10)
    g, data, frames = _oracle_stream_with_records("qcif12_q8_tss")
    for cut in (len(data) // 3, len(data) // 2 + 7, len(data) - 5):
        p = Parser(data[:cut])
        n = 0
        while p.next_picture() is not None:
            n += 1
        assert 1 <= n <= len(frames)
        p.close()
    rng = np.random.default_rng(1)
    for trial in range(20):                              # bit errors: must terminate without reading out of bounds
        bad = bytearray(data[:20000])
        for pos in rng.integers(40, len(bad), 30):
            bad[pos] ^= 1 << int(rng.integers(0, 8))
        p = Parser(bytes(bad))
        n = 0
        while p.next_picture() is not None and n < 100:
            n += 1
        p.close()


def _stream_with_vector(it, gob, m, mvx, mvy):
    """one intra picture, then a picture whose only macroblock is type 4 (MC, no coefficients) with the given vector"""
    from p64_b200.encoder import BitWriter, MB_DTYPE
    bw = BitWriter(it)
    lv = np.zeros((6, 64), np.int8); lv[:, 0] = 100
    ngob = 3 if it == y4m.IT_QCIF else 12
    bw.picture_header(0)
    for g in range(ngob):
        bw.gob_header(g, 8)
        for k in range(33):
            rec = np.zeros((), MB_DTYPE); rec["mtype"], rec["cbp"], rec["quant"] = 0, 0x3f, 8
            bw.mb(k, rec, lv)
    bw.picture_header(1)
    bw.gob_header(gob, 8)
    rec = np.zeros((), MB_DTYPE); rec["mtype"], rec["cbp"], rec["quant"], rec["mvx"], rec["mvy"] = 4, 0x3f, 8, mvx, mvy
    bw.mb(m, rec, np.zeros((6, 64), np.int8))
    bw.picture_header(2)
    bw.finish()
    return bw.data()


@pytest.mark.parametrize("gob,m,mvx,mvy,legal", [(0, 0, -16, -16, False), (0, 0, -1, 0, False), (0, 0, 0, -1, False), (0, 0, 15, 15, True),
                                               (2, 32, 1, 0, False), (2, 32, 0, 1, False), (2, 32, -16, -16, True), (1, 16, -16, 15, True)])
def test_parser_rejects_vectors_that_leave_the_picture(gob, m, mvx, mvy, legal):
    """a vector taken from the bit stream must keep the 16x16 prediction inside the picture: anything else is a corrupt
    stream (the picture ends there), never a record that makes the device read outside the frame store"""
    from p64_b200.encoder import Parser
    p = Parser(_stream_with_vector(y4m.IT_QCIF, gob, m, mvx, mvy))
    assert p.next_picture() is not None
    pic = p.next_picture()
    assert pic is not None
    mbs = pic[0]
    if legal:
        assert mbs["reserved"][gob * 33 + m] == 1 and (int(mbs["mvx"][gob * 33 + m]), int(mbs["mvy"][gob * 33 + m])) == (mvx, mvy)
    else:
        assert not mbs["reserved"].any()
    p.close()


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_device_decode_clamps_hostile_vectors():
    """records handed straight to p64b_ctx_decode_frames (no parser in front) with vectors far outside the picture, for the
    first and the last stream of a batch: the prediction block is clamped into the plane -- no fault, deterministic output"""
    from p64_b200.encoder import DeviceContext, MB_DTYPE
    it, S = y4m.IT_QCIF, 3
    w, h = y4m.DIMS[it]
    ctx = DeviceContext(it, S)
    nmb = ctx.geom["num_mb"]
    rng = np.random.default_rng(3)
    mbs = np.zeros((S, nmb), MB_DTYPE); lv = np.zeros((S, nmb, 6, 64), np.int8)
    mbs["mtype"], mbs["cbp"], mbs["quant"], mbs["reserved"] = 0, 0x3f, 8, 1
    lv[..., 0] = rng.integers(1, 127, (S, nmb, 6))
    ctx.decode_frames(mbs, lv)
    ref = [ctx.recon(s).copy() for s in range(S)]
    for mvx, mvy in [(-16, -16), (15, 15), (-16, 15), (15, -16)]:
        mbs["mtype"], mbs["mvx"], mbs["mvy"] = 4, mvx, mvy
        ctx.decode_frames(mbs, np.zeros_like(lv))
        for s in range(S):
            got = ctx.recon(s)
            Y, prevY = got[:w * h].reshape(h, w), ref[s][:w * h].reshape(h, w)
            for row, col in [(0, 0), (h // 16 - 1, w // 16 - 1), (4, 5)]:
                for c in range(4):
                    bx, by = col * 16 + (c & 1) * 8, row * 16 + (c >> 1) * 8
                    x, y = min(max(bx + mvx, 0), w - 8), min(max(by + mvy, 0), h - 8)
                    assert np.array_equal(Y[by:by + 8, bx:bx + 8], prevY[y:y + 8, x:x + 8]), (mvx, mvy, s, row, col, c)
        ctx.decode_frames(ref_records(mbs), lv)            # restore the same reference picture for the next vector
        assert all(np.array_equal(ctx.recon(s), ref[s]) for s in range(S))
    ctx.close()


def ref_records(mbs):
    out = mbs.copy()
    out["mtype"], out["mvx"], out["mvy"] = 0, 0, 0
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in PLAIN if n != "qcif140_q10_tss"])
def test_decoder_output_equals_reference_decoder(name):
    """streams of the reference ENCODER, decoded here, against the md5s of the frames the reference DECODER wrote"""
    from p64_b200.encoder import Decoder, Encoder
    g, clip = golden_clip(name)
    enc = Encoder(g["image_type"], 1, **golden_kwargs(g))
    for fr in clip:
        enc.encode(fr[None])
    enc.finish()
    data = enc.data(0)
    enc.close()
    assert hashlib.md5(data).hexdigest() == g["md5"]                 # = the reference encoder's bytes
    dec = Decoder(data)
    frames = dec.frames()
    dec.close()
    assert dec.image_type == g["image_type"] and len(frames) == g["decoded_frames"]
    step = max(1, len(frames) // 6)
    assert [hashlib.md5(f.tobytes()).hexdigest() for f in frames[::step]] == g["recon_md5"]
    assert hashlib.md5(frames[-1].tobytes()).hexdigest() == g["last_recon_md5"]


@pytest.mark.gpu
def test_decoder_handles_what_the_encoder_never_emits():
    """hand-written streams: skipped macroblocks (MBA > 1), a missing GOB, MQUANT types 1/3/6/9, stuffing, a
    temporal-reference gap (picture repeated) -- decoded vs a model built from the oracle's inverse half"""
    from p64_b200.encoder import BitWriter, Decoder, MB_DTYPE
    it = y4m.IT_QCIF
    w, h = y4m.DIMS[it]
    clip = y4m.synth_clip(it, 3, seed=77)
    orc = O.Encoder(it)
    rng = np.random.default_rng(9)
    bw = BitWriter(it)
    expect = []
    prev = np.zeros(w * h * 3 // 2, np.uint8)
    trs = [0, 1, 4]                                                    # gap: picture 1 is written 3 times
    for f, fr in enumerate(clip):
        recs, lv = orc.encode_frame(fr, 10, O.ME_TSS, 15)
        mbs, lv8 = recs_to_mb(recs), levels_to_i8(lv)
        full = orc.recon().copy()
        bw.picture_header(trs[f])
        cur = prev.copy()
        for gob in range(3):
            if f == 2 and gob == 1:
                continue                                               # a GOB that is not there at all
            bw.gob_header(gob, 10)
            last = -1
            for m in range(33):
                i = gob * 33 + m
                if f > 0 and rng.random() < 0.3:
                    continue                                           # skipped macroblock
                rec = mbs[i].copy()
                # (MQUANT variants keep the same quantiser, so the oracle's reconstruction still applies)
                if rng.random() < 0.5 and rec["mtype"] in (0, 2, 5, 8):
                    rec["mtype"] += 1
                if m - last > 1 and rng.random() < 0.5:
                    bw.put(0b00000001111, 11)                          # MBA stuffing
                _bits_mb_with_gap(bw, m, last, rec, lv8[i])
                last = m
                _copy_mb(cur, full, it, gob, m)
            # MV predictor / LastIntra of the oracle are irrelevant here: the stream carries explicit vectors
        expect.append(cur.copy())
        prev = cur
        # keep the oracle's reference in step with what the DECODER will have (skips change the prediction source)
        orc.set_recon(cur)
    bw.picture_header(5)
    bw.finish()
    dec = Decoder(bw.data())
    frames = dec.frames()
    dec.close()
    want = [expect[0], expect[1], expect[1], expect[1], expect[2]]
    assert len(frames) == len(want)
    for k, (a, b) in enumerate(zip(frames, want)):
        assert np.array_equal(a, b), k


def _copy_mb(dst, src, it, gob, m):
    w, h = y4m.DIMS[it]
    if it == y4m.IT_QCIF:
        col, row = m % 11, gob * 3 + m // 11
    else:
        col, row = (gob & 1) * 11 + m % 11, (gob >> 1) * 3 + m // 11
    Y = slice(0, w * h)
    d, s = dst[Y].reshape(h, w), src[Y].reshape(h, w)
    d[row * 16:row * 16 + 16, col * 16:col * 16 + 16] = s[row * 16:row * 16 + 16, col * 16:col * 16 + 16]
    for pl in range(2):
        o = w * h + pl * (w * h // 4)
        d = dst[o:o + w * h // 4].reshape(h // 2, w // 2); s = src[o:o + w * h // 4].reshape(h // 2, w // 2)
        d[row * 8:row * 8 + 8, col * 8:col * 8 + 8] = s[row * 8:row * 8 + 8, col * 8:col * 8 + 8]


def _bits_mb_with_gap(bw, m, last, rec, lv):
    """BitWriter.mb() derives MBA from its own LastMBA, so a gap in m is written as MBA > 1 by itself"""
    bw.mb(m, rec, lv)


@pytest.mark.gpu
def test_cli_decodes_like_the_reference(tmp_path):
    from p64_b200 import build
    from p64_b200.encoder import Encoder
    g, clip = golden_clip("qcif12_q8_tss")
    enc = Encoder(g["image_type"], 1, **golden_kwargs(g))
    for fr in clip:
        enc.encode(fr[None])
    enc.finish()
    open(tmp_path / "s.p64", "wb").write(enc.data(0))
    enc.close()
    subprocess.run([build.build_cli(), "-d", "-y4m", "-s", str(tmp_path / "s.p64"), str(tmp_path / "dec")], check=True, stdout=subprocess.DEVNULL)
    raw = open(tmp_path / "dec.y4m", "rb").read()
    assert raw.startswith(b"YUV4MPEG2 W176 H144 C420jpeg Ip F30000:1001\n")
    _, _, dec = y4m.read_y4m(str(tmp_path / "dec.y4m"))
    assert len(dec) == g["decoded_frames"]
    assert hashlib.md5(dec[-1].tobytes()).hexdigest() == g["last_recon_md5"]
    if O.have_ref():                                                   # byte-compare against the reference decoder's own file
        O.ref_decode(str(tmp_path / "s.p64"), str(tmp_path / "ref"))
        assert open(tmp_path / "ref.y4m", "rb").read() == raw


def _short_p64():
    import os
    p = os.path.join(O.REF_DIR, "short.p64")
    if not (O.have_ref() and os.path.exists(p)):
        pytest.skip("oracle/_ref/short.p64 not present (copied there by oracle/build_ref.sh)")
    return p


def test_parser_reads_the_references_own_1993_stream():
    """SURVEY 4 KAT #1, parser half: short.p64 (7 NTSC pictures written by the ORIGINAL 1993 encoder -- a foreign stream:
    other search, MBA skips) parses into 7 pictures with consecutive temporal references, and writing the parsed records and
    levels back with the product's bit writer reproduces the file byte for byte (a parser that dropped or misread anything
    could not)."""
    from p64_b200.encoder import BitWriter, Parser
    data = open(_short_p64(), "rb").read()
    p = Parser(data)
    assert p.image_type == y4m.IT_NTSC
    bw = BitWriter(y4m.IT_NTSC)
    n, skipped = 0, 0
    while True:
        pic = p.next_picture()
        if pic is None:
            break
        mbs, lv, tr, rep = pic
        assert rep == 1
        bw.picture_header(tr)
        for gob in range(10):
            sent = [m for m in range(33) if mbs["reserved"][gob * 33 + m] & 1]
            if not sent:
                continue
            bw.gob_header(gob, int(mbs["quant"][gob * 33 + sent[0]]))
            for m in sent:
                bw.mb(m, mbs[gob * 33 + m], lv[gob * 33 + m])
            skipped += 33 - len(sent)
        n += 1
    p.close()
    assert n == 7
    bw.picture_header((tr + 1) % 32)
    bw.finish()
    out = bw.data()
    assert len(out) == len(data) and out == data, (len(out), len(data), skipped)


@pytest.mark.gpu
def test_decoder_known_answer_short_p64(tmp_path):
    """SURVEY 4 KAT #1: the reference's own stream decoded here == decoded by the reference decoder, frame by frame"""
    from p64_b200.encoder import Decoder
    path = _short_p64()
    O.ref_decode(path, str(tmp_path / "ref"))
    w, h, want = y4m.read_y4m(str(tmp_path / "ref.y4m"))
    dec = Decoder(open(path, "rb").read())
    got = dec.frames()
    dec.close()
    assert (w, h, len(want)) == (352, 240, 7) and len(got) == 7
    for k in range(7):
        assert np.array_equal(got[k], want[k]), k

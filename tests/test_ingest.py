"""Ingest: the Y4M reader and the chroma conversions the reference applies before the encoder sees a
frame (y4m_input.c:195-545).  CPU: the NumPy oracle against the committed golden vectors of the reference's OWN reader
(tests/golden/ingest.json) and, when present, against that reader live; the product's Y4M parser.  GPU: the device
conversion against the oracle, and whole streams against the reference encoder's goldens."""
import hashlib
import importlib.util
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, golden_clip, golden_kwargs, oracle_encode_stream
from oracle import oracle as O
from p64_b200 import y4m

HERE = os.path.dirname(os.path.abspath(__file__))
INGEST = json.load(open(os.path.join(HERE, "golden", "ingest.json")))
_spec = importlib.util.spec_from_file_location("make_ingest_golden", os.path.join(HERE, "golden", "make_ingest_golden.py"))
_gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_gen)
CHROMA_CASES = [n for n in GOLDEN if GOLDEN[n]["args"].get("chroma")]


def _payloads(g):
    pay = _gen.payloads(g["w"], g["h"], g["chroma"], g["seed"])
    assert hashlib.md5(b"".join(p.tobytes() for p in pay)).hexdigest() == g["payload_md5"], "payload generator drifted"
    return pay


@pytest.mark.parametrize("name", list(INGEST))
def test_oracle_conversion_matches_reference_reader_golden(name):
    g = INGEST[name]
    for p, want in zip(_payloads(g), g["frames_md5"]):
        got = O.y4m_payload_to_encoder_frame(p, g["w"], g["h"], g["chroma"])
        assert hashlib.md5(got.tobytes()).hexdigest() == want


def test_oracle_conversion_matches_live_reference_reader(tmp_path):
    if not O.have_ref():
        pytest.skip("compiled reference not present")
    rng = np.random.default_rng(11)
    for (w, h) in ((176, 144), (352, 240)):
        for chroma in _gen.TYPES:
            n = O.payload_bytes(w, h, chroma)
            pay = [rng.integers(0, 256, n).astype(np.uint8), rng.integers(100, 140, n).astype(np.uint8)]
            O.write_y4m_raw(str(tmp_path / "a.y4m"), w, h, pay, chroma)
            ref = O.ref_y4m_frames(str(tmp_path / "a.y4m"), w, h)
            mine = np.stack([O.y4m_payload_to_encoder_frame(p, w, h, chroma) for p in pay])
            assert np.array_equal(ref, mine), (w, h, chroma)


@pytest.mark.parametrize("name", [n for n in CHROMA_CASES if GOLDEN[n]["image_type"] == y4m.IT_QCIF])
def test_oracle_stream_with_converted_input_matches_reference_golden(name):
    g, clip = golden_clip(name)
    data, recons, ovfl = oracle_encode_stream(g["image_type"], clip, **golden_kwargs(g))
    assert hashlib.md5(data).hexdigest() == g["md5"]
    assert ovfl == g["overflows"]


def test_y4m_reader_parses_like_the_reference(tmp_path):
    from p64_b200._lib import P64Error
    from p64_b200.encoder import CHROMA, Y4mReader
    w, h = 176, 144
    rng = np.random.default_rng(5)
    for chroma in _gen.TYPES:
        n = O.payload_bytes(w, h, chroma)
        pay = [rng.integers(0, 256, n).astype(np.uint8) for _ in range(3)]
        O.write_y4m_raw(str(tmp_path / "a.y4m"), w, h, pay, chroma, frame_params=b" Ixyz")
        r = Y4mReader(str(tmp_path / "a.y4m"))
        assert (r.info.width, r.info.height, r.info.fps_n, r.info.fps_d) == (w, h, 30000, 1001)
        assert r.info.chroma == CHROMA[chroma] and r.info.frame_bytes == n and chr(r.info.interlace) == "p"
        for p in pay:
            assert np.array_equal(r.read_frame(), p)
        assert r.read_frame() is None
        r.close()
    # header rules of y4m_input.c:75-134, 556-586
    def hdr(text, body=b""):
        open(tmp_path / "b.y4m", "wb").write(text + body)
        return Y4mReader(str(tmp_path / "b.y4m"))
    r = hdr(b"YUV4MPEG2 W176 H144 F25:1 Xfoo A128:117\n")           # no C tag -> "420"; no I tag -> '?'; unknown tag ignored
    assert (r.info.chroma, chr(r.info.interlace), r.info.par_n, r.info.par_d) == (0, "?", 128, 117)
    assert r.info.frame_bytes == 38016 and r.read_frame() is None
    for bad in (b"YUV4MPEG2 W176 F25:1\n", b"YUV4MPEG2 W176 H144\n", b"YUV4MPEG2 W176 H144 F25:1 It\n",
                b"YUV4MPEG2 W176 H144 F25:1 C420foo\n", b"RIFFxxxxAVI \n", b"YUV4MPEG2 W176 H144 F25\n"):
        with pytest.raises(P64Error):
            hdr(bad)
    r = hdr(b"YUV4MPEG2 W176 H144 F25:1 Cmono\n", b"FRAME\n" + bytes(176 * 144) + b"FRAMX\n")
    assert r.read_frame() is not None
    with pytest.raises(P64Error):
        r.read_frame()                                               # "Loss of framing"
    r = hdr(b"YUV4MPEG2 W176 H144 F25:1 Cmono\n", b"FRAME\n" + bytes(100))
    with pytest.raises(P64Error):
        r.read_frame()                                               # truncated frame
    with pytest.raises(P64Error):
        Y4mReader(str(tmp_path / "missing.y4m"))
    from p64_b200 import _lib
    L = _lib.lib()
    for it in (0, 1, 2):
        w, h = y4m.DIMS[it]
        for chroma in _gen.TYPES:
            assert L.p64b_raw_frame_bytes(it, CHROMA[chroma]) == O.payload_bytes(w, h, chroma)
    assert L.p64b_raw_frame_bytes(1, 99) < 0


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("it", [y4m.IT_QCIF, y4m.IT_CIF, y4m.IT_NTSC])
def test_device_conversion_matches_oracle(it):
    from p64_b200.encoder import DeviceContext
    w, h = y4m.DIMS[it]
    for k, chroma in enumerate(_gen.TYPES):
        S = 3
        pays = []
        for s in range(S):
            pays += _gen.payloads(w, h, chroma, 700 + 10 * k + s)
        want = [O.y4m_payload_to_encoder_frame(p, w, h, chroma) for p in pays]
        ctx = DeviceContext(it, S)
        try:
            ctx.set_input_chroma(chroma)
            for f in range(3):
                got = ctx.convert_frames(np.stack(pays[f::3]))
                for s in range(S):
                    assert np.array_equal(got[s], want[3 * s + f]), (chroma, f, s)
        finally:
            ctx.close()
    key = f"{w}x{h}_420paldv"
    g = INGEST[key]
    ctx = DeviceContext(it, 1)
    try:
        ctx.set_input_chroma("420paldv")
        for p, md5 in zip(_payloads(g), g["frames_md5"]):                 # straight against the reference reader's golden
            assert hashlib.md5(ctx.convert_frames(p[None])[0].tobytes()).hexdigest() == md5
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("host_vlc", [False, True])
@pytest.mark.parametrize("name", CHROMA_CASES)
def test_stream_from_unconverted_y4m_payloads_matches_reference_golden(name, host_vlc):
    """the reference encoder reading a 420mpeg2 / 420paldv / 422 / 411 / 444 / 444alpha / mono Y4M file vs this encoder fed
    the file's unconverted payloads (conversion on the device), incl. a rate-controlled case"""
    from p64_b200.encoder import Encoder
    g, clip = golden_clip(name)
    enc = Encoder(g["image_type"], 1, host_vlc=host_vlc, **golden_kwargs(g))
    try:
        for fr in clip:
            enc.encode(fr[None])
        enc.finish()
        data = enc.data(0)
        assert len(data) == g["size"] and hashlib.md5(data).hexdigest() == g["md5"]
        assert enc.overflows(0) == g["overflows"]
    finally:
        enc.close()


@pytest.mark.gpu
def test_cli_reads_converted_chroma_types(tmp_path):
    import subprocess
    from p64_b200 import build
    cli = build.build_cli()
    for name in ("qcif6_q6_full31_420paldv", "qcif5_q8_tss_411"):
        g, clip = golden_clip(name)
        a = g["args"]
        y4m.write_y4m(str(tmp_path / "c.y4m"), g["image_type"], clip, chroma=a["chroma"])
        cmd = [cli, "-y4m", "-QCIF", "-a", "0", "-b", str(g["n_frames"] - 1), "-q", str(a["q"])]
        if a.get("full_search"):
            cmd += ["--me", "full", "-i", str(a["search_limit"])]
        cmd += [str(tmp_path / "c"), "-s", str(tmp_path / "o.p64")]
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
        assert hashlib.md5(open(tmp_path / "o.p64", "rb").read()).hexdigest() == g["md5"]

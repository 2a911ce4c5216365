"""Whole-stream parity on the CPU: oracle hot path + the product's host VLC writer must reproduce the
reference encoder's .p64 byte for byte (golden md5s from tests/golden/make_golden.py), and -- when the compiled
reference is present -- match a live run of it."""
import hashlib
import os

import numpy as np
import pytest

from helpers import GOLDEN, golden_clip, golden_kwargs, oracle_encode_stream
from p64_b200 import y4m

FAST = ["cif8_q8_full15", "cif6_q3_tss", "cif6_q31_full31", "qcif12_q8_intra", "qcif12_q8_tss", "ntsc7_q8_tss",
        "ntsc7_q8_full31", "cif12_r128000_tss", "cif12_r64000_full31", "qcif20_r64000_tss",
        "qcif6_q8_tss_a3_k2_b14", "qcif5_r64000_a30_k3_b42", "cif4_r2000000_full31_a50_k3_f15"]


@pytest.mark.parametrize("name", FAST)
def test_oracle_plus_host_writer_matches_golden(name):
    g, clip = golden_clip(name)
    data, recons, ovfl = oracle_encode_stream(g["image_type"], clip, **golden_kwargs(g))
    assert len(data) == g["size"]
    assert hashlib.md5(data).hexdigest() == g["md5"]
    assert ovfl == g["overflows"]
    # closed loop: the encoder's reconstruction == the reference DECODER's output of the reference stream
    assert hashlib.md5(recons[-1].tobytes()).hexdigest() == g["last_recon_md5"]


def test_forced_intra_refresh_long_clip():
    g, clip = golden_clip("qcif140_q10_tss")
    data, recons, _ = oracle_encode_stream(g["image_type"], clip, **golden_kwargs(g))
    assert hashlib.md5(data).hexdigest() == g["md5"]


def test_matches_live_reference_run(orc, tmp_path):
    if not orc.have_ref():
        pytest.skip("compiled reference not present")
    it = y4m.IT_QCIF
    clip = y4m.synth_clip(it, 5, seed=2024, pan=(-2, 1))
    y4m.write_y4m(str(tmp_path / "c.y4m"), it, clip)
    for kw, okw in [(dict(q=6), dict(q=6)),
                    (dict(q=12, full_search=True, search_limit=31), dict(q=12, me_mode=1, search_limit=31)),
                    (dict(rate=96000), dict(rate=96000))]:
        orc.ref_encode(str(tmp_path / "c.y4m"), str(tmp_path / "o.p64"), it, 5, **kw)
        ref = open(tmp_path / "o.p64", "rb").read()
        mine, _, _ = oracle_encode_stream(it, clip, **okw)
        assert mine == ref


def test_short_p64_intra_kat(orc, tmp_path):
    """SURVEY 4 KAT #2: re-encoding the frames decoded from the reference's own short.p64 reproduces the first
    17 323 bytes of short.p64 and the frame-0 bit count of short.trace:3 (138 588 bits). Needs the reference tree."""
    ref_dir = os.environ.get("P64_REFERENCE", "/root/reference")
    if not (orc.have_ref() and os.path.exists(os.path.join(ref_dir, "short.p64"))):
        pytest.skip("reference tree not present")
    orc.ref_decode(os.path.join(ref_dir, "short.p64"), str(tmp_path / "short"))
    w, h, frames = y4m.read_y4m(str(tmp_path / "short.y4m"))
    assert (w, h, len(frames)) == (352, 240, 7)
    data, _, _ = oracle_encode_stream(y4m.IT_NTSC, frames[:1], q=8)
    golden = open(os.path.join(ref_dir, "short.p64"), "rb").read()
    assert data[:17323] == golden[:17323]
    # frame 0 = picture header .. last MB; the trailing picture header adds 32+9 bits (NTSC PSPARE) + padding
    assert len(data) * 8 >= 138588
    # ... and the statistics block of short.trace:3-8 for that frame, through the product's bit writer counters
    import ctypes as C
    from helpers import levels_to_i8, recs_to_mb
    from p64_b200._lib import FrameCounters, lib
    from p64_b200.encoder import BitWriter
    enc, bw = orc.Encoder(y4m.IT_NTSC), BitWriter(y4m.IT_NTSC)
    bw.picture_header(0)
    recs, lv = enc.encode_frame(frames[0], 8)
    mbs, lv8 = recs_to_mb(recs), levels_to_i8(lv)
    for gob in range(10):
        bw.gob_header(gob, 8)
        for m in range(33):
            bw.mb(m, mbs[gob * 33 + m], lv8[gob * 33 + m])
    c = FrameCounters()
    lib().p64b_bits_counters(bw.h, C.byref(c))
    assert bw.tell() == 138588
    assert (c.mb_attribute_bits, c.mv_bits, c.eob_bits) == (1650, 0, 3960)
    assert (c.y_bits, c.u_bits, c.v_bits) == (123214, 3805, 5658)
    assert c.macro_type_freq[0] == 330 and sum(c.macro_type_freq[:]) == 330

"""Frame statistics (stat.c:52-130 and PrintFrameStatistics p64.c:1299-1332): the device's exact integer
sums against NumPy, and the whole per-frame block the `p64b -l 1` command prints against the block the UNMODIFIED
reference encoder printed for the same clip (tests/golden/stats.json, made by tests/golden/make_stats_golden.py)."""
import json
import os
import subprocess

import numpy as np
import pytest

from helpers import golden_kwargs
from oracle import oracle as O
from p64_b200 import y4m

HERE = os.path.dirname(os.path.abspath(__file__))
STATS = json.load(open(os.path.join(HERE, "golden", "stats.json")))


def test_host_bit_writer_counts_like_the_reference():
    """CPU: oracle hot path + the product's host bit writer reproduce the reference's per-frame bit counters"""
    import ctypes as C
    from helpers import levels_to_i8, recs_to_mb
    from p64_b200._lib import FrameCounters, lib
    from p64_b200.encoder import BitWriter
    g = STATS["qcif4_q8_tss"]
    clip = y4m.synth_clip(g["image_type"], g["n_frames"], g["seed"])
    enc, bw = O.Encoder(g["image_type"]), BitWriter(g["image_type"])
    for f, fr in enumerate(clip):
        lib().p64b_bits_counters_reset(bw.h)
        bw.picture_header(f % 32)
        recs, lv = enc.encode_frame(fr, 8, O.ME_TSS, 15)
        mbs, lv8 = recs_to_mb(recs), levels_to_i8(lv)
        for gob in range(enc.ngob):
            bw.gob_header(gob, 8)
            for m in range(33):
                bw.mb(m, mbs[gob * 33 + m], lv8[gob * 33 + m])
        c = FrameCounters()
        lib().p64b_bits_counters(bw.h, C.byref(c))
        block = g["frames"][f]
        assert "MB Attribute Bits: %6d  MV Bits: %6d   EOB Bits: %6d" % (c.mb_attribute_bits, c.mv_bits, c.eob_bits) == block[2]
        assert "Y Bits: %7d  U Bits: %7d  V Bits: %7d  Total Bits: %7d" % (c.y_bits, c.u_bits, c.v_bits, c.y_bits + c.u_bits + c.v_bits) == block[3]
        assert "Macro Freq: " + "".join("%5d" % x for x in c.macro_type_freq) == block[6]
        assert "Y     Freq: " + "".join("%5d" % x for x in c.y_type_freq) == block[7]
        assert "UV    Freq: " + "".join("%5d" % x for x in c.uv_type_freq) == block[8]
        n6 = enc.ngob * 33 * 6
        assert "MV StepSize: %f  MV NumberNonZero: %f  MV NumberZero: %f" % (c.q_sum / c.q_use, c.number_nz / n6, (n6 * 64 - c.number_nz) / n6) == block[4]


def test_stat_from_sums_matches_stat_c_formulas():
    """CPU: the derived quantities for hand-made sums, incl. the 99.99 / -99.99 branches (stat.c:104-119)"""
    import ctypes as C
    from p64_b200._lib import PlaneStats, Stat, lib
    rng = np.random.default_rng(3)
    for case in range(4):
        src = rng.integers(0, 256, 25344).astype(np.int64)
        rec = src.copy() if case == 1 else np.clip(src + rng.integers(-3, 4, src.size), 0, 255)
        if case == 2:
            src[:] = 0
        if case == 3:
            src[:] = 7; rec = src + 1
        ps = PlaneStats()
        ps.n, ps.sum_src, ps.sum_rec = src.size, int(src.sum()), int(rec.sum())
        ps.sum_sq_err, ps.sum_sq_src = int(((rec - src) ** 2).sum()), int((src * src).sum())
        for v, k in zip(*np.unique(rec, return_counts=True)):
            ps.hist[int(v)] = int(k)
        st = Stat()
        lib().p64b_stat_from_sums(C.byref(ps), C.byref(st))
        top, sq, rsq, rv = float(src.size), float(ps.sum_sq_err), float(ps.sum_sq_src), float(ps.sum_src)
        assert st.mean == rec.sum() / top and st.mse == sq / top
        if sq:
            assert st.snr == (10 * np.log10(rsq / sq) if rsq else -99.99)
            mr = rsq - rv * rv / top
            assert st.mrsnr == (10 * np.log10(mr / sq) if mr else -99.99)
            assert abs(st.psnr - 10 * np.log10(65025.0 * top / sq)) < 1e-12
        else:
            assert (st.snr, st.mrsnr, st.psnr) == (99.99, 99.99, 99.99)
        p = np.bincount(rec, minlength=256) / top
        assert abs(st.entropy + (p[p > 0] * np.log(p[p > 0])).sum() / np.log(2.0)) < 1e-12


def _ref_stat(src, rec, w, h):
    """the reference's own StatisticsMem (stat.c:73-130) through oracle/_ref/libp64ref.so (ref_stat_shim.c)"""
    import ctypes as C
    L = C.CDLL(os.path.join(O.REF_DIR, "libp64ref.so"))
    a, b = np.ascontiguousarray(src, np.uint8), np.ascontiguousarray(rec, np.uint8)
    out = (C.c_double * 6)()
    L.ref_statistics_mem(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_int(w), C.c_int(h), out)
    return dict(zip(("mean", "mse", "mrsnr", "snr", "psnr", "entropy"), out))


def _have_ref_stat():
    import ctypes as C
    try:
        return O.have_ref() and hasattr(C.CDLL(os.path.join(O.REF_DIR, "libp64ref.so")), "ref_statistics_mem")
    except OSError:
        return False


@pytest.mark.skipif(not _have_ref_stat(), reason="oracle/_ref/libp64ref.so without the stat.c shim")
def test_stat_from_sums_matches_the_references_stat_c():
    """CPU: p64b_stat_from_sums fed with exact integer sums == the reference's StatisticsMem on the same planes, bit for bit
    (same operations in the same order on exactly representable integers), incl. identical planes, a black source and a
    constant offset"""
    import ctypes as C
    from p64_b200._lib import PlaneStats, Stat, lib
    rng = np.random.default_rng(11)
    w, h = 176, 144
    for case in range(6):
        src = rng.integers(0, 256, w * h).astype(np.int64)
        rec = np.clip(src + rng.integers(-case - 1, case + 2, src.size), 0, 255)
        if case == 1:
            rec = src.copy()
        if case == 2:
            src[:] = 0
        if case == 3:
            src[:] = 7; rec = src + 1
        ps = PlaneStats()
        ps.n, ps.sum_src, ps.sum_rec = src.size, int(src.sum()), int(rec.sum())
        ps.sum_sq_err, ps.sum_sq_src = int(((rec - src) ** 2).sum()), int((src * src).sum())
        for v, k in zip(*np.unique(rec, return_counts=True)):
            ps.hist[int(v)] = int(k)
        st = Stat()
        lib().p64b_stat_from_sums(C.byref(ps), C.byref(st))
        want = _ref_stat(src, rec, w, h)
        assert {k: getattr(st, k) for k in want} == want, case


@pytest.mark.gpu
@pytest.mark.skipif(not _have_ref_stat(), reason="oracle/_ref/libp64ref.so without the stat.c shim")
@pytest.mark.parametrize("it", [y4m.IT_QCIF, y4m.IT_CIF])
def test_device_statistics_match_the_references_stat_c(it):
    """GPU: plane_stats_kernel + p64b_stat_from_sums against the reference's own StatisticsMem run on the same source and
    reconstruction planes (stat.c:73-130), every plane of every stream"""
    import ctypes as C
    from p64_b200._lib import Stat, lib
    from p64_b200.encoder import DeviceContext, make_step
    S = 2
    w, h = y4m.DIMS[it]
    clips = [y4m.synth_clip(it, 3, seed=70 + s, pan=(1 - s, s)) for s in range(S)]
    ctx = DeviceContext(it, S)
    try:
        for f in range(3):
            src = np.stack([c[f] for c in clips])
            ctx.encode_frames(make_step(f == 0, 6 + 11 * f, 0, 15), src)
            st = ctx.statistics()
            for s in range(S):
                rec = ctx.recon(s)
                off = 0
                for pl, (pw, ph) in enumerate(((w, h), (w // 2, h // 2), (w // 2, h // 2))):
                    n = pw * ph
                    got = Stat()
                    lib().p64b_stat_from_sums(C.byref(st[s][pl]), C.byref(got))
                    want = _ref_stat(src[s, off:off + n], rec[off:off + n], pw, ph)
                    assert {k: getattr(got, k) for k in want} == want, (f, s, pl)
                    off += n
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("it", [y4m.IT_QCIF, y4m.IT_CIF, y4m.IT_NTSC])
def test_device_sums_match_numpy(it):
    from p64_b200.encoder import DeviceContext, make_step
    S = 3
    w, h = y4m.DIMS[it]
    clips = [y4m.synth_clip(it, 3, seed=40 + s, pan=(s, 1 - s)) for s in range(S)]
    ctx = DeviceContext(it, S)
    try:
        for f in range(3):
            src = np.stack([c[f] for c in clips])
            ctx.encode_frames(make_step(f == 0, 4 + 9 * f, 1, 31), src)
            st = ctx.statistics()
            for s in range(S):
                rec = ctx.recon(s).astype(np.int64)
                off = 0
                for pl, n in enumerate((w * h, w * h // 4, w * h // 4)):
                    a, b = src[s, off:off + n].astype(np.int64), rec[off:off + n]
                    p = st[s][pl]
                    assert (p.n, p.sum_src, p.sum_rec) == (n, a.sum(), b.sum()), (f, s, pl)
                    assert (p.sum_sq_err, p.sum_sq_src) == (((b - a) ** 2).sum(), (a * a).sum()), (f, s, pl)
                    assert np.array_equal(np.array(p.hist[:]), np.bincount(b, minlength=256)), (f, s, pl)
                    off += n
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(STATS))
def test_cli_loud_block_matches_reference_log(name, tmp_path):
    from p64_b200 import build
    g = STATS[name]
    a = g["args"]
    clip = y4m.synth_clip(g["image_type"], g["n_frames"], g["seed"])
    y4m.write_y4m(str(tmp_path / "c.y4m"), g["image_type"], clip)
    cmd = [build.build_cli(), "-y4m", O.FLAG[g["image_type"]], "-a", "0", "-b", str(g["n_frames"] - 1), "-l", "1"]
    if a.get("q"):
        cmd += ["-q", str(a["q"])]
    if a.get("rate"):
        cmd += ["-r", str(a["rate"])]
    if a.get("full_search"):
        cmd += ["--me", "full", "-i", str(a["search_limit"])]
    if a.get("intra_only"):
        cmd += ["--intra-only"]
    out = subprocess.run(cmd + [str(tmp_path / "c"), "-s", str(tmp_path / "o.p64")], check=True, stdout=subprocess.PIPE).stdout.decode()
    blocks, cur = [], None
    for line in out.splitlines():
        if line.startswith("START>Frame"):
            cur = []
        if cur is not None:
            cur.append(line.rstrip())
        if line.startswith("END>Frame") and cur is not None:
            blocks.append(cur); cur = None
    assert blocks == g["frames"]

"""Pins the C oracle (oracle/p64_oracle.c) against the reference's OWN compiled objects
(oracle/_ref/libp64ref.so = me.c, chendct.c, transform.c, io.c built from /root/reference)."""
import numpy as np
import pytest

from p64_b200 import y4m


def _rand_blocks(rng, n, lo, hi):
    b = rng.integers(lo, hi + 1, size=(n, 64)).astype(np.int32)
    b[0] = lo; b[1] = hi; b[2, ::2] = lo; b[2, 1::2] = hi; b[3] = 0
    return b


def test_fdct_matches_reference(orc, reflib):
    rng = np.random.default_rng(1)
    for blk in _rand_blocks(rng, 1500, -255, 255):
        assert np.array_equal(orc.fdct(blk), reflib.chen_dct(blk))
    for blk in _rand_blocks(rng, 300, 0, 255):
        assert np.array_equal(orc.fdct(blk), reflib.chen_dct(blk))


def test_idct_matches_reference(orc, reflib):
    rng = np.random.default_rng(2)
    for blk in _rand_blocks(rng, 1000, -7967, 7967):
        assert np.array_equal(orc.idct(blk), reflib.chen_idct(blk))
    for k in range(500):   # sparse, like real dequantised blocks
        blk = np.zeros(64, np.int32)
        idx = rng.integers(0, 64, size=rng.integers(1, 8))
        blk[idx] = rng.integers(-2040, 2041, size=idx.size)
        assert np.array_equal(orc.idct(blk), reflib.chen_idct(blk))


@pytest.mark.parametrize("q", list(range(1, 32)))
def test_quantisers_match_reference(orc, reflib, q):
    rng = np.random.default_rng(100 + q)
    for blk in _rand_blocks(rng, 40, -1023, 1023):
        blk = blk.copy(); blk[0] = rng.integers(-3000, 2048)
        blk = orc.bound_dct(blk)
        assert np.array_equal(blk, reflib.bound_dct(blk))
        qi, qp = orc.quant_intra(blk, q), orc.quant_inter(blk, q)
        assert np.array_equal(qi, reflib.quant_intra(blk, q))
        assert np.array_equal(qp, reflib.quant_inter(blk, q))
        assert np.array_equal(orc.iquant_intra(qi, q), reflib.iquant_intra(qi, q))
        assert np.array_equal(orc.iquant_inter(qp, q), reflib.iquant_inter(qp, q))


def test_zigzag_matches_reference(orc, reflib):
    a = np.arange(64, dtype=np.int32)
    assert np.array_equal(orc.zigzag(a), reflib.zigzag(a))
    assert np.array_equal(orc.izigzag(a), reflib.izigzag(a))
    assert list(orc.zigzag(a)[:10]) == [0, 1, 8, 16, 9, 2, 3, 10, 17, 24]
    assert np.array_equal(orc.izigzag(orc.zigzag(a)), a)


def test_loop_filter_matches_reference(orc, reflib):
    rng = np.random.default_rng(3)
    plane = rng.integers(0, 256, size=(72, 88)).astype(np.uint8)
    for _ in range(200):
        x, y = int(rng.integers(0, 80)), int(rng.integers(0, 64))
        assert np.array_equal(orc.loop_filter(plane, x, y), reflib.loop_filter(plane, x, y))
    # rounding identity used by the device kernel: two-stage == (S16+8)>>4 interior, (S4+2)>>2 edges
    p = plane.astype(np.int64)
    for (x, y) in [(0, 0), (13, 7), (80, 64)]:
        b = p[y:y + 8, x:x + 8]
        h = b * 4
        h[:, 1:7] = b[:, :6] + 2 * b[:, 1:7] + b[:, 2:]
        v = h * 4
        v[1:7] = h[:6] + 2 * h[1:7] + h[2:]
        assert np.array_equal(((v + 8) >> 4).ravel(), orc.loop_filter(plane, x, y))


def test_compensation_addressing_matches_reference(orc, reflib):
    """io.c:200-496: luma vector as is, chroma vector /2 truncating toward zero, filter per 8x8 block."""
    rng = np.random.default_rng(4)
    plane = rng.integers(0, 256, size=(72, 88)).astype(np.uint8)
    blk = rng.integers(0, 256, size=64).astype(np.int32)
    for mvx, mvy in [(-3, 5), (7, -7), (-15, -1), (1, 1), (0, 0), (-1, -1)]:
        for half in (False, True):
            for filt in (False, True):
                got = reflib.sub_compensate(plane, 4, 3, mvx, mvy, blk, half=half, filt=filt)
                dx, dy = (int(mvx / 2), int(mvy / 2)) if half else (mvx, mvy)
                x, y = 4 * 8 + dx, 3 * 8 + dy
                pred = orc.loop_filter(plane, x, y) if filt else plane[y:y + 8, x:x + 8].astype(np.int32).ravel()
                assert np.array_equal(got, blk - pred)


@pytest.mark.parametrize("kind", ["shift", "flat", "ramp", "noise", "levels4", "edge"])
def test_me_tss_matches_reference(orc, reflib, kind):
    rng = np.random.default_rng(5)
    it = y4m.IT_QCIF
    w, h = y4m.DIMS[it]
    if kind == "shift":
        ref, cur = y4m.random_pair(it, 11, shift=(5, -3), noise=3)
    elif kind == "flat":
        ref = np.full((h, w), 77, np.uint8); cur = np.full((h, w), 77, np.uint8)
    elif kind == "ramp":
        ref = (np.add.outer(np.arange(h), np.arange(w)) % 256).astype(np.uint8); cur = np.roll(ref, 2, axis=1)
    elif kind == "noise":
        ref = rng.integers(0, 256, (h, w)).astype(np.uint8); cur = rng.integers(0, 256, (h, w)).astype(np.uint8)
    elif kind == "levels4":
        ref = (rng.integers(0, 4, (h // 16, w // 16)).repeat(16, 0).repeat(16, 1) * 60).astype(np.uint8)
        cur = np.roll(ref, (3, -6), axis=(0, 1))
    else:
        ref, cur = y4m.random_pair(it, 12, shift=(-15, 15), noise=0)
    assert np.array_equal(orc.me_frame(ref, cur, orc.ME_TSS), reflib.motion_estimation(ref, cur))


@pytest.mark.parametrize("limit", [15, 31, 8])
def test_me_full_matches_reference(orc, reflib, limit):
    it = y4m.IT_QCIF
    for seed, shift in [(21, (6, -2)), (22, (-14, 13)), (23, (0, 0))]:
        ref, cur = y4m.random_pair(it, seed, shift=shift, noise=2)
        got = orc.me_frame(ref, cur, orc.ME_FULL, limit)
        want = reflib.motion_estimation(ref, cur, full=True, search_limit=limit)
        assert np.array_equal(got[:, :4], want[:, :4])   # MX,MY,MV,OMV (stats are file-static in me.c:44-46)
    flat = np.full((144, 176), 9, np.uint8)
    assert np.array_equal(orc.me_frame(flat, flat, orc.ME_FULL, limit)[:, :4],
                          reflib.motion_estimation(flat, flat, full=True, search_limit=limit)[:, :4])


def test_sad_surface_consistent_with_searches(orc):
    ref, cur = y4m.random_pair(y4m.IT_QCIF, 31, shift=(4, 4), noise=2)
    full = orc.me_frame(ref, cur, orc.ME_FULL, 31)
    for mb in (0, 5, 10, 50, 98):
        s = orc.sad_surface(ref, cur, mb % 11, mb // 11).astype(np.int64)
        assert s[15, 15] == full[mb, 3]
        legal = s[:30, :30].copy(); legal[legal < 0] = 1 << 30     # FastBME -i 31 covers [-15,14]
        assert min(legal.min(), s[15, 15]) == full[mb, 2]

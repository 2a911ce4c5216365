"""GPU parity proper: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs
(bit-exact: this is integer / byte / index work) and against the committed golden streams of the reference."""
import hashlib

import numpy as np
import pytest

from helpers import GOLDEN, golden_clip, golden_kwargs, levels_to_i8, oracle_encode_stream, recs_to_mb
from oracle import oracle as O
from p64_b200 import y4m
from p64_b200.encoder import DeviceContext, Encoder, encode_clip, make_step

pytestmark = pytest.mark.gpu

ME_FIELDS = ("mx", "my", "val", "oval", "var", "varor", "mwor")


def _me_array(rec):
    return np.stack([rec[f] for f in ME_FIELDS], axis=1)


def _run_me(it, ref, cur, mode, limit):
    """device ME of `cur` against reference luma `ref`: load ref as the reconstruction via an intra frame of a
    flat-chroma picture coded at q=1?  No -- simpler and exact: use the dev entry point through torch memory."""
    import torch
    ctx = DeviceContext(it, 1)
    try:
        w, h = y4m.DIMS[it]
        r = torch.from_numpy(np.ascontiguousarray(ref)).cuda()
        c = torch.from_numpy(np.ascontiguousarray(cur)).cuda()
        nmb = ctx.geom["num_mb"]
        out = torch.zeros(nmb * 8, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        ctx.set_cuda_stream(torch.cuda.current_stream().cuda_stream)
        ctx.motion_estimation_dev(r.data_ptr(), c.data_ptr(), 1, mode, limit, out.data_ptr())
        torch.cuda.synchronize()
        return out.cpu().numpy().reshape(nmb, 8)[:, :7]
    finally:
        ctx.close()


@pytest.mark.parametrize("it", [y4m.IT_QCIF, y4m.IT_CIF, y4m.IT_NTSC])
@pytest.mark.parametrize("mode,limit", [(0, 15), (1, 31), (1, 15), (1, 8), (1, 1)])
def test_me_matches_oracle(it, mode, limit):
    w, h = y4m.DIMS[it]
    rng = np.random.default_rng(7)
    cases = [y4m.random_pair(it, 11, shift=(5, -3), noise=3), y4m.random_pair(it, 12, shift=(-15, 15), noise=0),
             y4m.random_pair(it, 13, shift=(14, -14), noise=6),
             (np.full((h, w), 77, np.uint8), np.full((h, w), 77, np.uint8)),                      # all ties
             ((np.add.outer(np.arange(h), np.arange(w)) % 256).astype(np.uint8),) * 2,           # ramp: many ties
             (rng.integers(0, 256, (h, w)).astype(np.uint8), rng.integers(0, 256, (h, w)).astype(np.uint8)),
             (np.zeros((h, w), np.uint8), np.full((h, w), 255, np.uint8))]                        # max SAD 65280
    lv4 = (rng.integers(0, 4, (h // 16, w // 16)).repeat(16, 0).repeat(16, 1) * 60).astype(np.uint8)
    cases.append((lv4, np.roll(lv4, (3, -6), axis=(0, 1))))
    for ref, cur in cases:
        got = _run_me(it, ref, cur, mode, limit)
        want = O.me_frame(ref, cur, mode, limit)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("it", [y4m.IT_QCIF, y4m.IT_CIF])
def test_sad_surface_matches_oracle(it):
    """every legal position of every macroblock's 31x31 SAD surface (me.c:92-176, legality me.c:292-293)"""
    import ctypes as C
    import torch
    from p64_b200 import _lib
    w, h = y4m.DIMS[it]
    rng = np.random.default_rng(3)
    ref = rng.integers(0, 256, (h, w)).astype(np.uint8)
    cur = np.roll(ref, (2, -7), axis=(0, 1)) ^ (rng.integers(0, 8, (h, w)).astype(np.uint8))
    ctx = DeviceContext(it, 1)
    try:
        nmb = (w // 16) * (h // 16)
        r = torch.from_numpy(ref).cuda(); c = torch.from_numpy(cur).cuda()
        out = torch.zeros(nmb * 8, dtype=torch.int32, device="cuda")
        surf = torch.zeros(nmb * 961, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        _lib.check(ctx.L.p64b_ctx_sad_surface_dev(ctx.h, C.c_void_p(r.data_ptr()), C.c_void_p(c.data_ptr()), 1,
                                                  C.c_void_p(out.data_ptr()), C.c_void_p(surf.data_ptr())))
        torch.cuda.synchronize()
        sf = surf.cpu().numpy().reshape(nmb, 31, 31)
        for mb in range(nmb):
            assert np.array_equal(sf[mb], O.sad_surface(ref, cur, mb % (w // 16), mb // (w // 16))), mb
        assert np.array_equal(out.cpu().numpy().reshape(nmb, 8)[:, :7], O.me_frame(ref, cur, 0, 15))
    finally:
        ctx.close()


@pytest.mark.parametrize("it,nf,seed", [(y4m.IT_QCIF, 6, 4321), (y4m.IT_CIF, 5, 1234), (y4m.IT_NTSC, 4, 77)])
@pytest.mark.parametrize("mode,limit,q,intra", [(0, 15, 8, False), (1, 31, 8, False), (1, 15, 3, False),
                                                (0, 15, 31, False), (0, 15, 1, False), (0, 15, 5, True),
                                                (1, 31, 16, False)])
def test_frame_records_levels_recon_match_oracle(it, nf, seed, mode, limit, q, intra):
    clip = y4m.synth_clip(it, nf, seed)
    ctx = DeviceContext(it, 1)
    orc = O.Encoder(it)
    try:
        for f, fr in enumerate(clip):
            mbs, lv = ctx.encode_frames(make_step(f == 0, q, mode, limit, intra), fr[None])
            recs, olv = orc.encode_frame(fr, q, mode, limit, intra)
            if f:
                assert np.array_equal(_me_array(ctx.me_records(0)), orc.me_records()), f"ME frame {f}"
            want = recs_to_mb(recs)
            for k in ("mtype", "cbp", "mvx", "mvy", "quant"):
                assert np.array_equal(mbs[0][k], want[k]), f"{k} frame {f}"
            assert np.array_equal(lv[0], levels_to_i8(olv)), f"levels frame {f}"
            assert np.array_equal(ctx.recon(0), orc.recon()), f"recon frame {f}"
            assert np.array_equal(ctx.last_intra(0), orc.last_intra()), f"LastIntra frame {f}"
            nz = (np.abs(olv).sum(axis=2) != 0)
            want_nz = (nz * (1 << (5 - np.arange(6)))).sum(axis=1)
            assert np.array_equal(mbs[0]["nzmask"], want_nz)
    finally:
        ctx.close()


def test_adversarial_content_matches_oracle():
    """saturated / flat / checkerboard frames: clamps (transform.c:460-537), DC-only blocks, zero residuals."""
    it = y4m.IT_QCIF
    w, h = y4m.DIMS[it]
    n = w * h * 3 // 2
    rng = np.random.default_rng(5)
    cb = (np.indices((h, w)).sum(0) % 2 * 255).astype(np.uint8)
    frames = [np.full(n, 0, np.uint8), np.full(n, 255, np.uint8), np.full(n, 128, np.uint8),
              np.concatenate([cb.ravel(), np.full(n - w * h, 128, np.uint8)]),
              np.concatenate([255 - cb.ravel(), np.full(n - w * h, 0, np.uint8)]),
              rng.integers(0, 256, n).astype(np.uint8), rng.integers(0, 256, n).astype(np.uint8),
              np.full(n, 128, np.uint8), np.full(n, 129, np.uint8), np.full(n, 129, np.uint8)]
    for q, mode in [(1, 0), (8, 1), (31, 0), (2, 1)]:
        ctx = DeviceContext(it, 1); orc = O.Encoder(it)
        try:
            for f, fr in enumerate(frames):
                mbs, lv = ctx.encode_frames(make_step(f == 0, q, mode, 31), fr[None])
                recs, olv = orc.encode_frame(fr, q, mode, 31)
                want = recs_to_mb(recs)
                for k in ("mtype", "cbp", "mvx", "mvy"):
                    assert np.array_equal(mbs[0][k], want[k]), (k, f, q)
                assert np.array_equal(lv[0], levels_to_i8(olv)), (f, q)
                assert np.array_equal(ctx.recon(0), orc.recon()), (f, q)
        finally:
            ctx.close()


@pytest.mark.parametrize("host_vlc", [False, True])
@pytest.mark.parametrize("name", [n for n in GOLDEN if n != "qcif140_q10_tss"])
def test_stream_bytes_match_reference_golden(name, host_vlc):
    """every case runs twice: headers + VLC (and, for the -r cases, the rate control: per-GOB GQUANT, overflow overrides)
    on the device -- the default -- and on the host through the per-GOB calls"""
    g, clip = golden_clip(name)
    enc = Encoder(g["image_type"], 1, host_vlc=host_vlc, **golden_kwargs(g))
    try:
        for fr in clip:
            enc.encode(fr[None])
        enc.finish()
        data = enc.data(0)
        assert len(data) == g["size"]
        assert hashlib.md5(data).hexdigest() == g["md5"]
        assert enc.overflows(0) == g["overflows"]
        import ctypes as C
        from p64_b200 import _lib
        rec = np.zeros(enc.geom["frame_bytes"], np.uint8)
        _lib.check(_lib.lib().p64b_ctx_download_recon(enc.context_handle(), 0, C.c_void_p(rec.ctypes.data)))
        assert hashlib.md5(rec.tobytes()).hexdigest() == g["last_recon_md5"]      # closed loop vs the reference DECODER
    finally:
        enc.close()


def test_long_clip_forced_intra_refresh():
    g, clip = golden_clip("qcif140_q10_tss")
    data = encode_clip(g["image_type"], clip, **golden_kwargs(g))
    assert hashlib.md5(data).hexdigest() == g["md5"]


def test_live_reference_binary_if_present(tmp_path):
    """oracle/_ref travels to the GPU box: byte-compare against a live run of the unmodified reference."""
    if not O.have_ref():
        pytest.skip("compiled reference not present")
    it = y4m.IT_CIF
    clip = y4m.synth_clip(it, 6, seed=31337, pan=(-4, 3))
    y4m.write_y4m(str(tmp_path / "c.y4m"), it, clip)
    for kw, mine in [(dict(q=7), dict(q=7)), (dict(q=9, full_search=True, search_limit=31), dict(q=9, me_mode=1, search_limit=31)),
                     (dict(rate=200000), dict(rate=200000)), (dict(q=8, intra_only=True), dict(q=8, force_intra=True))]:
        O.ref_encode(str(tmp_path / "c.y4m"), str(tmp_path / "o.p64"), it, 6, **kw)
        assert encode_clip(it, clip, **mine) == open(tmp_path / "o.p64", "rb").read()


def test_multi_stream_batch_equals_single_streams():
    it = y4m.IT_QCIF
    S, nf = 5, 6
    clips = [y4m.synth_clip(it, nf, seed=100 + s, pan=(s - 2, 2 - s)) for s in range(S)]
    for kw in (dict(q=8, me_mode=1, search_limit=31), dict(rate=64000), dict(q=4)):
        enc = Encoder(it, S, **kw)
        try:
            for f in range(nf):
                enc.encode(np.stack([c[f] for c in clips]))
            enc.finish()
            for s in range(S):
                assert enc.data(s) == encode_clip(it, clips[s], **kw), (s, kw)
        finally:
            enc.close()


def test_ragged_and_bad_arguments():
    from p64_b200._lib import P64Error
    with pytest.raises(P64Error):
        DeviceContext(7, 1)
    with pytest.raises(P64Error):
        DeviceContext(y4m.IT_CIF, 0)
    ctx = DeviceContext(y4m.IT_QCIF, 1)
    try:
        fr = np.zeros((1, ctx.geom["frame_bytes"]), np.uint8)
        with pytest.raises(P64Error):
            ctx.encode_frames(make_step(True, 0), fr)            # quantiser 0
        with pytest.raises(P64Error):
            ctx.encode_frames(make_step(True, 8, 0, 99), fr)     # search limit
        with pytest.raises(P64Error):
            ctx.encode_gob(make_step(True, 8), 0, [8])           # outside frame_begin/frame_end
    finally:
        ctx.close()


def test_full_size_batch_properties():
    """BASELINE-size batch (256 CIF streams): every stream of a batch of identical inputs gives identical output,
    and the reconstruction equals the oracle's for a sampled stream."""
    it = y4m.IT_CIF
    S = 256
    clip = y4m.synth_clip(it, 3, 1234)
    ctx = DeviceContext(it, S); orc = O.Encoder(it)
    try:
        for f, fr in enumerate(clip):
            mbs, lv = ctx.encode_frames(make_step(f == 0, 8, 1, 31), np.broadcast_to(fr, (S, fr.size)))
            recs, olv = orc.encode_frame(fr, 8, 1, 31)
            assert (mbs == mbs[0]).all() and (lv == lv[0]).all()
            assert np.array_equal(lv[S - 1], levels_to_i8(olv))
        assert np.array_equal(ctx.recon(S - 1), orc.recon())
        assert np.array_equal(ctx.recon(0), orc.recon())
    finally:
        ctx.close()


@pytest.mark.parametrize("rate", [0, 384000])
def test_full_size_batch_round_trip_all_streams(rate):
    """BASELINE configs[4]-shaped batch: 256 CIF streams with DIFFERENT content (the bench's clip bank and phases), 10
    frames, pipelined device-side entropy coding (and, rate != 0, the device-side rate control).  Size-independent
    properties for every one of the 256 streams: (1) the stream parses, picture by picture, into as many pictures as were
    coded; (2) decoding it (host parser + the device's inverse half, all streams as one batch) reproduces the ENCODER's
    reconstruction of the last frame bit for bit -- encoder and decoder stay in lock step; (3) streams fed the same clip
    at the same phase are byte-identical; and for a sample of streams (4) the bytes equal a single-stream encode."""
    import bench
    from p64_b200.encoder import BitWriter, Parser
    it, S, nf = y4m.IT_CIF, 256, 10
    frames = bench.make_sources(range(S), nf)                         # [nf, S, frame_bytes]
    iq = 8 if not rate else min(max(10000000 // rate, 1), 31)
    ctx = DeviceContext(it, S)
    try:
        if rate:
            ctx.set_rate_control(rate)
        out = [bytearray() for _ in range(S)]
        tickets = []
        src = [np.ascontiguousarray(frames[f]) for f in range(nf)]
        for f in range(nf):
            if f >= 3:
                chunks, carry, clen, pos = ctx.wait_bits(tickets[f - 3])
                for s in range(S):
                    out[s] += chunks[s]
            tickets.append(ctx.submit_bits(make_step(f == 0, iq, 1, 31), f % 32, src[f].ctypes.data))
        for t in tickets[-3:]:
            chunks, carry, clen, pos = ctx.wait_bits(t)
            for s in range(S):
                out[s] += chunks[s]
        streams = []
        for s in range(S):
            bw = BitWriter(it)
            if clen[s]:
                bw.put(int(carry[s]) >> (32 - int(clen[s])), int(clen[s]))
            bw.picture_header(nf % 32)
            bw.finish()
            streams.append(bytes(out[s]) + bw.data())
            assert pos[s] == 8 * len(out[s]) + clen[s]
        enc_recon = [ctx.recon(s) for s in range(S)]
    finally:
        ctx.close()
    # (3) same clip, same phase -> same bytes (stream s plays clip s % 8 at phase (s // 8) % 8)
    for s in range(64, S):
        assert streams[s] == streams[s - 64], s
    assert len({hashlib.md5(x).hexdigest() for x in streams[:64]}) == 64
    # (1) + (2): parse all streams, decode them as one batch
    parsers = [Parser(x) for x in streams]
    dec = DeviceContext(it, S)
    try:
        for f in range(nf):
            pics = [p.next_picture() for p in parsers]
            assert all(pic is not None and pic[2] == f % 32 and pic[3] == 1 for pic in pics), f
            dec.decode_frames(np.stack([pic[0] for pic in pics]), np.stack([pic[1] for pic in pics]))
        assert all(p.next_picture() is None for p in parsers)
        for s in range(S):
            assert np.array_equal(dec.recon(s), enc_recon[s]), s
    finally:
        dec.close()
        for p in parsers:
            p.close()
    # (4) a sample against single-stream encodes
    for s in (0, 37, 63, 255):
        want = encode_clip(it, frames[:, s], q=0 if rate else 8, rate=rate, me_mode=1, search_limit=31)
        assert streams[s] == want, s


def test_cli_matches_reference_golden(tmp_path):
    """the p64b command line (reference flags) writes the reference's bytes"""
    import subprocess
    from p64_b200 import build
    cli = build.build_cli()
    for name, extra in [("cif8_q8_full15", ["--me", "full"]), ("qcif12_q8_tss", []), ("qcif12_q8_intra", ["--intra-only"]),
                        ("cif12_r128000_tss", []), ("qcif6_q8_tss_a3_k2_b14", []), ("qcif5_r64000_a30_k3_b42", []),
                        ("cif4_r2000000_full31_a50_k3_f15", ["--me", "full"])]:
        g, clip = golden_clip(name)
        a = g["args"]
        if a.get("start"):                                            # the file holds the StartFrame frames that -a skips
            clip = y4m.synth_clip(g["image_type"], g["n_frames"] + a["start"], g["seed"])
        y4m.write_y4m(str(tmp_path / "c.y4m"), g["image_type"], clip)
        cmd = [cli, "-y4m", y4m.FLAG[g["image_type"]], "-a", str(a.get("start", 0)), "-b", str(a.get("last", g["n_frames"] - 1))]
        cmd += ["-k", str(a["frame_skip"])] if a.get("frame_skip") else []
        cmd += ["-f", str(a["frame_rate"])] if a.get("frame_rate") else []
        cmd += (["-q", str(a["q"])] if "q" in a else []) + (["-r", str(a["rate"])] if "rate" in a else [])
        cmd += (["-i", str(a["search_limit"])] if a.get("search_limit") else []) + extra
        cmd += [str(tmp_path / "c"), "-s", str(tmp_path / "o.p64")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        data = open(tmp_path / "o.p64", "rb").read()
        assert hashlib.md5(data).hexdigest() == g["md5"], name
        assert f"Number of buffer overflows: {g['overflows']}" in r.stdout


@pytest.mark.parametrize("devices", [None, "0,0", "0-0", "0,0,0,0,0"])
def test_cli_batch_of_files_equals_single_encodes(tmp_path, devices):
    """several prefixes on one p64b command line = one batch of streams; every output equals the reference's golden -- also
    with the batch partitioned over a device list (--devices; the test box has one GPU, so the list repeats it; five entries
    for three streams leave two partitions empty) and --balance-links"""
    import subprocess
    from p64_b200 import build
    cli = build.build_cli()
    names = ["qcif12_q8_tss", "qcif20_r64000_tss"]
    extra = (["--devices", devices] + (["--balance-links"] if devices == "0,0" else [])) if devices else []
    for name in names:
        g, clip = golden_clip(name)
        rate = g["args"].get("rate")
        variants = []
        for k in range(3):                                  # the golden clip and two perturbed copies
            c = clip.copy()
            if k:
                c[:, ::7 * k] ^= 1
            y4m.write_y4m(str(tmp_path / f"{name}_{k}.y4m"), g["image_type"], c)
            variants.append(c)
        cmd = [cli, "-y4m", "-QCIF", "-a", "0", "-b", str(g["n_frames"] - 1)] + (["-r", str(rate)] if rate else ["-q", str(g["args"]["q"])])
        r = subprocess.run(cmd + extra + [str(tmp_path / f"{name}_{k}") for k in range(3)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert hashlib.md5(open(tmp_path / f"{name}_0.p64", "rb").read()).hexdigest() == g["md5"]
        for k in (1, 2):
            want = encode_clip(g["image_type"], variants[k], **golden_kwargs(g))
            assert open(tmp_path / f"{name}_{k}.p64", "rb").read() == want, (name, k)


def test_pipelined_submit_wait_equals_synchronous():
    import ctypes as C
    from p64_b200 import _lib
    it = y4m.IT_QCIF
    S, nf = 3, 7
    clips = [y4m.synth_clip(it, nf, seed=300 + s) for s in range(S)]
    sync = DeviceContext(it, S); pipe = DeviceContext(it, S)
    try:
        L = _lib.lib()
        fb, nmb = sync.geom["frame_bytes"], sync.geom["num_mb"]
        want = [sync.encode_frames(make_step(f == 0, 8, 1, 31), np.stack([c[f] for c in clips])) for f in range(nf)]
        src = [np.stack([c[f] for c in clips]).copy() for f in range(nf)]
        outs = [(np.zeros((S, nmb), want[0][0].dtype), np.zeros((S, nmb, 6, 64), np.int8)) for _ in range(nf)]
        tickets = [pipe.submit(make_step(f == 0, 8, 1, 31), src[f].ctypes.data, outs[f][0].ctypes.data, outs[f][1].ctypes.data)
                   for f in range(3)]
        for f in range(3, nf):
            pipe.wait(tickets[f - 3])
            tickets.append(pipe.submit(make_step(False, 8, 1, 31), src[f].ctypes.data, outs[f][0].ctypes.data, outs[f][1].ctypes.data))
        for t in tickets:
            pipe.wait(t)
        for f in range(nf):
            assert np.array_equal(outs[f][0], want[f][0]) and np.array_equal(outs[f][1], want[f][1]), f
        assert np.array_equal(pipe.recon(S - 1), sync.recon(S - 1))
    finally:
        sync.close(); pipe.close()


@pytest.mark.parametrize("it", [y4m.IT_QCIF, y4m.IT_CIF, y4m.IT_NTSC])
@pytest.mark.parametrize("q", [1, 2, 8, 31])
def test_device_vlc_equals_host_vlc_on_adversarial_content(it, q):
    """noise at small quantisers: escapes everywhere, frames larger than the download budget (slow path of
    p64b_ctx_wait_bits), maximal block lengths; flat frames: empty blocks, type-4 fallbacks"""
    w, h = y4m.DIMS[it]
    n = w * h * 3 // 2
    rng = np.random.default_rng(17 + q)
    frames = [rng.integers(0, 256, n).astype(np.uint8), rng.integers(0, 256, n).astype(np.uint8),
              np.full(n, 128, np.uint8), np.full(n, 128, np.uint8), rng.integers(100, 140, n).astype(np.uint8),
              np.full(n, 0, np.uint8), np.full(n, 255, np.uint8), rng.integers(0, 256, n).astype(np.uint8)]
    for mode, limit in ((0, 15), (1, 31)):
        got = encode_clip(it, np.stack(frames), q=q, me_mode=mode, search_limit=limit)
        want = encode_clip(it, np.stack(frames), q=q, me_mode=mode, search_limit=limit, host_vlc=True)
        assert got == want, (it, q, mode, len(got), len(want))


def test_device_vlc_pipelined_multi_stream():
    """p64b_ctx_submit_bits with three steps in flight, several streams: chunks + final carry reassemble the
    bytes the sequence encoder writes"""
    from p64_b200.encoder import BitWriter
    it = y4m.IT_QCIF
    S, nf = 5, 9
    clips = [y4m.synth_clip(it, nf, seed=500 + s, pan=(s - 2, 1 - s)) for s in range(S)]
    want = [encode_clip(it, clips[s], q=6, me_mode=1, search_limit=31, host_vlc=True) for s in range(S)]
    ctx = DeviceContext(it, S)
    try:
        src = [np.stack([c[f] for c in clips]).copy() for f in range(nf)]
        out = [b"" for _ in range(S)]
        tickets = []
        last = None
        for f in range(nf):
            if f >= 3:
                chunks, carry, clen, pos = ctx.wait_bits(tickets[f - 3])
                out = [o + c for o, c in zip(out, chunks)]
            tickets.append(ctx.submit_bits(make_step(f == 0, 6, 1, 31), f % 32, src[f].ctypes.data))
        for t in tickets[-3:]:
            chunks, carry, clen, pos = ctx.wait_bits(t)
            out = [o + c for o, c in zip(out, chunks)]
        for s in range(S):
            bw = BitWriter(it)
            if clen[s]:
                bw.put(int(carry[s]) >> (32 - int(clen[s])), int(clen[s]))
            bw.picture_header(nf % 32)
            bw.finish()
            assert out[s] + bw.data() == want[s], s
            assert pos[s] == 8 * len(out[s]) + clen[s]
    finally:
        ctx.close()


@pytest.mark.parametrize("it,rate,nf", [(y4m.IT_QCIF, 64000, 10), (y4m.IT_QCIF, 48000, 8), (y4m.IT_CIF, 128000, 6),
                                        (y4m.IT_CIF, 1500000, 5), (y4m.IT_NTSC, 96000, 5)])
def test_device_rate_control_multi_stream_pipelined(it, rate, nf):
    """Rate control on the device (p64b_ctx_set_rate_control + submit_bits, three steps in flight): streams with different
    content take different GQUANT / overflow paths; each must equal the host-side rate control of the sequence encoder
    run on that stream alone (which the golden tests pin to the reference's bytes)."""
    from p64_b200._lib import P64Error
    from p64_b200.encoder import BitWriter
    S = 4
    w, h = y4m.DIMS[it]
    n = w * h * 3 // 2
    rng = np.random.default_rng(rate + nf)
    clips = [y4m.synth_clip(it, nf, seed=900 + s, pan=(2 * s - 3, s - 1)) for s in range(S - 1)]
    clips.append(np.stack([rng.integers(0, 256, n).astype(np.uint8) if f % 3 != 2 else np.full(n, 90, np.uint8) for f in range(nf)]))
    want, want_ovf = [], []
    for s in range(S):
        enc = Encoder(it, 1, rate=rate, me_mode=1, search_limit=31, host_vlc=True)
        for fr in clips[s]:
            enc.encode(fr[None])
        enc.finish()
        want.append(enc.data(0)); want_ovf.append(enc.overflows(0))
        enc.close()
    iq = min(max(10000000 // rate, 1), 31)
    ctx = DeviceContext(it, S)
    try:
        ctx.set_rate_control(rate)
        src = [np.stack([c[f] for c in clips]).copy() for f in range(nf)]
        out = [b"" for _ in range(S)]
        tickets = []
        for f in range(nf):
            if f >= 3:
                chunks, carry, clen, pos = ctx.wait_bits(tickets[f - 3])
                out = [o + c for o, c in zip(out, chunks)]
            tickets.append(ctx.submit_bits(make_step(f == 0, iq, 1, 31), f % 32, src[f].ctypes.data))
        for t in tickets[-3:]:
            chunks, carry, clen, pos = ctx.wait_bits(t)
            out = [o + c for o, c in zip(out, chunks)]
        for s in range(S):
            bw = BitWriter(it)
            if clen[s]:
                bw.put(int(carry[s]) >> (32 - int(clen[s])), int(clen[s]))
            bw.picture_header(nf % 32)
            bw.finish()
            assert out[s] + bw.data() == want[s], (s, len(out[s]), len(want[s]))
            assert int(ctx.last_overflows[s]) == want_ovf[s], s
        with pytest.raises(P64Error):
            ctx.set_rate_control(rate)          # only before the first frame
    finally:
        ctx.close()
    assert sum(want_ovf) > 0 or rate >= 1000000


def test_me_microbenchmark_size_1024_pairs():
    """BASELINE configs[3]: 1024 CIF frame pairs in one motion-estimation call (405 504 macroblocks through the work
    queue).  Pairs cycle through 8 distinct seeded pairs: every copy must give the same records, and the distinct ones
    must equal the oracle's -- for the exhaustive search and for the three-step search."""
    import torch
    it = y4m.IT_CIF
    w, h = y4m.DIMS[it]
    base = [y4m.random_pair(it, 40 + k, shift=(k - 4, 3 - k), noise=k) for k in range(8)]
    n_pairs = 1024
    ref = torch.from_numpy(np.stack([base[k % 8][0] for k in range(n_pairs)])).cuda()
    cur = torch.from_numpy(np.stack([base[k % 8][1] for k in range(n_pairs)])).cuda()
    ctx = DeviceContext(it, 1)
    try:
        nmb = ctx.geom["num_mb"]
        out = torch.zeros(n_pairs * nmb * 8, dtype=torch.int32, device="cuda")
        ctx.set_cuda_stream(torch.cuda.current_stream().cuda_stream)
        for mode, limit in ((1, 31), (0, 15)):
            out.zero_()
            torch.cuda.synchronize()
            ctx.motion_estimation_dev(ref.data_ptr(), cur.data_ptr(), n_pairs, mode, limit, out.data_ptr())
            torch.cuda.synchronize()
            got = out.cpu().numpy().reshape(n_pairs, nmb, 8)[:, :, :7]
            for k in range(8):
                assert np.array_equal(got[k], O.me_frame(base[k][0], base[k][1], mode, limit)), (mode, k)
            assert np.array_equal(got, np.tile(got[:8], (n_pairs // 8, 1, 1))), mode
    finally:
        ctx.close()


@pytest.mark.parametrize("kw", [dict(q=8, me_mode=1, search_limit=31), dict(rate=64000), dict(q=6, host_vlc=True),
                                dict(rate=96000, host_vlc=True, me_mode=1, search_limit=15)])
def test_multi_device_partition_is_invisible(kw):
    """SURVEY 8(e) inside the product: p64b_enc_params.devices partitions the streams over device contexts (one worker thread
    and one p64b_ctx each, no exchange).  The bytes of every stream must not depend on the partition: 7 streams through one
    context == through 3 partitions (3+2+2 streams; the test box has one GPU, so the list names it three times) == through 8
    partitions (one of them empty)."""
    it, S, nf = y4m.IT_QCIF, 7, 6
    clips = [y4m.synth_clip(it, nf, seed=300 + s, pan=(s - 3, 2 - s)) for s in range(S)]
    outs = []
    for devices, balance in ((None, False), ([0, 0, 0], False), ([0] * 8, False), ([0, 0, 0], True)):
        enc = Encoder(it, S, devices=devices, balance_links=balance, **kw)
        parts = enc.partitions()
        if devices is None:
            assert parts == [(0, 0, S)]
        elif balance:      # blocks follow the measured link shares: contiguous, complete, none empty
            assert len(parts) == 3 and parts[0][1] == 0 and all(n >= 1 for _, _, n in parts) and sum(n for _, _, n in parts) == S
            assert all(parts[k][1] + parts[k][2] == parts[k + 1][1] for k in range(2))
        elif len(devices) == 3:
            assert parts == [(0, 0, 3), (0, 3, 2), (0, 5, 2)]
        else:
            assert len(parts) == 7 and all(n == 1 for _, _, n in parts)
        for f in range(nf):
            enc.encode(np.stack([c[f] for c in clips]))
        enc.finish()
        outs.append(([enc.data(s) for s in range(S)], [enc.overflows(s) for s in range(S)], [enc.first_frame_bits(s) for s in range(S)]))
        enc.close()
    assert outs[0] == outs[1] == outs[2] == outs[3]
    assert len(set(outs[0][0])) == S                       # the streams really differ


def test_multi_device_errors_surface():
    from p64_b200._lib import P64Error
    with pytest.raises(P64Error):
        Encoder(y4m.IT_QCIF, 2, q=8, devices=[0, 99])      # no such device: the partition's error text reaches the caller


def test_resident_bits_step_is_the_same_step():
    """p64b_ctx_encode_bits_dev (source already on the device, output left there: bench.py's `value_with_vlc`) runs the same
    kernels as p64b_ctx_submit_bits: after it, the stream state (carry bits, bit position, reconstruction, LastIntra) is what
    the host-buffer call leaves, so the following frames' bytes are identical."""
    import torch
    it, S = y4m.IT_QCIF, 3
    clips = [y4m.synth_clip(it, 3, seed=610 + s, pan=(s, -s)) for s in range(S)]
    frames = [np.stack([c[f] for c in clips]).copy() for f in range(3)]
    outs = []
    for resident_first in (False, True):
        ctx = DeviceContext(it, S)
        try:
            chunks = []
            for f in range(3):
                step = make_step(f == 0, 7, 1, 31)
                if resident_first and f == 0:
                    d = torch.from_numpy(frames[0]).cuda()
                    ctx.encode_bits_dev(step, 0, d.data_ptr())
                    torch.cuda.synchronize()
                    continue
                t = ctx.submit_bits(step, f % 32, frames[f].ctypes.data)
                c, carry, clen, pos = ctx.wait_bits(t)
                chunks.append((c, list(carry), list(clen), list(pos)))
            outs.append((chunks[-2:], [ctx.recon(s).tobytes() for s in range(S)], [ctx.last_intra(s).tobytes() for s in range(S)]))
        finally:
            ctx.close()
    assert outs[0] == outs[1]

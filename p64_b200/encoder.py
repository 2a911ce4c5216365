"""Python mirror of the C ABI objects: DeviceContext (p64b_ctx_*), BitWriter (p64b_bits_*), Encoder (p64b_enc_*).

Names and argument meaning follow the reference: a context plays the role of the frame stores CFS/OFS
(p64.c:69-72) plus the hot-path routines; BitWriter is marker.c/codec.c/stream.c; Encoder is
p64EncodeSequence (p64.c:524-613).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import MB, ME, BitsOut, EncParams, PlaneStats, RateControl, Stat, Step, Y4mInfo, check

CHROMA = {"420jpeg": 0, "420": 0, "420mpeg2": 1, "420paldv": 2, "422": 3, "411": 4, "444": 5, "444alpha": 6, "mono": 7}

MB_DTYPE = np.dtype([("mtype", "u1"), ("cbp", "u1"), ("mvx", "i1"), ("mvy", "i1"), ("quant", "u1"),
                     ("nzmask", "u1"), ("reserved", "u2")])
ME_DTYPE = np.dtype([(n, "i4") for n in ("mx", "my", "val", "oval", "var", "varor", "mwor", "pad")])


def geometry(image_type: int):
    L = _lib.lib()
    return dict(width=L.p64b_width(image_type), height=L.p64b_height(image_type),
                frame_bytes=L.p64b_frame_bytes(image_type), num_gob=L.p64b_num_gob(image_type),
                num_mb=L.p64b_num_mb(image_type))


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


def make_step(first_frame, gquant=8, me_mode=0, search_limit=15, force_intra=False) -> Step:
    return Step(int(first_frame), int(me_mode), int(search_limit), int(force_intra), int(gquant))


class DeviceContext:
    def __init__(self, image_type: int, n_streams: int = 1, device: int = 0):
        self.L = _lib.lib()
        self.image_type, self.n_streams = image_type, n_streams
        self.geom = geometry(image_type)
        h = C.c_void_p()
        check(self.L.p64b_ctx_create(C.byref(h), device, image_type, n_streams))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.p64b_ctx_destroy(self.h)
            self.h = None

    __del__ = close

    def set_cuda_stream(self, handle: int):
        check(self.L.p64b_ctx_set_cuda_stream(self.h, C.c_void_p(handle)))

    def _outputs(self, mb_per_stream):
        mbs = np.zeros((self.n_streams, mb_per_stream), MB_DTYPE)
        levels = np.zeros((self.n_streams, mb_per_stream, 6, 64), np.int8)
        return mbs, levels

    def encode_frames(self, step: Step, src: np.ndarray):
        """src uint8 [n_streams, frame_bytes] (host) -> (mbs [S,nmb], levels int8 [S,nmb,6,64])."""
        src = np.ascontiguousarray(src, np.uint8).reshape(self.n_streams, getattr(self, "raw_frame_bytes", self.geom["frame_bytes"]))
        mbs, levels = self._outputs(self.geom["num_mb"])
        check(self.L.p64b_ctx_encode_frames(self.h, C.byref(step), _ptr(src), _ptr(mbs), _ptr(levels)))
        return mbs, levels

    def submit(self, step: Step, src_ptr: int, mbs_ptr: int, levels_ptr: int) -> int:
        """Pipelined host path (raw pinned host pointers); returns a ticket for wait()."""
        t = C.c_int64()
        check(self.L.p64b_ctx_submit(self.h, C.byref(step), _ptr(src_ptr), _ptr(mbs_ptr), _ptr(levels_ptr), C.byref(t)))
        return t.value

    def set_input_chroma(self, chroma):
        """Host sources become unconverted Y4M payloads; the reader's chroma conversion (y4m_input.c:195-545) runs on the
        device.  `chroma`: a Y4M C-tag name ("420mpeg2", ...) or a P64B_CHROMA_* id."""
        cid = CHROMA[chroma] if isinstance(chroma, str) else int(chroma)
        check(self.L.p64b_ctx_set_input_chroma(self.h, cid))
        self.raw_frame_bytes = int(self.L.p64b_raw_frame_bytes(self.image_type, cid))

    def convert_frames(self, raw: np.ndarray) -> np.ndarray:
        """raw uint8 [n_streams, raw_frame_bytes] -> uint8 [n_streams, frame_bytes] (what ReadIob would install)"""
        nb = getattr(self, "raw_frame_bytes", self.geom["frame_bytes"])
        raw = np.ascontiguousarray(raw, np.uint8).reshape(self.n_streams, nb)
        out = np.zeros((self.n_streams, self.geom["frame_bytes"]), np.uint8)
        check(self.L.p64b_ctx_convert_frames(self.h, _ptr(raw), _ptr(out)))
        return out

    def set_rate_control(self, rate: int, frame_rate=(30000, 1001), frame_skip: int = 1, qdfact: int = 0, qoffs: int = 1):
        """Rate control on the device for submit_bits (p64.c:233-237, 458-481, 776-783); QDFact defaults to Rate/320."""
        r = RateControl(int(rate), int(frame_rate[0]), int(frame_rate[1]), int(frame_skip),
                        int(qdfact) if qdfact else max(int(rate) // 320, 1), int(qoffs))
        check(self.L.p64b_ctx_set_rate_control(self.h, C.byref(r)))

    def submit_bits(self, step: Step, temporal_reference: int, src_ptr: int) -> int:
        """device-side entropy coding: enqueue one frame step of every stream; `src_ptr` = host address (pinned)"""
        t = C.c_int64()
        check(self.L.p64b_ctx_submit_bits(self.h, C.byref(step), int(temporal_reference), C.c_void_p(src_ptr), C.byref(t)))
        return t.value

    def encode_bits_dev(self, step: Step, temporal_reference: int, src_dev: int):
        """the bits step with device-resident source frames, output left on the device (measurements only)"""
        check(self.L.p64b_ctx_encode_bits_dev(self.h, C.byref(step), int(temporal_reference), _ptr(src_dev)))

    def wait_bits(self, ticket: int):
        """-> (chunks: list of bytes per stream, carry, carry_len, bit_position) of that step"""
        o = BitsOut()
        check(self.L.p64b_ctx_wait_bits(self.h, ticket, C.byref(o)))
        S = self.n_streams
        off = np.ctypeslib.as_array(o.offset, (S + 1,)); nb = np.ctypeslib.as_array(o.nbytes, (S,))
        data = np.ctypeslib.as_array(o.data, (max(int(o.total_bytes), 1),))
        chunks = [data[off[s]:off[s] + nb[s]].tobytes() for s in range(S)]
        self.last_gquant = np.ctypeslib.as_array(o.gquant, (S,)).copy()
        self.last_overflows = np.ctypeslib.as_array(o.overflows, (S,)).copy()
        return (chunks, np.ctypeslib.as_array(o.carry, (S,)).copy(), np.ctypeslib.as_array(o.carry_len, (S,)).copy(),
                np.ctypeslib.as_array(o.bit_position, (S,)).copy())

    def wait(self, ticket: int):
        check(self.L.p64b_ctx_wait(self.h, ticket))

    def encode_frames_dev(self, step: Step, src_dev: int, mbs_dev: int, levels_dev: int):
        check(self.L.p64b_ctx_encode_frames_dev(self.h, C.byref(step), _ptr(src_dev), _ptr(mbs_dev), _ptr(levels_dev)))

    def frame_begin(self, step: Step, src: np.ndarray):
        src = np.ascontiguousarray(src, np.uint8).reshape(self.n_streams, self.geom["frame_bytes"])
        self._src_keepalive = src
        check(self.L.p64b_ctx_frame_begin(self.h, C.byref(step), _ptr(src)))

    def encode_gob(self, step: Step, gob: int, quant):
        quant = np.ascontiguousarray(quant, np.uint8).reshape(self.n_streams)
        mbs, levels = self._outputs(33)
        check(self.L.p64b_ctx_encode_gob(self.h, C.byref(step), gob, _ptr(quant), _ptr(mbs), _ptr(levels)))
        return mbs, levels

    def frame_end(self, overflow=None):
        if overflow is not None:
            overflow = np.ascontiguousarray(overflow, np.uint8).reshape(self.n_streams, self.geom["num_mb"])
        check(self.L.p64b_ctx_frame_end(self.h, _ptr(overflow)))

    def motion_estimation_dev(self, ref_dev: int, cur_dev: int, n_pairs: int, me_mode: int, search_limit: int,
                              out_dev: int):
        check(self.L.p64b_ctx_motion_estimation_dev(self.h, _ptr(ref_dev), _ptr(cur_dev), n_pairs, me_mode,
                                                    search_limit, _ptr(out_dev)))

    def me_records(self, stream: int = 0) -> np.ndarray:
        out = np.zeros(self.geom["num_mb"], ME_DTYPE)
        check(self.L.p64b_ctx_me_records(self.h, stream, _ptr(out)))
        return out

    def recon(self, stream: int = 0) -> np.ndarray:
        out = np.zeros(self.geom["frame_bytes"], np.uint8)
        check(self.L.p64b_ctx_download_recon(self.h, stream, _ptr(out)))
        return out

    def last_intra(self, stream: int = 0) -> np.ndarray:
        out = np.zeros(self.geom["num_mb"], np.uint8)
        check(self.L.p64b_ctx_last_intra(self.h, stream, _ptr(out)))
        return out

    def decode_frames(self, mbs: np.ndarray, levels: np.ndarray):
        """decoder's inverse half for one picture of every stream (records with reserved bit 0 = transmitted)"""
        mbs = np.ascontiguousarray(mbs, MB_DTYPE).reshape(self.n_streams, self.geom["num_mb"])
        levels = np.ascontiguousarray(levels, np.int8).reshape(self.n_streams, self.geom["num_mb"], 6, 64)
        check(self.L.p64b_ctx_decode_frames(self.h, _ptr(mbs), _ptr(levels)))

    def statistics(self):
        """Statistics of the last coded frame (stat.c:52-130): -> ((PlaneStats * 3) * n_streams) of exact integer sums"""
        arr = ((PlaneStats * 3) * self.n_streams)()
        check(self.L.p64b_ctx_statistics(self.h, C.cast(arr, C.c_void_p)))
        return arr

    def me_executed(self, reset: bool = False) -> int:
        """packed SAD operations the ME kernel's sweeps have executed so far (device counter)"""
        n = C.c_uint64()
        check(self.L.p64b_ctx_me_executed(self.h, C.byref(n), int(reset)))
        return int(n.value)

    def profile(self, enable: bool):
        check(self.L.p64b_ctx_profile(self.h, int(enable)))

    def wait_bits_raw(self, ticket: int) -> BitsOut:
        o = BitsOut()
        check(self.L.p64b_ctx_wait_bits(self.h, ticket, C.byref(o)))
        return o

    def profile_read(self):
        """-> {"me": (ms_total, launches), "mb": (...), "vlc": (ms_total, steps)}"""
        ms = (C.c_double * 3)()
        n = (C.c_int32 * 3)()
        check(self.L.p64b_ctx_profile_read(self.h, ms, n))
        return {"me": (ms[0], n[0]), "mb": (ms[1], n[1]), "vlc": (ms[2], n[2])}

    @property
    def launches(self) -> int:
        return int(self.L.p64b_ctx_launches(self.h))


class BitWriter:
    def __init__(self, image_type: int):
        self.L = _lib.lib()
        self.h = C.c_void_p(self.L.p64b_bits_create(image_type))
        if not self.h:
            raise ValueError("bad image type")

    def close(self):
        if getattr(self, "h", None):
            self.L.p64b_bits_destroy(self.h)
            self.h = None

    __del__ = close

    def picture_header(self, tr: int): self.L.p64b_bits_picture_header(self.h, tr)
    def gob_header(self, gob: int, gquant: int): self.L.p64b_bits_gob_header(self.h, gob, gquant)

    def mb(self, mdu: int, rec: np.ndarray, levels: np.ndarray):
        """rec: MB_DTYPE scalar/0-d array; levels: int8 [6,64]"""
        rec = np.ascontiguousarray(rec, MB_DTYPE)
        levels = np.ascontiguousarray(levels, np.int8)
        self.L.p64b_bits_mb(self.h, mdu, _ptr(rec), _ptr(levels))

    def put(self, value: int, nbits: int): self.L.p64b_bits_put(self.h, int(value), int(nbits))
    def tell(self) -> int: return int(self.L.p64b_bits_tell(self.h))
    def finish(self) -> int: return int(self.L.p64b_bits_finish(self.h))

    def data(self) -> bytes:
        n = C.c_size_t()
        p = self.L.p64b_bits_data(self.h, C.byref(n))
        return C.string_at(p, n.value)


def stat_from_sums(sums: PlaneStats) -> Stat:
    st = Stat()
    _lib.lib().p64b_stat_from_sums(C.byref(sums), C.byref(st))
    return st


def format_statistics(planes) -> list:
    """the `Comp:` lines Statistics() prints (stat.c:60-62) for one stream's (PlaneStats * 3)"""
    out = []
    for i in range(3):
        s = stat_from_sums(planes[i])
        out.append("Comp: %d  MRSNR: %2.2f  SNR: %2.2f  PSNR: %2.2f  MSE: %4.2f  Entropy: %1.2f" %
                   (i, s.mrsnr, s.snr, s.psnr, s.mse, s.entropy))
    return out


def default_params() -> EncParams:
    p = EncParams()
    _lib.lib().p64b_enc_default_params(C.byref(p))
    return p


class Encoder:
    """Batch sequence encoder: `encode(frames)` once per frame time, then `finish()`; `data(s)` = stream bytes."""

    def __init__(self, image_type: int, n_streams: int = 1, *, q: int = 0, rate: int = 0, me_mode: int = 0,
                 search_limit: int = 15, force_intra: bool = False, start_frame: int = 0, device: int = 0,
                 frame_rate=(30000, 1001), frame_skip: int = 1, vlc_threads: int = 0, host_vlc: bool = False,
                 input_chroma="420jpeg", last_frame=None, devices=None, balance_links=False):
        """devices: list of GPU indices -- the streams are partitioned over them in contiguous blocks (one context and one
        host worker thread per GPU, no exchange); None = the single `device`."""
        self.L = _lib.lib()
        p = default_params()
        if devices is not None:
            p.n_devices = len(devices)
            for k, d in enumerate(devices):
                p.devices[k] = int(d)
            p.balance_links = int(balance_links)
        p.image_type, p.n_streams, p.device, p.start_frame = image_type, n_streams, device, start_frame
        p.initial_quant, p.rate, p.me_mode, p.search_limit = q, rate, me_mode, search_limit
        p.force_intra, p.frame_rate, p.frame_rate_div, p.frame_skip = int(force_intra), frame_rate[0], frame_rate[1], frame_skip
        p.vlc_threads = vlc_threads
        p.host_vlc = int(host_vlc)
        p.input_chroma = CHROMA[input_chroma] if isinstance(input_chroma, str) else int(input_chroma)
        p.last_frame = 0 if last_frame is None else int(last_frame) + 1      # the reference's -b (p64.c:600-602)
        self.n_streams = n_streams
        self.geom = geometry(image_type)
        self.src_bytes = int(self.L.p64b_raw_frame_bytes(image_type, p.input_chroma))
        h = C.c_void_p()
        check(self.L.p64b_enc_create(C.byref(h), C.byref(p)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.p64b_enc_destroy(self.h)
            self.h = None

    __del__ = close

    def encode(self, frames: np.ndarray):
        frames = np.ascontiguousarray(frames, np.uint8).reshape(self.n_streams, self.src_bytes)
        check(self.L.p64b_enc_encode(self.h, _ptr(frames)))

    def staging(self) -> np.ndarray:
        """the encoder's own pinned upload buffer for the NEXT frame, uint8 [n_streams, src_bytes]: fill it in place and pass it
        to encode() -- no staging copy then.  It changes after every encode()."""
        p = self.L.p64b_enc_staging(self.h)
        buf = (C.c_uint8 * (self.n_streams * self.src_bytes)).from_address(p)
        return np.frombuffer(buf, np.uint8).reshape(self.n_streams, self.src_bytes)

    def finish(self):
        check(self.L.p64b_enc_finish(self.h))

    def data(self, stream: int = 0) -> bytes:
        n = C.c_size_t()
        p = self.L.p64b_enc_data(self.h, stream, C.byref(n))
        return C.string_at(p, n.value)

    def overflows(self, stream: int = 0) -> int: return int(self.L.p64b_enc_overflows(self.h, stream))
    def first_frame_bits(self, stream: int = 0) -> int: return int(self.L.p64b_enc_first_frame_bits(self.h, stream))

    def context_handle(self):
        return C.c_void_p(self.L.p64b_enc_ctx(self.h))

    def partitions(self):
        """-> [(device, first_stream, n_streams)] of the encoder's device partitions"""
        out = []
        for k in range(int(self.L.p64b_enc_partitions(self.h))):
            d, f, n = C.c_int(), C.c_int(), C.c_int()
            check(self.L.p64b_enc_partition(self.h, k, C.byref(d), C.byref(f), C.byref(n)))
            out.append((d.value, f.value, n.value))
        return out


class Y4mReader:
    """p64b_y4m_*: the ingest path's YUV4MPEG2 reader (header / FRAME parsing as y4m_input.c:75-134, 556-586, 708-740)."""

    def __init__(self, path: str):
        self.L = _lib.lib()
        h = C.c_void_p()
        check(self.L.p64b_y4m_open(C.byref(h), path.encode()))
        self.h = h
        self.info = Y4mInfo()
        check(self.L.p64b_y4m_get_info(self.h, C.byref(self.info)))

    def read_frame(self):
        """-> uint8 [frame_bytes] payload, or None at end of file"""
        buf = np.empty(int(self.info.frame_bytes), np.uint8)
        r = self.L.p64b_y4m_read_frame(self.h, _ptr(buf))
        if r < 0:
            check(r)
        return buf if r == 1 else None

    def close(self):
        if getattr(self, "h", None):
            self.L.p64b_y4m_close(self.h)
            self.h = None

    __del__ = close


class Parser:
    """p64b_parser_*: the decoder's sequential half alone (no device) -- pictures as records + levels"""

    def __init__(self, data: bytes):
        self.L = _lib.lib()
        self._data = np.frombuffer(data, np.uint8).copy()       # must outlive the parser
        h = C.c_void_p()
        check(self.L.p64b_parser_create(C.byref(h), _ptr(self._data), len(self._data)))
        self.h = h
        self.image_type = int(self.L.p64b_parser_image_type(self.h))
        self.num_mb = int(self.L.p64b_num_mb(self.image_type))

    def next_picture(self):
        """-> (mbs [nmb], levels int8 [nmb,6,64], temporal_reference, repeat) or None at the end"""
        mbs = np.zeros(self.num_mb, MB_DTYPE)
        lv = np.zeros((self.num_mb, 6, 64), np.int8)
        tr, rep = C.c_int(), C.c_int()
        r = self.L.p64b_parser_next_picture(self.h, _ptr(mbs), _ptr(lv), C.byref(tr), C.byref(rep))
        if r < 0:
            check(r)
        return (mbs, lv, tr.value, rep.value) if r == 1 else None

    def close(self):
        if getattr(self, "h", None):
            self.L.p64b_parser_destroy(self.h)
            self.h = None

    __del__ = close


class Decoder:
    """p64b_dec_*: p64DecodeSequence (p64.c:1022-1126) -- host parser + the device's inverse half"""

    def __init__(self, data: bytes, device: int = 0):
        self.L = _lib.lib()
        self._data = np.frombuffer(data, np.uint8).copy()
        h = C.c_void_p()
        check(self.L.p64b_dec_create(C.byref(h), device, _ptr(self._data), len(self._data)))
        self.h = h
        self.image_type = int(self.L.p64b_dec_image_type(self.h))
        self.frame_bytes = int(self.L.p64b_frame_bytes(self.image_type))

    def frames(self):
        """every frame p64DecodeSequence writes, in order (a picture repeated for temporal-reference gaps)"""
        out = []
        while True:
            buf = np.zeros(self.frame_bytes, np.uint8)
            rep = C.c_int()
            r = self.L.p64b_dec_next_picture(self.h, _ptr(buf), C.byref(rep))
            if r < 0:
                check(r)
            if r != 1:
                return out
            out += [buf] * rep.value

    def close(self):
        if getattr(self, "h", None):
            self.L.p64b_dec_destroy(self.h)
            self.h = None

    __del__ = close


def encode_clip(image_type: int, clip: np.ndarray, **kw) -> bytes:
    """One stream: clip uint8 [n_frames, frame_bytes] -> .p64 bytes."""
    enc = Encoder(image_type, 1, **kw)
    try:
        for fr in clip:
            enc.encode(fr[None])
        enc.finish()
        return enc.data(0)
    finally:
        enc.close()

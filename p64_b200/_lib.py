"""ctypes binding of the C ABI (include/p64_b200.h).  Loading fails loudly: there is no fallback path."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("P64B_LIB") or os.path.join(HERE, "libp64b200.so")   # P64B_LIB: experiments only


class MB(C.Structure):
    """p64b_mb"""
    _fields_ = [("mtype", C.c_uint8), ("cbp", C.c_uint8), ("mvx", C.c_int8), ("mvy", C.c_int8),
                ("quant", C.c_uint8), ("nzmask", C.c_uint8), ("reserved", C.c_uint16)]


class ME(C.Structure):
    """p64b_me"""
    _fields_ = [(n, C.c_int32) for n in ("mx", "my", "val", "oval", "var", "varor", "mwor", "pad")]


class Step(C.Structure):
    """p64b_step"""
    _fields_ = [("first_frame", C.c_int32), ("me_mode", C.c_int32), ("search_limit", C.c_int32),
                ("force_intra", C.c_int32), ("gquant", C.c_int32), ("reserved", C.c_int32 * 3)]


class EncParams(C.Structure):
    """p64b_enc_params"""
    _fields_ = [(n, C.c_int32) for n in ("image_type", "n_streams", "device", "start_frame", "initial_quant", "rate",
                                         "frame_rate", "frame_rate_div", "frame_skip", "me_mode", "search_limit",
                                         "force_intra", "vlc_threads", "host_vlc", "input_chroma", "last_frame", "n_devices")] + \
               [("devices", C.c_int32 * 16), ("balance_links", C.c_int32)]


class Y4mInfo(C.Structure):
    """p64b_y4m_info"""
    _fields_ = [(n, C.c_int32) for n in ("width", "height", "fps_n", "fps_d", "par_n", "par_d", "chroma", "interlace")] + \
               [("frame_bytes", C.c_int64)]


class BitsOut(C.Structure):
    """p64b_bits_out"""
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("offset", C.POINTER(C.c_uint32)), ("nbytes", C.POINTER(C.c_uint32)),
                ("carry", C.POINTER(C.c_uint32)), ("carry_len", C.POINTER(C.c_uint32)),
                ("bit_position", C.POINTER(C.c_uint64)), ("total_bytes", C.c_size_t), ("downloaded_bytes", C.c_size_t),
                ("gquant", C.POINTER(C.c_uint32)), ("overflows", C.POINTER(C.c_uint32))]


class PlaneStats(C.Structure):
    """p64b_plane_stats"""
    _fields_ = [(n, C.c_uint64) for n in ("n", "sum_src", "sum_rec", "sum_sq_err", "sum_sq_src")] + [("hist", C.c_uint32 * 256)]


class Stat(C.Structure):
    """p64b_stat"""
    _fields_ = [(n, C.c_double) for n in ("mean", "mse", "snr", "mrsnr", "psnr", "entropy")]


class FrameCounters(C.Structure):
    """p64b_frame_counters"""
    _fields_ = [(n, C.c_int32) for n in ("mb_attribute_bits", "mv_bits", "eob_bits", "y_bits", "u_bits", "v_bits",
                                         "number_nz", "q_sum", "q_use")] + \
               [(n, C.c_int32 * 10) for n in ("macro_type_freq", "y_type_freq", "uv_type_freq")] + \
               [(n, C.c_int32) for n in ("total_bits", "last_bits", "buffer_contents", "buffer_size")]


class RateControl(C.Structure):
    """p64b_rate_control"""
    _fields_ = [(n, C.c_int32) for n in ("rate", "frame_rate", "frame_rate_div", "frame_skip", "qdfact", "qoffs")] + \
               [("reserved", C.c_int32 * 2)]


# name -> (restype, argtypes); every symbol include/p64_b200.h declares
_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t
SIGNATURES = {
    "p64b_last_error": (C.c_char_p, []),
    "p64b_version": (_i, []),
    "p64b_width": (_i, [_i]), "p64b_height": (_i, [_i]), "p64b_frame_bytes": (_i, [_i]),
    "p64b_num_gob": (_i, [_i]), "p64b_num_mb": (_i, [_i]),
    "p64b_ctx_create": (_i, [C.POINTER(_vp), _i, _i, _i]),
    "p64b_ctx_destroy": (None, [_vp]),
    "p64b_ctx_streams": (_i, [_vp]),
    "p64b_ctx_set_cuda_stream": (_i, [_vp, _vp]),
    "p64b_host_alloc": (_vp, [_sz]),
    "p64b_host_free": (None, [_vp]),
    "p64b_ctx_encode_frames": (_i, [_vp, C.POINTER(Step), _vp, _vp, _vp]),
    "p64b_ctx_submit": (_i, [_vp, C.POINTER(Step), _vp, _vp, _vp, C.POINTER(C.c_int64)]),
    "p64b_ctx_wait": (_i, [_vp, C.c_int64]),
    "p64b_ctx_encode_frames_dev": (_i, [_vp, C.POINTER(Step), _vp, _vp, _vp]),
    "p64b_ctx_submit_bits": (_i, [_vp, C.POINTER(Step), _i, _vp, C.POINTER(C.c_int64)]),
    "p64b_ctx_wait_bits": (_i, [_vp, C.c_int64, C.POINTER(BitsOut)]),
    "p64b_ctx_encode_bits_dev": (_i, [_vp, C.POINTER(Step), _i, _vp]),
    "p64b_ctx_set_rate_control": (_i, [_vp, C.POINTER(RateControl)]),
    "p64b_raw_frame_bytes": (_i, [_i, _i]),
    "p64b_ctx_set_input_chroma": (_i, [_vp, _i]),
    "p64b_ctx_convert_frames": (_i, [_vp, _vp, _vp]),
    "p64b_ctx_decode_frames": (_i, [_vp, _vp, _vp]),
    "p64b_dec_create": (_i, [C.POINTER(_vp), _i, _vp, _sz]),
    "p64b_dec_destroy": (None, [_vp]),
    "p64b_dec_image_type": (_i, [_vp]),
    "p64b_dec_next_picture": (_i, [_vp, _vp, C.POINTER(_i)]),
    "p64b_parser_create": (_i, [C.POINTER(_vp), _vp, _sz]),
    "p64b_parser_destroy": (None, [_vp]),
    "p64b_parser_image_type": (_i, [_vp]),
    "p64b_parser_next_picture": (_i, [_vp, _vp, _vp, C.POINTER(_i), C.POINTER(_i)]),
    "p64b_y4m_open": (_i, [C.POINTER(_vp), C.c_char_p]),
    "p64b_y4m_close": (None, [_vp]),
    "p64b_y4m_get_info": (_i, [_vp, C.POINTER(Y4mInfo)]),
    "p64b_y4m_read_frame": (_i, [_vp, _vp]),
    "p64b_y4m_payload_bytes": (C.c_int64, [_i, _i, _i]),
    "p64b_enc_staging": (_vp, [_vp]),
    "p64b_ctx_frame_begin": (_i, [_vp, C.POINTER(Step), _vp]),
    "p64b_ctx_encode_gob": (_i, [_vp, C.POINTER(Step), _i, _vp, _vp, _vp]),
    "p64b_ctx_frame_end": (_i, [_vp, _vp]),
    "p64b_ctx_motion_estimation_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "p64b_ctx_sad_surface_dev": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "p64b_ctx_me_records": (_i, [_vp, _i, _vp]),
    "p64b_ctx_download_recon": (_i, [_vp, _i, _vp]),
    "p64b_ctx_last_intra": (_i, [_vp, _i, _vp]),
    "p64b_ctx_statistics": (_i, [_vp, _vp]),
    "p64b_stat_from_sums": (None, [C.POINTER(PlaneStats), C.POINTER(Stat)]),
    "p64b_ctx_launches": (C.c_int64, [_vp]),
    "p64b_ctx_second_copies": (C.c_int64, [_vp]),
    "p64b_ctx_me_executed": (_i, [_vp, C.POINTER(C.c_uint64), _i]),
    "p64b_ctx_profile": (_i, [_vp, _i]),
    "p64b_ctx_profile_read": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "p64b_measure_sad_peak": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "p64b_measure_h2d": (_i, [_i, _vp, _sz, _i, C.POINTER(C.c_double)]),
    "p64b_measure_link": (_i, [_i, C.POINTER(_vp), _i, _sz, _vp, _sz, _i, _i, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "p64b_host_alloc_flags": (_vp, [_sz, _i]),
    "p64b_debug_oob": (_i, [_i, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "p64b_debug_pool_selftest": (_i, [_i, _i, _i]),
    "p64b_probe_links": (_i, [C.POINTER(C.c_int32), _i, C.POINTER(C.c_double)]),
    "p64b_bits_create": (_vp, [_i]),
    "p64b_bits_destroy": (None, [_vp]),
    "p64b_bits_picture_header": (None, [_vp, _i]),
    "p64b_bits_gob_header": (None, [_vp, _i, _i]),
    "p64b_bits_mb": (None, [_vp, _i, _vp, _vp]),
    "p64b_bits_put": (None, [_vp, C.c_uint32, _i]),
    "p64b_bits_tell": (C.c_int64, [_vp]),
    "p64b_bits_finish": (_sz, [_vp]),
    "p64b_bits_data": (C.POINTER(C.c_uint8), [_vp, C.POINTER(_sz)]),
    "p64b_bits_reset": (None, [_vp]),
    "p64b_bits_counters": (None, [_vp, C.POINTER(FrameCounters)]),
    "p64b_bits_counters_reset": (None, [_vp]),
    "p64b_enc_frame_counters": (_i, [_vp, _i, C.POINTER(FrameCounters)]),
    "p64b_enc_default_params": (None, [C.POINTER(EncParams)]),
    "p64b_enc_create": (_i, [C.POINTER(_vp), C.POINTER(EncParams)]),
    "p64b_enc_destroy": (None, [_vp]),
    "p64b_enc_encode": (_i, [_vp, _vp]),
    "p64b_enc_finish": (_i, [_vp]),
    "p64b_enc_data": (C.POINTER(C.c_uint8), [_vp, _i, C.POINTER(_sz)]),
    "p64b_enc_ctx": (_vp, [_vp]),
    "p64b_enc_partitions": (_i, [_vp]),
    "p64b_enc_partition": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "p64b_enc_overflows": (C.c_int64, [_vp, _i]),
    "p64b_enc_first_frame_bits": (C.c_int64, [_vp, _i]),
}

_lib = None


def lib() -> C.CDLL:
    """The loaded C-ABI library. Raises if it has not been built (`python -m p64_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m p64_b200.build` "
                               "(p64_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the library does not export it
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


class P64Error(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        raise P64Error(f"p64_b200 error {rc}: {lib().p64b_last_error().decode(errors='replace')}")

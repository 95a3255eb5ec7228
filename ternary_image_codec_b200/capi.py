"""ctypes binding of libt3c.so (include/t3c.h) and a thin numpy/torch-friendly host mirror of the
reference interface (same function names as old/include/ternary_image_codec_v6_min.hpp).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present, the
constructors raise.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libt3c.so")

REF_EXACT, FIXED = 0, 1
RAW_MODE = 0xFF
P1_RS26_24, P2_RS26_22, P3_RS26_20, P4_RS26_18, P5_RS26_22_2D = 0, 1, 2, 3, 4
UEP_LUMA_PRIORITY = (2, 1, 1, 2, 1, 1, 2, 1, 1)  # uep_luma_priority, OLD:68-72
PIXEL_DTYPE = np.dtype([("Yq", "<u2"), ("Cbq", "<i2"), ("Crq", "<i2")])  # PixelYCbCrQuant, OLD:670-674

OK, ERR_ARG, ERR_CUDA, ERR_CAPACITY, ERR_NODEVICE, ERR_UNSUPPORTED = range(6)


class T3CError(RuntimeError):
    pass


class Config(C.Structure):
    """t3c_config == EncoderConfig (OLD:862-873) == DecoderConfigSeen (OLD:874-884)."""
    _fields_ = [
        ("profile", C.c_uint8), ("uep", C.c_uint8 * 9),
        ("tile_w", C.c_uint16), ("tile_h", C.c_uint16),
        ("seed_a", C.c_uint32), ("seed_b", C.c_uint32), ("seed_s0", C.c_uint32),
        ("beacon_period", C.c_uint32), ("beacon_slot", C.c_uint8), ("beacon_enabled", C.c_uint8),
        ("subword", C.c_uint8), ("centered", C.c_uint8), ("coset", C.c_uint8), ("pad_", C.c_uint8 * 3),
        ("superframe_words", C.c_uint32),
    ]

    def copy(self) -> "Config":
        return Config.from_buffer_copy(bytes(self))


def make_config(profile=P2_RS26_22, uep=1, tile=(0, 0), seed=(1, 1, 1), beacon=(0, 0, False), superframe_words=8192,
                subword=27, centered=True, coset=0) -> Config:
    """EncoderConfig with the reference's defaults; ``uep`` is one index (uep_uniform) or nine."""
    c = Config()
    c.profile = profile
    u = [uep] * 9 if isinstance(uep, int) else list(uep)
    for i in range(9):
        c.uep[i] = u[i]
    c.tile_w, c.tile_h = tile
    c.seed_a, c.seed_b, c.seed_s0 = seed
    c.beacon_period, c.beacon_slot, c.beacon_enabled = beacon[0], beacon[1], 1 if beacon[2] else 0
    c.superframe_words = superframe_words
    c.subword, c.centered, c.coset = subword, 1 if centered else 0, coset
    return c


_EXPORTS = [
    "t3c_create", "t3c_destroy", "t3c_set_host_registration", "t3c_stream_create", "t3c_stream_destroy", "t3c_stream_lanes",
    "t3c_stream_encode_rgb8", "t3c_stream_decode_rgb8", "t3c_last_error", "t3c_version", "t3c_config_default", "t3c_stream", "t3c_sync",
    "t3c_kernel_launches", "t3c_profile_words", "t3c_rgb_to_quant", "t3c_quant_to_rgb", "t3c_pack_pixels",
    "t3c_unpack_pixels", "t3c_words_to_bytes", "t3c_rs_encode_blocks", "t3c_rs_decode_blocks", "t3c_interleave2d",
    "t3c_header_emit", "t3c_header_parse", "t3c_header_pack", "t3c_header_check", "t3c_header_unpack", "t3c_crc3_rem12",
    "t3c_scramble_symbols", "t3c_beacon_symbol", "t3c_gf27_tables", "t3c_t3v_index_build", "t3c_encode_profile", "t3c_decode_profile", "t3c_decode_profile_fixed",
    "t3c_encode_frames_rgb8", "t3c_decode_frames_rgb8", "t3c_rgb_to_quant_dev", "t3c_quant_to_rgb_dev",
    "t3c_pack_pixels_dev", "t3c_unpack_pixels_dev", "t3c_rs_encode_blocks_dev", "t3c_rs_decode_blocks_dev",
    "t3c_encode_profile_dev", "t3c_decode_profile_fixed_dev", "t3c_encode_frames_rgb8_dev",
    "t3c_decode_frames_rgb8_dev", "t3c_fast_path_available", "t3c_super_path_available", "t3c_debug_counters", "t3c_super_plan_describe",
    "t3c_subword_stream", "t3c_words_from_subword_stream", "t3c_base243_pack", "t3c_base243_unpack", "t3c_words_to_base243",
    "t3c_v6new_pack_pixels", "t3c_v6new_unpack_pixels", "t3c_subword_stream_dev", "t3c_words_from_subword_stream_dev",
    "t3c_base243_pack_dev", "t3c_base243_unpack_dev", "t3c_words_to_base243_dev", "t3c_v6new_pack_pixels_dev", "t3c_v6new_unpack_pixels_dev",
    "t3c_resize_rgb_nn", "t3c_blit_center_rgb", "t3c_extract_center_q", "t3c_resize_rgb_nn_dev", "t3c_blit_center_rgb_dev", "t3c_extract_center_q_dev",
    "t3c_v6new_image_to_words", "t3c_v6new_words_to_image", "t3c_crc32", "t3c_t3v_frame_record", "t3c_t3v_read_frame", "t3c_t3v_header", "t3c_t3v_frame_records_dev", "t3c_t3v_read_frames_dev",
]

_lib = None


def load_library() -> C.CDLL:
    """dlopen libt3c.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise T3CError(f"{LIB_PATH} is missing: run `python -m ternary_image_codec_b200._build` "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, sz, u8p, i32 = C.c_void_p, C.c_size_t, C.c_void_p, C.c_int
    cfgp = C.POINTER(Config)
    szp = C.POINTER(C.c_size_t)
    L.t3c_create.argtypes = [i32, C.POINTER(vp)]
    L.t3c_destroy.argtypes = [vp]
    L.t3c_set_host_registration.argtypes = [vp, i32]
    L.t3c_stream_create.argtypes = [C.POINTER(i32), i32, C.POINTER(vp)]
    L.t3c_stream_destroy.argtypes = [vp]
    L.t3c_stream_destroy.restype = None
    L.t3c_stream_lanes.argtypes = [vp]
    L.t3c_stream_encode_rgb8.argtypes = [vp, cfgp, i32, vp, sz, sz, sz, vp, sz, szp]
    L.t3c_stream_decode_rgb8.argtypes = [vp, cfgp, vp, sz, sz, sz, sz, sz, vp, vp, szp]
    L.t3c_destroy.restype = None
    L.t3c_last_error.argtypes = [vp]
    L.t3c_last_error.restype = C.c_char_p
    L.t3c_config_default.argtypes = [cfgp]
    L.t3c_config_default.restype = None
    L.t3c_stream.argtypes = [vp]
    L.t3c_stream.restype = vp
    L.t3c_sync.argtypes = [vp]
    L.t3c_kernel_launches.argtypes = [vp]
    L.t3c_kernel_launches.restype = C.c_uint64
    L.t3c_profile_words.argtypes = [cfgp, sz]
    L.t3c_profile_words.restype = sz
    L.t3c_fast_path_available.argtypes = [cfgp]
    L.t3c_super_path_available.argtypes = [cfgp]
    L.t3c_rgb_to_quant.argtypes = [vp, u8p, sz, vp]
    L.t3c_quant_to_rgb.argtypes = [vp, vp, sz, u8p]
    L.t3c_pack_pixels.argtypes = [vp, vp, sz, u8p, szp]
    L.t3c_unpack_pixels.argtypes = [vp, u8p, sz, vp]
    L.t3c_words_to_bytes.argtypes = [vp, u8p, sz, u8p]
    L.t3c_rs_encode_blocks.argtypes = [vp, i32, i32, u8p, sz, u8p]
    L.t3c_rs_decode_blocks.argtypes = [vp, i32, i32, u8p, sz, u8p, u8p]
    L.t3c_interleave2d.argtypes = [vp, u8p, sz, C.c_uint16, C.c_uint16, i32]
    L.t3c_header_emit.argtypes = [vp, cfgp, i32, u8p, u8p]
    L.t3c_header_parse.argtypes = [vp, i32, u8p, sz, cfgp, C.POINTER(i32)]
    # L0 / L1 names of the reference's public surface (single items); t3c_header / t3c_gf27 travel as opaque buffers here
    L.t3c_header_pack.argtypes = [vp, vp, u8p]
    L.t3c_header_check.argtypes = [vp, u8p, C.POINTER(i32)]
    L.t3c_header_unpack.argtypes = [vp, u8p, vp]
    L.t3c_crc3_rem12.argtypes = [vp, u8p, sz, u8p]
    L.t3c_scramble_symbols.argtypes = [vp, u8p, sz, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), i32]
    L.t3c_beacon_symbol.argtypes = [vp, i32, C.c_uint32, C.c_uint32, u8p]
    L.t3c_gf27_tables.argtypes = [vp, vp]
    L.t3c_t3v_index_build.argtypes = [vp, vp, sz, C.c_uint64, u8p, szp]
    L.t3c_encode_profile.argtypes = [vp, cfgp, i32, u8p, sz, u8p, sz, szp]
    L.t3c_decode_profile.argtypes = [vp, cfgp, u8p, sz, u8p, sz, szp, C.POINTER(i32)]
    L.t3c_decode_profile_fixed.argtypes = [vp, cfgp, sz, u8p, sz, u8p, sz, szp, C.POINTER(i32), szp]
    L.t3c_encode_frames_rgb8.argtypes = [vp, cfgp, i32, u8p, sz, sz, u8p, sz, szp]
    L.t3c_decode_frames_rgb8.argtypes = [vp, cfgp, u8p, sz, sz, sz, sz, u8p, u8p, szp, szp]
    L.t3c_rgb_to_quant_dev.argtypes = [vp, vp, sz, vp, vp]
    L.t3c_quant_to_rgb_dev.argtypes = [vp, vp, sz, vp, vp]
    L.t3c_pack_pixels_dev.argtypes = [vp, vp, sz, vp, vp]
    L.t3c_unpack_pixels_dev.argtypes = [vp, vp, sz, vp, vp]
    L.t3c_rs_encode_blocks_dev.argtypes = [vp, i32, i32, vp, sz, vp, vp]
    L.t3c_rs_decode_blocks_dev.argtypes = [vp, i32, i32, vp, sz, vp, vp, vp]
    L.t3c_encode_profile_dev.argtypes = [vp, cfgp, i32, vp, sz, vp, sz, vp]
    L.t3c_decode_profile_fixed_dev.argtypes = [vp, cfgp, sz, vp, sz, vp, sz, vp, vp]
    L.t3c_encode_frames_rgb8_dev.argtypes = [vp, cfgp, i32, vp, sz, sz, vp, sz, vp]
    L.t3c_decode_frames_rgb8_dev.argtypes = [vp, cfgp, vp, sz, sz, sz, sz, vp, vp, vp]
    L.t3c_subword_stream.argtypes = [vp, u8p, sz, i32, u8p]
    L.t3c_words_from_subword_stream.argtypes = [vp, u8p, sz, i32, C.c_uint8, u8p, szp]
    L.t3c_base243_pack.argtypes = [vp, u8p, sz, u8p, szp]
    L.t3c_base243_unpack.argtypes = [vp, u8p, sz, u8p, sz, szp, C.POINTER(i32)]
    L.t3c_words_to_base243.argtypes = [vp, u8p, sz, i32, u8p, szp]
    L.t3c_v6new_pack_pixels.argtypes = [vp, vp, sz, vp, i32]
    L.t3c_v6new_unpack_pixels.argtypes = [vp, vp, sz, vp, i32]
    L.t3c_subword_stream_dev.argtypes = [vp, vp, sz, i32, vp, vp]
    L.t3c_words_from_subword_stream_dev.argtypes = [vp, vp, sz, i32, C.c_uint8, vp, vp]
    L.t3c_base243_pack_dev.argtypes = [vp, vp, sz, vp, vp]
    L.t3c_base243_unpack_dev.argtypes = [vp, vp, sz, vp, vp]
    L.t3c_words_to_base243_dev.argtypes = [vp, vp, sz, i32, vp, vp]
    L.t3c_v6new_pack_pixels_dev.argtypes = [vp, vp, sz, vp, vp]
    L.t3c_v6new_unpack_pixels_dev.argtypes = [vp, vp, sz, vp, vp]
    u32 = C.c_uint32
    L.t3c_resize_rgb_nn.argtypes = [vp, u8p, i32, i32, u8p, i32, i32]
    L.t3c_blit_center_rgb.argtypes = [vp, u8p, i32, i32, u8p, i32, i32]
    L.t3c_extract_center_q.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    L.t3c_resize_rgb_nn_dev.argtypes = [vp, vp, i32, i32, vp, i32, i32, vp]
    L.t3c_blit_center_rgb_dev.argtypes = [vp, vp, i32, i32, vp, i32, i32, vp]
    L.t3c_extract_center_q_dev.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    L.t3c_v6new_image_to_words.argtypes = [vp, u8p, i32, i32, i32, i32, vp, sz, szp, C.POINTER(i32)]
    L.t3c_v6new_words_to_image.argtypes = [vp, vp, sz, i32, i32, i32, u8p, C.POINTER(i32)]
    L.t3c_crc32.argtypes = [vp, u8p, sz, C.POINTER(u32)]
    L.t3c_t3v_frame_record.argtypes = [vp, u8p, sz, u8p, szp]
    L.t3c_t3v_read_frame.argtypes = [vp, u8p, sz, u8p, sz, szp, C.POINTER(i32)]
    L.t3c_t3v_header.argtypes = [vp, u8p, i32, i32, i32, i32, u32, u32, C.POINTER(u32), u32, u32, u32, i32]
    L.t3c_t3v_frame_records_dev.argtypes = [vp, vp, sz, sz, sz, vp, sz, vp]
    L.t3c_t3v_read_frames_dev.argtypes = [vp, vp, sz, sz, sz, vp, sz, vp, vp]
    for name in _EXPORTS:
        getattr(L, name)  # AttributeError here = header and library out of sync
    _lib = L
    return L


def exported_symbols():
    return list(_EXPORTS)


def profile_words(cfg: Config, n_raw_words: int) -> int:
    return int(load_library().t3c_profile_words(C.byref(cfg), n_raw_words))


def fast_path_available(cfg: Config) -> bool:
    return bool(load_library().t3c_fast_path_available(C.byref(cfg)))


def super_plan(cfg: Config, n_raw_words: int, decode: bool = False, words: bool = False):
    """host-only: the super-tile plan for one super-frame, or None; dict with M, UN, n_tiles, k, ncw, npass, smem, map, pass_kv"""
    L = load_library()
    out = (C.c_uint32 * 16)()
    mp = np.zeros(3 * 64 * 32, np.uint16)
    kv = np.zeros(3 * 64, np.uint8)
    L.t3c_super_plan_describe.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    if not L.t3c_super_plan_describe(C.byref(cfg), n_raw_words, int(decode), int(words), out, mp.ctypes.data, kv.ctypes.data):
        return None
    v = list(out)
    nk = v[3]
    return dict(M=v[0], UN=v[1], n_tiles=v[2], k=v[4:4 + nk], ncw=v[8:8 + nk], npass=v[12:15], smem=v[15], map=mp.reshape(3, 64, 32), pass_kv=kv.reshape(3, 64))


def super_path_available(cfg: Config) -> bool:
    """the super-tile kernels (per-band k, 2D, beacon) take this config"""
    return bool(load_library().t3c_super_path_available(C.byref(cfg)))


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint8)


class Codec:
    """One context per device (EncoderContext + DecoderContext of the reference, OLD:885-916)."""

    def __init__(self, device: int = 0, arith: int = REF_EXACT):
        self.lib = load_library()
        self.h = C.c_void_p()
        self.arith = arith
        self.cfg = make_config()                 # EncoderContext::cfg
        self.cfg_last_seen = make_config()       # DecoderContext::cfg_last_seen
        st = self.lib.t3c_create(device, C.byref(self.h))
        if st == ERR_NODEVICE:
            raise T3CError("t3c_create: no CUDA device -- this library has no CPU fallback")
        if st != OK:
            raise T3CError(f"t3c_create failed with status {st}")

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.t3c_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != OK:
            raise T3CError(f"status {st}: {self.lib.t3c_last_error(self.h).decode()}")

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.t3c_kernel_launches(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.t3c_stream(self.h) or 0)

    def sync(self):
        self._ck(self.lib.t3c_sync(self.h))

    # ---------------- data formats either side of the path ----------------
    def extract_subword_stream_from_words(self, words, N) -> np.ndarray:
        w = _u8(words)
        nw = w.size // 9
        out = np.zeros(nw * N, np.uint8)
        self._ck(self.lib.t3c_subword_stream(self.h, _p(w), nw, N, _p(out)))
        return out

    def build_words_from_subword_stream(self, trits, N, fill=0) -> np.ndarray:
        t = _u8(trits)
        out = np.zeros(((t.size + N - 1) // N + 1, 9), np.uint8)
        n = C.c_size_t()
        self._ck(self.lib.t3c_words_from_subword_stream(self.h, _p(t), t.size, N, fill, _p(out), C.byref(n)))
        return out[:n.value].copy()

    def ut_to_base243(self, trits) -> np.ndarray:
        t = _u8(trits)
        out = np.zeros(4 + (t.size + 4) // 5 + 8, np.uint8)
        n = C.c_size_t()
        self._ck(self.lib.t3c_base243_pack(self.h, _p(t), t.size, _p(out), C.byref(n)))
        return out[:n.value].copy()

    def base243_to_ut(self, data, cap=None):
        d = _u8(data)
        cap = 5 * d.size + 8 if cap is None else cap
        out = np.zeros(cap, np.uint8)
        n, ok = C.c_size_t(), C.c_int()
        self._ck(self.lib.t3c_base243_unpack(self.h, _p(d), d.size, _p(out), cap, C.byref(n), C.byref(ok)))
        return bool(ok.value), out[:min(n.value, cap)].copy()

    def words_to_base243(self, words, N) -> np.ndarray:
        w = _u8(words)
        nw = w.size // 9
        out = np.zeros(4 + (nw * N + 4) // 5 + 8, np.uint8)
        n = C.c_size_t()
        self._ck(self.lib.t3c_words_to_base243(self.h, _p(w), nw, N, _p(out), C.byref(n)))
        return out[:n.value].copy()

    def v6new_encode_raw_pixels_to_words(self, px, subword=0):
        """NEW-generation encode_raw_pixels_to_words[_subword]; returns (ok, words) like the reference's bool"""
        px = np.ascontiguousarray(px, dtype=PIXEL_DTYPE)
        out = np.zeros(px.size, np.uint32)
        st = self.lib.t3c_v6new_pack_pixels(self.h, _p(px), px.size, _p(out), subword)
        if st == ERR_ARG:
            return False, out
        self._ck(st)
        return True, out

    def v6new_decode_raw_words_to_pixels(self, words, subword=0):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        px = np.zeros(w.size, PIXEL_DTYPE)
        st = self.lib.t3c_v6new_unpack_pixels(self.h, _p(w), w.size, _p(px), subword)
        if st == ERR_ARG:
            return False, px
        self._ck(st)
        return True, px

    # --- image-bridge geometry and the NEW-generation image <-> words pipelines (include/io_image.hpp)
    def resize_rgb_nn(self, src, dw, dh):
        a = np.ascontiguousarray(src, dtype=np.uint8)
        out = np.zeros((dh, dw, 3), np.uint8)
        self._ck(self.lib.t3c_resize_rgb_nn(self.h, _p(a), a.shape[1], a.shape[0], _p(out), dw, dh))
        return out

    def blit_center_rgb(self, src, cw, ch):
        a = np.ascontiguousarray(src, dtype=np.uint8)
        out = np.zeros((ch, cw, 3), np.uint8)
        self._ck(self.lib.t3c_blit_center_rgb(self.h, _p(a), a.shape[1], a.shape[0], _p(out), cw, ch))
        return out

    def extract_center_q(self, full, fw, fh, sw, sh):
        f = np.ascontiguousarray(full, dtype=PIXEL_DTYPE)
        out = np.zeros(sw * sh, PIXEL_DTYPE)
        self._ck(self.lib.t3c_extract_center_q(self.h, _p(f), fw, fh, sw, sh, _p(out)))
        return out

    def v6new_image_to_words(self, rgb, sub, centered):
        a = np.ascontiguousarray(rgb, dtype=np.uint8)
        cap = 7680 * 4320
        out = np.zeros(cap, np.uint32)
        n, ok = C.c_size_t(), C.c_int()
        self._ck(self.lib.t3c_v6new_image_to_words(self.h, _p(a), a.shape[1], a.shape[0], sub, int(centered), _p(out), cap, C.byref(n), C.byref(ok)))
        return bool(ok.value), out[:n.value].copy()

    def v6new_words_to_image(self, words, sub, w, h):
        wd = np.ascontiguousarray(words, dtype=np.uint32)
        out = np.zeros((h, w, 3), np.uint8)
        ok = C.c_int()
        self._ck(self.lib.t3c_v6new_words_to_image(self.h, _p(wd), wd.size, sub, w, h, _p(out), C.byref(ok)))
        return bool(ok.value), out

    # --- .t3v container records (old/include/t3v_io.hpp)
    def crc32(self, data):
        d = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        out = C.c_uint32()
        self._ck(self.lib.t3c_crc32(self.h, _p(d), d.size, C.byref(out)))
        return int(out.value)

    def t3v_write_frame(self, words):
        w = np.ascontiguousarray(words, dtype=np.uint8).reshape(-1, 9)
        rec = np.zeros(8 + 9 * w.shape[0], np.uint8)
        n = C.c_size_t()
        self._ck(self.lib.t3c_t3v_frame_record(self.h, _p(w), w.shape[0], _p(rec), C.byref(n)))
        return rec[:n.value]

    def t3v_read_frame(self, rec):
        r = np.ascontiguousarray(rec, dtype=np.uint8).reshape(-1)
        cap = max(r.size, 9) // 9 + 1
        out = np.zeros((cap, 9), np.uint8)
        n, ok = C.c_size_t(), C.c_int()
        self._ck(self.lib.t3c_t3v_read_frame(self.h, _p(r), r.size, _p(out), cap, C.byref(n), C.byref(ok)))
        return bool(ok.value), out[:n.value].copy()

    def t3v_header(self, profile, subword_code, centered, coset, width, height, aw, fps_num=0, fps_den=1, frame_count=1, file_type=0):
        out = np.zeros(54, np.uint8)
        a = (C.c_uint32 * 4)(*aw)
        self._ck(self.lib.t3c_t3v_header(self.h, _p(out), profile, subword_code, int(centered), coset, width, height, a, fps_num, fps_den, frame_count, file_type))
        return out

    def t3v_frame_records_dev(self, d_words, n_words, stride_words, n_frames, d_records, record_pitch, stream=0):
        self._ck(self.lib.t3c_t3v_frame_records_dev(self.h, self._dp(d_words), n_words, stride_words, n_frames, self._dp(d_records), record_pitch, stream))

    def t3v_read_frames_dev(self, d_records, record_pitch, n_frames, n_words, d_words, stride_words, d_ok, stream=0):
        self._ck(self.lib.t3c_t3v_read_frames_dev(self.h, self._dp(d_records), record_pitch, n_frames, n_words, self._dp(d_words), stride_words, self._dp(d_ok), stream))

    def subword_stream_dev(self, d_words, n_words, N, d_trits, stream=0):
        self._ck(self.lib.t3c_subword_stream_dev(self.h, self._dp(d_words), n_words, N, self._dp(d_trits), stream))

    def words_to_base243_dev(self, d_words, n_words, N, d_out, stream=0):
        self._ck(self.lib.t3c_words_to_base243_dev(self.h, self._dp(d_words), n_words, N, self._dp(d_out), stream))

    def base243_pack_dev(self, d_trits, n_trits, d_out, stream=0):
        self._ck(self.lib.t3c_base243_pack_dev(self.h, self._dp(d_trits), n_trits, self._dp(d_out), stream))

    def base243_unpack_dev(self, d_payload, n_trits, d_trits, stream=0):
        self._ck(self.lib.t3c_base243_unpack_dev(self.h, self._dp(d_payload), n_trits, self._dp(d_trits), stream))

    def v6new_pack_pixels_dev(self, d_px, n_px, d_words, stream=0):
        self._ck(self.lib.t3c_v6new_pack_pixels_dev(self.h, self._dp(d_px), n_px, self._dp(d_words), stream))

    def v6new_unpack_pixels_dev(self, d_words, n_words, d_px, stream=0):
        self._ck(self.lib.t3c_v6new_unpack_pixels_dev(self.h, self._dp(d_words), n_words, self._dp(d_px), stream))

    # ---------------- K1 ----------------
    def rgb_to_quant_stream(self, rgb) -> np.ndarray:
        rgb = _u8(rgb)
        n = rgb.size // 3
        out = np.zeros(n, PIXEL_DTYPE)
        self._ck(self.lib.t3c_rgb_to_quant(self.h, _p(rgb), n, _p(out)))
        return out

    def quant_stream_to_rgb(self, px) -> np.ndarray:
        px = np.ascontiguousarray(px, dtype=PIXEL_DTYPE)
        out = np.zeros((px.size, 3), np.uint8)
        self._ck(self.lib.t3c_quant_to_rgb(self.h, _p(px), px.size, _p(out)))
        return out

    def encode_raw_pixels_to_words(self, px) -> np.ndarray:
        px = np.ascontiguousarray(px, dtype=PIXEL_DTYPE)
        out = np.zeros(((px.size + 1) // 2, 9), np.uint8)
        nw = C.c_size_t()
        self._ck(self.lib.t3c_pack_pixels(self.h, _p(px), px.size, _p(out), C.byref(nw)))
        return out[:nw.value]

    def decode_raw_words_to_pixels(self, words) -> np.ndarray:
        words = _u8(words)
        n = words.size // 9
        out = np.zeros(2 * n, PIXEL_DTYPE)
        self._ck(self.lib.t3c_unpack_pixels(self.h, _p(words), n, _p(out)))
        return out

    def words_to_bytes(self, words) -> np.ndarray:
        words = _u8(words)
        out = np.zeros(words.size, np.uint8)
        self._ck(self.lib.t3c_words_to_bytes(self.h, _p(words), words.size // 9, _p(out)))
        return out

    # ---------------- block codecs ----------------
    def rs_encode_blocks(self, k, data, arith=None) -> np.ndarray:
        data = _u8(data)
        n = data.size // k
        out = np.zeros((n, 26), np.uint8)
        self._ck(self.lib.t3c_rs_encode_blocks(self.h, k, self.arith if arith is None else arith, _p(data), n, _p(out)))
        return out

    def rs_decode_blocks(self, k, blocks, arith=None):
        io = np.array(blocks, dtype=np.uint8, copy=True).reshape(-1, 26)
        n = io.shape[0]
        out = np.zeros((n, k), np.uint8)
        ok = np.zeros(n, np.uint8)
        self._ck(self.lib.t3c_rs_decode_blocks(self.h, k, self.arith if arith is None else arith, _p(io), n, _p(out), _p(ok)))
        return io, out, ok

    def interleave2D_boustrophedon(self, syms, w, h, inverse=False) -> np.ndarray:
        a = np.array(syms, dtype=np.uint8, copy=True)
        self._ck(self.lib.t3c_interleave2d(self.h, _p(a), a.size, w, h, 1 if inverse else 0))
        return a

    def header_emit(self, cfg, arith=None):
        h27, c52 = np.zeros(27, np.uint8), np.zeros(52, np.uint8)
        self._ck(self.lib.t3c_header_emit(self.h, C.byref(cfg), self.arith if arith is None else arith, _p(h27), _p(c52)))
        return h27, c52

    def header_parse(self, words, arith=None):
        words = _u8(words)
        cfg = make_config()
        ok = C.c_int()
        self._ck(self.lib.t3c_header_parse(self.h, self.arith if arith is None else arith, _p(words), words.size // 9,
                                           C.byref(cfg), C.byref(ok)))
        return bool(ok.value), cfg

    # ---------------- profile codec ----------------
    def encode_profile_from_raw(self, raw_words, cfg=None, arith=None) -> np.ndarray:
        cfg = self.cfg if cfg is None else cfg
        raw = _u8(raw_words)
        n = raw.size // 9
        cap = profile_words(cfg, n)
        out = np.zeros((cap, 9), np.uint8)
        n_out = C.c_size_t()
        self._ck(self.lib.t3c_encode_profile(self.h, C.byref(cfg), self.arith if arith is None else arith, _p(raw), n,
                                             _p(out), cap, C.byref(n_out)))
        return out[:n_out.value]

    def decode_profile_to_raw(self, words):
        """The reference decoder as shipped; mutates ``self.cfg_last_seen``.  Returns (ok, words)."""
        w = _u8(words)
        n = w.size // 9
        out = np.zeros((n + 8, 9), np.uint8)
        n_out, ok = C.c_size_t(), C.c_int()
        self._ck(self.lib.t3c_decode_profile(self.h, C.byref(self.cfg_last_seen), _p(w), n, _p(out), n + 8, C.byref(n_out), C.byref(ok)))
        return bool(ok.value), out[:n_out.value].copy()

    def decode_profile_fixed(self, words, cfg=None, n_raw_words=0):
        """Consistent decoder (FIXED).  Returns (ok, words, n_corrected)."""
        cfg = self.cfg if cfg is None else cfg
        w = _u8(words)
        n = w.size // 9
        out = np.zeros((n + 8, 9), np.uint8)
        n_out, ok, nc = C.c_size_t(), C.c_int(), C.c_size_t()
        self._ck(self.lib.t3c_decode_profile_fixed(self.h, C.byref(cfg), n_raw_words, _p(w), n, _p(out), n + 8, C.byref(n_out),
                                                   C.byref(ok), C.byref(nc)))
        return bool(ok.value), out[:n_out.value].copy(), nc.value

    # ---------------- fused frames ----------------
    def encode_frames_rgb8(self, rgb_frames, cfg=None, arith=None) -> np.ndarray:
        """rgb_frames: [F, n_px, 3] uint8 -> [F, words_per_frame, 9]."""
        cfg = self.cfg if cfg is None else cfg
        rgb = _u8(rgb_frames)
        if rgb.ndim == 2:
            rgb = rgb[None]
        F, n_px = rgb.shape[0], rgb.shape[1]
        wpf = profile_words(cfg, (n_px + 1) // 2)
        out = np.zeros((F, wpf, 9), np.uint8)
        got = C.c_size_t()
        self._ck(self.lib.t3c_encode_frames_rgb8(self.h, C.byref(cfg), self.arith if arith is None else arith, _p(rgb), n_px, F,
                                                 _p(out), wpf, C.byref(got)))
        assert got.value == wpf
        return out

    def decode_frames_rgb8(self, words_frames, n_px, cfg=None):
        """[F, words_per_frame, 9] -> (ok[F], rgb[F, px_recovered, 3], n_corrected)."""
        cfg = self.cfg if cfg is None else cfg
        w = _u8(words_frames)
        if w.ndim == 2:
            w = w[None]
        F, wpf = w.shape[0], w.shape[1]
        rgb = np.zeros((F, n_px, 3), np.uint8)
        ok = np.zeros(F, np.uint8)
        rec, nc = C.c_size_t(), C.c_size_t()
        self._ck(self.lib.t3c_decode_frames_rgb8(self.h, C.byref(cfg), _p(w), wpf, wpf, F, n_px, _p(rgb), _p(ok), C.byref(rec), C.byref(nc)))
        return ok.astype(bool), rgb[:, :rec.value].copy(), nc.value

    # ---------------- device-pointer calls (torch tensors / raw pointers) ----------------
    @staticmethod
    def _dp(t) -> int:
        return t if isinstance(t, int) else t.data_ptr()

    def encode_frames_rgb8_dev(self, d_rgb, n_px, n_frames, d_out, stride_words, cfg=None, arith=None, stream=0):
        cfg = self.cfg if cfg is None else cfg
        self._ck(self.lib.t3c_encode_frames_rgb8_dev(self.h, C.byref(cfg), self.arith if arith is None else arith, self._dp(d_rgb), n_px,
                                                     n_frames, self._dp(d_out), stride_words, stream))

    def decode_frames_rgb8_dev(self, d_in, words_per_frame, stride_words, n_frames, n_px, d_rgb, d_status, cfg=None, stream=0):
        cfg = self.cfg if cfg is None else cfg
        self._ck(self.lib.t3c_decode_frames_rgb8_dev(self.h, C.byref(cfg), self._dp(d_in), words_per_frame, stride_words, n_frames, n_px,
                                                     self._dp(d_rgb), self._dp(d_status), stream))

    def pack_pixels_dev(self, d_px, n_px, d_words, stream=0):
        self._ck(self.lib.t3c_pack_pixels_dev(self.h, self._dp(d_px), n_px, self._dp(d_words), stream))

    def unpack_pixels_dev(self, d_words, n_words, d_px, stream=0):
        self._ck(self.lib.t3c_unpack_pixels_dev(self.h, self._dp(d_words), n_words, self._dp(d_px), stream))

    def rgb_to_quant_dev(self, d_rgb, n_px, d_px, stream=0):
        self._ck(self.lib.t3c_rgb_to_quant_dev(self.h, self._dp(d_rgb), n_px, self._dp(d_px), stream))

    def quant_to_rgb_dev(self, d_px, n_px, d_rgb, stream=0):
        self._ck(self.lib.t3c_quant_to_rgb_dev(self.h, self._dp(d_px), n_px, self._dp(d_rgb), stream))

    def encode_profile_dev(self, d_raw, n_words, d_out, cap_words, cfg=None, arith=None, stream=0):
        cfg = self.cfg if cfg is None else cfg
        self._ck(self.lib.t3c_encode_profile_dev(self.h, C.byref(cfg), self.arith if arith is None else arith, self._dp(d_raw), n_words,
                                                 self._dp(d_out), cap_words, stream))

    def decode_profile_fixed_dev(self, d_in, n_words, n_raw_words, d_out, cap_words, d_status, cfg=None, stream=0):
        cfg = self.cfg if cfg is None else cfg
        self._ck(self.lib.t3c_decode_profile_fixed_dev(self.h, C.byref(cfg), n_raw_words, self._dp(d_in), n_words, self._dp(d_out),
                                                       cap_words, self._dp(d_status), stream))

    def rs_encode_blocks_dev(self, k, d_data, n, d_out, arith=None, stream=0):
        self._ck(self.lib.t3c_rs_encode_blocks_dev(self.h, k, self.arith if arith is None else arith, self._dp(d_data), n, self._dp(d_out), stream))

    def rs_decode_blocks_dev(self, k, d_inout, n, d_out, d_ok, arith=None, stream=0):
        self._ck(self.lib.t3c_rs_decode_blocks_dev(self.h, k, self.arith if arith is None else arith, self._dp(d_inout), n, self._dp(d_out),
                                                   self._dp(d_ok), stream))


class Stream:
    """A multi-device stream (t3c_stream_*, include/t3c.h): one context and one host thread per lane, frame f -> lane f mod n_lanes,
    results in frame order.  ``devices`` may list a device more than once (several lanes on one GPU)."""

    def __init__(self, devices):
        self.lib = load_library()
        self.h = C.c_void_p()
        devs = (C.c_int32 * len(devices))(*devices)
        st = self.lib.t3c_stream_create(devs, len(devices), C.byref(self.h))
        if st == ERR_NODEVICE:
            raise T3CError("t3c_stream_create: no CUDA device -- this library has no CPU fallback")
        if st != OK:
            raise T3CError(f"t3c_stream_create failed with status {st}")

    @property
    def lanes(self):
        return int(self.lib.t3c_stream_lanes(self.h))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.t3c_stream_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode_rgb8(self, rgb: np.ndarray, cfg: Config, arith: int, first_frame: int = 0, out: np.ndarray | None = None):
        """rgb: (n_frames, n_px, 3) uint8 -> (n_frames, words_per_frame, 9) uint8"""
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        n_frames, n_px = rgb.shape[0], rgb.shape[1]
        wpf = profile_words(cfg, (n_px + 1) // 2)
        if out is None:
            out = np.empty((n_frames, wpf, 9), np.uint8)
        got = C.c_size_t()
        st = self.lib.t3c_stream_encode_rgb8(self.h, C.byref(cfg), arith, rgb.ctypes.data, n_px, n_frames, first_frame, out.ctypes.data, out.shape[1], C.byref(got))
        if st != OK:
            raise T3CError(f"t3c_stream_encode_rgb8: status {st}")
        return out[:, :got.value]

    def decode_rgb8(self, enc: np.ndarray, n_px: int, cfg: Config, first_frame: int = 0):
        """enc: (n_frames, words_per_frame, 9) -> (ok[n_frames], rgb (n_frames, n_px, 3), n_corrected)"""
        enc = np.ascontiguousarray(enc, dtype=np.uint8)
        n_frames, wpf = enc.shape[0], enc.shape[1]
        rgb = np.empty((n_frames, n_px, 3), np.uint8)
        ok = np.zeros(n_frames, np.uint8)
        nc = C.c_size_t()
        st = self.lib.t3c_stream_decode_rgb8(self.h, C.byref(cfg), enc.ctypes.data, wpf, wpf, n_frames, first_frame, n_px, rgb.ctypes.data, ok.ctypes.data, C.byref(nc))
        if st != OK:
            raise T3CError(f"t3c_stream_decode_rgb8: status {st}")
        return ok, rgb, int(nc.value)

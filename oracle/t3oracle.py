"""ctypes front-end for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Loads ``oracle/libt3oracle.so`` (plain-C restatement, ``Oracle``) and, when present,
``oracle/_ref/libt3ref{,_fixed}.so`` (the reference itself, ``Reference``).  Only tests/,
``bench.py``'s cpu_baseline / ``--impl reference`` legs and ``__graft_entry__.smoke()``
may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
RAW_MODE = 0xFF
P1, P2, P3, P4, P5 = 0, 1, 2, 3, 4
K_OF_UEP = (24, 22, 20, 18)


class Cfg(C.Structure):
    """Field-for-field mirror of ``t3o_cfg`` (= EncoderConfig, OLD:862-873)."""
    _fields_ = [
        ("profile", C.c_uint8), ("uep", C.c_uint8 * 9),
        ("tile_w", C.c_uint16), ("tile_h", C.c_uint16),
        ("seed_a", C.c_uint32), ("seed_b", C.c_uint32), ("seed_s0", C.c_uint32),
        ("beacon_period", C.c_uint32), ("beacon_slot", C.c_uint8), ("beacon_enabled", C.c_uint8),
        ("subword", C.c_uint8), ("centered", C.c_uint8), ("coset", C.c_uint8), ("pad_", C.c_uint8 * 3),
        ("superframe_words", C.c_uint32),
    ]

    def copy(self) -> "Cfg":
        c = Cfg()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(Cfg))
        return c

    def astuple(self):
        return (self.profile, tuple(self.uep), self.tile_w, self.tile_h, self.seed_a, self.seed_b, self.seed_s0,
                self.beacon_period, self.beacon_slot, self.beacon_enabled, self.subword, self.centered, self.coset)


def make_cfg(profile=P2, uep=1, tile=(0, 0), seed=(1, 1, 1), beacon=(0, 0, False), superframe_words=8192,
             subword=27, centered=True, coset=0) -> Cfg:
    """EncoderConfig with the reference defaults (EncoderContext(): uniform UEP index 1 => k=22)."""
    c = Cfg()
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


UEP_LUMA = (2, 1, 1, 2, 1, 1, 2, 1, 1)  # uep_luma_priority, OLD:68-72

PIXEL_DTYPE = np.dtype([("Yq", "<u2"), ("Cbq", "<i2"), ("Crq", "<i2")])


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def build(force: bool = False) -> None:
    """Compile the C restatement (always) and the reference shim (when /root/reference exists)."""
    if force or not os.path.exists(os.path.join(HERE, "libt3oracle.so")) or \
            os.path.getmtime(os.path.join(HERE, "libt3oracle.so")) < os.path.getmtime(os.path.join(HERE, "t3_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.exists("/root/reference/old/include/ternary_image_codec_v6_min.hpp"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


class _Lib:
    """Common numpy-level API over either library; ``pfx`` is the symbol prefix (t3o_/t3r_)."""

    def __init__(self, path: str, pfx: str, fixed: int | None):
        self.lib = C.CDLL(path)
        self.pfx = pfx
        self.fixed = fixed  # None: the library takes a `fixed` argument (oracle); else baked in (reference builds)
        L = self.lib
        sz, u8p = C.c_size_t, C.POINTER(C.c_uint8)
        for name in ("encode_profile", "pack_pixels", "profile_words_bound", "bytes_to_words", "encode_rgb", "words_from_subword_stream", "base243_pack",
                     "t3v_frame_record"):
            if hasattr(L, pfx + name):
                getattr(L, pfx + name).restype = sz
        if hasattr(L, pfx + "crc32"):
            getattr(L, pfx + "crc32").restype = C.c_uint32
        for name in ("gf_add", "gf_sub", "gf_mul", "gf_inv", "gf_pow_alpha", "scramble_symbol", "descramble_symbol", "beacon_symbol"):
            if hasattr(L, pfx + name):
                getattr(L, pfx + name).restype = C.c_uint8
        del u8p

    def f(self, name):
        return getattr(self.lib, self.pfx + name)

    # --- RS ---
    def rs_gen(self, k):
        g = np.zeros(9, np.uint8)
        n = self.f("rs_gen")(C.c_int(k), _ptr(g))
        return g[:n].copy()

    def rs_encode_blocks(self, k, data, fixed=0):
        data, dp = _u8(data)
        n = data.size // k
        out = np.zeros((n, 26), np.uint8)
        if self.fixed is None:
            self.f("rs_encode_blocks")(C.c_int(k), C.c_int(fixed), dp, C.c_size_t(n), _ptr(out))
        else:
            assert fixed == self.fixed
            self.f("rs_encode_blocks")(C.c_int(k), dp, C.c_size_t(n), _ptr(out))
        return out

    def rs_decode_blocks(self, k, blocks, fixed=0):
        """returns (corrected inout [n,26], out [n,k] (zeros when !ok), ok [n])"""
        io = np.array(blocks, dtype=np.uint8, copy=True).reshape(-1, 26)
        n = io.shape[0]
        out = np.zeros((n, k), np.uint8)
        ok = np.zeros(n, np.uint8)
        if self.fixed is None:
            self.f("rs_decode_blocks")(C.c_int(k), C.c_int(fixed), _ptr(io), C.c_size_t(n), _ptr(out), _ptr(ok))
        else:
            assert fixed == self.fixed
            self.f("rs_decode_blocks")(C.c_int(k), _ptr(io), C.c_size_t(n), _ptr(out), _ptr(ok))
        return io, out, ok

    # --- header ---
    def crc12(self, trits):
        t, tp = _u8(trits)
        out = np.zeros(12, np.uint8)
        self.f("crc12")(tp, C.c_size_t(t.size), _ptr(out))
        return out

    def header_pack(self, cfg, frame_seq=0, band_map_hash=0):
        out = np.zeros(27, np.uint8)
        self.f("header_pack")(C.byref(cfg), C.c_uint32(frame_seq), C.c_uint32(band_map_hash), _ptr(out))
        return out

    def header_check(self, sym27):
        s, sp = _u8(sym27)
        return bool(self.f("header_check")(sp))

    def header_unpack(self, sym27):
        s, sp = _u8(sym27)
        cfg = Cfg()
        fs, bh, mg, ver = C.c_uint32(), C.c_uint32(), C.c_uint16(), C.c_uint8()
        self.f("header_unpack")(sp, C.byref(cfg), C.byref(fs), C.byref(bh), C.byref(mg), C.byref(ver))
        return cfg, fs.value, bh.value, mg.value, ver.value

    # --- pixels ---
    def pack_pixels(self, px):
        px = np.ascontiguousarray(px, dtype=PIXEL_DTYPE)
        nw = (px.size + 1) // 2
        out = np.zeros((nw, 9), np.uint8)
        self.f("pack_pixels")(px.ctypes.data_as(C.c_void_p), C.c_size_t(px.size), _ptr(out))
        return out

    def unpack_pixels(self, words):
        w, wp = _u8(words)
        nw = w.size // 9
        px = np.zeros(2 * nw, PIXEL_DTYPE)
        self.f("unpack_pixels")(wp, C.c_size_t(nw), px.ctypes.data_as(C.c_void_p))
        return px

    # --- sub-word streams + base-243 (same symbols in the oracle and the reference shim) ---
    def subword_stream(self, words, N):
        w, wp = _u8(words)
        nw = w.size // 9
        out = np.zeros(nw * N, np.uint8)
        self.f("subword_stream")(wp, C.c_size_t(nw), C.c_int(N), _ptr(out))
        return out

    def words_from_subword_stream(self, trits, N, fill=0):
        t, tp = _u8(trits)
        out = np.zeros(((t.size + N - 1) // N + 1, 9), np.uint8)
        n = self.f("words_from_subword_stream")(tp, C.c_size_t(t.size), C.c_int(N), C.c_uint8(fill), _ptr(out))
        return out[:n].copy()

    def base243_pack(self, trits):
        t, tp = _u8(trits)
        out = np.zeros(4 + (t.size + 4) // 5 + 8, np.uint8)
        n = self.f("base243_pack")(tp, C.c_size_t(t.size), _ptr(out))
        return out[:n].copy()

    def base243_unpack(self, data, cap=None):
        d, dp = _u8(data)
        cap = 5 * d.size + 8 if cap is None else cap
        out = np.zeros(cap, np.uint8)
        n = C.c_size_t()
        ok = self.f("base243_unpack")(dp, C.c_size_t(d.size), _ptr(out), C.c_size_t(cap), C.byref(n))
        return bool(ok), out[:min(n.value, cap)].copy()

    # --- image-bridge geometry (oracle only; the reference's live in ReferenceNew) ---
    def resize_rgb_nn(self, src, dw, dh):
        a = np.ascontiguousarray(src, np.uint8)
        sh, sw = a.shape[0], a.shape[1]
        out = np.zeros((dh, dw, 3), np.uint8)
        self.lib.t3o_resize_rgb_nn(_ptr(a), C.c_int(sw), C.c_int(sh), _ptr(out), C.c_int(dw), C.c_int(dh))
        return out

    def blit_center_rgb(self, src, cw, ch):
        a = np.ascontiguousarray(src, np.uint8)
        out = np.zeros((ch, cw, 3), np.uint8)
        self.lib.t3o_blit_center_rgb(_ptr(a), C.c_int(a.shape[1]), C.c_int(a.shape[0]), _ptr(out), C.c_int(cw), C.c_int(ch))
        return out

    def extract_center_q(self, full, fw, fh, sw, sh):
        f = np.ascontiguousarray(full, PIXEL_DTYPE)
        out = np.zeros(sw * sh, PIXEL_DTYPE)
        self.lib.t3o_extract_center_q(f.ctypes.data_as(C.c_void_p), C.c_int(fw), C.c_int(fh), C.c_int(sw), C.c_int(sh), out.ctypes.data_as(C.c_void_p))
        return out

    def v6new_image_to_words(self, rgb, sub, centered):
        """image_to_words_subword after the load (include/io_image.hpp:238-301), composed from the restated pieces"""
        if sub not in V6NEW_STD_RES or rgb.size == 0:
            return False, np.zeros(0, np.uint32)
        tw, th = V6NEW_STD_RES[sub]
        work = rgb if (rgb.shape[1], rgb.shape[0]) == (tw, th) else self.resize_rgb_nn(rgb, tw, th)
        if centered and sub != 27:
            work = self.blit_center_rgb(work, 7680, 4320)
        return True, self.v6new_pack_pixels(self.rgb_to_quant(work.reshape(-1, 3)))

    def v6new_words_to_image(self, words, sub, w, h):
        """words_to_image_subword before the write (:304-338)"""
        if sub not in V6NEW_STD_RES:
            return False, np.zeros((h, w, 3), np.uint8)
        q = self.v6new_unpack_pixels(words)
        tw, th = V6NEW_STD_RES[sub]
        if q.size != w * h and q.size == 7680 * 4320 and sub != 27:
            q = self.extract_center_q(q, 7680, 4320, tw, th)
        out = np.zeros((h * w, 3), np.uint8)
        n = min(q.size, w * h)
        if n:
            out[:n] = self.quant_to_rgb(q[:n])
        return True, out.reshape(h, w, 3)

    # --- .t3v container records (same symbols in the oracle and the reference shim) ---
    def crc32(self, data):
        d, dp = _u8(data)
        return int(self.f("crc32")(dp, C.c_size_t(d.size)))

    def t3v_frame_record(self, words):
        w, wp = _u8(words)
        nw = w.size // 9
        out = np.zeros(8 + 9 * nw, np.uint8)
        n = self.f("t3v_frame_record")(wp, C.c_uint32(nw), _ptr(out))
        return out[:n].copy()

    def t3v_read_frame(self, rec):
        r, rp = _u8(rec)
        out = np.zeros((max(r.size, 9) // 9 + 1, 9), np.uint8)
        n = C.c_uint32()
        ok = self.f("t3v_read_frame")(rp, C.c_size_t(r.size), _ptr(out), C.byref(n))
        return bool(ok), out[:n.value].copy()

    def rgb_to_quant(self, rgb):
        r, rp = _u8(rgb)
        n = r.size // 3
        px = np.zeros(n, PIXEL_DTYPE)
        self.f("rgb_to_quant")(rp, C.c_size_t(n), px.ctypes.data_as(C.c_void_p))
        return px

    def quant_to_rgb(self, px):
        px = np.ascontiguousarray(px, dtype=PIXEL_DTYPE)
        out = np.zeros((px.size, 3), np.uint8)
        self.f("quant_to_rgb")(px.ctypes.data_as(C.c_void_p), C.c_size_t(px.size), _ptr(out))
        return out

    # --- permutation / scrambler ---
    def interleave2d(self, sy, w, h, inverse=False):
        a = np.array(sy, dtype=np.uint8, copy=True)
        self.f("deinterleave2d" if inverse else "interleave2d")(_ptr(a), C.c_size_t(a.size), C.c_uint(w), C.c_uint(h))
        return a

    def scramble_stream(self, sy, a, b, s0, inverse=False):
        out = np.zeros(len(sy), np.uint8)
        st = C.c_uint32(s0 % 3)
        fn = self.f("descramble_symbol" if inverse else "scramble_symbol")
        for i, s in enumerate(sy):
            out[i] = fn(C.c_uint8(int(s)), C.c_uint32(a), C.c_uint32(b), C.byref(st))
        return out

    def beacon_symbol(self, profile, seq, health=0):
        return int(self.f("beacon_symbol")(C.c_uint8(profile), C.c_uint16(seq), C.c_uint8(health)))

    # --- profile codec ---
    def encode_profile(self, cfg, raw_words, fixed=0, cap=None):
        raw, rp = _u8(raw_words)
        n = raw.size // 9
        if cap is None:
            cap = 2 * n + 64
        out = np.zeros((cap, 9), np.uint8)
        if self.fixed is None:
            r = self.f("encode_profile")(C.byref(cfg), C.c_int(fixed), rp, C.c_size_t(n), _ptr(out), C.c_size_t(cap))
        else:
            assert fixed == self.fixed
            r = self.f("encode_profile")(C.byref(cfg), rp, C.c_size_t(n), _ptr(out), C.c_size_t(cap))
        assert r != C.c_size_t(-1).value, "capacity too small"
        return out[:r].copy()

    def encode_rgb(self, cfg, rgb, fixed=0):
        r_, rp = _u8(rgb)
        n = r_.size // 3
        cap = n + 64
        out = np.zeros((cap, 9), np.uint8)
        if self.fixed is None:
            r = self.f("encode_rgb")(C.byref(cfg), C.c_int(fixed), rp, C.c_size_t(n), _ptr(out), C.c_size_t(cap))
        else:
            assert fixed == self.fixed
            r = self.f("encode_rgb")(C.byref(cfg), rp, C.c_size_t(n), _ptr(out), C.c_size_t(cap))
        return out[:r].copy()


class Oracle(_Lib):
    def __init__(self):
        build()
        super().__init__(os.path.join(HERE, "libt3oracle.so"), "t3o_", None)
        self.lib.t3o_gf_log.restype = C.c_int

    def v6new_pack_pixels(self, px):
        px = np.ascontiguousarray(px, dtype=PIXEL_DTYPE)
        out = np.zeros(px.size, np.uint32)
        self.lib.t3o_v6new_pack_pixels(px.ctypes.data_as(C.c_void_p), C.c_size_t(px.size), out.ctypes.data_as(C.c_void_p))
        return out

    def v6new_unpack_pixels(self, words):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        px = np.zeros(w.size, PIXEL_DTYPE)
        self.lib.t3o_v6new_unpack_pixels(w.ctypes.data_as(C.c_void_p), C.c_size_t(w.size), px.ctypes.data_as(C.c_void_p))
        return px

    def words_bound(self, cfg, n_words):
        return int(self.lib.t3o_profile_words_bound(C.byref(cfg), C.c_size_t(n_words)))

    def decode_profile_ref(self, seen, words):
        """returns (ok, out_words, seen_after)"""
        w, wp = _u8(words)
        n = w.size // 9
        seen = seen.copy()
        out = np.zeros((n + 8, 9), np.uint8)
        n_out = C.c_size_t()
        ok = self.lib.t3o_decode_profile_ref(C.byref(seen), wp, C.c_size_t(n), _ptr(out), C.c_size_t(n + 8), C.byref(n_out))
        return bool(ok), out[:n_out.value].copy(), seen

    def decode_profile_fixed(self, cfg, words, n_raw_words=0):
        """returns (ok, out_words, n_corrected)"""
        w, wp = _u8(words)
        n = w.size // 9
        out = np.zeros((n + 8, 9), np.uint8)
        n_out, ncorr = C.c_size_t(), C.c_size_t()
        ok = self.lib.t3o_decode_profile_fixed(C.byref(cfg), C.c_size_t(n_raw_words), wp, C.c_size_t(n), _ptr(out),
                                               C.c_size_t(n + 8), C.byref(n_out), C.byref(ncorr))
        return bool(ok), out[:n_out.value].copy(), ncorr.value

    def decode_rgb_fixed(self, cfg, words, n_px):
        w, wp = _u8(words)
        n = w.size // 9
        rgb = np.zeros((n_px, 3), np.uint8)
        npx, ncorr = C.c_size_t(), C.c_size_t()
        ok = self.lib.t3o_decode_rgb_fixed(C.byref(cfg), C.c_size_t(n_px), wp, C.c_size_t(n), _ptr(rgb), C.byref(npx), C.byref(ncorr))
        return bool(ok), rgb[:npx.value].copy(), ncorr.value

    def gf_tables(self):
        L = self.lib
        exp = np.array([L.t3o_gf_pow_alpha(C.c_int(i)) for i in range(78)], np.uint8)
        log = np.array([L.t3o_gf_log(C.c_uint8(i)) for i in range(27)], np.int16)
        mul = np.array([[L.t3o_gf_mul(C.c_uint8(a), C.c_uint8(b)) for b in range(27)] for a in range(27)], np.uint8).reshape(-1)
        inv = np.array([L.t3o_gf_inv(C.c_uint8(i)) for i in range(27)], np.uint8)
        return exp, log, mul, inv


class Reference(_Lib):
    """The reference itself (oracle/_ref/, built from /root/reference); fixed=False: as shipped."""

    def __init__(self, fixed: bool = False):
        build()
        path = os.path.join(HERE, "_ref", "libt3ref_fixed.so" if fixed else "libt3ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        super().__init__(path, "t3r_", 1 if fixed else 0)

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", "libt3ref.so")) or \
            os.path.exists("/root/reference/old/include/ternary_image_codec_v6_min.hpp")

    def decode_profile_ref(self, seen, words):
        w, wp = _u8(words)
        n = w.size // 9
        seen = seen.copy()
        out = np.zeros((n + 8, 9), np.uint8)
        n_out = C.c_size_t()
        ok = self.lib.t3r_decode_profile(C.byref(seen), wp, C.c_size_t(n), _ptr(out), C.c_size_t(n + 8), C.byref(n_out))
        return bool(ok), out[:n_out.value].copy(), seen

    def gf_tables(self):
        exp, log, mul, inv = np.zeros(78, np.uint8), np.zeros(27, np.int16), np.zeros(729, np.uint8), np.zeros(27, np.uint8)
        prim = C.c_uint8()
        self.lib.t3r_gf_tables(_ptr(exp), log.ctypes.data_as(C.c_void_p), _ptr(mul), _ptr(inv), C.byref(prim))
        return exp, log, mul, inv

    def selftests(self):
        a, b = C.c_int(), C.c_int()
        self.lib.t3r_selftests(C.byref(a), C.byref(b))
        return bool(a.value), bool(b.value)

    def words_to_bytes(self, words):
        w, wp = _u8(words)
        out = np.zeros(w.size, np.uint8)
        self.lib.t3r_words_to_bytes(wp, C.c_size_t(w.size // 9), _ptr(out))
        return out

    def bytes_to_words(self, b):
        b, bp = _u8(b)
        out = np.zeros((b.size // 9 + 1, 9), np.uint8)
        n = self.lib.t3r_bytes_to_words(bp, C.c_size_t(b.size), _ptr(out))
        return out[:n].copy()


V6NEW_STD_RES = {27: (7680, 4320), 24: (3840, 2160), 21: (1920, 1080), 18: (1280, 720), 15: (960, 540)}   # std_res_for, NEW:55-64


class ReferenceNew:
    """The reference's NEW-generation core (one pixel -> one 32-bit word), oracle/_ref/libt3ref_new.so; SURVEY 8(f).3."""

    def __init__(self):
        build()
        path = os.path.join(HERE, "_ref", "libt3ref_new.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)

    def pack_pixels(self, px, subword=0):
        px = np.ascontiguousarray(px, dtype=PIXEL_DTYPE)
        out = np.zeros(px.size, np.uint32)
        ok = self.lib.t3n_pack_pixels(px.ctypes.data_as(C.c_void_p), C.c_size_t(px.size), out.ctypes.data_as(C.c_void_p), C.c_int(subword))
        return bool(ok), out

    def unpack_pixels(self, words, subword=0):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        px = np.zeros(w.size, PIXEL_DTYPE)
        ok = self.lib.t3n_unpack_pixels(w.ctypes.data_as(C.c_void_p), C.c_size_t(w.size), px.ctypes.data_as(C.c_void_p), C.c_int(subword))
        return bool(ok), px

    # --- the image bridge (include/io_image.hpp) through in-memory stb stubs
    def resize_rgb_nn(self, src, dw, dh):
        a = np.ascontiguousarray(src, np.uint8)
        out = np.zeros((dh, dw, 3), np.uint8)
        self.lib.t3n_resize_rgb_nn(a.ctypes.data_as(C.c_void_p), C.c_int(a.shape[1]), C.c_int(a.shape[0]), out.ctypes.data_as(C.c_void_p), C.c_int(dw), C.c_int(dh))
        return out

    def blit_center_rgb(self, src, cw, ch):
        a = np.ascontiguousarray(src, np.uint8)
        out = np.zeros((ch, cw, 3), np.uint8)
        self.lib.t3n_blit_center_rgb(a.ctypes.data_as(C.c_void_p), C.c_int(a.shape[1]), C.c_int(a.shape[0]), out.ctypes.data_as(C.c_void_p), C.c_int(cw), C.c_int(ch))
        return out

    def extract_center_q(self, full, fw, fh, sw, sh):
        f = np.ascontiguousarray(full, PIXEL_DTYPE)
        out = np.zeros(sw * sh, PIXEL_DTYPE)
        self.lib.t3n_extract_center_q.restype = C.c_size_t
        n = self.lib.t3n_extract_center_q(f.ctypes.data_as(C.c_void_p), C.c_int(fw), C.c_int(fh), C.c_int(sw), C.c_int(sh), out.ctypes.data_as(C.c_void_p))
        return out[:n]

    def image_to_words_subword(self, rgb, sub, centered):
        a = np.ascontiguousarray(rgb, np.uint8)
        cap = 7680 * 4320
        out = np.zeros(cap, np.uint32)
        self.lib.t3n_image_to_words_subword.restype = C.c_longlong
        n = self.lib.t3n_image_to_words_subword(a.ctypes.data_as(C.c_void_p), C.c_int(a.shape[1]), C.c_int(a.shape[0]), C.c_int(sub), C.c_int(int(centered)),
                                                out.ctypes.data_as(C.c_void_p), C.c_size_t(cap))
        return (n >= 0), out[:max(n, 0)].copy()

    def words_to_image_subword(self, words, sub, w, h):
        wd = np.ascontiguousarray(words, np.uint32)
        out = np.zeros((h, w, 3), np.uint8)
        ok = self.lib.t3n_words_to_image_subword(wd.ctypes.data_as(C.c_void_p), C.c_size_t(wd.size), C.c_int(sub), C.c_int(w), C.c_int(h), out.ctypes.data_as(C.c_void_p))
        return bool(ok), out



# --------------------------------------------------------------------------------------
# Deterministic synthetic inputs (counter based: identical on CPU and GPU), SURVEY 8(d).
# --------------------------------------------------------------------------------------
def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def h(seed: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        return splitmix64(np.uint64(seed) * np.uint64(0xD1342543DE82EF95) + idx.astype(np.uint64))


def synth_rgb(seed: int, n_px: int) -> np.ndarray:
    return (h(seed, np.arange(3 * n_px, dtype=np.uint64)) & np.uint64(0xFF)).astype(np.uint8).reshape(n_px, 3)


def synth_checker(w: int, hgt: int, cell: int = 8) -> np.ndarray:
    """8x8 checkerboard of (32,200,64)/(200,32,220), src/minitest_codec.cpp:31-42."""
    yy, xx = np.mgrid[0:hgt, 0:w]
    m = ((xx // cell + yy // cell) % 2).astype(bool)
    out = np.empty((hgt, w, 3), np.uint8)
    out[~m] = (32, 200, 64)
    out[m] = (200, 32, 220)
    return out.reshape(-1, 3)


def synth_quant(seed: int, n_px: int) -> np.ndarray:
    r = h(seed, np.arange(3 * n_px, dtype=np.uint64)).reshape(n_px, 3)
    px = np.zeros(n_px, PIXEL_DTYPE)
    px["Yq"] = (r[:, 0] % np.uint64(243)).astype(np.uint16)
    px["Cbq"] = (r[:, 1] % np.uint64(81)).astype(np.int16) - 40
    px["Crq"] = (r[:, 2] % np.uint64(81)).astype(np.int16) - 40
    return px


def inject_errors(words: np.ndarray, cfg: Cfg, n_raw_words: int, seed: int, gf_add, exact_t: bool = False):
    """Corrupt e_c <= t_b symbols in every body codeword of an encoder output (no beacon/any k).

    Positions and magnitudes are counter based: for codeword c, e_c = h(seed,c) % (t+1) (or t when
    exact_t), distinct positions, symbol replaced by gf_add(sym, 1 + h%26).  Works on the wire format
    (52 header symbols, band-major body, optional beacon expansion).  Returns (corrupted, n_errors).
    """
    flat = np.array(words, dtype=np.uint8, copy=True).reshape(-1)
    n_s = (26 * n_raw_words + 2) // 3
    period, slot = cfg.beacon_period, cfg.beacon_slot
    has_b = bool(cfg.beacon_enabled and period > 0 and slot < 9)
    cw_index = 0
    total = 0
    for b in range(9):
        k = K_OF_UEP[cfg.uep[b] % 4]
        t = (26 - k) // 2
        s_b = (n_s - b + 8) // 9 if n_s > b else 0
        ncw = s_b // k
        if ncw == 0:
            continue
        c_idx = np.arange(cw_index, cw_index + ncw, dtype=np.uint64)
        e = np.full(ncw, t, np.int64) if exact_t else (h(seed, c_idx) % np.uint64(t + 1)).astype(np.int64)
        for j in range(t):
            sel = e > j
            if not sel.any():
                continue
            # distinct positions: j-th element of a per-codeword pseudo-random permutation start/stride
            start = (h(seed + 1, c_idx) % np.uint64(26)).astype(np.int64)
            stride = np.array([1, 3, 5, 7, 9, 11, 15, 17, 19, 21, 23, 25], np.int64)[(h(seed + 2, c_idx) % np.uint64(12)).astype(np.int64)]
            pos = (start + j * stride) % 26
            mag = (h(seed + 3 + j, c_idx) % np.uint64(26)).astype(np.int64) + 1
            p = 26 * c_idx.astype(np.int64) + pos  # pre-beacon body index
            if has_b:
                q = beacon_expand_index(p, period, slot)
            else:
                q = p
            idx = (52 + q)[sel]
            flat[idx] = gf_add[flat[idx].astype(np.int64) % 27, mag[sel]]
            total += int(sel.sum())
        cw_index += ncw
    return flat.reshape(-1, 9), total


def beacon_expand_index(p, period, slot):
    """pre-beacon body index p -> index q in the beacon-expanded body (A.5): every `period` words one
    word carries the beacon in `slot` and only 8 body symbols."""
    per = 9 * period - 1
    blk, rem = p // per, p % per
    first = rem < 8
    wordoff = np.where(first, 0, 1 + (rem - 8) // 9)
    slotidx = np.where(first, np.where(rem < slot, rem, rem + 1), (rem - 8) % 9)
    return 9 * (blk * period + wordoff) + slotidx


def gf_add_table() -> np.ndarray:
    a = np.arange(27)
    t = np.zeros((27, 27), np.uint8)
    for x in a:
        for y in a:
            t[x, y] = ((x % 3 + y % 3) % 3) + 3 * (((x // 3) % 3 + (y // 3) % 3) % 3) + 9 * (((x // 9) + (y // 9)) % 3)
    return t

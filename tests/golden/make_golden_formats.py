"""Generates tests/golden/golden_formats_v1.npz from the REFERENCE ITSELF (oracle/_ref/libt3ref*.so and libt3ref_new.so,
compiled from /root/reference by oracle/Makefile) for the formats either side of the hot path:
.t3v frame records and header, sub-word trit streams, base-243 payloads, the NEW generation's raw words and image
bridge.  Run in the build container:

    python tests/golden/make_golden_formats.py

tests/test_golden.py replays the fixtures against the C oracle (CPU), tests/test_gpu_parity.py against the CUDA path
(GPU box, where /root/reference does not exist).
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import t3oracle as T  # noqa: E402

SUBWORDS = (27, 24, 21, 18, 15)
T3V_WORDS = (0, 1, 15, 3641, 7300)          # around one 128-byte segment, one 32 KiB tile, two tiles
HEADERS = ((1, 0, 1, 0, 240, 0), (4, 2, 1, 1, 1, 1), (5, 4, 0, 2, 7, 0))   # profile, subword code, centered, coset, frames, file type


def main():
    ref, new = T.Reference(False), T.ReferenceNew()
    r = np.random.default_rng(20261019)
    out = {}
    # --- f.1: .t3v records (old/include/t3v_io.hpp:128-160) and header (:97-119)
    for nw in T3V_WORDS:
        words = r.integers(0, 27, size=(nw, 9), dtype=np.uint8)
        if nw == 15:
            words = r.integers(0, 256, size=(nw, 9), dtype=np.uint8)        # bytes >= 27 are stored % 27
        out[f"t3v{nw}_words"] = words
        out[f"t3v{nw}_record"] = ref.t3v_frame_record(words)
    aw = (3, 1, 4, 1)
    for i, (prof, code, cen, coset, fc, ft) in enumerate(HEADERS):
        buf = np.zeros(256, np.uint8)
        back = C.c_int()
        ref.lib.t3r_t3v_header.restype = C.c_size_t
        sub_mode = SUBWORDS[code] if code < 5 else 27
        n = ref.lib.t3r_t3v_header(buf.ctypes.data_as(C.c_void_p), prof, sub_mode, cen, coset, 7680, 4320, (C.c_uint32 * 4)(*aw), 30000, 1001, fc, ft, C.byref(back))
        out[f"t3vhdr{i}"] = buf[:n].copy()
    # --- f.2: sub-word streams and base-243 payloads (ternary_packing.hpp, ternary_image_codec_v6_min.hpp)
    words = r.integers(0, 256, size=(40, 9), dtype=np.uint8)
    out["sub_words"] = words
    for N in SUBWORDS:
        t = ref.subword_stream(words, N)
        out[f"sub{N}_stream"] = t
        out[f"sub{N}_pack"] = ref.base243_pack(t)
        short = t[: 5 * N + 3]
        out[f"sub{N}_rebuilt_fill2"] = ref.words_from_subword_stream(short, N, 2)
    wild = r.integers(0, 256, 203, dtype=np.uint8)                          # trits >= 3: the reference's own arithmetic applies
    out["wild_trits"] = wild
    out["wild_pack"] = ref.base243_pack(wild)
    out["wild_rebuilt21"] = ref.words_from_subword_stream(wild, 21, 1)
    ok, back_t = ref.base243_unpack(out["wild_pack"])
    out["wild_unpack_ok"], out["wild_unpack"] = np.array([ok]), back_t
    # --- f.3 / f.4: NEW generation raw words and the image bridge (include/io_image.hpp)
    px = T.synth_quant(11, 4096)
    out["new_px"] = px.view(np.uint8).reshape(-1, px.dtype.itemsize)
    for sub in (0, 15, 21):
        ok, w = new.pack_pixels(px, sub)
        out[f"new_pack{sub}_ok"], out[f"new_pack{sub}"] = np.array([ok]), w
    img = r.integers(0, 256, (41, 67, 3), dtype=np.uint8)
    out["img"] = img
    out["img_resize_960x540"] = new.resize_rgb_nn(img, 960, 540)
    out["img_resize_17x31"] = new.resize_rgb_nn(img, 17, 31)
    out["img_blit_101x77"] = new.blit_center_rgb(img, 101, 77)
    q = T.synth_quant(3, 60 * 40)
    out["q_60x40"] = q.view(np.uint8).reshape(-1, q.dtype.itemsize)
    out["q_center_30x20"] = new.extract_center_q(q, 60, 40, 30, 20).view(np.uint8).reshape(-1, q.dtype.itemsize)
    for sub, cen in ((15, True), (15, False), (18, True)):
        ok, w = new.image_to_words_subword(img, sub, cen)
        assert ok
        out[f"img_words_{sub}_{int(cen)}"] = w
        ok, back_img = new.words_to_image_subword(w, sub, 100, 50)
        assert ok
        out[f"img_back_{sub}_{int(cen)}_100x50"] = back_img
    path = os.path.join(HERE, "golden_formats_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()

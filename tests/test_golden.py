"""Replays the committed golden vectors (generated FROM THE REFERENCE by tests/golden/make_golden.py)
against the C oracle.  CPU only; needs neither /root/reference nor oracle/_ref."""
import os

import numpy as np
import pytest

import t3oracle as T

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
NAMES = sorted(k[4:] for k in G.files if k.startswith("cfg_"))


def cfg_of(name):
    return T.Cfg.from_buffer_copy(G["cfg_" + name].tobytes())


@pytest.mark.parametrize("k", (24, 22, 20, 18))
def test_rs_golden(oracle, k):
    d = G[f"rs{k}_data"]
    assert np.array_equal(oracle.rs_encode_blocks(k, d, 0), G[f"rs{k}_enc_ref"])
    assert np.array_equal(oracle.rs_encode_blocks(k, d, 1), G[f"rs{k}_enc_fix"])
    for tag, fx in (("ref", 0), ("fix", 1)):
        io, out, ok = oracle.rs_decode_blocks(k, G[f"rs{k}_dec_in"], fx)
        assert np.array_equal(ok, G[f"rs{k}_dec_{tag}_ok"])
        assert np.array_equal(io, G[f"rs{k}_dec_{tag}_io"])
        assert np.array_equal(out, G[f"rs{k}_dec_{tag}_out"])


def test_bridge_and_packing_golden(oracle):
    q = oracle.rgb_to_quant(G["rgb"])
    assert np.array_equal(q.view(np.uint8).reshape(-1, 6), G["quant"])
    assert np.array_equal(oracle.quant_to_rgb(q), G["rgb_back"])
    assert np.array_equal(oracle.pack_pixels(q), G["raw_words"])
    assert np.array_equal(oracle.unpack_pixels(G["raw_words"]).view(np.uint8).reshape(-1, 6), G["unpacked"])


@pytest.mark.parametrize("name", NAMES)
def test_pipeline_golden(oracle, name):
    cfg = cfg_of(name)
    raw = G["pipe_raw"]
    assert np.array_equal(oracle.header_pack(cfg), G["hdr_" + name])
    assert np.array_equal(oracle.encode_profile(cfg, raw, 0), G["enc_ref_" + name])
    assert np.array_equal(oracle.encode_profile(cfg, raw, 1), G["enc_fix_" + name])
    assert np.array_equal(oracle.encode_rgb(cfg, G["rgb"], 0), G["encrgb_ref_" + name])
    ok, words, seen = oracle.decode_profile_ref(T.make_cfg(), G["dec_ref_in_" + name])
    assert ok == bool(G["dec_ref_ok_" + name][0])
    assert np.array_equal(words, G["dec_ref_out_" + name])
    assert bytes(seen) == G["dec_ref_seen_" + name].tobytes()
    ok, out, ncorr = oracle.decode_profile_fixed(cfg, G["enc_fix_" + name], n_raw_words=raw.shape[0])
    assert ok and ncorr == 0 and np.array_equal(out, raw[:out.shape[0]]) and out.shape[0] > 700


# ------------------------------------------------------------------ formats either side of the path
GF = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_formats_v1.npz"))   # tests/golden/make_golden_formats.py
T3V_WORDS = sorted(int(k[3:-6]) for k in GF.files if k.startswith("t3v") and k.endswith("_words"))
HEADERS = ((1, 0, 1, 0, 240, 0), (4, 2, 1, 1, 1, 1), (5, 4, 0, 2, 7, 0))                  # as in make_golden_formats.py


def test_formats_golden_t3v(oracle):
    import ctypes as C
    for nw in T3V_WORDS:
        words, rec = GF[f"t3v{nw}_words"], GF[f"t3v{nw}_record"]
        assert np.array_equal(oracle.t3v_frame_record(words), rec), nw
        ok, back = oracle.t3v_read_frame(rec)
        assert ok and np.array_equal(back, words % 27)
    for i, (prof, code, cen, coset, fc, ft) in enumerate(HEADERS):
        o = np.zeros(54, np.uint8)
        oracle.lib.t3o_t3v_header(o.ctypes.data_as(C.c_void_p), prof, code, cen, coset, 7680, 4320, (C.c_uint32 * 4)(3, 1, 4, 1), 30000, 1001, fc, ft)
        assert np.array_equal(o, GF[f"t3vhdr{i}"]), i


def test_formats_golden_subword_and_base243(oracle):
    words = GF["sub_words"]
    for N in (27, 24, 21, 18, 15):
        t = oracle.subword_stream(words, N)
        assert np.array_equal(t, GF[f"sub{N}_stream"])
        assert np.array_equal(oracle.base243_pack(t), GF[f"sub{N}_pack"])
        assert np.array_equal(oracle.words_from_subword_stream(t[:5 * N + 3], N, 2), GF[f"sub{N}_rebuilt_fill2"])
    wild = GF["wild_trits"]
    assert np.array_equal(oracle.base243_pack(wild), GF["wild_pack"])
    assert np.array_equal(oracle.words_from_subword_stream(wild, 21, 1), GF["wild_rebuilt21"])
    ok, u = oracle.base243_unpack(GF["wild_pack"])
    assert ok == bool(GF["wild_unpack_ok"][0]) and np.array_equal(u, GF["wild_unpack"])


def test_formats_golden_new_generation(oracle):
    px = GF["new_px"].reshape(-1).view(T.PIXEL_DTYPE)
    assert np.array_equal(oracle.v6new_pack_pixels(px), GF["new_pack0"]) and np.array_equal(GF["new_pack0"], GF["new_pack15"])
    img = GF["img"]
    assert np.array_equal(oracle.resize_rgb_nn(img, 960, 540), GF["img_resize_960x540"])
    assert np.array_equal(oracle.resize_rgb_nn(img, 17, 31), GF["img_resize_17x31"])
    assert np.array_equal(oracle.blit_center_rgb(img, 101, 77), GF["img_blit_101x77"])
    q = GF["q_60x40"].reshape(-1).view(T.PIXEL_DTYPE)
    assert np.array_equal(oracle.extract_center_q(q, 60, 40, 30, 20).view(np.uint8).reshape(-1, q.dtype.itemsize), GF["q_center_30x20"])
    for sub, cen in ((15, True), (15, False), (18, True)):
        ok, w = oracle.v6new_image_to_words(img, sub, cen)
        assert ok and np.array_equal(w, GF[f"img_words_{sub}_{int(cen)}"])
        ok, back = oracle.v6new_words_to_image(w, sub, 100, 50)
        assert ok and np.array_equal(back, GF[f"img_back_{sub}_{int(cen)}_100x50"])

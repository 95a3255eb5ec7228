"""Pins the C restatement (oracle/t3_oracle.c) against the reference itself.

The reference ships no golden vectors, so the oracle is pinned differentially
against oracle/_ref/libt3ref.so (the reference compiled as shipped) and libt3ref_fixed.so (same
sources + the 3-line repair of SURVEY Appendix B), stage by stage and through the whole pipeline,
plus the known answers captured in SURVEY Appendix C.  CPU only.
"""
import ctypes as C

import numpy as np
import pytest

import t3oracle as T

KS = (24, 22, 20, 18)


def rng(seed):
    return np.random.default_rng(seed)


# ---------------------------------------------------------------- field + generator polynomials
def test_gf_tables_match_reference(oracle, ref):
    for a, b in zip(oracle.gf_tables(), ref.gf_tables()):
        assert np.array_equal(a, b)


def test_gf_known_answers(oracle):
    exp, log, mul, inv = oracle.gf_tables()
    assert list(exp[:26]) == [1, 3, 9, 5, 15, 23, 13, 17, 20, 4, 12, 14, 11, 2, 6, 18, 7, 21, 16, 26, 22, 10, 8, 24, 25, 19]
    assert list(inv) == [0, 1, 2, 19, 21, 24, 11, 12, 15, 25, 23, 6, 7, 22, 18, 8, 20, 26, 14, 3, 16, 4, 13, 10, 5, 9, 17]
    assert list(log) == [-1, 0, 13, 1, 9, 3, 14, 16, 22, 2, 21, 12, 10, 6, 11, 4, 18, 7, 15, 25, 8, 17, 20, 5, 23, 24, 19]


def test_generators(oracle, ref):
    want = {24: [5, 24, 1], 22: [12, 24, 15, 16, 1], 20: [10, 4, 19, 13, 16, 10, 1], 18: [12, 14, 24, 23, 1, 8, 23, 12, 1]}
    for k in KS:
        assert list(oracle.rs_gen(k)) == want[k]
        assert list(ref.rs_gen(k)) == want[k]


# ---------------------------------------------------------------- RS block codec
@pytest.mark.parametrize("k", KS)
def test_rs_encode_matches_reference(oracle, ref, ref_fixed, k):
    data = rng(k).integers(0, 27, size=(20000, k), dtype=np.uint8)
    data[0] = (5 * np.arange(k) + 7) % 27  # selftest_rs_unit pattern, OLD:1186
    data[1] = 0
    assert np.array_equal(oracle.rs_encode_blocks(k, data, fixed=0), ref.rs_encode_blocks(k, data, fixed=0))
    assert np.array_equal(oracle.rs_encode_blocks(k, data, fixed=1), ref_fixed.rs_encode_blocks(k, data, fixed=1))


def test_rs_encode_known_answers(oracle):
    ref_par = {24: [3, 18], 22: [23, 3, 22, 16], 20: [20, 20, 20, 7, 26, 16], 18: [19, 19, 26, 23, 6, 5, 11, 26]}
    fix_par = {24: [1, 1], 22: [9, 5, 5, 9], 20: [25, 18, 1, 19, 11, 15], 18: [6, 23, 24, 1, 5, 7, 13, 19]}
    for k in KS:
        d = ((5 * np.arange(k) + 7) % 27).astype(np.uint8)
        assert list(oracle.rs_encode_blocks(k, d, 0)[0, k:]) == ref_par[k]
        assert list(oracle.rs_encode_blocks(k, d, 1)[0, k:]) == fix_par[k]


def _decode_inputs(oracle, k, n, seed):
    """Three input classes of SURVEY 3.3: garbage, true codewords +- errors, shipped-encoder outputs +- errors."""
    r = rng(seed)
    t = (26 - k) // 2
    add = T.gf_add_table()
    data = r.integers(0, 27, size=(n, k), dtype=np.uint8)
    blocks = [r.integers(0, 27, size=(n, 26), dtype=np.uint8)]
    for fixed in (1, 0):
        cw = oracle.rs_encode_blocks(k, data, fixed)
        for e in range(0, t + 3):
            c = cw.copy()
            for row in range(n):
                pos = r.choice(26, size=min(e, 26), replace=False)
                c[row, pos] = add[c[row, pos], r.integers(1, 27, size=pos.size)]
            blocks.append(c)
    return np.concatenate(blocks)


@pytest.mark.parametrize("k", KS)
def test_rs_decode_matches_reference(oracle, ref, ref_fixed, k):
    blocks = _decode_inputs(oracle, k, 1500, 100 + k)
    for fixed, R in ((0, ref), (1, ref_fixed)):
        io_o, out_o, ok_o = oracle.rs_decode_blocks(k, blocks, fixed)
        io_r, out_r, ok_r = R.rs_decode_blocks(k, blocks, fixed)
        assert np.array_equal(ok_o, ok_r)
        assert np.array_equal(io_o, io_r)
        assert np.array_equal(out_o, out_r)
        assert 0 < ok_o.sum() < ok_o.size or fixed == 0  # both outcomes exercised


@pytest.mark.parametrize("k", KS)
def test_rs_fixed_corrects_up_to_t(oracle, k):
    r = rng(k)
    t = (26 - k) // 2
    add = T.gf_add_table()
    data = r.integers(0, 27, size=(3000, k), dtype=np.uint8)
    cw = oracle.rs_encode_blocks(k, data, 1)
    for e in range(t + 1):
        c = cw.copy()
        for row in range(c.shape[0]):
            pos = r.choice(26, size=e, replace=False)
            c[row, pos] = add[c[row, pos], r.integers(1, 27, size=e)]
        io, out, ok = oracle.rs_decode_blocks(k, c, 1)
        assert ok.all() and np.array_equal(out, data) and np.array_equal(io, cw)


def test_reference_selftests_status(ref, ref_fixed):
    # shipped code prints RS:FAIL API:FAIL; the arithmetic repair turns RS into OK only.
    assert ref.selftests() == (False, False)
    assert ref_fixed.selftests() == (True, False)


# ---------------------------------------------------------------- header / CRC
def _random_cfg(r, wild=False):
    hi = 2 ** 32 if wild else 27
    return T.make_cfg(profile=int(r.choice([0, 1, 2, 3, 4])), uep=[int(x) for x in r.integers(0, 4, 9)],
                      tile=(int(r.integers(0, 70)), int(r.integers(0, 70))) if r.random() < 0.7 else (0, 0),
                      seed=tuple(int(x) for x in r.integers(0, hi, 3)),
                      beacon=(int(r.integers(0, 40)), int(r.integers(0, 12)), bool(r.integers(0, 2))),
                      superframe_words=int(r.integers(0, 100000)), subword=int(r.choice([27, 24, 21, 18, 15])),
                      centered=bool(r.integers(0, 2)), coset=int(r.integers(0, 3)))


def test_header_pack_check_unpack(oracle, ref):
    assert list(oracle.header_pack(T.make_cfg(uep=0))) == \
        [0, 6, 1, 1, 0, 0, 0, 0, 0, 1, 1, 1, 9, 0, 0, 0, 0, 0, 0, 0, 8, 15, 3, 0, 0, 0, 19]  # SURVEY A.6 KAT
    r = rng(7)
    for i in range(400):
        cfg = _random_cfg(r, wild=(i % 2 == 1))
        fs, bh = int(r.integers(0, 2 ** 32)), int(r.integers(0, 2 ** 32))
        so, sr = oracle.header_pack(cfg, fs, bh), ref.header_pack(cfg, fs, bh)
        assert np.array_equal(so, sr)
        assert oracle.header_check(so) and ref.header_check(so)
        uo, ur = oracle.header_unpack(so), ref.header_unpack(so)
        assert uo[0].astuple() == ur[0].astuple() and uo[1:] == ur[1:]
        bad = so.copy()
        bad[int(r.integers(0, 27))] = (bad[int(r.integers(0, 27))] + 1 + int(r.integers(0, 25))) % 27
        assert oracle.header_check(bad) == ref.header_check(bad)
    for _ in range(200):  # arbitrary symbol vectors
        s = r.integers(0, 27, 27, dtype=np.uint8)
        assert oracle.header_check(s) == ref.header_check(s)
        assert oracle.header_unpack(s)[0].astuple() == ref.header_unpack(s)[0].astuple()
    for n in (0, 1, 5, 69, 100):
        tr = r.integers(0, 3, n, dtype=np.uint8)
        assert np.array_equal(oracle.crc12(tr), ref.crc12(tr))


def test_luma_priority_header_roundtrip_is_garbled(oracle):
    # bug B7: MSB-first pack, LSB-first unpack => 211211211 comes back as 112112112
    cfg = T.make_cfg(uep=T.UEP_LUMA)
    got = oracle.header_unpack(oracle.header_pack(cfg))[0]
    assert tuple(got.uep) == (1, 1, 2, 1, 1, 2, 1, 1, 2)


# ---------------------------------------------------------------- pixel packing, interleave, scrambler
def test_pack_unpack_pixels(oracle, ref):
    r = rng(3)
    px = T.synth_quant(4, 4099)  # odd count: last word pairs with the default pixel
    assert np.array_equal(oracle.pack_pixels(px), ref.pack_pixels(px))
    wild = np.zeros(3000, T.PIXEL_DTYPE)  # out-of-range values wrap through %3 digits
    wild["Yq"] = r.integers(0, 65536, 3000)
    wild["Cbq"] = r.integers(-32768, 32768, 3000)
    wild["Crq"] = r.integers(-32768, 32768, 3000)
    assert np.array_equal(oracle.pack_pixels(wild), ref.pack_pixels(wild))
    words = r.integers(0, 27, size=(5000, 9), dtype=np.uint8)
    assert np.array_equal(oracle.unpack_pixels(words), ref.unpack_pixels(words))
    wordsb = r.integers(0, 256, size=(2000, 9), dtype=np.uint8)  # bytes >= 27 read as their low 3 trits
    assert np.array_equal(oracle.unpack_pixels(wordsb), ref.unpack_pixels(wordsb))
    kat = np.zeros(2, T.PIXEL_DTYPE)
    i = np.arange(4)
    kat4 = np.zeros(4, T.PIXEL_DTYPE)
    kat4["Yq"], kat4["Cbq"], kat4["Crq"] = (7 * i) % 243, (3 * i) % 81 - 40, (5 * i) % 81 - 40
    w = oracle.pack_pixels(kat4)
    assert list(w[0]) == [0, 0, 0, 0, 21, 0, 3, 15, 0] and list(w[1]) == [14, 0, 2, 10, 9, 2, 9, 18, 1]
    del kat


@pytest.mark.parametrize("w,h", [(1, 1), (2, 3), (7, 5), (26, 26), (26, 3), (64, 64), (5, 1), (1, 9), (300, 2)])
def test_interleave2d(oracle, ref, w, h):
    r = rng(w * 131 + h)
    for n in (0, 1, w * h - 1, w * h, w * h + 1, 3 * w * h + w + 1, 2501):
        if n < 0:
            continue
        sy = r.integers(0, 27, n, dtype=np.uint8)
        a, b = oracle.interleave2d(sy, w, h), ref.interleave2d(sy, w, h)
        assert np.array_equal(a, b)
        assert np.array_equal(oracle.interleave2d(a, w, h, inverse=True), sy)
        assert np.array_equal(ref.interleave2d(a, w, h, inverse=True), sy)


def test_scrambler(oracle, ref):
    r = rng(11)
    for a, b, s0 in [(1, 1, 1), (2, 1, 1), (0, 2, 1), (1, 0, 2), (5, 7, 11), (2 ** 32 - 1, 2 ** 32 - 2, 2), (2 ** 31 + 3, 2 ** 30 + 1, 5)]:
        sy = r.integers(0, 27, 200, dtype=np.uint8)
        so, sr = oracle.scramble_stream(sy, a, b, s0), ref.scramble_stream(sy, a, b, s0)
        assert np.array_equal(so, sr)
        assert np.array_equal(oracle.scramble_stream(so, a, b, s0, inverse=True), sy)
        assert np.array_equal(ref.scramble_stream(so, a, b, s0, inverse=True), sy)
    assert oracle.beacon_symbol(1, 8192 % 5) == 11 == ref.beacon_symbol(1, 8192 % 5)


# ---------------------------------------------------------------- bridge
def test_bridge_exhaustive_rgb(oracle, ref):
    # every 4th RGB triple on CPU here (4.2M); the full 2^24 sweep runs in the GPU parity test.
    idx = np.arange(0, 1 << 24, 4, dtype=np.uint32)
    rgb = np.stack([(idx >> 16) & 255, (idx >> 8) & 255, idx & 255], axis=1).astype(np.uint8)
    assert np.array_equal(oracle.rgb_to_quant(rgb), ref.rgb_to_quant(rgb))


def test_bridge_all_quant_values(oracle, ref):
    yq, cb, cr = np.meshgrid(np.arange(243), np.arange(-40, 41), np.arange(-40, 41), indexing="ij")
    px = np.zeros(yq.size, T.PIXEL_DTYPE)
    px["Yq"], px["Cbq"], px["Crq"] = yq.ravel(), cb.ravel(), cr.ravel()
    assert np.array_equal(oracle.quant_to_rgb(px), ref.quant_to_rgb(px))


# ---------------------------------------------------------------- whole pipeline
CONFIGS = [
    dict(),                                                                # EncoderContext defaults: P2, k=22
    dict(profile=T.P2, uep=T.UEP_LUMA),                                    # selftest_api_roundtrip config
    dict(profile=T.P3, uep=2),                                             # "RS(26,20)" (bug B9)
    dict(profile=T.P1, uep=0), dict(profile=T.P4, uep=3),
    dict(profile=T.P5, tile=(7, 5), beacon=(4, 2, True), uep=T.UEP_LUMA, seed=(2, 1, 1)),
    dict(profile=T.P5, tile=(26, 3), beacon=(26, 8, True), uep=3, seed=(1, 2, 0)),
    dict(profile=T.P5, tile=(26, 26), beacon=(26, 2, True), uep=T.UEP_LUMA, seed=(2, 1, 1), coset=1),
    dict(profile=T.P2, tile=(64, 64), beacon=(83, 2, True)),               # old/src/main.cpp:17
    dict(profile=T.P5, tile=(300, 7), uep=(0, 1, 2, 3, 0, 1, 2, 3, 0), beacon=(1, 0, True)),
    dict(profile=T.P3, uep=2, beacon=(5, 11, True)),                       # band_slot > 8: no beacon emitted, words still padded
    dict(profile=T.P3, uep=2, seed=(2 ** 32 - 1, 2 ** 31 + 5, 7)),         # uint32 wrap in the scrambler LCG
    dict(profile=T.RAW_MODE),
]


@pytest.mark.parametrize("ci", range(len(CONFIGS)))
def test_encode_profile_matches_reference(oracle, ref, ref_fixed, ci):
    cfg = T.make_cfg(**CONFIGS[ci])
    r = rng(1000 + ci)
    for n in (0, 1, 2, 3, 5, 26, 27, 64, 777, 1000, 8192):
        raw = r.integers(0, 27, size=(n, 9), dtype=np.uint8)
        if n == 64:
            raw = r.integers(0, 256, size=(n, 9), dtype=np.uint8)  # bytes >= 27: low 3 trits are used
        eo, er = oracle.encode_profile(cfg, raw, 0), ref.encode_profile(cfg, raw, 0)
        assert eo.shape == er.shape and np.array_equal(eo, er), (ci, n)
        assert oracle.words_bound(cfg, n) == er.shape[0]
        assert np.array_equal(oracle.encode_profile(cfg, raw, 1), ref_fixed.encode_profile(cfg, raw, 1))


def test_encode_known_answer_selftest_config(oracle):
    i = np.arange(64)
    px = np.zeros(64, T.PIXEL_DTYPE)
    px["Yq"], px["Cbq"], px["Crq"] = (7 * i) % 243, (3 * i) % 81 - 40, (5 * i) % 81 - 40
    out = oracle.encode_profile(T.make_cfg(profile=T.P2, uep=T.UEP_LUMA), oracle.pack_pixels(px), 0)
    assert out.shape[0] == 32
    want = [[0, 6, 1, 1, 22, 22, 22, 0, 0], [1, 1, 1, 9, 0, 0, 0, 0, 0], [21, 13, 26, 21, 26, 25, 17, 23, 0],
            [0, 4, 24, 21, 0, 0, 0, 4, 0], [0, 0, 0, 0, 0, 0, 0, 0, 4], [15, 17, 12, 6, 11, 18, 16, 26, 4]]
    assert out[:6].tolist() == want


def _valid_header_stream(oracle, cfg, body_words, rnd):
    """A stream whose header is built from TRUE RS(26,18) codewords so that the shipped decoder gets past it."""
    hp = oracle.header_pack(cfg)
    a = oracle.rs_encode_blocks(18, hp[:18], 1)[0]
    b = oracle.rs_encode_blocks(18, np.concatenate([hp[18:], np.zeros(9, np.uint8)]), 1)[0]
    head = np.concatenate([a, b, rnd.integers(0, 27, 2, dtype=np.uint8)])
    return np.concatenate([head.reshape(6, 9), body_words])


@pytest.mark.parametrize("ci", range(len(CONFIGS) - 1))
def test_decode_ref_exact_matches_reference(oracle, ref, ci):
    cfg = T.make_cfg(**CONFIGS[ci])
    r = rng(2000 + ci)
    seen0 = T.make_cfg()
    # (i) the reference encoder's own output: the shipped decoder rejects it at the header (bug B1)
    raw = r.integers(0, 27, size=(500, 9), dtype=np.uint8)
    enc = ref.encode_profile(cfg, raw, 0)
    ro, rr = oracle.decode_profile_ref(seen0, enc), ref.decode_profile_ref(seen0, enc)
    assert ro[0] == rr[0] is False and ro[1].size == rr[1].size == 0 and ro[2].astuple() == rr[2].astuple()
    # (ii) valid header, body of true (repaired) codewords laid out the way the DEcoder reads them, +- errors,
    #      and random bodies (mostly `false` for small r, garbled `true` otherwise)
    for trial, nbody in enumerate((0, 5, 26, 27, 130, 260, 263)):
        body = r.integers(0, 27, size=(nbody, 9), dtype=np.uint8)
        if trial >= 3:
            fo = oracle.encode_profile(cfg, r.integers(0, 27, size=(3 * nbody, 9), dtype=np.uint8), 1)
            body = fo[6:6 + nbody]
        s = _valid_header_stream(oracle, cfg, body, r)
        ro, rr = oracle.decode_profile_ref(seen0, s), ref.decode_profile_ref(seen0, s)
        assert ro[0] == rr[0], (ci, trial)
        assert ro[1].shape == rr[1].shape and np.array_equal(ro[1], rr[1]), (ci, trial)
        assert ro[2].astuple() == rr[2].astuple()
    # (iii) random words and short inputs
    for n in (0, 3, 6, 40):
        s = r.integers(0, 27, size=(n, 9), dtype=np.uint8)
        ro, rr = oracle.decode_profile_ref(seen0, s), ref.decode_profile_ref(seen0, s)
        assert ro[0] == rr[0] and np.array_equal(ro[1], rr[1]) and ro[2].astuple() == rr[2].astuple()
    # RAW passthrough is keyed on the PREVIOUS header (stateful), OLD:998
    seen_raw = T.make_cfg(profile=T.RAW_MODE)
    ro, rr = oracle.decode_profile_ref(seen_raw, raw), ref.decode_profile_ref(seen_raw, raw)
    assert ro[0] and rr[0] and np.array_equal(ro[1], raw) and np.array_equal(rr[1], raw)


def test_decode_ref_exact_true_path_executes(oracle, ref):
    """At least one stream must get through the shipped decoder with `true` so that the body path is compared."""
    cfg = T.make_cfg(profile=T.P4, uep=3)  # k=18: random blocks often "decode"
    r = rng(5)
    hits = 0
    for _ in range(40):
        s = _valid_header_stream(oracle, cfg, r.integers(0, 27, size=(26, 9), dtype=np.uint8), r)
        ro, rr = oracle.decode_profile_ref(T.make_cfg(), s), ref.decode_profile_ref(T.make_cfg(), s)
        assert ro[0] == rr[0] and np.array_equal(ro[1], rr[1])
        hits += int(rr[0])
    # clean all-zero body decodes to all-zero words
    s = _valid_header_stream(oracle, cfg, np.zeros((52, 9), np.uint8), r)
    s[6:] = ref.scramble_stream(np.zeros(52 * 9, np.uint8), 1, 1, 1).reshape(-1, 9)
    ro, rr = oracle.decode_profile_ref(T.make_cfg(), s), ref.decode_profile_ref(T.make_cfg(), s)
    assert ro[0] and rr[0] and np.array_equal(ro[1], rr[1]) and rr[1].shape[0] > 0 and not rr[1].any()


# ---------------------------------------------------------------- FIXED mode: consistent decoder
@pytest.mark.parametrize("ci", range(len(CONFIGS) - 1))
def test_fixed_roundtrip_and_error_correction(oracle, ref_fixed, ci):
    cfg = T.make_cfg(**CONFIGS[ci])
    r = rng(3000 + ci)
    add = T.gf_add_table()
    for n in (0, 1, 3, 64, 777, 4000):
        raw = r.integers(0, 27, size=(n, 9), dtype=np.uint8)
        raw[:, 8] %= 9  # T[26]=0, as produced by pack_two_pixels
        enc = ref_fixed.encode_profile(cfg, raw, 1)  # wire format = the (repaired) reference encoder's
        ok, out, ncorr = oracle.decode_profile_fixed(cfg, enc, n_raw_words=n)
        assert ok and ncorr == 0
        assert np.array_equal(out, raw[:out.shape[0]])
        # only the tail (< 9*(k_max-1)+26 symbols, bug B8) may be lost
        assert n - out.shape[0] <= (9 * 24 * 3 + 26 * 9) // 26 + 2 + (cfg.tile_w if cfg.profile == 4 else 0)
        if not (cfg.profile == 4 and cfg.tile_w and cfg.tile_h):
            ok2, out2, _ = oracle.decode_profile_fixed(cfg, enc, n_raw_words=0)  # N_w inferred from N_out
            assert ok2 and np.array_equal(out2, out)
        if n >= 64 and not (cfg.beacon_enabled and cfg.beacon_slot > 8):
            bad, nerr = T.inject_errors(enc, cfg, n, seed=3, gf_add=add)
            ok3, out3, ncorr3 = oracle.decode_profile_fixed(cfg, bad, n_raw_words=n)
            assert ok3 and np.array_equal(out3, out) and ncorr3 == nerr and nerr > 0
            bad_t, nerr_t = T.inject_errors(enc, cfg, n, seed=4, gf_add=add, exact_t=True)
            ok4, out4, ncorr4 = oracle.decode_profile_fixed(cfg, bad_t, n_raw_words=n)
            assert ok4 and np.array_equal(out4, out) and ncorr4 == nerr_t


def test_fused_rgb_chain(oracle, ref, ref_fixed):
    rgb = T.synth_rgb(1, 512 * 64)
    cfg = T.make_cfg(profile=T.P3, uep=2)
    assert np.array_equal(oracle.encode_rgb(cfg, rgb, 0), ref.encode_rgb(cfg, rgb, 0))
    enc = ref_fixed.encode_rgb(cfg, rgb, 1)
    assert np.array_equal(oracle.encode_rgb(cfg, rgb, 1), enc)
    ok, back, _ = oracle.decode_rgb_fixed(cfg, enc, rgb.shape[0])
    assert ok
    want = ref.quant_to_rgb(ref.rgb_to_quant(rgb))[:back.shape[0]]
    assert np.array_equal(back, want)
    assert np.abs(back.astype(int) - rgb[:back.shape[0]].astype(int)).max() <= 8  # lossy quantisation only
    assert rgb.shape[0] - back.shape[0] < 600


def test_tpack_words_bytes(ref):
    w = rng(1).integers(0, 256, size=(100, 9), dtype=np.uint8)
    b = ref.words_to_bytes(w)
    assert np.array_equal(b, (w % 27).reshape(-1))
    assert np.array_equal(ref.bytes_to_words(b), (w % 27))
    assert ref.bytes_to_words(b[:-1]).shape[0] == 0


# ------------------------------------------------------------------ formats either side of the path
SUBWORDS = (27, 24, 21, 18, 15)


def _wild_trits(r, n):
    t = r.integers(0, 3, n, dtype=np.uint8)
    if n:
        t[r.integers(0, n, max(1, n // 50))] = r.integers(3, 256, max(1, n // 50))  # UTrit is a byte: out-of-range values wrap through pack3
    return t


@pytest.mark.parametrize("N", SUBWORDS)
def test_subword_streams_and_base243_match_reference(oracle, ref, N):
    r = rng(800 + N)
    for nw in (0, 1, 2, 7, 1000):
        words = r.integers(0, 256, size=(nw, 9), dtype=np.uint8)      # bytes >= 27 read as their low three trits
        a = oracle.subword_stream(words, N)
        assert np.array_equal(a, ref.subword_stream(words, N)) and a.size == nw * N
        for n in sorted({0, 1, N - 1, N, N + 1, 5 * N + 3, a.size}):
            if n > a.size:
                continue
            for fill in (0, 1, 2, 77):
                t = _wild_trits(r, n) if fill == 77 else a[:n]
                assert np.array_equal(oracle.words_from_subword_stream(t, N, fill), ref.words_from_subword_stream(t, N, fill))
            for t in (a[:n], _wild_trits(r, n)):
                p = oracle.base243_pack(t)
                assert np.array_equal(p, ref.base243_pack(t)) and p.size == 4 + (n + 4) // 5
                for blob in (p, p[:max(0, p.size - 1)], p[:3], np.concatenate([p, r.integers(0, 256, 3, dtype=np.uint8)])):
                    ok_o, u_o = oracle.base243_unpack(blob)
                    ok_r, u_r = ref.base243_unpack(blob)
                    assert ok_o == ok_r and np.array_equal(u_o, u_r)
        # known answer: the first N trits of a valid word are the digits of its symbols
        if nw:
            w0 = words[0] % 27
            digits = np.array([(w0[i // 3] // 3 ** (i % 3)) % 3 for i in range(27)], np.uint8)
            assert np.array_equal(a[:N], digits[:N])


def test_v6new_raw_path_matches_reference(oracle, ref_new):
    r = rng(900)
    px = np.zeros(20001, T.PIXEL_DTYPE)
    px["Yq"], px["Cbq"], px["Crq"] = r.integers(0, 65536, px.size), r.integers(-32768, 32768, px.size), r.integers(-32768, 32768, px.size)
    px[:6000] = T.synth_quant(7, 6000)
    ok, w = ref_new.pack_pixels(px)
    assert ok and np.array_equal(w, oracle.v6new_pack_pixels(px))
    assert w[:6000].max() < 3 ** 13 and np.array_equal(oracle.v6new_unpack_pixels(w[:6000]).view(np.uint8), px[:6000].view(np.uint8))
    wild = r.integers(0, 2 ** 32, 20000, dtype=np.uint32)
    ok, p = ref_new.unpack_pixels(wild)
    assert ok and np.array_equal(p.view(np.uint8), oracle.v6new_unpack_pixels(wild).view(np.uint8))
    for sub, good in ((27, True), (24, True), (21, True), (18, True), (15, True), (7, False), (26, False)):
        ok, w2 = ref_new.pack_pixels(px[:100], sub)
        assert ok is good and (not good or np.array_equal(w2, w[:100]))     # the sub-word argument only gates validity (v6_min)
        assert ref_new.unpack_pixels(w[:100], sub)[0] is good
    # known answer: Y + 243 (Cb+40 + 81 (Cr+40))
    one = np.zeros(1, T.PIXEL_DTYPE)
    one["Yq"], one["Cbq"], one["Crq"] = 242, 40, 40
    assert int(oracle.v6new_pack_pixels(one)[0]) == 3 ** 13 - 1


# ------------------------------------------------------------------ .t3v container records
def test_t3v_records_match_reference_and_zlib(oracle, ref):
    import zlib
    r = rng(950)
    for n in (0, 1, 3, 4, 5, 1000, 65537):
        data = r.integers(0, 256, n, dtype=np.uint8)
        assert oracle.crc32(data) == ref.crc32(data) == zlib.crc32(data.tobytes())
    assert oracle.crc32(np.frombuffer(b"123456789", np.uint8)) == 0xCBF43926           # the CRC-32 check value
    for nw in (0, 1, 2, 7, 1000, 20011):
        words = r.integers(0, 256, size=(nw, 9), dtype=np.uint8)                        # bytes >= 27 are stored % 27
        rec = oracle.t3v_frame_record(words)
        assert np.array_equal(rec, ref.t3v_frame_record(words)) and rec.size == 8 + 9 * nw
        assert int.from_bytes(rec[:4].tobytes(), "little") == nw and np.array_equal(rec[4:4 + 9 * nw], (words % 27).reshape(-1))
        want_crc = zlib.crc32(rec[4:4 + 9 * nw].tobytes()) ^ ((zlib.crc32(rec[:4].tobytes()) * 16777619) & 0xFFFFFFFF)
        assert int.from_bytes(rec[-4:].tobytes(), "little") == want_crc
        for blob in (rec, rec[:-1], rec[:3], np.concatenate([rec, rec[:5]])):
            ok_o, w_o = oracle.t3v_read_frame(blob)
            ok_r, w_r = ref.t3v_read_frame(blob)
            assert ok_o == ok_r and np.array_equal(w_o, w_r)
        if nw:
            bad = rec.copy()
            bad[4 + int(r.integers(0, 9 * nw))] ^= 1
            assert oracle.t3v_read_frame(bad)[0] is False and ref.t3v_read_frame(bad)[0] is False
            ok_o, w_o = oracle.t3v_read_frame(rec)
            assert ok_o and np.array_equal(w_o, words % 27)


def test_t3v_header_matches_reference(oracle, ref):
    aw = (C.c_uint32 * 4)(140, 0, 6280, 4320)
    for prof, sub_mode, sub_code, cen, coset, ft, fc in ((1, 27, 0, 1, 0, 0, 1), (4, 15, 4, 0, 2, 1, 240), (2, 21, 2, 1, 1, 1, 0)):
        o = np.zeros(54, np.uint8)
        oracle.lib.t3o_t3v_header(o.ctypes.data_as(C.c_void_p), prof, sub_code, cen, coset, 7680, 4320, aw, 30000, 1001, fc, ft)
        rr = np.zeros(64, np.uint8)
        back = C.c_int()
        ref.lib.t3r_t3v_header.restype = C.c_size_t
        n = ref.lib.t3r_t3v_header(rr.ctypes.data_as(C.c_void_p), prof, sub_mode, cen, coset, 7680, 4320, aw, 30000, 1001, fc, ft, C.byref(back))
        assert n == 54 and back.value == 1 and np.array_equal(o, rr[:54])


# ------------------------------------------------------------------ image-bridge geometry (NEW generation)
def test_image_bridge_geometry_matches_reference(oracle, ref_new):
    r = rng(970)
    for (sh, sw), (dh, dw) in (((37, 53), (540, 960)), ((1, 1), (7, 5)), ((300, 200), (31, 17)), ((64, 64), (64, 64)), ((5, 9), (1, 1))):
        img = r.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(oracle.resize_rgb_nn(img, dw, dh), ref_new.resize_rgb_nn(img, dw, dh))
    for (sh, sw), (ch, cw) in (((37, 53), (77, 101)), ((540, 960), (541, 961)), ((10, 8), (4, 8)), ((3, 3), (3, 3)), ((9, 2), (1, 7))):
        img = r.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(oracle.blit_center_rgb(img, cw, ch), ref_new.blit_center_rgb(img, cw, ch))      # src taller than the canvas: rows dropped
    for (fh, fw), (sh, sw) in (((40, 60), (20, 30)), ((40, 60), (40, 60)), ((10, 60), (14, 30)), ((7, 9), (1, 1))):
        q = T.synth_quant(3, fw * fh)
        a, b = oracle.extract_center_q(q, fw, fh, sw, sh), ref_new.extract_center_q(q, fw, fh, sw, sh)
        assert a.size == b.size == sw * sh and np.array_equal(a.view(np.uint8), b.view(np.uint8))           # window taller than the frame: zero rows


def test_image_to_words_pipelines_match_reference(oracle, ref_new):
    r = rng(971)
    img = r.integers(0, 256, (41, 67, 3), dtype=np.uint8)
    for sub, cen in ((15, True), (15, False), (18, True), (27, True), (7, True)):
        ok_r, w_r = ref_new.image_to_words_subword(img, sub, cen)
        ok_o, w_o = oracle.v6new_image_to_words(img, sub, cen)
        assert ok_r == ok_o and np.array_equal(w_r, w_o), (sub, cen)
        if ok_r:
            tw, th = T.V6NEW_STD_RES[sub]
            for (w, h) in ((tw, th), (100, 50)):
                ok1, i1 = ref_new.words_to_image_subword(w_r, sub, w, h)
                ok2, i2 = oracle.v6new_words_to_image(w_o, sub, w, h)
                assert ok1 and ok2 and np.array_equal(i1, i2), (sub, cen, w, h)
    exact = r.integers(0, 256, (540, 960, 3), dtype=np.uint8)                                               # already the target size: no resize
    ok_r, w_r = ref_new.image_to_words_subword(exact, 15, False)
    ok_o, w_o = oracle.v6new_image_to_words(exact, 15, False)
    assert ok_r and ok_o and w_r.size == 960 * 540 and np.array_equal(w_r, w_o)

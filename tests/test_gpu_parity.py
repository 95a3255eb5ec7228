"""GPU parity tests: the CUDA path, called through the C ABI (libt3c.so), against the CPU oracle on the
same seeded inputs and against the golden vectors generated from the reference.  Bit-exact everywhere:
every quantity on this path is an integer/byte (the float32 bridge must match to the last bit too).
"""
import os

import numpy as np
import pytest

import t3oracle as T

pytestmark = pytest.mark.gpu

KS = (24, 22, 20, 18)
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


@pytest.fixture(scope="module")
def t3():
    import ternary_image_codec_b200 as m
    return m


@pytest.fixture(scope="module")
def codec(t3):
    c = t3.Codec(0)
    yield c
    c.close()


def rng(s):
    return np.random.default_rng(s)


def both(kw):
    """same config for the oracle struct and the product struct"""
    import ternary_image_codec_b200 as m
    return T.make_cfg(**kw), m.make_config(**kw)


CONFIGS = [
    dict(),
    dict(profile=T.P2, uep=T.UEP_LUMA),
    dict(profile=T.P3, uep=2),
    dict(profile=T.P1, uep=0), dict(profile=T.P4, uep=3),
    dict(profile=T.P5, tile=(7, 5), beacon=(4, 2, True), uep=T.UEP_LUMA, seed=(2, 1, 1)),
    dict(profile=T.P5, tile=(26, 3), beacon=(26, 8, True), uep=3, seed=(1, 2, 0)),
    dict(profile=T.P5, tile=(26, 26), beacon=(26, 2, True), uep=T.UEP_LUMA, seed=(2, 1, 1), coset=1),
    dict(profile=T.P2, tile=(64, 64), beacon=(83, 2, True)),
    dict(profile=T.P5, tile=(300, 7), uep=(0, 1, 2, 3, 0, 1, 2, 3, 0), beacon=(1, 0, True)),
    dict(profile=T.P3, uep=2, beacon=(5, 11, True)),
    dict(profile=T.P3, uep=2, seed=(2 ** 32 - 1, 2 ** 31 + 5, 7)),
    dict(profile=T.P3, uep=2, seed=(3, 5, 2)),      # a%3==0: constant scrambler after one step
    dict(profile=T.RAW_MODE),
]


# ------------------------------------------------------------------ K1
def test_bridge_exhaustive_2_24(codec, oracle):
    idx = np.arange(1 << 24, dtype=np.uint32)
    rgb = np.stack([(idx >> 16) & 255, (idx >> 8) & 255, idx & 255], axis=1).astype(np.uint8)
    got = codec.rgb_to_quant_stream(rgb)
    want = oracle.rgb_to_quant(rgb)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


def test_bridge_back_all_quant_values_and_wild(codec, oracle):
    yq, cb, cr = np.meshgrid(np.arange(243), np.arange(-40, 41), np.arange(-40, 41), indexing="ij")
    px = np.zeros(yq.size, T.PIXEL_DTYPE)
    px["Yq"], px["Cbq"], px["Crq"] = yq.ravel(), cb.ravel(), cr.ravel()
    assert np.array_equal(codec.quant_stream_to_rgb(px), oracle.quant_to_rgb(px))
    r = rng(1)
    wild = np.zeros(200000, T.PIXEL_DTYPE)  # out-of-range inputs clamp exactly like the reference
    wild["Yq"], wild["Cbq"], wild["Crq"] = r.integers(0, 65536, wild.size), r.integers(-32768, 32768, wild.size), r.integers(-32768, 32768, wild.size)
    assert np.array_equal(codec.quant_stream_to_rgb(wild), oracle.quant_to_rgb(wild))


@pytest.mark.parametrize("n", [0, 1, 2, 3, 2047, 2048, 2049, 4096, 4097, 100001])
def test_pack_unpack_pixels(codec, oracle, n):
    px = T.synth_quant(4, n)
    w = codec.encode_raw_pixels_to_words(px)
    assert np.array_equal(w, oracle.pack_pixels(px))
    assert np.array_equal(codec.decode_raw_words_to_pixels(w).view(np.uint8), oracle.unpack_pixels(w).view(np.uint8))
    r = rng(n)
    wild = np.zeros(n, T.PIXEL_DTYPE)
    wild["Yq"], wild["Cbq"], wild["Crq"] = r.integers(0, 65536, n), r.integers(-32768, 32768, n), r.integers(-32768, 32768, n)
    assert np.array_equal(codec.encode_raw_pixels_to_words(wild), oracle.pack_pixels(wild))
    words = r.integers(0, 256, size=((n + 1) // 2, 9), dtype=np.uint8)  # bytes >= 27 read as their low three trits
    assert np.array_equal(codec.decode_raw_words_to_pixels(words).view(np.uint8), oracle.unpack_pixels(words).view(np.uint8))
    assert np.array_equal(codec.words_to_bytes(words), (words % 27).reshape(-1))


def test_golden_bridge_and_packing(codec):
    q = codec.rgb_to_quant_stream(G["rgb"])
    assert np.array_equal(q.view(np.uint8).reshape(-1, 6), G["quant"])
    assert np.array_equal(codec.quant_stream_to_rgb(q), G["rgb_back"])
    assert np.array_equal(codec.encode_raw_pixels_to_words(q), G["raw_words"])
    assert np.array_equal(codec.decode_raw_words_to_pixels(G["raw_words"]).view(np.uint8).reshape(-1, 6), G["unpacked"])


# ------------------------------------------------------------------ RS block codec
@pytest.mark.parametrize("k", KS)
def test_rs_encode_blocks(codec, oracle, t3, k):
    data = rng(k).integers(0, 27, size=(50000, k), dtype=np.uint8)
    data[0] = (5 * np.arange(k) + 7) % 27
    data[1] = 0
    for arith in (t3.REF_EXACT, t3.FIXED):
        assert np.array_equal(codec.rs_encode_blocks(k, data, arith), oracle.rs_encode_blocks(k, data, arith))
    assert np.array_equal(codec.rs_encode_blocks(k, G[f"rs{k}_data"], 0), G[f"rs{k}_enc_ref"])
    assert np.array_equal(codec.rs_encode_blocks(k, G[f"rs{k}_data"], 1), G[f"rs{k}_enc_fix"])
    assert codec.rs_encode_blocks(k, np.zeros((0, k), np.uint8)).shape == (0, 26)


def decode_inputs(oracle, k, n, seed):
    r = rng(seed)
    t = (26 - k) // 2
    add = T.gf_add_table()
    data = r.integers(0, 27, size=(n, k), dtype=np.uint8)
    blocks = [r.integers(0, 27, size=(n, 26), dtype=np.uint8)]
    for fixed in (1, 0):
        cw = oracle.rs_encode_blocks(k, data, fixed)
        for e in range(0, t + 3):
            c = cw.copy()
            pos = np.argsort(r.random((n, 26)), axis=1)[:, :e]
            mag = r.integers(1, 27, size=(n, e))
            rows = np.arange(n)[:, None]
            c[rows, pos] = add[c[rows, pos], mag]
            blocks.append(c)
    return np.concatenate(blocks)


@pytest.mark.parametrize("k", KS)
def test_rs_decode_blocks(codec, oracle, t3, k):
    blocks = decode_inputs(oracle, k, 8000, 100 + k)
    for arith in (t3.REF_EXACT, t3.FIXED):
        io_g, out_g, ok_g = codec.rs_decode_blocks(k, blocks, arith)
        io_o, out_o, ok_o = oracle.rs_decode_blocks(k, blocks, arith)
        assert np.array_equal(ok_g, ok_o)
        assert np.array_equal(io_g, io_o)
        assert np.array_equal(out_g, out_o)
    for tag, arith in (("ref", 0), ("fix", 1)):
        io, out, ok = codec.rs_decode_blocks(k, G[f"rs{k}_dec_in"], arith)
        assert np.array_equal(ok, G[f"rs{k}_dec_{tag}_ok"]) and np.array_equal(io, G[f"rs{k}_dec_{tag}_io"]) and np.array_equal(out, G[f"rs{k}_dec_{tag}_out"])


def test_selftest_rs_unit_inputs(codec, oracle, t3):
    """selftest_rs_unit (OLD:1172-1207): t errors on (5i+7)%27; passes with the repaired arithmetic,
    and reproduces the reference's wrong answer bit for bit as shipped."""
    r = rng(1)
    add = T.gf_add_table()
    for k in KS:
        t = (26 - k) // 2
        d = ((5 * np.arange(k) + 7) % 27).astype(np.uint8)
        for arith in (t3.REF_EXACT, t3.FIXED):
            code = codec.rs_encode_blocks(k, d, arith)[0].copy()
            pos = r.choice(26, size=t, replace=False)
            code[pos] = add[code[pos], r.integers(1, 27, size=t)]
            io, out, ok = codec.rs_decode_blocks(k, code, arith)
            io_o, out_o, ok_o = oracle.rs_decode_blocks(k, code, arith)
            assert np.array_equal(out, out_o) and np.array_equal(ok, ok_o)
            if arith == t3.FIXED:
                assert ok[0] and np.array_equal(out[0], d)


@pytest.mark.parametrize("w,h", [(1, 1), (2, 3), (7, 5), (26, 26), (64, 64), (300, 2), (5, 1)])
def test_interleave2d(codec, oracle, w, h):
    r = rng(w * 131 + h)
    for n in (0, 1, w * h - 1, w * h, w * h + 1, 3 * w * h + w + 1, 100003):
        sy = r.integers(0, 27, n, dtype=np.uint8)
        a = codec.interleave2D_boustrophedon(sy, w, h)
        assert np.array_equal(a, oracle.interleave2d(sy, w, h))
        assert np.array_equal(codec.interleave2D_boustrophedon(a, w, h, inverse=True), sy)


# ------------------------------------------------------------------ header
def test_header_emit_and_parse(codec, oracle, t3):
    r = rng(7)
    for i in range(60):
        kw = dict(profile=int(r.choice([0, 1, 2, 3, 4])), uep=[int(x) for x in r.integers(0, 4, 9)],
                  tile=(int(r.integers(0, 70)), int(r.integers(0, 70))),
                  seed=tuple(int(x) for x in r.integers(0, 2 ** 32 if i % 2 else 27, 3)),
                  beacon=(int(r.integers(0, 40)), int(r.integers(0, 12)), bool(r.integers(0, 2))),
                  subword=int(r.choice([27, 24, 21, 18, 15])), centered=bool(r.integers(0, 2)), coset=int(r.integers(0, 3)))
        oc, gc = both(kw)
        for arith in (0, 1):
            h27, c52 = codec.header_emit(gc, arith)
            hp = oracle.header_pack(oc)
            assert np.array_equal(h27, hp)
            a = oracle.rs_encode_blocks(18, hp[:18], arith)[0]
            b = oracle.rs_encode_blocks(18, np.concatenate([hp[18:], np.zeros(9, np.uint8)]), arith)[0]
            assert np.array_equal(c52, np.concatenate([a, b]))
        # parse: true RS(26,18) codewords (+ up to 4 symbol errors in FIXED mode)
        _, c52 = codec.header_emit(gc, 1)
        words = np.concatenate([c52, np.zeros(2, np.uint8)]).reshape(6, 9)
        ok, got = codec.header_parse(words, arith=0)
        want = oracle.header_unpack(oracle.header_pack(oc))[0]
        assert ok and bytes(got)[:40] == bytes(want)[:40]
        bad = words.copy().reshape(-1)
        for p in r.choice(26, size=4, replace=False):
            bad[p] = (bad[p] + 1 + int(r.integers(0, 25))) % 27
        ok2, got2 = codec.header_parse(bad.reshape(6, 9), arith=1)
        assert ok2 and bytes(got2)[:40] == bytes(want)[:40]
    ok, _ = codec.header_parse(np.zeros((5, 9), np.uint8))
    assert not ok


# ------------------------------------------------------------------ profile encoder
@pytest.mark.parametrize("ci", range(len(CONFIGS)))
def test_encode_profile(codec, oracle, t3, ci):
    oc, gc = both(CONFIGS[ci])
    r = rng(1000 + ci)
    for n in (0, 1, 2, 3, 5, 26, 27, 64, 777, 1000, 8192, 30011):
        raw = r.integers(0, 27, size=(n, 9), dtype=np.uint8)
        if n in (64, 8192):  # out-of-alphabet bytes read as their low three trits (small: general kernels, large: tiled kernels)
            raw = r.integers(0, 256, size=(n, 9), dtype=np.uint8)
        for arith in (t3.REF_EXACT, t3.FIXED):
            got = codec.encode_profile_from_raw(raw, gc, arith)
            want = oracle.encode_profile(oc, raw, arith)
            assert got.shape == want.shape and np.array_equal(got, want), (ci, n, arith)


@pytest.mark.parametrize("name", sorted(k[4:] for k in G.files if k.startswith("cfg_")))
def test_golden_pipeline(codec, t3, name):
    gc = t3.Config.from_buffer_copy(G["cfg_" + name].tobytes())
    raw = G["pipe_raw"]
    assert np.array_equal(codec.encode_profile_from_raw(raw, gc, 0), G["enc_ref_" + name])
    assert np.array_equal(codec.encode_profile_from_raw(raw, gc, 1), G["enc_fix_" + name])
    assert np.array_equal(codec.encode_frames_rgb8(G["rgb"], gc, 0)[0], G["encrgb_ref_" + name])
    codec.cfg_last_seen = t3.make_config()
    ok, words = codec.decode_profile_to_raw(G["dec_ref_in_" + name])
    assert ok == bool(G["dec_ref_ok_" + name][0])
    assert np.array_equal(words, G["dec_ref_out_" + name])
    assert bytes(codec.cfg_last_seen) == G["dec_ref_seen_" + name].tobytes()
    ok, out, nc = codec.decode_profile_fixed(G["enc_fix_" + name], gc, n_raw_words=raw.shape[0])
    assert ok and nc == 0 and np.array_equal(out, raw[:out.shape[0]]) and out.shape[0] > 700


# ------------------------------------------------------------------ reference decoder as shipped
def valid_header_stream(oracle, oc, body_words, rnd):
    hp = oracle.header_pack(oc)
    a = oracle.rs_encode_blocks(18, hp[:18], 1)[0]
    b = oracle.rs_encode_blocks(18, np.concatenate([hp[18:], np.zeros(9, np.uint8)]), 1)[0]
    head = np.concatenate([a, b, rnd.integers(0, 27, 2, dtype=np.uint8)])
    return np.concatenate([head.reshape(6, 9), body_words])


@pytest.mark.parametrize("ci", range(len(CONFIGS) - 1))
def test_decode_profile_ref_exact(codec, oracle, t3, ci):
    oc, gc = both(CONFIGS[ci])
    r = rng(2000 + ci)

    def check(stream, seen_kw=None):
        codec.cfg_last_seen = t3.make_config(**(seen_kw or {}))
        ok_g, out_g = codec.decode_profile_to_raw(stream)
        ok_o, out_o, seen_o = oracle.decode_profile_ref(T.make_cfg(**(seen_kw or {})), stream)
        assert ok_g == ok_o
        assert out_g.shape == out_o.shape and np.array_equal(out_g, out_o)
        assert bytes(codec.cfg_last_seen) == bytes(seen_o)
        return ok_g

    raw = r.integers(0, 27, size=(500, 9), dtype=np.uint8)
    assert check(oracle.encode_profile(oc, raw, 0)) is False        # (i) the shipped encoder's output: rejected at the header
    n_true = 0
    for trial, nbody in enumerate((0, 5, 26, 27, 130, 260, 263, 2600)):
        body = r.integers(0, 27, size=(nbody, 9), dtype=np.uint8)
        if trial >= 3:
            body = oracle.encode_profile(oc, r.integers(0, 27, size=(3 * nbody, 9), dtype=np.uint8), 1)[6:6 + nbody]
        n_true += check(valid_header_stream(oracle, oc, body, r))    # (ii) valid header: the body path executes
    for n in (0, 3, 6, 40):
        check(r.integers(0, 27, size=(n, 9), dtype=np.uint8))       # (iii) random words / short inputs
    assert check(raw, dict(profile=T.RAW_MODE))                     # RAW passthrough keyed on the previous header
    if ci == 4:
        assert n_true > 0


# ------------------------------------------------------------------ FIXED: consistent decode
@pytest.mark.parametrize("ci", range(len(CONFIGS) - 1))
def test_fixed_roundtrip_with_errors(codec, oracle, t3, ci):
    oc, gc = both(CONFIGS[ci])
    r = rng(3000 + ci)
    add = T.gf_add_table()
    for n in (0, 1, 3, 64, 777, 4000, 30011):
        raw = r.integers(0, 27, size=(n, 9), dtype=np.uint8)
        raw[:, 8] %= 9
        enc = codec.encode_profile_from_raw(raw, gc, t3.FIXED)
        ok, out, nc = codec.decode_profile_fixed(enc, gc, n_raw_words=n)
        ok_o, out_o, nc_o = oracle.decode_profile_fixed(oc, enc, n_raw_words=n)
        assert ok and ok_o and nc == nc_o == 0
        assert np.array_equal(out, out_o) and np.array_equal(out, raw[:out.shape[0]])
        if not (gc.profile == 4 and gc.tile_w and gc.tile_h):
            ok2, out2, _ = codec.decode_profile_fixed(enc, gc, n_raw_words=0)
            assert ok2 and np.array_equal(out2, out)
        if n >= 64 and not (gc.beacon_enabled and gc.beacon_slot > 8):
            for exact_t in (False, True):
                bad, nerr = T.inject_errors(enc, oc, n, seed=3 + exact_t, gf_add=add, exact_t=exact_t)
                ok3, out3, nc3 = codec.decode_profile_fixed(bad, gc, n_raw_words=n)
                assert ok3 and np.array_equal(out3, out) and nc3 == nerr and nerr > 0
            # more than t errors in one codeword: same verdict and output as the oracle
            worse = bad.copy().reshape(-1)
            worse[52:52 + 12] = (worse[52:52 + 12] + 1) % 27
            ok4, out4, _ = codec.decode_profile_fixed(worse.reshape(-1, 9), gc, n_raw_words=n)
            ok5, out5, _ = oracle.decode_profile_fixed(oc, worse.reshape(-1, 9), n_raw_words=n)
            assert ok4 == ok5 and np.array_equal(out4, out5)


# ------------------------------------------------------------------ fused frames
def test_decode_frames_more_than_one_mailbox(codec, oracle, t3):
    """the host-buffer decode call takes any number of frames (its status mailbox holds 32: longer batches go through in pieces)"""
    oc, gc = both(CONFIGS[1])
    shape, F = (40, 27), 70
    n_px = shape[0] * shape[1]
    frames = np.stack([T.synth_rgb(100 + f, n_px) for f in range(F)])
    enc = codec.encode_frames_rgb8(frames, gc, t3.FIXED)
    add = T.gf_add_table()
    bad, tot = enc.copy(), 0
    for f in (0, 31, 32, 33, 69):
        bad[f], ne = T.inject_errors(enc[f], oc, (n_px + 1) // 2, seed=5 + f, gf_add=add)
        tot += ne
    ok, rgb, nc = codec.decode_frames_rgb8(bad, n_px, gc)
    assert ok.all() and nc == tot
    for f in range(F):
        assert np.array_equal(rgb[f], oracle.decode_rgb_fixed(oc, enc[f], n_px)[1]), f


@pytest.mark.parametrize("ci", [0, 1, 2, 3, 4, 5, 7, 9, 11, 12])
@pytest.mark.parametrize("shape", [(64, 64), (512, 512), (130, 77)])
def test_fused_frames_rgb8(codec, oracle, t3, ci, shape):
    oc, gc = both(CONFIGS[ci])
    n_px = shape[0] * shape[1]
    frames = np.stack([T.synth_rgb(1, n_px), T.synth_checker(shape[0], shape[1]), T.synth_rgb(9, n_px)])
    add = T.gf_add_table()
    for arith in (t3.REF_EXACT, t3.FIXED):
        got = codec.encode_frames_rgb8(frames, gc, arith)
        for f in range(frames.shape[0]):
            assert np.array_equal(got[f], oracle.encode_rgb(oc, frames[f], arith)), (ci, shape, arith, f)
    enc = codec.encode_frames_rgb8(frames, gc, t3.FIXED)
    ok, rgb, nc = codec.decode_frames_rgb8(enc, n_px, gc)
    assert ok.all() and nc == 0
    for f in range(frames.shape[0]):
        ok_o, rgb_o, _ = oracle.decode_rgb_fixed(oc, enc[f], n_px)
        assert ok_o and np.array_equal(rgb[f], rgb_o)
        assert np.array_equal(rgb[f], oracle.quant_to_rgb(oracle.rgb_to_quant(frames[f]))[:rgb.shape[1]])
    if not (gc.beacon_enabled and gc.beacon_slot > 8):
        bad = enc.copy()
        tot = 0
        for f in range(frames.shape[0]):
            bad[f], ne = T.inject_errors(enc[f], oc, (n_px + 1) // 2, seed=11 + f, gf_add=add)
            tot += ne
        ok2, rgb2, nc2 = codec.decode_frames_rgb8(bad, n_px, gc)
        assert ok2.all() and nc2 == tot and np.array_equal(rgb2, rgb)


# ------------------------------------------------------------------ BASELINE.json full sizes (8K)
def _dev_roundtrip_8k(codec, t3, gc, n_frames=1, corrupt=None):
    """device-resident 8K encode -> (optional corruption) -> decode; returns torch tensors"""
    import torch
    n_px = 7680 * 4320
    wpf = t3.profile_words(gc, n_px // 2)
    stride = (wpf + 15) & ~15
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(2)
    rgb = torch.randint(0, 256, (n_frames, n_px * 3), dtype=torch.uint8, device=dev, generator=g)
    enc = torch.zeros(n_frames, stride * 9, dtype=torch.uint8, device=dev)
    back = torch.zeros(n_frames, n_px * 3, dtype=torch.uint8, device=dev)
    status = torch.zeros(2 * n_frames, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    codec.encode_frames_rgb8_dev(rgb, n_px, n_frames, enc, stride, gc, t3.FIXED, s)
    if corrupt is not None:
        corrupt(enc, wpf)
    codec.decode_frames_rgb8_dev(enc, wpf, stride, n_frames, n_px, back, status, gc, s)
    q = torch.empty(n_frames, n_px * 6, dtype=torch.uint8, device=dev)
    want = torch.empty_like(back)
    for f in range(n_frames):
        codec.rgb_to_quant_dev(rgb[f], n_px, q[f], s)
        codec.quant_to_rgb_dev(q[f], n_px, want[f], s)
    torch.cuda.synchronize()
    return rgb, enc, back, want, status.cpu().numpy(), wpf


def test_8k_rs26_20_encode_matches_oracle_on_slices_and_roundtrips(codec, oracle, t3):
    """BASELINE configs[1]: one 8K RGB8 frame, RS(26,20), 1D.  Whole-frame properties on the device
    (decode(encode(x)) == dequant(quant(x)), clean flags, zero corrections) and oracle parity on slices of
    the frame's head, which pins the tile/offset arithmetic at full size."""
    import torch
    kw = dict(profile=T.P3, uep=2)
    oc, gc = both(kw)
    rgb, enc, back, want, st, wpf = _dev_roundtrip_8k(codec, t3, gc)
    assert wpf == 20766726 and list(st) == [1, 0]
    assert torch.equal(back, want)
    # header + first codewords of every band vs the oracle's encode of the whole frame is too slow for the
    # CPU; the wire format is band-major, so encode a 1/64 frame prefix on the CPU and compare the codewords
    # it fully determines: codeword c of band b depends on stream symbols < 9*k*(c+1) only.
    n_sub = 7680 * 4320 // 64
    sub = oracle.encode_rgb(oc, rgb[0, :3 * n_sub].cpu().numpy().reshape(-1, 3), 1).reshape(-1)
    full = enc[0].cpu().numpy()
    assert np.array_equal(full[:52], sub[:52])                       # same header
    ncw_sub = (sub.size - 52) // 26 // 9                             # codewords per band in the sub-frame
    ncw_full = 798720
    for b in range(9):
        a = full[52 + 26 * ncw_full * b: 52 + 26 * (ncw_full * b + ncw_sub - 2)]
        w = sub[52 + 26 * ncw_sub * b: 52 + 26 * (ncw_sub * b + ncw_sub - 2)]
        # scrambler phase differs between the two layouts (body offsets differ): compare descrambled symbols
        pa = (np.arange(a.size) + 26 * ncw_full * b) % 3
        pw = (np.arange(w.size) + 26 * ncw_sub * b) % 3
        st_tab = np.array([2, 0, 1])  # seed {1,1,1}: st_p = (p+2) % 3
        sub_tab = T.gf_add_table()
        neg13 = {0: 0, 1: 26, 2: 13}
        da = np.array([sub_tab[x, neg13[int(s)]] for x, s in zip(a[:2600], st_tab[pa[:2600]])])
        dw = np.array([sub_tab[x, neg13[int(s)]] for x, s in zip(w[:2600], st_tab[pw[:2600]])])
        assert np.array_equal(da, dw), b


def test_8k_injected_errors_are_corrected(codec, t3):
    """errors up to t=3 in a spread of codewords of an 8K RS(26,20) frame: same pixels as the clean decode"""
    import torch
    kw = dict(profile=T.P3, uep=2)
    _, gc = both(kw)
    n_bad = 200000

    def corrupt(enc, wpf):
        gen = torch.Generator(device=enc.device)
        gen.manual_seed(7)
        cw = torch.randperm(7188480, device=enc.device, generator=gen)[:n_bad]
        for j in range(3):                                  # three distinct positions per chosen codeword
            pos = 52 + 26 * cw + (5 + 7 * j)
            enc[0, pos] = (enc[0, pos] + 1 + j) % 27        # any other alphabet value is a symbol error
    rgb, enc, back, want, st, wpf = _dev_roundtrip_8k(codec, t3, gc, corrupt=corrupt)
    assert st[0] == 1 and st[1] == 3 * n_bad
    assert torch.equal(back, want)


# ------------------------------------------------------------------ BASELINE.json full sizes, whole frame against the oracle
# One 8K frame costs the single-threaded C oracle ~20 s to encode and ~25 s to decode; the oracle calls of a test run side by
# side on host threads (ctypes releases the GIL).  `slow`, but inside -m gpu: these are the parity tests proper for
# BASELINE configs 1 and 2 -- every byte of every codeword of the frame, not a prefix.
N8K = 7680 * 4320
CFG1 = dict(profile=T.P3, uep=2)
CFG2 = dict(profile=T.P5, tile=(26, 26), beacon=(26, 2, True), uep=T.UEP_LUMA, seed=(2, 1, 1), coset=1)


def _pool(jobs):
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        futs = [ex.submit(j) for j in jobs]
        return [f.result() for f in futs]


@pytest.fixture(scope="module")
def frame8k():
    return T.synth_rgb(2, N8K)                                   # SURVEY 8(d) config 1 input: seed 2


def _whole_frame_encode(codec, oracle, t3, kw, rgb):
    oc, gc = both(kw)
    got = [codec.encode_frames_rgb8(rgb[None], gc, a)[0] for a in (t3.REF_EXACT, t3.FIXED)]
    oracle.rs_gen(20)  # the oracle's only global state is its lazily built GF(27) tables: built before the threads start
    want = _pool([lambda: oracle.encode_rgb(oc, rgb, 0), lambda: oracle.encode_rgb(oc, rgb, 1)])
    for g_, w_, name in zip(got, want, ("REF-EXACT", "FIXED")):
        assert g_.shape == w_.shape, name
        assert np.array_equal(g_, w_), name
    return got[1]


def _whole_frame_decode(codec, oracle, t3, kw, enc_fixed):
    oc, gc = both(kw)
    add = T.gf_add_table()
    bad_some, n_some = T.inject_errors(enc_fixed, oc, N8K // 2, seed=3, gf_add=add)               # e_c = h(3, c) % (t + 1)
    bad_all, n_all = T.inject_errors(enc_fixed, oc, N8K // 2, seed=5, gf_add=add, exact_t=True)   # exactly t in every codeword
    oracle.rs_gen(20)
    want = _pool([lambda: oracle.decode_rgb_fixed(oc, bad_some, N8K), lambda: oracle.decode_rgb_fixed(oc, bad_all, N8K)])
    for bad, nerr, (ok_o, rgb_o, nc_o), name in ((bad_some, n_some, want[0], "0..t"), (bad_all, n_all, want[1], "t")):
        ok, rgb, nc = codec.decode_frames_rgb8(bad[None], N8K, gc)
        assert ok_o and ok.all(), name
        assert nc == nc_o == nerr, (name, nc, nc_o, nerr)
        assert rgb.shape[1] == rgb_o.shape[0] and np.array_equal(rgb[0], rgb_o), name


@pytest.mark.slow
def test_8k_config1_whole_frame_matches_oracle(codec, oracle, t3, frame8k):
    """BASELINE configs[1] (8K RGB8, RS(26,20), 1D) through k_encode_v5 / k_decode_v5: the whole frame byte-exact against the oracle's
    encode in both arithmetic modes (OLD:1043-1169), and the whole-frame decode of the FIXED encoding with injected symbol errors
    (0..t per codeword, and exactly t in every codeword) against the oracle's consistent decoder: pixels, ok flag, corrected count."""
    assert t3.fast_path_available(both(CFG1)[1])
    enc = _whole_frame_encode(codec, oracle, t3, CFG1, frame8k)
    assert enc.shape[0] == 20766726
    _whole_frame_decode(codec, oracle, t3, CFG1, enc)


@pytest.mark.slow
def test_8k_config2_whole_frame_matches_oracle(codec, oracle, t3, frame8k):
    """BASELINE configs[2] (8K, 2D 26x26 boustrophedon + luma-priority UEP + coset C1 + beacon(26,2)) through k_encode_super /
    k_decode_super (+ the general kernels on the ragged end): whole-frame encode parity in both arithmetic modes and whole-frame
    decode with injected errors up to t per codeword, against the oracle."""
    gc = both(CFG2)[1]
    assert not t3.fast_path_available(gc) and t3.super_path_available(gc)
    enc = _whole_frame_encode(codec, oracle, t3, CFG2, frame8k)
    _whole_frame_decode(codec, oracle, t3, CFG2, enc)


def test_8k_config2_device_roundtrip_clean(codec, t3):
    """BASELINE configs[2], device-resident clean round trip (no errors, no oracle: decode(encode(x)) == dequant(quant(x)) on the
    pixels the encoder keeps); the oracle comparison with injected errors is test_8k_config2_whole_frame_matches_oracle."""
    import torch
    kw = dict(profile=T.P5, tile=(26, 26), beacon=(26, 2, True), uep=T.UEP_LUMA, seed=(2, 1, 1), coset=1)
    _, gc = both(kw)
    assert not t3.fast_path_available(gc)
    rgb, enc, back, want, st, wpf = _dev_roundtrip_8k(codec, t3, gc)
    assert st[0] == 1 and st[1] == 0
    npx_ok = 7680 * 4320 - 2000                              # the encoder drops < k symbols per band (bug B8)
    assert torch.equal(back[0, :3 * npx_ok], want[0, :3 * npx_ok])


def test_8k_raw_mode_pack_unpack(codec, t3):
    """BASELINE configs[3]: RAW mode, 8K PixelYCbCrQuant <-> Word27 (involution on valid pixels)"""
    import torch
    n_px = 7680 * 4320
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    px = torch.empty(n_px, 3, dtype=torch.int16, device=dev)
    px[:, 0] = torch.randint(0, 243, (n_px,), device=dev, generator=g, dtype=torch.int16)
    px[:, 1:] = torch.randint(-40, 41, (n_px, 2), device=dev, generator=g, dtype=torch.int16)
    words = torch.empty(n_px // 2 * 9, dtype=torch.uint8, device=dev)
    back = torch.empty_like(px)
    s = torch.cuda.current_stream().cuda_stream
    codec.pack_pixels_dev(px, n_px, words, s)
    codec.unpack_pixels_dev(words, n_px // 2, back, s)
    torch.cuda.synchronize()
    assert torch.equal(px, back) and int(words.max()) <= 26
    assert int(words.view(-1, 9)[:, 8].max()) <= 8           # T[26] = 0


def test_multi_frame_stream_device(codec, oracle, t3):
    """configs[4] in miniature: a stream of frames encoded in one batched call == per-frame oracle encodes"""
    import torch
    kw = dict(profile=T.P3, uep=2)
    oc, gc = both(kw)
    n_px, F = 640 * 360, 6
    frames = np.stack([T.synth_rgb(5 + f, n_px) for f in range(F)])
    got = codec.encode_frames_rgb8(frames, gc, t3.REF_EXACT)
    for f in range(F):
        assert np.array_equal(got[f], oracle.encode_rgb(oc, frames[f], 0))


@pytest.mark.parametrize("stride_pad,kw,shape,F", [(0, dict(profile=T.P3, uep=2), (1920, 1080), 3), (16, dict(profile=T.P3, uep=2), (1920, 1080), 3),
                                                   (3, dict(profile=T.P3, uep=2), (1920, 1080), 3), (0, dict(profile=T.P1, uep=0), (3840, 2160), 2)])
def test_fused_regular_frames_tensor_copies(codec, oracle, t3, stride_pad, kw, shape, F):
    """1080p frames at k = 20 and 4K frames at k = 24 are `regular` (every band holds the same number of codewords, band pitch a multiple of 16 bytes): the v5
    kernels then move the nine runs of a mini-tile by one 3-D tensor copy each way (UTMASTG / UTMALDG).  Three frames in one device
    call -- frame strides that keep the batch regular (0, 16 words) and one that does not (3 words: the per-run bulk copies) -- and the
    chunked host pipelines (tile ranges that start inside a frame): encode parity with the oracle, t errors per codeword corrected,
    decode parity."""
    import torch
    oc, gc = both(kw)
    n_px = shape[0] * shape[1]
    frames = np.stack([T.synth_rgb(70 + f, n_px) for f in range(F)])
    want = [oracle.encode_rgb(oc, frames[f], t3.FIXED) for f in range(F)]
    wpf = t3.profile_words(gc, n_px // 2)
    assert want[0].shape[0] == wpf
    stride = ((wpf + 15) & ~15) + stride_pad
    dev = torch.device("cuda", 0)
    rgb = torch.from_numpy(frames.reshape(F, -1)).to(dev)
    enc = torch.zeros(F, stride * 9, dtype=torch.uint8, device=dev)
    back = torch.zeros(F, n_px * 3, dtype=torch.uint8, device=dev)
    status = torch.zeros(2 * F, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    codec.encode_frames_rgb8_dev(rgb, n_px, F, enc, stride, gc, t3.FIXED, s)
    torch.cuda.synchronize()
    got = enc.cpu().numpy()
    for f in range(F):
        assert np.array_equal(got[f, :wpf * 9].reshape(wpf, 9), want[f]), (stride_pad, f)
        assert not got[f, wpf * 9:].any()                       # nothing written between the frames
    add = T.gf_add_table()
    bad = got.copy()
    tot = 0
    for f in range(F):
        b, ne = T.inject_errors(want[f], oc, n_px // 2, seed=3 + f, gf_add=add)
        bad[f, :wpf * 9] = b.reshape(-1)
        tot += ne
    codec.decode_frames_rgb8_dev(torch.from_numpy(bad).to(dev), wpf, stride, F, n_px, back, status, gc, s)
    torch.cuda.synchronize()
    st = status.cpu().numpy()
    assert all(st[2 * f] == 1 for f in range(F)) and int(st[1::2].sum()) == tot
    rgb_back = back.cpu().numpy().reshape(F, n_px, 3)
    for f in range(F):
        ok_o, rgb_o, _ = oracle.decode_rgb_fixed(oc, want[f], n_px)
        assert ok_o and np.array_equal(rgb_back[f], rgb_o), (stride_pad, f)
    if stride_pad == 0:                                          # the host-buffer calls: chunked pipelines, tile ranges inside a frame
        h = codec.encode_frames_rgb8(frames, gc, t3.FIXED)
        for f in range(F):
            assert np.array_equal(h[f], want[f]), f
        ok, rgb_h, nc = codec.decode_frames_rgb8(h, n_px, gc)
        assert ok.all() and nc == 0 and np.array_equal(rgb_h, rgb_back)


@pytest.mark.parametrize("uep", [2, 1, 0])
def test_fused_fast_path_all_2_24_colours_match_general_kernels(codec, t3, uep):
    """Every RGB colour once (a 4096x4096 frame) through the fused tiled kernels, against the stage-by-stage
    general kernels on the device (those are pinned to the oracle exhaustively by test_bridge_exhaustive_2_24
    and test_encode_profile): identical profile words, identical decoded RGB."""
    import torch
    dev = torch.device("cuda", 0)
    s = torch.cuda.current_stream().cuda_stream
    _, gc = both(dict(profile=T.P3, uep=uep))
    assert t3.fast_path_available(gc)
    n_px = 1 << 24
    idx = torch.arange(n_px, dtype=torch.int32, device=dev)
    rgb = torch.stack([(idx >> 16) & 255, (idx >> 8) & 255, idx & 255], dim=1).to(torch.uint8).contiguous().view(-1)
    wpf = t3.profile_words(gc, n_px // 2)
    enc = torch.zeros(wpf * 9, dtype=torch.uint8, device=dev)
    codec.encode_frames_rgb8_dev(rgb, n_px, 1, enc, wpf, gc, t3.FIXED, s)
    q = torch.empty(n_px * 6, dtype=torch.uint8, device=dev)
    raw = torch.empty(n_px // 2 * 9, dtype=torch.uint8, device=dev)
    enc2 = torch.zeros_like(enc)
    codec.rgb_to_quant_dev(rgb, n_px, q, s)
    codec.pack_pixels_dev(q, n_px, raw, s)
    codec.encode_profile_dev(raw, n_px // 2, enc2, wpf, gc, t3.FIXED, s)
    torch.cuda.synchronize()
    assert torch.equal(enc, enc2)
    back = torch.zeros(n_px * 3, dtype=torch.uint8, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    codec.decode_frames_rgb8_dev(enc, wpf, wpf, 1, n_px, back, status, gc, s)
    want = torch.zeros_like(back)
    codec.quant_to_rgb_dev(q, n_px, want, s)
    torch.cuda.synchronize()
    assert status.tolist() == [1, 0]
    n_ok = 3 * (n_px - 2000)  # the encoder drops < k symbols per band at the very end (bug B8)
    assert torch.equal(back[:n_ok], want[:n_ok])


def test_fused_fast_path_out_of_alphabet_and_uncorrectable(codec, oracle, t3):
    """bytes >= 27 in a received stream read as their low three trits (unpack3, OLD:28-31), and a codeword with
    more than t errors makes the frame fail: same verdicts and pixels as the oracle's consistent decoder"""
    oc, gc = both(dict(profile=T.P3, uep=2))
    n_px = 4 * 540 + 77
    rgb = T.synth_rgb(5, n_px)
    enc = codec.encode_frames_rgb8(rgb[None], gc, t3.FIXED)[0]
    wild = enc.copy().reshape(-1)
    r = rng(3)
    pos = 52 + r.choice(wild.size - 60, 300, replace=False)
    wild[pos] = wild[pos] + 27 * r.integers(1, 9, pos.size).astype(np.uint8)   # same symbol mod 27
    ok, back, nc = codec.decode_frames_rgb8(wild.reshape(1, -1, 9), n_px, gc)
    ok_o, back_o, nc_o = oracle.decode_rgb_fixed(oc, wild.reshape(-1, 9), n_px)
    assert ok.all() and ok_o and nc == nc_o == 0 and np.array_equal(back[0], back_o)
    bad = enc.copy().reshape(-1)
    bad[52 + 26 * 7: 52 + 26 * 7 + 5] = (bad[52 + 26 * 7: 52 + 26 * 7 + 5] + 1) % 27     # 5 errors > t = 3
    ok2, _, _ = codec.decode_frames_rgb8(bad.reshape(1, -1, 9), n_px, gc)
    ok2_o, _, _ = oracle.decode_rgb_fixed(oc, bad.reshape(-1, 9), n_px)
    assert bool(ok2[0]) == bool(ok2_o)


@pytest.mark.parametrize("kw", [dict(profile=T.P3, uep=2), dict(profile=T.P2, uep=1), dict(profile=T.P1, uep=0), dict(profile=T.P4, uep=3),
                                dict(profile=T.P2, uep=T.UEP_LUMA)])
def test_consistent_decoder_rejects_more_than_t_errors(codec, oracle, t3, kw):
    """t + 1 symbol errors in a codeword: decode_block (OLD:611-624) would accept most such blocks with 0..t bogus corrections; the
    consistent decoder (ours, no reference to match) accepts a block only when Berlekamp-Massey's L <= t equals the number of distinct
    roots of sigma.  Frames that each hold ONE such codeword (in band 0): same verdict as the oracle frame by frame, and the
    share of rejected frames is what bounded-distance decoding predicts."""
    oc, gc = both(kw)
    n_px = 8 * 5940 + 3
    rgb = T.synth_rgb(17, n_px)
    enc = codec.encode_frames_rgb8(rgb[None], gc, t3.FIXED)[0]
    add = T.gf_add_table()
    k0 = T.K_OF_UEP[oc.uep[0] % 4]
    t0 = (26 - k0) // 2
    n_s = (26 * ((n_px + 1) // 2) + 2) // 3
    ncw0 = ((n_s + 8) // 9) // k0                                  # codewords of band 0: body symbols [0, 26 * ncw0)
    r = rng(23)
    n_frames, rejected = 24, 0
    bad = np.repeat(enc[None], n_frames, axis=0)
    for f in range(n_frames):
        flat = bad[f].reshape(-1)
        start = 52 + 26 * int(r.integers(0, ncw0))
        for q in start + r.choice(26, t0 + 1, replace=False):
            flat[q] = add[flat[q] % 27, int(r.integers(1, 27))]
    ok, back, _ = codec.decode_frames_rgb8(bad, n_px, gc)
    for f in range(n_frames):
        ok_o, back_o, _ = oracle.decode_rgb_fixed(oc, bad[f], n_px)
        assert bool(ok[f]) == bool(ok_o), f
        if ok_o:
            assert np.array_equal(back[f], back_o), f
        rejected += not ok_o
    # a block with t + 1 errors decodes (to a wrong codeword) when its syndrome is one of the sum_{e<=t} C(26,e) 26^e correctable ones
    # of 27^r: 93 % for k = 24, 41 % for k = 22, 12 % for k = 20, 2.4 % for k = 18
    assert rejected >= {24: 0, 22: 6, 20: 15, 18: 19}[k0], rejected


@pytest.mark.parametrize("kw", [dict(profile=T.P3, uep=2), dict(profile=T.P2, uep=1)])
def test_host_pipeline_chunked_frames_match_oracle(codec, oracle, t3, kw):
    """frames large enough (>= 512 full mini-tiles) for the chunked H2D / kernel / D2H pipeline of the host-buffer
    calls: same words as the oracle, round trip with injected errors, odd pixel count and two frames"""
    oc, gc = both(kw)
    n_px = 640 * 480 + 1
    frames = np.stack([T.synth_rgb(21, n_px), T.synth_rgb(22, n_px)])
    for arith in (t3.REF_EXACT, t3.FIXED):
        got = codec.encode_frames_rgb8(frames, gc, arith)
        for f in range(2):
            assert np.array_equal(got[f], oracle.encode_rgb(oc, frames[f], arith)), (arith, f)
    enc = codec.encode_frames_rgb8(frames, gc, t3.FIXED)
    add = T.gf_add_table()
    bad = enc.copy()
    tot = 0
    for f in range(2):
        bad[f], ne = T.inject_errors(enc[f], oc, (n_px + 1) // 2, seed=31 + f, gf_add=add)
        tot += ne
    ok, rgb, nc = codec.decode_frames_rgb8(bad, n_px, gc)
    assert ok.all() and nc == tot
    for f in range(2):
        ok_o, rgb_o, _ = oracle.decode_rgb_fixed(oc, enc[f], n_px)
        assert ok_o and np.array_equal(rgb[f], rgb_o)


@pytest.mark.parametrize("kw,n_px,nf", [(dict(profile=T.P3, uep=2), 540 * 70 + 33, 2), (dict(profile=T.P2, uep=1), 594 * 40, 1),
                                        (dict(profile=T.P1, uep=0), 648 * 35 + 1, 3), (dict(profile=T.P4, uep=3), 486 * 9 + 5, 2)])
def test_fused_multi_frame_odd_sizes(codec, oracle, t3, kw, n_px, nf):
    """several frames of odd pixel count in one call: frames start on odd byte offsets (RGB) -- every alignment of the
    tiled kernels' 18-byte pixel units and of the nine body runs; encode parity, error correction, decode parity"""
    oc, gc = both(kw)
    frames = np.stack([T.synth_rgb(40 + f, n_px) for f in range(nf)])
    for arith in (t3.REF_EXACT, t3.FIXED):
        enc = codec.encode_frames_rgb8(frames, gc, arith)
        for f in range(nf):
            assert np.array_equal(enc[f], oracle.encode_rgb(oc, frames[f], arith)), (arith, f)
    add = T.gf_add_table()
    bad = enc.copy()
    tot = 0
    for f in range(nf):
        bad[f], ne = T.inject_errors(enc[f], oc, (n_px + 1) // 2, seed=5 + f, gf_add=add)
        tot += ne
    ok, rgb, nc = codec.decode_frames_rgb8(bad, n_px, gc)
    assert ok.all() and nc == tot
    for f in range(nf):
        ok_o, rgb_o, _ = oracle.decode_rgb_fixed(oc, enc[f], n_px)
        assert ok_o and np.array_equal(rgb[f], rgb_o)


def test_two_devices_in_one_process(oracle, t3):
    """one context per GPU inside one process (t3c_shim::context(device) in the C++ headers): kernel attributes are per device"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    oc, gc = both(dict(profile=T.P3, uep=2))
    n_px = 540 * 300 + 7
    rgb = T.synth_rgb(77, n_px)
    want = oracle.encode_rgb(oc, rgb, 1)
    px = T.synth_quant(3, 40000)
    for dev in (1, 0, 1):
        c = t3.Codec(dev)
        assert np.array_equal(c.encode_frames_rgb8(rgb[None], gc, t3.FIXED)[0], want)
        ok, back, _ = c.decode_frames_rgb8(want[None], n_px, gc)
        assert ok.all() and np.array_equal(back[0], oracle.decode_rgb_fixed(oc, want, n_px)[1])
        assert np.array_equal(c.encode_raw_pixels_to_words(px), oracle.pack_pixels(px))
        c.close()


# ------------------------------------------------------------------ SURVEY 8(f).2 / 8(f).3
def test_words_to_base243_every_subword_length(codec, oracle):
    """the fused kernel is a template on N (20 words per thread, compile-time trit pieces): every N, sizes around the 20-word groups and
    the 128-thread CTAs, bytes >= 27 (unpack3 reduces them mod 27, OLD:28-31)"""
    r = rng(77)
    for N in range(1, 28):
        for nw in (19, 20, 21, 2559, 2560, 2561 + 20 * N):
            words = r.integers(0, 256 if nw % 2 else 27, size=(nw, 9), dtype=np.uint8)
            assert np.array_equal(codec.words_to_base243(words, N), oracle.base243_pack(oracle.subword_stream(words, N))), (N, nw)


@pytest.mark.parametrize("N", (27, 24, 21, 18, 15, 1, 5))
def test_subword_streams_and_base243(codec, oracle, N):
    r = rng(600 + N)
    for nw in (0, 1, 3, 50, 4097, 100003):
        words = r.integers(0, 256, size=(nw, 9), dtype=np.uint8)
        a = codec.extract_subword_stream_from_words(words, N)
        assert np.array_equal(a, oracle.subword_stream(words, N))
        assert np.array_equal(codec.words_to_base243(words, N), oracle.base243_pack(a))
        for n in sorted({0, 1, N, N + 1, 5 * N + 3, a.size}):
            if n > a.size:
                continue
            t = a[:n].copy()
            if n > 10:
                t[r.integers(0, n, 5)] = r.integers(3, 256, 5)          # bytes that are not trits wrap like pack3 / the byte cast
            for fill in (0, 2, 200):
                assert np.array_equal(codec.build_words_from_subword_stream(t, N, fill), oracle.words_from_subword_stream(t, N, fill))
            p = codec.ut_to_base243(t)
            assert np.array_equal(p, oracle.base243_pack(t))
            for blob in (p, p[:max(0, p.size - 1)], p[:2], np.concatenate([p, np.array([7, 200], np.uint8)])):
                ok_g, u_g = codec.base243_to_ut(blob)
                ok_o, u_o = oracle.base243_unpack(blob)
                assert ok_g == ok_o and np.array_equal(u_g, u_o)


def test_v6new_raw_path(codec, oracle):
    r = rng(700)
    for n in (0, 1, 3, 4, 5, 4099, 200001):
        px = np.zeros(n, T.PIXEL_DTYPE)
        px["Yq"], px["Cbq"], px["Crq"] = r.integers(0, 65536, n), r.integers(-32768, 32768, n), r.integers(-32768, 32768, n)
        px[: n // 2] = T.synth_quant(9, n // 2)
        ok, w = codec.v6new_encode_raw_pixels_to_words(px)
        assert ok and np.array_equal(w, oracle.v6new_pack_pixels(px))
        wild = r.integers(0, 2 ** 32, n, dtype=np.uint32)
        for words in (w, wild):
            ok, p = codec.v6new_decode_raw_words_to_pixels(words)
            assert ok and np.array_equal(p.view(np.uint8), oracle.v6new_unpack_pixels(words).view(np.uint8))
    assert codec.v6new_encode_raw_pixels_to_words(px, 24)[0] and not codec.v6new_encode_raw_pixels_to_words(px, 7)[0]
    assert codec.v6new_decode_raw_words_to_pixels(w, 15)[0] and not codec.v6new_decode_raw_words_to_pixels(w, 26)[0]


# ------------------------------------------------------------------ super-tile kernels (per-band k, 2D tiles, beacon): k_super.cuh
SUPER_CONFIGS = [
    dict(profile=T.P5, tile=(26, 26), beacon=(26, 2, True), uep=T.UEP_LUMA, seed=(2, 1, 1), coset=1),   # BASELINE config 2
    dict(profile=T.P5, tile=(13, 7), uep=(0, 1, 0, 1, 0, 1, 0, 1, 0), seed=(1, 1, 1)),                   # k = 24/22, two rows per unit
    dict(profile=T.P5, tile=(2, 5), uep=(2, 2, 0, 0, 2, 2, 0, 0, 2), beacon=(3, 0, True), seed=(1, 2, 2)),  # k = 20/24, shortest period, slot 0
    dict(profile=T.P5, tile=(26, 5), uep=1, beacon=(255, 8, True), seed=(2, 2, 1)),                      # uniform k, odd tile height, longest period
    dict(profile=T.P2, uep=(0, 1, 2, 0, 1, 2, 0, 1, 2), beacon=(7, 4, True)),                            # three k values, 1D
    dict(profile=T.P5, tile=(1, 9), uep=T.UEP_LUMA, beacon=(26, 11, True)),                              # w = 1: identity rows; slot > 8: words completed, no beacon
    dict(profile=T.P3, uep=2, beacon=(26, 2, True), seed=(0, 1, 2)),
    dict(profile=T.P5, tile=(26, 4), uep=(3, 2, 3, 2, 3, 2, 3, 2, 3), beacon=(9, 5, True), seed=(1, 1, 0)),   # k = 18/20: eight parity symbols (dense plane layout)
    dict(profile=T.P4, uep=3, beacon=(26, 2, True)),                                                            # uniform k = 18 with a beacon
]


@pytest.mark.parametrize("ci", range(len(SUPER_CONFIGS)))
def test_super_tile_kernels_words(codec, oracle, t3, ci):
    oc, gc = both(SUPER_CONFIGS[ci])
    assert t3.super_path_available(gc)
    r = rng(5000 + ci)
    add = T.gf_add_table()
    for n in (3000, 8192, 30011, 70001):
        raw = r.integers(0, 27, size=(n, 9), dtype=np.uint8)
        raw[:, 8] %= 9
        wild = raw.copy()
        wild[r.integers(0, n, 50), r.integers(0, 9, 50)] = r.integers(27, 256, 50)
        for arith in (t3.REF_EXACT, t3.FIXED):
            assert np.array_equal(codec.encode_profile_from_raw(wild, gc, arith), oracle.encode_profile(oc, wild, arith)), (ci, n, arith)
        enc = codec.encode_profile_from_raw(raw, gc, t3.FIXED)
        assert np.array_equal(enc, oracle.encode_profile(oc, raw, t3.FIXED))
        ok, out, nc = codec.decode_profile_fixed(enc, gc, n_raw_words=n)
        ok_o, out_o, nc_o = oracle.decode_profile_fixed(oc, enc, n_raw_words=n)
        assert ok and ok_o and nc == nc_o == 0 and np.array_equal(out, out_o) and np.array_equal(out, raw[:out.shape[0]])
        if not (gc.beacon_enabled and gc.beacon_slot > 8):
            for exact_t in (False, True):
                bad, nerr = T.inject_errors(enc, oc, n, seed=5 + exact_t, gf_add=add, exact_t=exact_t)
                ok3, out3, nc3 = codec.decode_profile_fixed(bad, gc, n_raw_words=n)
                assert ok3 and np.array_equal(out3, out) and nc3 == nerr and nerr > 0
            worse = bad.copy().reshape(-1)
            worse[52 + 26 * 700:52 + 26 * 700 + 12] = (worse[52 + 26 * 700:52 + 26 * 700 + 12] + 1) % 27   # more than t errors inside a super-tile
            worse[60:70] = r.integers(27, 256, 10)                                                          # out-of-alphabet bytes
            ok4, out4, _ = codec.decode_profile_fixed(worse.reshape(-1, 9), gc, n_raw_words=n)
            ok5, out5, _ = oracle.decode_profile_fixed(oc, worse.reshape(-1, 9), n_raw_words=n)
            assert ok4 == ok5 and np.array_equal(out4, out5)


@pytest.mark.parametrize("ci", range(len(SUPER_CONFIGS)))
@pytest.mark.parametrize("n_px,nf", [(512 * 512, 2), (130 * 77 * 3 + 1, 3), (5940 * 4, 1)])
def test_super_tile_kernels_rgb_frames(codec, oracle, t3, ci, n_px, nf):
    oc, gc = both(SUPER_CONFIGS[ci])
    frames = np.stack([T.synth_rgb(21 + f, n_px) for f in range(nf)])
    add = T.gf_add_table()
    for arith in (t3.REF_EXACT, t3.FIXED):
        got = codec.encode_frames_rgb8(frames, gc, arith)
        for f in range(nf):
            assert np.array_equal(got[f], oracle.encode_rgb(oc, frames[f], arith)), (ci, n_px, arith, f)
    enc = codec.encode_frames_rgb8(frames, gc, t3.FIXED)
    ok, rgb, nc = codec.decode_frames_rgb8(enc, n_px, gc)
    assert ok.all() and nc == 0
    for f in range(nf):
        ok_o, rgb_o, _ = oracle.decode_rgb_fixed(oc, enc[f], n_px)
        assert ok_o and np.array_equal(rgb[f], rgb_o)
    if not (gc.beacon_enabled and gc.beacon_slot > 8):
        bad = enc.copy()
        tot = 0
        for f in range(nf):
            bad[f], ne = T.inject_errors(enc[f], oc, (n_px + 1) // 2, seed=31 + f, gf_add=add, exact_t=bool(f & 1))
            tot += ne
        ok2, rgb2, nc2 = codec.decode_frames_rgb8(bad, n_px, gc)
        assert ok2.all() and nc2 == tot and np.array_equal(rgb2, rgb)


# ------------------------------------------------------------------ .t3v container records
def test_t3v_records_and_crc32(codec, oracle):
    import zlib
    r = rng(960)
    for n in (0, 1, 3, 4, 5, 127, 128, 129, 32767, 32768, 32769, 1000003, 64 * 32768, 40000001):   # ... whole tiles only; more than 1024 tiles
        data = r.integers(0, 256, n, dtype=np.uint8)
        assert codec.crc32(data) == zlib.crc32(data.tobytes()) == oracle.crc32(data), n
    for nw in (0, 1, 2, 14, 15, 3640, 3641, 3642, 20011, 300007):     # around one segment (128 B), one tile (32 KiB) and many tiles
        words = r.integers(0, 27, size=(nw, 9), dtype=np.uint8)
        if nw in (2, 20011):
            words = r.integers(0, 256, size=(nw, 9), dtype=np.uint8)    # bytes >= 27 are stored % 27
        rec = codec.t3v_write_frame(words)
        assert np.array_equal(rec, oracle.t3v_frame_record(words)), nw
        ok, back = codec.t3v_read_frame(rec)
        ok_o, back_o = oracle.t3v_read_frame(rec)
        assert ok and ok_o and np.array_equal(back, back_o) and np.array_equal(back, words % 27)
        for blob in (rec[:-1], rec[:3], np.concatenate([rec, rec[:5]])):
            ok2, w2 = codec.t3v_read_frame(blob)
            ok3, w3 = oracle.t3v_read_frame(blob)
            assert ok2 == ok3 and np.array_equal(w2, w3)
        if nw:
            bad = rec.copy()
            bad[4 + int(r.integers(0, 9 * nw))] ^= 0x10
            assert codec.t3v_read_frame(bad)[0] is False and oracle.t3v_read_frame(bad)[0] is False
            bad = rec.copy()
            bad[-2] ^= 1
            assert codec.t3v_read_frame(bad)[0] is False
    aw = (140, 0, 6280, 4320)
    o = np.zeros(54, np.uint8)
    import ctypes as C
    oracle.lib.t3o_t3v_header(o.ctypes.data_as(C.c_void_p), 4, 2, 1, 1, 7680, 4320, (C.c_uint32 * 4)(*aw), 30000, 1001, 240, 1)
    assert np.array_equal(codec.t3v_header(4, 2, True, 1, 7680, 4320, aw, 30000, 1001, 240, 1), o)


def test_t3v_records_batched_device(codec, oracle):
    import torch
    dev = torch.device("cuda", 0)
    n_words, nf = 50001, 3
    stride = (n_words + 3) & ~3
    pitch = (8 + 9 * n_words + 15) & ~15
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    words = torch.randint(0, 27, (nf, stride * 9), dtype=torch.uint8, device=dev, generator=g)
    rec = torch.zeros(nf, pitch, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    codec.t3v_frame_records_dev(words, n_words, stride, nf, rec, pitch, s)
    back = torch.zeros_like(words)
    okf = torch.zeros(nf, dtype=torch.uint8, device=dev)
    rec[1, 4 + 777] ^= 3                                                  # frame 1 is damaged in transit
    codec.t3v_read_frames_dev(rec, pitch, nf, n_words, back, stride, okf, s)
    torch.cuda.synchronize()
    assert okf.cpu().tolist() == [1, 0, 1]
    rec[1, 4 + 777] ^= 3
    for f in range(nf):
        w = words[f, :9 * n_words].cpu().numpy().reshape(-1, 9)
        assert np.array_equal(rec[f, :8 + 9 * n_words].cpu().numpy(), oracle.t3v_frame_record(w))
        if f != 1:
            assert torch.equal(back[f, :9 * n_words], words[f, :9 * n_words])


# ------------------------------------------------------------------ image-bridge geometry and NEW-generation pipelines
def test_image_bridge_geometry(codec, oracle, t3):
    r = rng(980)
    for (sh, sw), (dh, dw) in (((37, 53), (540, 960)), ((1, 1), (7, 5)), ((300, 200), (31, 17)), ((64, 64), (64, 64)), ((5, 9), (1, 1)), ((1080, 1920), (2160, 3840))):
        img = r.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(codec.resize_rgb_nn(img, dw, dh), oracle.resize_rgb_nn(img, dw, dh)), ((sh, sw), (dh, dw))
    for (sh, sw), (ch, cw) in (((37, 53), (77, 101)), ((540, 960), (541, 961)), ((10, 8), (4, 8)), ((3, 3), (3, 3)), ((9, 2), (1, 7)), ((540, 960), (4320, 7680))):
        img = r.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(codec.blit_center_rgb(img, cw, ch), oracle.blit_center_rgb(img, cw, ch)), ((sh, sw), (ch, cw))
    with pytest.raises(t3.T3CError):
        codec.blit_center_rgb(np.zeros((2, 9, 3), np.uint8), 8, 8)        # wider than the canvas: the reference would write past the row
    for (fh, fw), (sh, sw) in (((40, 60), (20, 30)), ((40, 60), (40, 60)), ((10, 60), (14, 30)), ((7, 9), (1, 1)), ((4320, 7680), (540, 960))):
        q = T.synth_quant(3, fw * fh)
        assert np.array_equal(codec.extract_center_q(q, fw, fh, sw, sh).view(np.uint8), oracle.extract_center_q(q, fw, fh, sw, sh).view(np.uint8))


def test_v6new_image_pipelines(codec, oracle):
    r = rng(981)
    img = r.integers(0, 256, (41, 67, 3), dtype=np.uint8)
    for sub, cen in ((15, True), (15, False), (21, False), (27, True), (7, True)):
        ok_g, w_g = codec.v6new_image_to_words(img, sub, cen)
        ok_o, w_o = oracle.v6new_image_to_words(img, sub, cen)
        assert ok_g == ok_o and np.array_equal(w_g, w_o), (sub, cen)
        if ok_g:
            tw, th = T.V6NEW_STD_RES[sub]
            for (w, h) in ((tw, th), (100, 50), (tw + 3, th)):
                ok1, i1 = codec.v6new_words_to_image(w_g, sub, w, h)
                ok2, i2 = oracle.v6new_words_to_image(w_o, sub, w, h)
                assert ok1 and ok2 and np.array_equal(i1, i2), (sub, cen, w, h)
    assert codec.v6new_words_to_image(np.zeros(4, np.uint32), 7, 2, 2)[0] is False


# ------------------------------------------------------------------ golden vectors generated from the reference itself
GF = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_formats_v1.npz"))   # tests/golden/make_golden_formats.py


def test_golden_formats_t3v(codec):
    for nw in sorted(int(k[3:-6]) for k in GF.files if k.startswith("t3v") and k.endswith("_words")):
        words, rec = GF[f"t3v{nw}_words"], GF[f"t3v{nw}_record"]
        assert np.array_equal(codec.t3v_write_frame(words), rec), nw
        ok, back = codec.t3v_read_frame(rec)
        assert ok and np.array_equal(back, words % 27)
        if rec.size > 8:
            assert codec.crc32(rec[4:-4]) == __import__("zlib").crc32(rec[4:-4].tobytes())
    for i, (prof, code, cen, coset, fc, ft) in enumerate(((1, 0, 1, 0, 240, 0), (4, 2, 1, 1, 1, 1), (5, 4, 0, 2, 7, 0))):
        assert np.array_equal(codec.t3v_header(prof, code, cen, coset, 7680, 4320, (3, 1, 4, 1), 30000, 1001, fc, ft), GF[f"t3vhdr{i}"]), i


def test_golden_formats_subword_and_base243(codec):
    words = GF["sub_words"]
    for N in (27, 24, 21, 18, 15):
        t = codec.extract_subword_stream_from_words(words, N)
        assert np.array_equal(t, GF[f"sub{N}_stream"])
        assert np.array_equal(codec.ut_to_base243(t), GF[f"sub{N}_pack"])
        assert np.array_equal(codec.words_to_base243(words, N), GF[f"sub{N}_pack"])
        assert np.array_equal(codec.build_words_from_subword_stream(t[:5 * N + 3], N, 2), GF[f"sub{N}_rebuilt_fill2"])
    wild = GF["wild_trits"]
    assert np.array_equal(codec.ut_to_base243(wild), GF["wild_pack"])
    assert np.array_equal(codec.build_words_from_subword_stream(wild, 21, 1), GF["wild_rebuilt21"])
    ok, u = codec.base243_to_ut(GF["wild_pack"])
    assert ok == bool(GF["wild_unpack_ok"][0]) and np.array_equal(u, GF["wild_unpack"])


def test_golden_formats_new_generation(codec, t3):
    px = GF["new_px"].reshape(-1).view(t3.PIXEL_DTYPE)
    for sub in (0, 15, 21):
        ok, w = codec.v6new_encode_raw_pixels_to_words(px, sub)
        assert ok == bool(GF[f"new_pack{sub}_ok"][0]) and np.array_equal(w, GF[f"new_pack{sub}"])
    img = GF["img"]
    assert np.array_equal(codec.resize_rgb_nn(img, 960, 540), GF["img_resize_960x540"])
    assert np.array_equal(codec.resize_rgb_nn(img, 17, 31), GF["img_resize_17x31"])
    assert np.array_equal(codec.blit_center_rgb(img, 101, 77), GF["img_blit_101x77"])
    q = GF["q_60x40"].reshape(-1).view(t3.PIXEL_DTYPE)
    assert np.array_equal(codec.extract_center_q(q, 60, 40, 30, 20).view(np.uint8).reshape(-1, q.dtype.itemsize), GF["q_center_30x20"])
    for sub, cen in ((15, True), (15, False), (18, True)):
        ok, w = codec.v6new_image_to_words(img, sub, cen)
        assert ok and np.array_equal(w, GF[f"img_words_{sub}_{int(cen)}"])
        ok, back = codec.v6new_words_to_image(w, sub, 100, 50)
        assert ok and np.array_equal(back, GF[f"img_back_{sub}_{int(cen)}_100x50"])


def test_crc32_more_tiles_than_resident_warps(codec):
    """more 32 KiB tiles than the tile kernel's capped grid holds warps: the warps stride over the tiles"""
    import zlib
    n = 9500 * 32768 + 12345
    data = rng(961).integers(0, 256, n, dtype=np.uint8)
    assert codec.crc32(data) == zlib.crc32(data.tobytes())


# ------------------------------------------------------------------ multi-device streams (t3c_stream_*, BASELINE config 4)
@pytest.mark.parametrize("lanes", [[0], [0, 0, 0]])
def test_stream_api_frames_in_order_match_oracle(oracle, t3, lanes):
    """t3c_stream_encode_rgb8 / t3c_stream_decode_rgb8: frame f -> lane (first_frame + f) mod n_lanes, every frame equal to the oracle's
    encode of that frame, in frame order; decode with injected errors returns the oracle's pixels and the total corrected count"""
    import torch
    n_gpu = torch.cuda.device_count()
    devs = [d % n_gpu for d in (lanes if len(lanes) == 1 else list(range(len(lanes))))]
    kw = dict(profile=T.P3, uep=2)
    oc, gc = both(kw)
    n_px, n_frames = 540 * 40 + 17, 7
    frames = np.stack([T.synth_rgb(5 + f, n_px) for f in range(n_frames)])
    st = t3.Stream(devs)
    assert st.lanes == len(devs)
    enc = st.encode_rgb8(frames, gc, t3.FIXED, first_frame=3)
    add = T.gf_add_table()
    bad = enc.copy()
    tot = 0
    for f in range(n_frames):
        assert np.array_equal(enc[f], oracle.encode_rgb(oc, frames[f], 1)), f
        bad[f], ne = T.inject_errors(enc[f], oc, (n_px + 1) // 2, seed=40 + f, gf_add=add)
        tot += ne
    ok, rgb, nc = st.decode_rgb8(bad, n_px, gc, first_frame=1)
    assert ok.all() and nc == tot
    for f in range(n_frames):
        ok_o, rgb_o, _ = oracle.decode_rgb_fixed(oc, bad[f], n_px)
        assert ok_o and np.array_equal(rgb[f, :rgb_o.shape[0]], rgb_o), f
    st.close()

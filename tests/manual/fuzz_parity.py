"""Random configs x sizes: CUDA path (through the C ABI) against the CPU oracle -- raw-word API and RGB frames, both arithmetics,
consistent decode with injected errors.  python tests/manual/fuzz_parity.py [seconds] [seed]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import t3oracle as T
import ternary_image_codec_b200 as t3

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
r = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
codec, oracle = t3.Codec(0), T.Oracle()
add = T.gf_add_table()
t0 = time.time()
n_cases = n_super = n_fast = 0
while time.time() - t0 < budget:
    kw = dict(profile=int(r.integers(0, 5)))
    mode = r.integers(0, 4)
    kw["uep"] = int(r.integers(0, 4)) if mode == 0 else (T.UEP_LUMA if mode == 1 else tuple(int(x) for x in r.choice(r.choice(4, size=2, replace=False), size=9)))
    if r.random() < 0.5:
        kw["tile"] = (int(r.choice([1, 2, 13, 26, 26, 26, 7, 5, 64])), int(r.integers(1, 27)))
    if r.random() < 0.5:
        kw["beacon"] = (int(r.choice([1, 2, 3, 4, 9, 26, 26, 83, 255, 300])), int(r.integers(0, 12)), True)
    kw["seed"] = tuple(int(x) for x in r.integers(0, 5, 3))
    oc, gc = T.make_cfg(**kw), t3.make_config(**kw)
    n = int(r.choice([0, 1, 5, 300, 2999, 9000, 20011, 40000, 70001]))
    raw = r.integers(0, 27, size=(n, 9), dtype=np.uint8)
    if n:
        raw[:, 8] %= 9
    arith = int(r.integers(0, 2))
    wild = raw.copy()
    if n > 10 and r.random() < 0.3:
        wild[r.integers(0, n, 20), r.integers(0, 9, 20)] = r.integers(27, 256, 20)
    got, want = codec.encode_profile_from_raw(wild, gc, arith), oracle.encode_profile(oc, wild, arith)
    assert np.array_equal(got, want), ("encode words", kw, n, arith)
    enc = codec.encode_profile_from_raw(raw, gc, t3.FIXED)
    ok, out, nc = codec.decode_profile_fixed(enc, gc, n_raw_words=n)
    ok_o, out_o, nc_o = oracle.decode_profile_fixed(oc, enc, n_raw_words=n)
    assert ok == ok_o and nc == nc_o and np.array_equal(out, out_o), ("decode words", kw, n)
    if n >= 300 and not (gc.beacon_enabled and gc.beacon_slot > 8):
        bad, nerr = T.inject_errors(enc, oc, n, seed=int(r.integers(0, 1000)), gf_add=add, exact_t=bool(r.integers(0, 2)))
        if r.random() < 0.3:
            flat = bad.reshape(-1)
            p = 52 + int(r.integers(0, max(1, flat.size - 80)))
            flat[p:p + 14] = (flat[p:p + 14] + 1 + r.integers(0, 26, 14)) % 27          # beyond t: the verdict and the output must still agree
        ok3, out3, nc3 = codec.decode_profile_fixed(bad, gc, n_raw_words=n)
        ok4, out4, nc4 = oracle.decode_profile_fixed(oc, bad, n_raw_words=n)
        assert ok3 == ok4 and np.array_equal(out3, out4) and (not ok3 or nc3 == nc4), ("decode words with errors", kw, n, ok3, ok4, nc3, nc4)
    if kw["profile"] != 4 or True:
        n_px = int(r.choice([1, 77, 5940, 10010, 36864, 70001]))
        nf = int(r.integers(1, 3))
        frames = np.stack([T.synth_rgb(int(r.integers(0, 1000)), n_px) for _ in range(nf)])
        if gc.profile != t3.RAW_MODE:
            gotf = codec.encode_frames_rgb8(frames, gc, arith)
            for f in range(nf):
                assert np.array_equal(gotf[f], oracle.encode_rgb(oc, frames[f], arith)), ("encode rgb", kw, n_px, nf, arith, f)
            encf = codec.encode_frames_rgb8(frames, gc, t3.FIXED)
            okf, rgb, _ = codec.decode_frames_rgb8(encf, n_px, gc)
            for f in range(nf):
                ok_o, rgb_o, _ = oracle.decode_rgb_fixed(oc, encf[f], n_px)
                assert bool(okf[f]) == ok_o and np.array_equal(rgb[f], rgb_o), ("decode rgb", kw, n_px, nf, f)
    n_cases += 1
    n_super += int(t3.super_path_available(gc) and not t3.fast_path_available(gc))
    n_fast += int(t3.fast_path_available(gc))
print(f"fuzz ok: {n_cases} random configs ({n_fast} warp-tile family, {n_super} super-tile family) in {time.time() - t0:.0f} s")
codec.close()

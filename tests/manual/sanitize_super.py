"""Small super-tile kernel cases (per-band k, 2D, beacon) for compute-sanitizer: raw words and RGB frames, with errors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "oracle"))
import numpy as np
import t3oracle as T
import ternary_image_codec_b200 as t3

codec = t3.Codec(0)
oracle = T.Oracle()
add = T.gf_add_table()
r = np.random.default_rng(1)
for kw in (dict(profile=T.P5, tile=(26, 26), beacon=(26, 2, True), uep=T.UEP_LUMA, seed=(2, 1, 1), coset=1),
           dict(profile=T.P2, uep=(0, 1, 2, 0, 1, 2, 0, 1, 2), beacon=(7, 4, True)),
           dict(profile=T.P5, tile=(13, 7), uep=(0, 1, 0, 1, 0, 1, 0, 1, 0))):
    oc, gc = T.make_cfg(**kw), t3.make_config(**kw)
    assert t3.super_path_available(gc)
    n = 9001
    raw = r.integers(0, 27, size=(n, 9), dtype=np.uint8)
    raw[:, 8] %= 9
    enc = codec.encode_profile_from_raw(raw, gc, t3.FIXED)
    print("enc words", np.array_equal(enc, oracle.encode_profile(oc, raw, t3.FIXED)), flush=True)
    bad, nerr = T.inject_errors(enc, oc, n, seed=5, gf_add=add)
    ok, out, nc = codec.decode_profile_fixed(bad, gc, n_raw_words=n)
    print("dec words", ok, nc == nerr, np.array_equal(out, raw[:out.shape[0]]), flush=True)
    n_px, nf = 5940 * 3 + 77, 2
    frames = np.stack([T.synth_rgb(40 + f, n_px) for f in range(nf)])
    encf = codec.encode_frames_rgb8(frames, gc, t3.FIXED)
    print("enc rgb", all(np.array_equal(encf[f], oracle.encode_rgb(oc, frames[f], 1)) for f in range(nf)), flush=True)
    ok, rgb, nc = codec.decode_frames_rgb8(encf, n_px, gc)
    print("dec rgb", ok.all(), nc, all(np.array_equal(rgb[f], oracle.decode_rgb_fixed(oc, encf[f], n_px)[1]) for f in range(nf)), flush=True)
print("sanitize_super done")
codec.close()

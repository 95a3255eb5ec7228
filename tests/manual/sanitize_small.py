"""Small fused encode/decode cases for compute-sanitizer (memcheck / racecheck): multi-frame, odd sizes, errors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "oracle"))
import numpy as np
import t3oracle as T
import ternary_image_codec_b200 as t3

codec = t3.Codec(0)
oracle = T.Oracle()
for kw, n_px, nf in ((dict(profile=T.P3, uep=2), 540 * 70 + 33, 2), (dict(profile=T.P2, uep=1), 594 * 40, 1), (dict(profile=T.P1, uep=0), 648 * 35 + 1, 2)):
    oc, gc = T.make_cfg(**kw), t3.make_config(**kw)
    frames = np.stack([T.synth_rgb(40 + f, n_px) for f in range(nf)])
    enc = codec.encode_frames_rgb8(frames, gc, t3.FIXED)
    for f in range(nf):
        assert np.array_equal(enc[f], oracle.encode_rgb(oc, frames[f], 1))
    bad = enc.copy()
    tot = 0
    add = T.gf_add_table()
    for f in range(nf):
        bad[f], ne = T.inject_errors(enc[f], oc, (n_px + 1) // 2, seed=5 + f, gf_add=add)
        tot += ne
    ok, rgb, nc = codec.decode_frames_rgb8(bad, n_px, gc)
    assert ok.all() and nc == tot
    for f in range(nf):
        ok_o, rgb_o, _ = oracle.decode_rgb_fixed(oc, enc[f], n_px)
        assert np.array_equal(rgb[f], rgb_o)
print("sanitize_small ok")
codec.close()

"""Generates tests/golden/golden_v1.npz from the REFERENCE ITSELF (oracle/_ref/libt3ref*.so, compiled
from /root/reference by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

The fixtures are small seeded input/output pairs for every stage of the hot path plus whole-pipeline
outputs for the config matrix; tests/test_golden.py replays them against the C oracle (CPU) and
tests/test_gpu_parity.py against the CUDA path (GPU box, where /root/reference does not exist).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import t3oracle as T  # noqa: E402

CONFIGS = {
    "default_p2_k22": dict(),
    "selftest_p2_luma": dict(profile=T.P2, uep=T.UEP_LUMA),
    "p3_k20_1d": dict(profile=T.P3, uep=2),
    "p1_k24": dict(profile=T.P1, uep=0),
    "p4_k18": dict(profile=T.P4, uep=3),
    "p5_tile7x5_beacon4_2_luma": dict(profile=T.P5, tile=(7, 5), beacon=(4, 2, True), uep=T.UEP_LUMA, seed=(2, 1, 1)),
    "p5_tile26x26_beacon26_2_luma_c1": dict(profile=T.P5, tile=(26, 26), beacon=(26, 2, True), uep=T.UEP_LUMA, seed=(2, 1, 1), coset=1),
    "p5_tile300x7_mixedk_beacon1_0": dict(profile=T.P5, tile=(300, 7), uep=(0, 1, 2, 3, 0, 1, 2, 3, 0), beacon=(1, 0, True)),
    "p2_tile64_beacon83_2_main_cpp": dict(profile=T.P2, tile=(64, 64), beacon=(83, 2, True)),
}


def main():
    ref, fix = T.Reference(False), T.Reference(True)
    out = {}
    r = np.random.default_rng(20261018)
    for k in (24, 22, 20, 18):
        d = r.integers(0, 27, size=(64, k), dtype=np.uint8)
        d[0] = (5 * np.arange(k) + 7) % 27
        out[f"rs{k}_data"] = d
        out[f"rs{k}_enc_ref"] = ref.rs_encode_blocks(k, d, 0)
        out[f"rs{k}_enc_fix"] = fix.rs_encode_blocks(k, d, 1)
        add = T.gf_add_table()
        blocks = [r.integers(0, 27, size=(64, 26), dtype=np.uint8)]
        for base in (out[f"rs{k}_enc_ref"], out[f"rs{k}_enc_fix"]):
            for e in range(0, (26 - k) // 2 + 2):
                c = base.copy()
                for row in range(c.shape[0]):
                    pos = r.choice(26, size=e, replace=False)
                    c[row, pos] = add[c[row, pos], r.integers(1, 27, size=e)]
                blocks.append(c)
        blocks = np.concatenate(blocks)
        out[f"rs{k}_dec_in"] = blocks
        for tag, lib, fx in (("ref", ref, 0), ("fix", fix, 1)):
            io, o, ok = lib.rs_decode_blocks(k, blocks, fx)
            out[f"rs{k}_dec_{tag}_io"], out[f"rs{k}_dec_{tag}_out"], out[f"rs{k}_dec_{tag}_ok"] = io, o, ok
    # bridge + packing
    rgb = T.synth_rgb(1, 4096)
    out["rgb"] = rgb
    q = ref.rgb_to_quant(rgb)
    out["quant"] = q.view(np.uint8).reshape(-1, 6)
    out["rgb_back"] = ref.quant_to_rgb(q)
    out["raw_words"] = ref.pack_pixels(q)
    out["unpacked"] = ref.unpack_pixels(out["raw_words"]).view(np.uint8).reshape(-1, 6)
    # pipeline
    raw = r.integers(0, 27, size=(777, 9), dtype=np.uint8)
    raw[:, 8] %= 9
    out["pipe_raw"] = raw
    for name, kw in CONFIGS.items():
        cfg = T.make_cfg(**kw)
        out[f"cfg_{name}"] = np.frombuffer(bytes(cfg), np.uint8).copy()
        out[f"hdr_{name}"] = ref.header_pack(cfg)
        out[f"enc_ref_{name}"] = ref.encode_profile(cfg, raw, 0)
        out[f"enc_fix_{name}"] = fix.encode_profile(cfg, raw, 1)
        out[f"encrgb_ref_{name}"] = ref.encode_rgb(cfg, rgb, 0)
        # shipped decoder on a stream with a TRUE-codeword header and the fixed encoder's body words
        hp = ref.header_pack(cfg)
        a = fix.rs_encode_blocks(18, hp[:18], 1)[0]
        b = fix.rs_encode_blocks(18, np.concatenate([hp[18:], np.zeros(9, np.uint8)]), 1)[0]
        head = np.concatenate([a, b, np.array([3, 7], np.uint8)]).reshape(6, 9)
        stream = np.concatenate([head, out[f"enc_fix_{name}"][6:266]])
        ok, words, seen = ref.decode_profile_ref(T.make_cfg(), stream)
        out[f"dec_ref_in_{name}"] = stream
        out[f"dec_ref_ok_{name}"] = np.array([ok], np.uint8)
        out[f"dec_ref_out_{name}"] = words
        out[f"dec_ref_seen_{name}"] = np.frombuffer(bytes(seen), np.uint8).copy()
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()

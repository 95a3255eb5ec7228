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

"""Secondary workloads of BASELINE.json (configs[2..4]) and the formats either side of the path -- device-resident timing with
CUDA events on the launching stream.  A library: bench.py puts these numbers into the `secondary` object of its JSON line,
tools/bench_extra.py prints them one per line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ternary_image_codec_b200 as t3

N_PX = 7680 * 4320
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


class Ctx:
    def __init__(self, device_index=0, codec=None):
        self.dev = torch.device("cuda", device_index)
        self.codec = codec or t3.Codec(device_index, arith=t3.FIXED)
        self.S = torch.cuda.current_stream().cuda_stream


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))[n // 2]

def words8k(ctx, uep, name):
    """the reference's own API on an 8K frame: encode_profile_from_raw / consistent decode on raw Word27 words (device-resident)"""
    dev, codec, S = ctx.dev, ctx.codec, ctx.S
    cfg = t3.make_config(profile=t3.P3_RS26_20 if uep == 2 else 1, uep=uep)
    n_w = N_PX // 2
    g = torch.Generator(device=dev); g.manual_seed(6)
    raw = torch.randint(0, 27, (n_w, 9), dtype=torch.uint8, device=dev, generator=g)
    raw[:, 8] %= 9
    wpf = t3.profile_words(cfg, n_w)
    enc = torch.empty(wpf * 9, dtype=torch.uint8, device=dev)
    back = torch.zeros(n_w * 9, dtype=torch.uint8, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    te = timeit(lambda: codec.encode_profile_dev(raw, n_w, enc, wpf, cfg, t3.FIXED, S))
    td = timeit(lambda: codec.decode_profile_fixed_dev(enc, wpf, n_w, back, n_w, status, cfg, S))
    torch.cuda.synchronize()
    alg = 9 * n_w + 9 * wpf
    n_ok = (n_w - 300) * 9
    return {"workload": name, "encode_us": te * 1e3, "decode_us": td * 1e3, "encode_gbs": alg / te / 1e6, "decode_gbs": alg / td / 1e6,
                      "roundtrip": bool(torch.equal(back[:n_ok], raw.view(-1)[:n_ok])), "status": status.tolist(), "algorithmic_bytes": alg}


def raw8k(ctx):
    dev, codec, S = ctx.dev, ctx.codec, ctx.S  # see also tools/quick_raw.py (back-to-back launches between two events)
    g = torch.Generator(device=dev); g.manual_seed(4)
    NB = 3
    px = []
    for _ in range(NB):
        p = torch.empty(N_PX, 3, dtype=torch.int16, device=dev)
        p[:, 0] = torch.randint(0, 243, (N_PX,), device=dev, generator=g, dtype=torch.int16)
        p[:, 1:] = torch.randint(-40, 41, (N_PX, 2), device=dev, generator=g, dtype=torch.int16)
        px.append(p)
    words = [torch.empty(N_PX // 2 * 9, dtype=torch.uint8, device=dev) for _ in range(NB)]
    back = [torch.empty_like(px[0]) for _ in range(NB)]
    i = [0]
    def pack(): codec.pack_pixels_dev(px[i[0] % NB], N_PX, words[i[0] % NB], S); i[0] += 1
    def unpack(): codec.unpack_pixels_dev(words[i[0] % NB], N_PX // 2, back[i[0] % NB], S); i[0] += 1
    tp, tu = timeit(pack), timeit(unpack)
    alg = 6 * N_PX + 9 * (N_PX // 2)
    assert torch.equal(px[0], back[0])
    return {"workload": "raw8k: 8K PixelYCbCrQuant <-> Word27 (RAW mode, no RS)", "pack_us": tp * 1e3, "unpack_us": tu * 1e3,
                      "pack_gbs": alg / tp / 1e6, "unpack_gbs": alg / tu / 1e6, "pack_frac_of_measured_peak": alg / tp / 1e6 / PEAK,
                      "unpack_frac_of_measured_peak": alg / tu / 1e6 / PEAK, "mpix_per_s_pack_plus_unpack": N_PX / (tp + tu) / 1e3,
                      "algorithmic_bytes": alg}


def uep2d(ctx):
    dev, codec, S = ctx.dev, ctx.codec, ctx.S
    cfg = t3.make_config(profile=t3.P5_RS26_22_2D, tile=(26, 26), beacon=(26, 2, True), uep=t3.UEP_LUMA_PRIORITY, seed=(2, 1, 1), coset=1)
    wpf = t3.profile_words(cfg, N_PX // 2)
    g = torch.Generator(device=dev); g.manual_seed(3)
    rgb = torch.randint(0, 256, (N_PX * 3,), dtype=torch.uint8, device=dev, generator=g)
    enc = torch.empty(wpf * 9, dtype=torch.uint8, device=dev)
    back = torch.empty(N_PX * 3, dtype=torch.uint8, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    def E(): codec.encode_frames_rgb8_dev(rgb, N_PX, 1, enc, wpf, cfg, t3.FIXED, S)
    def D(): codec.decode_frames_rgb8_dev(enc, wpf, wpf, 1, N_PX, back, status, cfg, S)
    te = timeit(E, n=5, warm=2)
    td_clean = timeit(D, n=5, warm=2)
    # injected symbol errors (BASELINE config 2: "up to t"): in codeword c of band b (t_b = (26-k_b)/2), n errors at distinct positions
    # 3, 11, 19 (+ c mod 7), each symbol + (1 + c mod 26) mod 27; the beacon expansion (A.5) maps body index -> frame index
    ks = [24, 22, 20, 18]
    kb = [ks[u % 4] for u in t3.UEP_LUMA_PRIORITY]
    n_s = (26 * (N_PX // 2) + 2) // 3
    ncw = [((n_s - b + 8) // 9) // kb[b] for b in range(9)]
    P, slot = 26, 2
    clean = enc.clone()

    def inject(mode):
        enc.copy_(clean)
        body = enc[52:]
        base = 0
        total = 0
        for b in range(9):
            t_b = (26 - kb[b]) // 2
            c = torch.arange(ncw[b], device=dev, dtype=torch.int64)
            n_err = torch.full_like(c, t_b) if mode == "t" else (torch.ones_like(c) if mode == "one" else c % (t_b + 1))
            for j in range(t_b):
                sel = n_err > j
                cc = c[sel]
                p = 26 * (base + cc) + 3 + 8 * j + cc % 7
                blk, rem = p // (9 * P - 1), p % (9 * P - 1)
                o = torch.where(rem < 8, 9 * blk * P + torch.where(rem < slot, rem, rem + 1), 9 * (blk * P + 1 + (rem - 8) // 9) + (rem - 8) % 9)
                body[o] = ((body[o].to(torch.int64) + 1 + cc % 26) % 27).to(torch.uint8)
                total += int(sel.sum())
            base += ncw[b]
        return total

    res = {}
    for mode in ("one", "mixed", "t"):
        n_inj = inject(mode)
        res[mode] = {"us": timeit(D, n=3, warm=1) * 1e3, "injected": n_inj, "status": status.tolist()}
        assert status[0].item() == 1 and status[1].item() == n_inj, (mode, status.tolist(), n_inj)
    td_err = res["one"]["us"] / 1e3
    enc.copy_(clean)
    D(); torch.cuda.synchronize()
    alg = 3 * N_PX + 9 * wpf
    return {"workload": "uep2d: 8K, P5 2D 26x26 + luma UEP + coset C1 + beacon(26,2), super-tile kernels" if t3.super_path_available(cfg) else "uep2d (general kernels)", "encode_us": te * 1e3,
                      "decode_clean_us": td_clean * 1e3, "decode_with_errors_us": td_err * 1e3,
                      "decode_errors": {"one per codeword": res["one"], "0..t per codeword": res["mixed"], "t per codeword": res["t"]}, "status": status.tolist(),
                      "encode_gbs": alg / te / 1e6, "decode_clean_gbs": alg / td_clean / 1e6, "mpix_per_s_enc_plus_dec_clean": N_PX / (te + td_clean) / 1e3,
                      "profile_words": wpf}


def formats8k(ctx):
    """sub-word trit streams, base-243 packing and the NEW-generation 1-pixel words on 8K-sized inputs"""
    dev, codec, S = ctx.dev, ctx.codec, ctx.S
    n_words = N_PX // 2
    g = torch.Generator(device=dev); g.manual_seed(8)
    words = torch.randint(0, 27, (n_words * 9,), dtype=torch.uint8, device=dev, generator=g)
    N = 24
    trits = torch.empty(n_words * N, dtype=torch.uint8, device=dev)
    packed = torch.empty(4 + (n_words * N + 4) // 5 + 16, dtype=torch.uint8, device=dev)
    t_sub = timeit(lambda: codec.subword_stream_dev(words, n_words, N, trits, S))
    t_fused = timeit(lambda: codec.words_to_base243_dev(words, n_words, N, packed, S))
    px = torch.randint(0, 243, (N_PX, 3), dtype=torch.int16, device=dev, generator=g)
    px[:, 1:] = px[:, 1:] % 81 - 40
    w32 = torch.empty(N_PX, dtype=torch.int32, device=dev)
    back = torch.empty_like(px)
    t_pack = timeit(lambda: codec.lib.t3c_v6new_pack_pixels_dev(codec.h, px.data_ptr(), N_PX, w32.data_ptr(), S))
    t_unpack = timeit(lambda: codec.lib.t3c_v6new_unpack_pixels_dev(codec.h, w32.data_ptr(), N_PX, back.data_ptr(), S))
    assert torch.equal(px, back)
    res = {"workload": "formats8k: 16.6 M Word27 -> first 24 trits per word (1 B/trit) | fused -> base-243 bytes; 33.2 M pixels <-> NEW-generation 32-bit words",
           "subword_stream_us": t_sub * 1e3, "subword_stream_gbs": (9 + N) * n_words / t_sub / 1e6,
           "words_to_base243_us": t_fused * 1e3, "words_to_base243_gbs": (9 + N / 5) * n_words / t_fused / 1e6,
           "v6new_pack_us": t_pack * 1e3, "v6new_pack_gbs": 10 * N_PX / t_pack / 1e6, "v6new_unpack_us": t_unpack * 1e3, "v6new_unpack_gbs": 10 * N_PX / t_unpack / 1e6}
    return res


def t3v8k(ctx):
    """.t3v frame records of an 8K frame's profile words (RS(26,20): 20.8 M words): emit (n | payload % 27 | CRC) and check"""
    dev, codec, S = ctx.dev, ctx.codec, ctx.S
    cfg = t3.make_config(profile=t3.P3_RS26_20, uep=2)
    wpf = t3.profile_words(cfg, N_PX // 2)
    NB = 3
    stride = (wpf + 15) & ~15
    pitch = (8 + 9 * wpf + 15) & ~15
    g = torch.Generator(device=dev); g.manual_seed(6)
    words = torch.randint(0, 27, (NB, stride * 9), dtype=torch.uint8, device=dev, generator=g)
    # a record is n (4 bytes) | payload | CRC: its base sits at 12 mod 16 so that the payload -- all but 8 of its bytes -- moves in 16-byte accesses
    rec_flat = torch.zeros(NB * pitch + 16, dtype=torch.uint8, device=dev)
    rec = [rec_flat[12 + j * pitch:12 + (j + 1) * pitch] for j in range(NB)]
    back = torch.zeros_like(words)
    okf = torch.zeros(1, dtype=torch.uint8, device=dev)
    i = [0]
    def W(): codec.t3v_frame_records_dev(words[i[0] % NB], wpf, stride, 1, rec[i[0] % NB], pitch, S); i[0] += 1
    def R(): codec.t3v_read_frames_dev(rec[i[0] % NB], pitch, 1, wpf, back[i[0] % NB], stride, okf, S); i[0] += 1
    tw = timeit(W)
    tr = timeit(R)
    import zlib
    r0 = rec[0][:8 + 9 * wpf].cpu().numpy()
    want = zlib.crc32(r0[4:-4].tobytes()) ^ ((zlib.crc32(r0[:4].tobytes()) * 16777619) & 0xFFFFFFFF)
    assert int.from_bytes(r0[-4:].tobytes(), "little") == want and okf.item() == 1 and torch.equal(back[0, :9 * wpf], words[0, :9 * wpf])
    alg = 2 * 9 * wpf
    return {"workload": "t3v8k: .t3v frame record of an 8K frame's 20.8 M profile words (payload % 27 + CRC-32), emit and check; record base at 12 mod 16", "write_us": tw * 1e3,
                      "read_check_us": tr * 1e3, "write_gbs": alg / tw / 1e6, "read_gbs": alg / tr / 1e6, "write_frac_of_measured_peak": alg / tw / 1e6 / PEAK,
                      "read_frac_of_measured_peak": alg / tr / 1e6 / PEAK, "algorithmic_bytes": alg}


def stream240(ctx, frames_per_call=8, n_frames=240):
    dev, codec, S = ctx.dev, ctx.codec, ctx.S
    cfg = t3.make_config(profile=t3.P3_RS26_20, uep=2)
    wpf = t3.profile_words(cfg, N_PX // 2)
    stride = (wpf + 15) & ~15
    F = frames_per_call
    g = torch.Generator(device=dev); g.manual_seed(5)
    rgb = torch.randint(0, 256, (F, N_PX * 3), dtype=torch.uint8, device=dev, generator=g)
    enc = torch.empty(F, stride * 9, dtype=torch.uint8, device=dev)
    back = torch.empty(F, N_PX * 3, dtype=torch.uint8, device=dev)
    status = torch.zeros(2 * F, dtype=torch.int32, device=dev)
    calls = n_frames // F
    def step():
        codec.encode_frames_rgb8_dev(rgb, N_PX, F, enc, stride, cfg, t3.FIXED, S)
        codec.decode_frames_rgb8_dev(enc, wpf, stride, F, N_PX, back, status, cfg, S)
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(calls):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"workload": f"stream240: 240 synthetic 8K frames, RS(26,20) 1D, {F} frames per batched launch, 1 GPU (frame f -> GPU f mod G shards this linearly)",
                      "total_ms": ms, "frames_per_s": calls * F / ms * 1e3, "mpix_per_s": calls * F * N_PX / ms / 1e3, "frames": calls * F, "ok": bool((status[0::2] == 1).all())}




def bridge8k(ctx):
    """general-chain bridge kernels on an 8K frame: RGB8 -> PixelYCbCrQuant (3 + 6 B/px) and back (IMG:156-192)"""
    dev, codec, S = ctx.dev, ctx.codec, ctx.S
    g = torch.Generator(device=dev); g.manual_seed(9)
    NB = 3
    rgb = [torch.randint(0, 256, (N_PX * 3,), dtype=torch.uint8, device=dev, generator=g) for _ in range(NB)]
    q = [torch.empty(N_PX * 6, dtype=torch.uint8, device=dev) for _ in range(NB)]
    back = [torch.empty(N_PX * 3, dtype=torch.uint8, device=dev) for _ in range(NB)]
    i = [0]
    def F(): codec.rgb_to_quant_dev(rgb[i[0] % NB], N_PX, q[i[0] % NB], S); i[0] += 1
    def B(): codec.quant_to_rgb_dev(q[i[0] % NB], N_PX, back[i[0] % NB], S); i[0] += 1
    tf, tb = timeit(F), timeit(B)
    alg = 9 * N_PX
    return {"workload": "bridge8k: rgb_to_quant_stream / quant_stream_to_rgb on 33.2 M pixels (general chain)", "rgb_to_quant_us": tf * 1e3, "quant_to_rgb_us": tb * 1e3,
            "rgb_to_quant_gbs": alg / tf / 1e6, "quant_to_rgb_gbs": alg / tb / 1e6, "algorithmic_bytes": alg}


def entry(us, alg):
    """one `secondary` record: milliseconds, algorithmic bytes, achieved GB/s and its share of the measured HBM peak"""
    gbs = alg / us / 1e3
    return {"ms": us / 1e3, "algorithmic_bytes": int(alg), "gbs": gbs, "frac": gbs / PEAK}


def secondary_single_gpu(ctx):
    """BASELINE configs 2 and 3, the reference's raw-word API (default context, k = 22) and the formats either side of the path,
    as `secondary` records for bench.py"""
    out = {}
    u = uep2d(ctx)
    wpf = u["profile_words"]
    alg2 = 3 * N_PX + 9 * wpf
    out["config2_uep_2d_beacon"] = {
        "workload": u["workload"], "encode": entry(u["encode_us"], alg2), "decode_clean": entry(u["decode_clean_us"], alg2),
        "decode_one_error_per_codeword": entry(u["decode_errors"]["one per codeword"]["us"], alg2),
        "decode_0_to_t_errors_per_codeword": entry(u["decode_errors"]["0..t per codeword"]["us"], alg2),
        "decode_t_errors_per_codeword": entry(u["decode_errors"]["t per codeword"]["us"], alg2),
        "injected_symbol_errors": {k: v["injected"] for k, v in u["decode_errors"].items()}}
    r = raw8k(ctx)
    out["config3_raw_mode"] = {"workload": r["workload"], "pack": entry(r["pack_us"], r["algorithmic_bytes"]), "unpack": entry(r["unpack_us"], r["algorithmic_bytes"])}
    for uep, key in ((1, "words_api_default_context_k22"), (2, "words_api_k20")):
        w = words8k(ctx, uep, key)
        assert w["roundtrip"]
        out[key] = {"workload": "encode_profile_from_raw + consistent decode on 16.6 M raw Word27 (8K), device-resident", "encode": entry(w["encode_us"], w["algorithmic_bytes"]),
                    "decode": entry(w["decode_us"], w["algorithmic_bytes"])}
    b = bridge8k(ctx)
    out["bridge_general_chain"] = {"workload": b["workload"], "rgb_to_quant": entry(b["rgb_to_quant_us"], b["algorithmic_bytes"]),
                                   "quant_to_rgb": entry(b["quant_to_rgb_us"], b["algorithmic_bytes"])}
    f = formats8k(ctx)
    n_w = N_PX // 2
    out["formats"] = {"workload": f["workload"], "subword_stream_N24": entry(f["subword_stream_us"], (9 + 24) * n_w),
                      "words_to_base243_N24": entry(f["words_to_base243_us"], (9 + 24 / 5) * n_w),
                      "v6new_pack": entry(f["v6new_pack_us"], 10 * N_PX), "v6new_unpack": entry(f["v6new_unpack_us"], 10 * N_PX)}
    t = t3v8k(ctx)
    out["t3v_records"] = {"workload": t["workload"], "write": entry(t["write_us"], t["algorithmic_bytes"]), "read_check": entry(t["read_check_us"], t["algorithmic_bytes"])}
    return out

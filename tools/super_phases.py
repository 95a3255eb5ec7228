"""Debug build only (T3C_NVCC_EXTRA=-DT3C_SUPER_DEBUG python -m ternary_image_codec_b200._build --force): cycles CTA 0 spends between the
barriers of the super-tile kernels on an 8K frame of BASELINE config 2."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ternary_image_codec_b200 as t3

n_px = 7680 * 4320
cfg = t3.make_config(profile=t3.P5_RS26_22_2D, tile=(26, 26), beacon=(26, 2, True), uep=t3.UEP_LUMA_PRIORITY, seed=(2, 1, 1), coset=1)
codec = t3.Codec(0, arith=t3.FIXED)
lib = t3.load_library()
wpf = t3.profile_words(cfg, n_px // 2)
dev = torch.device("cuda", 0)
rgb = torch.randint(0, 256, (n_px * 3,), dtype=torch.uint8, device=dev)
enc = torch.empty(wpf * 9, dtype=torch.uint8, device=dev)
back = torch.empty(n_px * 3, dtype=torch.uint8, device=dev)
status = torch.zeros(2, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
buf = (C.c_uint32 * 32)()
for it in range(3):
    codec.encode_frames_rgb8_dev(rgb, n_px, 1, enc, wpf, cfg, t3.FIXED, s)
    codec.decode_frames_rgb8_dev(enc, wpf, wpf, 1, n_px, back, status, cfg, s)
    torch.cuda.synchronize()
    ok = lib.t3c_debug_counters(buf)
    v = list(buf)
    names = ["enc load", "enc phase A", "enc phase B", "enc phase C", "dec load", "dec squeeze", "dec phase B", "dec phase A", "dec store"]
    print(ok, {n: v[i] for i, n in enumerate(names)}, "enc sum", sum(v[:4]), "dec sum", sum(v[4:9]), "encC: chunks", v[9], "edges", v[10])

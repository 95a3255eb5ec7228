"""Short driver for ncu: a few super-tile encode/decode launches (BASELINE config 2) on one 8K frame (no timing)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ternary_image_codec_b200 as t3

n_px = 7680 * 4320
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = t3.make_config(profile=t3.P5_RS26_22_2D, tile=(26, 26), beacon=(26, 2, True), uep=t3.UEP_LUMA_PRIORITY, seed=(2, 1, 1), coset=1)
codec = t3.Codec(0, arith=t3.FIXED)
wpf = t3.profile_words(cfg, n_px // 2)
dev = torch.device("cuda", 0)
rgb = torch.randint(0, 256, (n_px * 3,), dtype=torch.uint8, device=dev)
enc = torch.empty(wpf * 9, dtype=torch.uint8, device=dev)
back = torch.empty(n_px * 3, dtype=torch.uint8, device=dev)
status = torch.zeros(2, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(iters):
    codec.encode_frames_rgb8_dev(rgb, n_px, 1, enc, wpf, cfg, t3.FIXED, s)
    codec.decode_frames_rgb8_dev(enc, wpf, wpf, 1, n_px, back, status, cfg, s)
torch.cuda.synchronize()
print("status", status.tolist())

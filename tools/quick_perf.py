"""Quick device-side timing of the fused kernels on one 8K frame (CUDA events, rotating buffers)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ternary_image_codec_b200 as t3

n_px = 7680 * 4320
k_idx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = t3.make_config(profile=t3.P3_RS26_20, uep=k_idx)
codec = t3.Codec(0, arith=t3.FIXED)
wpf = t3.profile_words(cfg, n_px // 2)
dev = torch.device("cuda", 0)
NB = 3
rgb = [torch.randint(0, 256, (n_px * 3,), dtype=torch.uint8, device=dev) for _ in range(NB)]
enc = [torch.empty(wpf * 9, dtype=torch.uint8, device=dev) for _ in range(NB)]
back = [torch.empty(n_px * 3, dtype=torch.uint8, device=dev) for _ in range(NB)]
status = torch.zeros(2 * NB, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
def E(i): codec.encode_frames_rgb8_dev(rgb[i], n_px, 1, enc[i], wpf, cfg, t3.FIXED, s)
def D(i): codec.decode_frames_rgb8_dev(enc[i], wpf, wpf, 1, n_px, back[i], status[2 * i:], cfg, s)
for i in range(NB): E(i); D(i)
torch.cuda.synchronize()
N = 10
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * N + 1)]
ev[0].record()
for i in range(N):
    E(i % NB); ev[2 * i + 1].record(); D((i + 1) % NB); ev[2 * i + 2].record()
torch.cuda.synchronize()
e = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(N))
d = sorted(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(N))
alg = 3 * n_px + 9 * wpf
print(f"k_idx={k_idx} encode median {e[N//2]*1e3:.1f} us ({alg/e[N//2]/1e6:.0f} GB/s)  decode median {d[N//2]*1e3:.1f} us ({alg/d[N//2]/1e6:.0f} GB/s)  status {status.tolist()}")
q = torch.empty(n_px * 6, dtype=torch.uint8, device=dev); chk = torch.empty(n_px * 3, dtype=torch.uint8, device=dev)
codec.rgb_to_quant_dev(rgb[0], n_px, q, s); codec.quant_to_rgb_dev(q, n_px, chk, s); torch.cuda.synchronize()
print("roundtrip ok:", bool(torch.equal(chk, back[0])))

"""Short driver for ncu: RAW-mode pack / unpack of one 8K frame of quantised pixels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ternary_image_codec_b200 as t3
N_PX = 7680 * 4320
dev = torch.device("cuda", 0)
codec = t3.Codec(0)
S = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=dev); g.manual_seed(4)
p = torch.empty(N_PX, 3, dtype=torch.int16, device=dev)
p[:, 0] = torch.randint(0, 243, (N_PX,), device=dev, generator=g, dtype=torch.int16)
p[:, 1:] = torch.randint(-40, 41, (N_PX, 2), device=dev, generator=g, dtype=torch.int16)
words = torch.empty(N_PX // 2 * 9, dtype=torch.uint8, device=dev)
back = torch.empty_like(p)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    codec.pack_pixels_dev(p, N_PX, words, S)
    codec.unpack_pixels_dev(words, N_PX // 2, back, S)
torch.cuda.synchronize()
print("equal", bool(torch.equal(p, back)))

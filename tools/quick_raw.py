"""Device-side timing of the RAW-mode kernels on one 8K frame: N back-to-back launches between two events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ternary_image_codec_b200 as t3
N_PX = 7680 * 4320
dev = torch.device("cuda", 0)
codec = t3.Codec(0)
S = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=dev); g.manual_seed(4)
NB = 3
px, words, back = [], [], []
for _ in range(NB):
    p = torch.empty(N_PX, 3, dtype=torch.int16, device=dev)
    p[:, 0] = torch.randint(0, 243, (N_PX,), device=dev, generator=g, dtype=torch.int16)
    p[:, 1:] = torch.randint(-40, 41, (N_PX, 2), device=dev, generator=g, dtype=torch.int16)
    px.append(p); words.append(torch.empty(N_PX // 2 * 9, dtype=torch.uint8, device=dev)); back.append(torch.empty_like(p))
def run(fn, n=30):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
alg = 6 * N_PX + 9 * (N_PX // 2)
tp = run(lambda i: codec.pack_pixels_dev(px[i % NB], N_PX, words[i % NB], S))
tu = run(lambda i: codec.unpack_pixels_dev(words[i % NB], N_PX // 2, back[i % NB], S))
print(f"pack {tp:.1f} us ({alg/tp/1e3:.0f} GB/s)  unpack {tu:.1f} us ({alg/tu/1e3:.0f} GB/s)  roundtrip {bool(torch.equal(px[0], back[0]))}")

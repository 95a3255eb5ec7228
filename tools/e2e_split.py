"""Host-buffer C ABI calls timed one by one on an 8K frame (bench.py's headline config): encode alone, decode alone, against the PCIe time of
their bytes.  python tools/e2e_split.py"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ternary_image_codec_b200 as t3  # noqa: E402

W, H = 7680, 4320
n_px = W * H
codec = t3.Codec(0, arith=t3.FIXED)
cfg = t3.make_config(profile=t3.P3_RS26_20, uep=2)
wpf = t3.profile_words(cfg, (n_px + 1) // 2)
h_rgb = torch.randint(0, 256, (1, n_px, 3), dtype=torch.uint8).pin_memory()
got = C.c_size_t()
L = codec.lib
h_enc = torch.empty((1, wpf, 9), dtype=torch.uint8).pin_memory()
h_back = torch.empty((1, n_px, 3), dtype=torch.uint8).pin_memory()
ok = np.zeros(1, np.uint8)
rec, nc = C.c_size_t(), C.c_size_t()


def enc():
    assert L.t3c_encode_frames_rgb8(codec.h, C.byref(cfg), t3.FIXED, h_rgb.data_ptr(), n_px, 1, h_enc.data_ptr(), wpf, C.byref(got)) == 0


def dec():
    assert L.t3c_decode_frames_rgb8(codec.h, C.byref(cfg), h_enc.data_ptr(), wpf, wpf, 1, n_px, h_back.data_ptr(), ok.ctypes.data_as(C.c_void_p),
                                    C.byref(rec), C.byref(nc)) == 0 and ok[0] == 1


def timed(f, n=8):
    f()
    f()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    return (time.perf_counter() - t0) / n * 1e3


d = torch.empty(9 * wpf, dtype=torch.uint8, device="cuda")
def h2d_words(): d.copy_(h_enc.view(-1), non_blocking=True); torch.cuda.synchronize()
def d2h_words(): h_enc.view(-1).copy_(d, non_blocking=True); torch.cuda.synchronize()
print(f"words per frame {wpf}: {9 * wpf / 1e6:.1f} MB, rgb {3 * n_px / 1e6:.1f} MB")
print(f"plain copies of the word buffer: H2D {timed(h2d_words):.3f} ms, D2H {timed(d2h_words):.3f} ms")
print(f"encode call {timed(enc):.3f} ms, decode call {timed(dec):.3f} ms")

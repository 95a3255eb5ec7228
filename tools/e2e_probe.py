"""Times the host-buffer C-ABI calls (pinned host memory) separately: encode, decode."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ternary_image_codec_b200 as t3
n_px = 7680 * 4320
cfg = t3.make_config(profile=t3.P3_RS26_20, uep=2)
codec = t3.Codec(0, arith=t3.FIXED)
wpf = t3.profile_words(cfg, n_px // 2)
h_rgb = torch.randint(0, 256, (n_px * 3,), dtype=torch.uint8).pin_memory()
h_enc = torch.empty(wpf * 9, dtype=torch.uint8).pin_memory()
h_back = torch.empty(n_px * 3, dtype=torch.uint8).pin_memory()
L = codec.lib
okb = np.zeros(1, np.uint8); got, rec, nc = C.c_size_t(), C.c_size_t(), C.c_size_t()
def enc(): assert L.t3c_encode_frames_rgb8(codec.h, C.byref(cfg), t3.FIXED, h_rgb.data_ptr(), n_px, 1, h_enc.data_ptr(), wpf, C.byref(got)) == 0
def dec(): assert L.t3c_decode_frames_rgb8(codec.h, C.byref(cfg), h_enc.data_ptr(), wpf, wpf, 1, n_px, h_back.data_ptr(), okb.ctypes.data_as(C.c_void_p), C.byref(rec), C.byref(nc)) == 0
def t(fn, n=5):
    fn(); t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
print(f"chunks={os.environ.get('T3C_PIPE_CHUNKS','8')} encode {t(enc):.2f} ms  decode {t(dec):.2f} ms  ok={okb[0]}")

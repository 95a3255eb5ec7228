"""How much does pageable host memory cost the host-buffer calls?  8K frame, encode + decode through the C ABI with (a) pinned torch
buffers, (b) plain numpy (pageable) buffers, (c) pageable buffers registered with cudaHostRegister (cost of the registration shown)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ternary_image_codec_b200 as t3

n_px = 7680 * 4320
cfg = t3.make_config(profile=t3.P3_RS26_20, uep=2)
codec = t3.Codec(0, arith=t3.FIXED)
wpf = t3.profile_words(cfg, n_px // 2)
L = codec.lib
cudart = C.CDLL("libcudart.so") if False else None


def run(rgb_ptr, enc_ptr, back_ptr, n=5):
    got = C.c_size_t(); okb = np.zeros(1, np.uint8); rec = C.c_size_t(); nc = C.c_size_t()
    def step():
        s1 = L.t3c_encode_frames_rgb8(codec.h, C.byref(cfg), t3.FIXED, rgb_ptr, n_px, 1, enc_ptr, wpf, C.byref(got))
        s2 = L.t3c_decode_frames_rgb8(codec.h, C.byref(cfg), enc_ptr, wpf, wpf, 1, n_px, back_ptr, okb.ctypes.data_as(C.c_void_p), C.byref(rec), C.byref(nc))
        assert s1 == 0 and s2 == 0 and okb[0] == 1
    step()
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    return (time.perf_counter() - t0) / n * 1e3


rng = np.random.default_rng(1)
src = rng.integers(0, 256, n_px * 3, dtype=np.uint8)
p_rgb = torch.from_numpy(src.copy()).pin_memory(); p_enc = torch.empty(wpf * 9, dtype=torch.uint8).pin_memory(); p_back = torch.empty(n_px * 3, dtype=torch.uint8).pin_memory()
ms_pinned = run(p_rgb.data_ptr(), p_enc.data_ptr(), p_back.data_ptr())
a_rgb = src.copy(); a_enc = np.empty(wpf * 9, np.uint8); a_back = np.empty(n_px * 3, np.uint8)
ms_pageable = run(a_rgb.ctypes.data, a_enc.ctypes.data, a_back.ctypes.data)
assert np.array_equal(a_back, p_back.numpy())
t0 = time.perf_counter()
for a in (a_rgb, a_enc, a_back):
    r = torch.cuda.cudart().cudaHostRegister(a.ctypes.data, a.nbytes, 0)
ms_reg = (time.perf_counter() - t0) * 1e3
ms_registered = run(a_rgb.ctypes.data, a_enc.ctypes.data, a_back.ctypes.data)
t0 = time.perf_counter()
for a in (a_rgb, a_enc, a_back):
    torch.cuda.cudart().cudaHostUnregister(a.ctypes.data)
ms_unreg = (time.perf_counter() - t0) * 1e3
print({"ms_per_step_pinned": ms_pinned, "ms_per_step_pageable": ms_pageable, "ms_per_step_registered": ms_registered,
       "register_ms_for_573MB": ms_reg, "unregister_ms": ms_unreg, "mpix_s_pinned": n_px / ms_pinned / 1e3, "mpix_s_pageable": n_px / ms_pageable / 1e3})

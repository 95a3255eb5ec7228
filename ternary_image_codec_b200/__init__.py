"""ternary_image_codec_b200 -- B200-native (sm_100a) encode/decode hot path of the Ternary Image Codec v6.

The product is ``libt3c.so`` (CUDA kernels behind the C ABI of ``include/t3c.h``); this package is the
Python host mirror used by the tests and ``bench.py``.  Importing it never loads ``oracle/``; using
it without the built library or without a GPU raises instead of falling back to the CPU.
"""
from .capi import (FIXED, P1_RS26_24, P2_RS26_22, P3_RS26_20, P4_RS26_18, P5_RS26_22_2D, PIXEL_DTYPE, RAW_MODE, REF_EXACT,  # noqa: F401
                   UEP_LUMA_PRIORITY, Codec, Config, Stream, T3CError, exported_symbols, fast_path_available, load_library, super_path_available, super_plan, make_config,
                   profile_words)

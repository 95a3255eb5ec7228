"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/t3c.h declares, its host-side geometry agrees with the oracle, and it refuses to run
without a GPU instead of falling back to a CPU path.  No compute calls are made here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import t3oracle as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import ternary_image_codec_b200 as t3
    from ternary_image_codec_b200 import _build
    _build.build()
    return t3.load_library()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "t3c.h")).read()
    return sorted(set(re.findall(r"\b(t3c_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/t3c.h but not exported by libt3c.so"
    out = subprocess.run(["nm", "-D", "--defined-only", lib._name], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (t3c_[a-z0-9_]+)", out))
    assert set(names) <= exported
    # nothing but the C ABI leaks out of the library
    assert all(s.startswith("t3c_") for s in re.findall(r" T (\S+)", out))


def test_python_binding_matches_header(lib):
    import ternary_image_codec_b200 as t3
    assert sorted(t3.exported_symbols()) == declared_symbols()


def test_config_layout_matches_oracle_struct():
    import ternary_image_codec_b200 as t3
    assert C.sizeof(t3.Config) == C.sizeof(T.Cfg) == 44
    for (na, ta), (nb, tb) in zip(t3.Config._fields_, T.Cfg._fields_):
        assert na == nb and getattr(t3.Config, na).offset == getattr(T.Cfg, nb).offset


def test_profile_words_matches_oracle(lib, oracle):
    import ternary_image_codec_b200 as t3
    r = np.random.default_rng(5)
    for _ in range(300):
        kw = dict(profile=int(r.choice([0, 1, 2, 3, 4])), uep=[int(x) for x in r.integers(0, 4, 9)],
                  tile=(int(r.integers(0, 40)), int(r.integers(0, 40))),
                  beacon=(int(r.integers(0, 30)), int(r.integers(0, 11)), bool(r.integers(0, 2))))
        n = int(r.choice([0, 1, 2, 3, 26, 27, 100, 777, 8192, 16588800, int(r.integers(0, 10 ** 6))]))
        assert t3.profile_words(t3.make_config(**kw), n) == oracle.words_bound(T.make_cfg(**kw), n), (kw, n)
    assert t3.profile_words(t3.make_config(profile=t3.P3_RS26_20, uep=2), 16588800) == 20766726
    assert t3.profile_words(t3.make_config(profile=t3.RAW_MODE), 12345) == 12345


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    import ternary_image_codec_b200 as t3
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(t3.T3CError, match="no CPU fallback"):
        t3.Codec(0)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "ternary_image_codec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "t3oracle" not in txt and "t3_oracle" not in txt and "libt3ref" not in txt, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        assert "oracle" not in open(os.path.join(ROOT, "include", f)).read().replace("no oracle", ""), f

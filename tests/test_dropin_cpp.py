"""Drop-in proof at the C++ level: tests/cpp/dropin_main.cpp is written against the reference's public
names only.  Built against the reference headers it runs the reference's CPU code (oracle/_ref/dropin_ref,
made by oracle/Makefile in the build container); built against this repo's include/ + libt3c.so it runs
on the B200.  Both must print the same lines."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_ref")
OURS = os.path.join(ROOT, "tests", "cpp", "_dropin_ours")


def build_ours():
    from ternary_image_codec_b200 import _build
    _build.build()
    pkg = os.path.join(ROOT, "ternary_image_codec_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp"),
                           "-L" + pkg, "-lt3c", "-Wl,-rpath," + pkg, "-o", OURS])


def test_dropin_caller_compiles_against_our_headers():
    """CPU: the reference-style caller compiles and links against include/ + libt3c.so unchanged."""
    build_ours()
    assert os.path.exists(OURS)


@pytest.mark.gpu
def test_dropin_caller_prints_the_same_as_the_reference_build():
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/dropin_ref not built (needs /root/reference at build time)")
    build_ours()
    want = subprocess.run([REF_BIN], capture_output=True, text=True, check=True).stdout.splitlines()
    got = subprocess.run([OURS], capture_output=True, text=True, check=True).stdout.splitlines()
    assert len(want) > 25
    assert got == want

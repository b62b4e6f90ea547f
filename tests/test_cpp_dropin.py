"""C++ drop-in boundary: the reference-named headers in include/ + libsobfu_b200.so.
CPU: both test programs were built by build() (the reference's own test/*.cpp compile UNCHANGED against our headers).
GPU: run them -- the reference's six asserting gtest cases and three solver smoke tests pass on our library."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "sobfu_b200", "_lib", "dropin_test")
THEIRS = os.path.join(ROOT, "oracle", "_ref", "sobfu_test_dropin")


def test_dropin_programs_are_built(built):
    assert os.path.exists(OURS)
    if os.path.isdir("/root/reference/test"):
        assert os.path.exists(THEIRS)
        syms = subprocess.run(["nm", "-D", "--undefined-only", THEIRS], capture_output=True, text=True).stdout
        assert "sobfu_b200_solver_estimate_psi" in syms and "sobfu_b200_jacobian" in syms    # reference tests -> our C ABI


@pytest.mark.gpu
def test_own_dropin_program_passes(built):
    r = subprocess.run([OURS], capture_output=True, text=True, timeout=600, env=dict(os.environ, SOBFU_B200_QUIET="1"))
    assert r.returncode == 0 and "2 tests ran, 0 failed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_reference_gtest_sources_pass_on_our_library(built):
    if not os.path.exists(THEIRS):
        pytest.skip("oracle/_ref/sobfu_test_dropin not built (needs /root/reference at build time)")
    r = subprocess.run([THEIRS], capture_output=True, text=True, timeout=900, env=dict(os.environ, SOBFU_B200_QUIET="1"))
    assert r.returncode == 0 and "9 tests ran, 0 failed" in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]

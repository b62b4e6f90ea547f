"""The C-ABI library loads without a GPU and exports every symbol include/sobfu_b200.h declares; the product does not
link or load the oracle."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "sobfu_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sobfu_b200_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(built):
    L = ctypes.CDLL(built)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "missing export: " + n


def test_ctypes_binding_covers_the_header(built):
    from sobfu_b200 import _capi
    bound = set(_capi.SIGNATURES) | set(_capi.OTHER_SYMBOLS)
    assert set(declared_symbols()) == bound


def test_product_library_does_not_reference_the_oracle(built):
    out = subprocess.run(["nm", "-D", built], capture_output=True, text=True).stdout
    assert "orc_" not in out and "ref_" not in out
    needed = subprocess.run(["readelf", "-d", built], capture_output=True, text=True).stdout
    assert "liboracle" not in needed and "libsobfu_ref" not in needed
    for root, _, files in os.walk(os.path.join(ROOT, "sobfu_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(root, f)).read()
                assert "pyoracle" not in src and "liboracle" not in src and "sobfu_oracle" not in src, f


def test_calls_fail_loudly_without_a_gpu(built):
    """no silent fallback: with no CUDA device the host classes raise"""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sobfu_b200 as sf
    with pytest.raises(sf.Sobfu200Error):
        sf.TsdfVolume(sf.Params(volume_dims=(16, 16, 16)))
    with pytest.raises(sf.Sobfu200Error):
        sf.Solver(sf.Params(volume_dims=(16, 16, 16), max_iter=1))


def test_taps_match_the_reference_table(built):
    """decompose_sobolev_filter (solver.cpp:160-262): host logic, no GPU needed"""
    import numpy as np
    from oracle import pyoracle as orc
    from sobfu_b200._capi import lib
    for s, lams in ((7, (0.05, 0.1, 0.2, 0.4)), (3, (0.1,)), (9, (0.05, 0.1)), (11, (0.1,))):
        for lam in lams:
            t = (ctypes.c_float * 16)()
            assert lib().sobfu_b200_sobolev_taps(s, ctypes.c_float(lam), t) == 0
            got = np.array(list(t)[:s], dtype=np.float32)
            assert np.array_equal(got, orc.sobolev_taps(s, lam))
            assert abs(float(got.sum()) - 1.0) < 1e-6 and np.array_equal(got, got[::-1])
    t = (ctypes.c_float * 16)()
    assert lib().sobfu_b200_sobolev_taps(7, ctypes.c_float(0.3), t) != 0
    assert b"solver.cpp" in lib().sobfu_b200_last_error()


def test_header_is_plain_c_and_links_from_c(built, tmp_path):
    """the drop-in boundary is a C ABI: include/sobfu_b200.h compiles as C99 (-pedantic), a C program links against the library
    and calls host-only entries (version, filter taps, slab partition) without a GPU"""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <sobfu_b200.h>
#include <stdio.h>
int main(void) {
    float taps[11];
    int z0 = -1, nz = -1;
    if (!sobfu_b200_version()) return 1;
    if (sobfu_b200_sobolev_taps(7, 0.1f, taps) != 0) return 2;
    if (sobfu_b200_sobolev_taps(7, 0.3f, taps) == 0 || !sobfu_b200_last_error()[0]) return 3;   /* not tabulated: refused, with a message */
    if (sobfu_b200_sobolev_taps_computed(7, 0.3f, taps) != 0) return 4;
    if (sobfu_b200_slab_range(256, 3, 8, &z0, &nz) != 0 || z0 != 96 || nz != 32) return 5;
    printf("%s %.6f\n", sobfu_b200_version(), taps[3]);
    return 0;
}
''')
    lib = os.path.dirname(built)
    exe = str(tmp_path / "abi")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe,
                           "-L" + lib, "-lsobfu_b200", "-Wl,-rpath," + lib, "-Wl,-rpath,/usr/local/cuda/lib64"])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)

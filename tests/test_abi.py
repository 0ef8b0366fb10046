"""CPU: the C-ABI shared library loads, exports every symbol include/c12381_cuda.h declares (and nothing the Python
binding does not know about), and — there being no CPU fallback — refuses to compute without a CUDA device."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "c12381_cuda.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"^C12381_API [^;(]*?\b(c12381_[a-z0-9_]+)\(", text, flags=re.M)))


@pytest.fixture(scope="module")
def lib():
    from crypto12381_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from crypto12381_b200 import build
        build.build(verbose=False)
    return _lib


def test_header_and_binding_agree(lib):
    syms = declared_symbols()
    assert len(syms) >= 45
    assert syms == sorted(lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (c12381_[a-z0-9_]+)$", out, flags=re.M))
    assert exported == set(declared_symbols())
    l = lib.lib()
    for name in declared_symbols():
        assert getattr(l, name) is not None


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    l = lib.lib()
    assert l.c12381_init(0) == lib.ENODEV
    assert b"no CPU fallback" in l.c12381_last_error()
    out = ctypes.create_string_buffer(49)
    assert l.c12381_g1_msm(bytes(96), bytes(32), 1, out) == lib.ENODEV
    assert l.c12381_pairing_product_batch(bytes(96), bytes(192), 1, 1, ctypes.create_string_buffer(576)) == lib.ENODEV
    from crypto12381_b200 import bridge
    with pytest.raises(lib.C12381Error):
        bridge.sum_of_products(bytes(96), bytes(32))


def test_host_argument_checks():
    from crypto12381_b200 import bridge
    with pytest.raises(ValueError):
        bridge._count(bytes(95), 96, "points")
    with pytest.raises(ValueError):
        bridge._pairs(bytes(96), bytes(192), 0)
    with pytest.raises(ValueError):
        bridge._pairs(bytes(96), bytes(192 * 2), 1)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under crypto12381_b200/ (or the entry points' product paths) imports,
    links or executes it, and the library has no dependency on the reference shim."""
    pkg = os.path.join(ROOT, "crypto12381_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "libref12381" not in text, os.path.join(dirpath, f)
    from crypto12381_b200 import _lib
    if os.path.exists(_lib.LIB_PATH):
        needed = subprocess.run(["readelf", "-d", _lib.LIB_PATH], capture_output=True, text=True).stdout
        assert "libref12381" not in needed and "miracl" not in needed.lower()

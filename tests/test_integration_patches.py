"""CPU: the header / CMake patches of integration/patches/ (INTEGRATION.md §2 option A, §5) still apply to the reference
they were cut against, and the CMake fragment builds the `crypto12381` target with the replacement bridge linked in.
Needs /root/reference (the build container); skipped on the GPU box, which does not have it."""
import glob
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PATCHES = sorted(glob.glob(os.path.join(ROOT, "integration", "patches", "*.patch")))

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference sources are only present in the build container")


def _copy(tmp_path):
    dst = tmp_path / "ref"
    (dst / "include").mkdir(parents=True)
    shutil.copytree(os.path.join(REF, "include", "crypto12381"), dst / "include" / "crypto12381")
    shutil.copy(os.path.join(REF, "CMakeLists.txt"), dst / "CMakeLists.txt")
    return dst


def test_patches_apply_cleanly(tmp_path):
    assert len(PATCHES) == 5
    dst = _copy(tmp_path)
    for p in PATCHES:
        r = subprocess.run(["patch", "-p1", "--fuzz=0", "-s", "-i", p], cwd=dst, capture_output=True, text=True)
        assert r.returncode == 0, (p, r.stdout, r.stderr)
    hdr = (dst / "include" / "crypto12381" / "miracl_core_interface.hpp").read_text()
    assert "void sum_of_products(point2& result, int n, point2* points, const big* numbers) noexcept;" in hdr
    assert "void pair_multi_ate(fp12& result, int n, point2* p2s, point1* p1s) noexcept;" in hdr
    g1 = (dst / "include" / "crypto12381" / "g1_point.hpp").read_text()
    assert "miracl_core::sum_of_products(" in g1 and "// miracl_core::sum_of_products" not in g1
    assert "class G2Pow" in (dst / "include" / "crypto12381" / "g2_point.hpp").read_text()
    assert "pair_multi_ate(" in (dst / "include" / "crypto12381" / "liner_pair.hpp").read_text()


def test_additive_declarations_match_the_replacement_bridge(tmp_path):
    """The patched bridge header and integration/miracl_core_interface_b200.cpp agree: the TU compiles against it."""
    dst = _copy(tmp_path)
    subprocess.run(["patch", "-p1", "-s", "-i", PATCHES[0]], cwd=dst, check=True)
    r = subprocess.run(["g++", "-std=c++23", "-fsyntax-only", "-I", str(dst / "include"), "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "integration", "miracl_core_interface_b200.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_cmake_fragment_builds_the_library(tmp_path):
    from crypto12381_b200 import _lib
    if shutil.which("cmake") is None or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cmake or the built CUDA library is missing")
    src = tmp_path / "src"
    shutil.copytree(REF, src, ignore=shutil.ignore_patterns(".git", "examples", "unit-tests"))
    subprocess.run(["patch", "-p1", "-s", "-i", PATCHES[-1]], cwd=src, check=True)
    b = tmp_path / "b"
    r = subprocess.run(["cmake", "-S", str(src), "-B", str(b), f"-DCRYPTO12381_B200={ROOT}", "-DBUILD_TESTING=OFF"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run(["cmake", "--build", str(b), "--target", "crypto12381", "-j8"], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    syms = subprocess.run(["nm", "-C", str(b / "libcrypto12381.a")], capture_output=True, text=True).stdout
    assert syms.count("refcpu_") >= 9                                   # the stock hot definitions, renamed aside
    assert "c12381_sum_of_products_miracl" in syms                       # ... and the replacement's calls into the CUDA library

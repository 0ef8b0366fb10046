"""GPU (-m gpu): the drop-in build of the reference bridge (integration/build/libcrypto12381_b200.so = the reference's
own src/miracl_core_interface.cpp + MIRACL-core, with the nine hot functions replaced by forwards into
libc12381_cuda.so) driven by integration/bridge_props.cpp — a bridge-level restatement of the reference's unit
tests plus differential checks against the reference's own MIRACL definitions kept under refcpu_*.
Built by `make -C integration` where /root/reference exists; the binaries travel to the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "integration", "build", "bridge_props")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(EXE), reason="integration/build/bridge_props not built (needs /root/reference)")
def test_reference_bridge_with_hot_path_on_gpu():
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout

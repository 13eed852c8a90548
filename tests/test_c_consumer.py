"""The boundary from the other side: include/*.h compile as plain C11 and a C program linked against libvindex_b200.so
runs (what the reference's SwiftPM system-library target / a cgo binding would consume).  No GPU: it must see every entry
point refuse with VIX_ERR_NO_DEVICE.  GPU: the same program runs a flat search and the reference's cpq_* encoder symbol."""
import os
import shutil
import subprocess

import pytest

from conftest import HAS_GPU, ROOT


def _build(tmp_path):
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if cc is None:
        pytest.skip("no C compiler")
    lib_dir = os.path.join(ROOT, "vectorindex_b200")
    if not os.path.exists(os.path.join(lib_dir, "libvindex_b200.so")):
        pytest.skip("libvindex_b200.so not built")
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call([cc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe, "-L", lib_dir, "-lvindex_b200",
                           "-Wl,-rpath," + lib_dir, "-lm"])
    return exe


@pytest.mark.skipif(HAS_GPU, reason="the no-device behaviour")
def test_c_program_links_and_sees_no_cpu_fallback(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "no device" in out.stdout


@pytest.mark.gpu
def test_c_program_runs_on_the_device(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "device: flat search" in out.stdout

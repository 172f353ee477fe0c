"""The boundary is a C ABI: include/pcdn_fanout.h must compile as plain C (what cgo, Rust bindgen or
JNI consume) and the shared library must be usable from a C program with no Python or torch in the
process.  No GPU needed: the C program drives a host-only engine."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_header_is_plain_c_and_library_links_from_c(pcdn, tmp_path):
    exe = tmp_path / "abi_smoke"
    libdir = os.path.dirname(pcdn.LIB_PATH)
    cmd = ["gcc", "-std=c11", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi", "abi_smoke.c"), "-o", str(exe),
           "-L", libdir, "-lpcdn_fanout", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_smoke ok" in r.stdout

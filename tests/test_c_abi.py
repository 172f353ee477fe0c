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


def _build_c(pcdn, tmp_path, name):
    exe = tmp_path / name
    libdir = os.path.dirname(pcdn.LIB_PATH)
    cmd = ["gcc", "-std=c11", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi", name + ".c"), "-o", str(exe),
           "-L", libdir, "-lpcdn_fanout", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_gpu_roundtrip_program_compiles(pcdn, tmp_path):
    _build_c(pcdn, tmp_path, "gpu_roundtrip")


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c_host_drives_the_engine_on_the_gpu(pcdn, tmp_path):
    """no Python, no torch in the process: a C host does state calls, routes two broadcasts and two
    direct messages, polls, reads the framed records (pcdn_read and in place with host rings)"""
    exe = _build_c(pcdn, tmp_path, "gpu_roundtrip")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "gpu_roundtrip ok" in r.stdout

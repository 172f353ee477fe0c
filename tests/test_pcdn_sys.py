"""The Rust `-sys` crate (pcdn-sys/) is generated from include/pcdn_fanout.h and committed uncompiled
(no Rust toolchain in this image): the generator must reproduce the committed file exactly, and every
function it declares must be exported by libpcdn_fanout.so with the arity the ctypes binding uses."""
import ctypes as C
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_crate_matches_header():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_pcdn_sys.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_crate_functions_exist_with_the_same_arity(pcdn):
    src = open(os.path.join(ROOT, "pcdn-sys", "src", "lib.rs")).read()
    fns = re.findall(r"pub fn (pcdn_\w+)\((.*?)\)(?: -> [^;]+)?;", src)
    assert len(fns) >= 50
    L = C.CDLL(pcdn.LIB_PATH)
    for name, params in fns:
        assert hasattr(L, name), name
        n = 0 if not params.strip() else len(params.split(","))
        assert n == len(pcdn.ABI[name][1]), (name, n, len(pcdn.ABI[name][1]))
    assert {f for f, _ in fns} == set(pcdn.ABI)
    # struct sizes the Rust side would compute (repr(C), same field order) == the ctypes mirrors
    for rust, ct in (("pcdn_config", pcdn.Config), ("pcdn_span", pcdn.Span), ("pcdn_batch_result", pcdn.BatchResult),
                     ("pcdn_stats", pcdn.Stats), ("pcdn_shard_desc", pcdn.ShardDesc), ("pcdn_egress_chunk", pcdn.EgressChunk)):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % rust, src, re.S).group(1)
        assert len(re.findall(r"pub \w+:", body)) == len(ct._fields_), rust

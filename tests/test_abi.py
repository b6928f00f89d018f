"""The C-ABI libraries load and export every symbol the headers declare; without a GPU the product refuses to run."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", text)))


def exported(path):
    out = subprocess.run(["nm", "-D", "--defined-only", path], check=True, capture_output=True, text=True).stdout
    return {line.split()[-1] for line in out.splitlines() if " T " in line}


@pytest.fixture(scope="module")
def built():
    subprocess.run(["make", "-C", ROOT, "lib", "host"], check=True, stdout=subprocess.DEVNULL)
    return os.path.join(ROOT, "zkcnn_b200", "lib")


def test_cuda_library_exports_the_header(built):
    syms = exported(os.path.join(built, "libzkcnn_b200.so"))
    want = declared("zkcnn_b200.h", "zk_")
    assert len(want) >= 45
    assert [s for s in want if s not in syms] == []


def test_host_library_exports_the_header(built):
    syms = exported(os.path.join(built, "libzkcnn_host.so"))
    want = declared("zkcnn_host.h", "zkh_")
    assert [s for s in want if s not in syms] == []


def test_product_library_is_cuda_for_sm_100a(built):
    lib = os.path.join(built, "libzkcnn_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    dll = ctypes.CDLL(lib)
    dll.zk_version.restype = ctypes.c_char_p
    assert b"sm_100a" in dll.zk_version()


def test_no_cpu_fallback(built):
    """on a box without a GPU the product library must fail loudly, never compute"""
    dll = ctypes.CDLL(os.path.join(built, "libzkcnn_b200.so"))
    if dll.zk_device_count() > 0:
        pytest.skip("a GPU is present")
    dll.zk_ctx_create.restype = ctypes.c_void_p
    dll.zk_last_error.restype = ctypes.c_char_p
    assert dll.zk_ctx_create(0) is None
    assert b"no CUDA device" in dll.zk_last_error()
    import zkcnn_b200
    with pytest.raises(zkcnn_b200.ZkError):
        zkcnn_b200.context()


def test_product_does_not_reference_the_oracle():
    """the oracle and the emulator are test infrastructure: nothing under zkcnn_b200/ may import, include or link them"""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zkcnn_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for pat in ("zkcnn_oracle", "oracle/", "libref_", "cuda_emu.cpp"):
                    if pat in text and not (pat == "oracle/" and f.endswith((".hpp", ".cpp", ".cuh", ".cu", ".py")) and
                                            all("oracle/" not in line or line.lstrip().startswith(("//", "#", "*", '"""')) or "//" in line
                                                for line in text.splitlines())):
                        bad.append((f, pat))
    assert bad == []

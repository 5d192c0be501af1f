"""CPU-only checks of the drop-in boundary: libicp_b200.so loads, exports every symbol that
include/icp_b200.h declares, the ctypes mirror of icp_b200_batch matches the header, and — with no
CUDA device in this container — compute entries fail loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import icp_rust_b200 as icp
from icp_rust_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "icp_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(icp_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    names = _declared_functions()
    assert len(names) >= 18
    L = _lib.cuda_lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/icp_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names


def test_exports_are_unmangled_c_symbols():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.CUDA_SO], capture_output=True, text=True, check=True).stdout
    exported = set(l.split()[-1] for l in out.splitlines() if " T " in l)
    for n in _lib.EXPORTS:
        assert n in exported


def test_abi_version_and_status_strings():
    L = _lib.cuda_lib()
    assert L.icp_b200_abi_version() == icp.ABI_VERSION == 4
    m = re.search(r"#define ICP_B200_ABI_VERSION (\d+)", open(HEADER).read())
    assert int(m.group(1)) == icp.ABI_VERSION
    assert L.icp_b200_status_string(0) == b"ok"
    assert b"empty" in L.icp_b200_status_string(icp.ERR_EMPTY_TARGET)
    assert L.icp_b200_status_string(12345) == b"unknown status"


def test_batch_struct_layout_matches_header():
    """Compile a 10-line C program against the header and compare sizeof/offsetof with ctypes."""
    fields = [f for f, _ in _lib.Batch._fields_]
    prog = "#include <stdio.h>\n#include <stddef.h>\n#include \"icp_b200.h\"\nint main(){printf(\"%zu\", sizeof(icp_b200_batch));" + \
        "".join(f'printf(" %zu", offsetof(icp_b200_batch, {f}));' for f in fields) + \
        'printf(" %zu", sizeof(icp_b200_call_info));return 0;}\n'
    d = os.path.join(ROOT, "tests", "simt", "_build")
    os.makedirs(d, exist_ok=True)
    src, exe = os.path.join(d, "abi_layout.c"), os.path.join(d, "abi_layout")
    open(src, "w").write(prog)
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", exe, src], check=True)
    vals = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert vals[0] == C.sizeof(_lib.Batch)
    for f, off in zip(fields, vals[1:-1]):
        assert getattr(_lib.Batch, f).offset == off, f
    assert vals[-1] == C.sizeof(_lib.CallInfo)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    L = _lib.cuda_lib()
    n = C.c_int(-1)
    rc = L.icp_b200_device_count(C.byref(n))
    assert rc == 3 and n.value == 0          # ICP_B200_ERR_NO_DEVICE
    with pytest.raises(icp.IcpError) as e:
        icp.Context(0)
    assert e.value.status == 3
    with pytest.raises(icp.IcpError):
        icp.Icp2d.new(np.zeros((4, 2))).estimate(np.zeros((4, 2)), icp.Transform.identity(), 1)


def test_null_arguments_are_rejected():
    L = _lib.cuda_lib()
    assert L.icp_b200_create(0, None) == 1   # ICP_B200_ERR_INVALID_ARGUMENT
    assert L.icp_b200_estimate_batch(None, None) == 1
    assert L.icp_b200_last_error(None) == b"null context"
    assert L.icp_b200_estimate_batch_submit(None, None) == 1
    assert L.icp_b200_estimate_batch_wait(None) == 1
    assert L.icp_b200_target_estimate_batch_submit(None, None) == 1
    L.icp_b200_destroy(None)                 # no-op
    L.icp_b200_target_destroy(None)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under icp_rust_b200/ or include/ may import,
    include or load it, and the shipped library must not contain its symbols."""
    for base in ("icp_rust_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "obj" in dirpath.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c", "Makefile")):
                    text = open(os.path.join(dirpath, f), errors="replace").read()
                    assert "import oracle" not in text and "icp_oracle" not in text and "orc_" not in text, os.path.join(dirpath, f)
    out = subprocess.run(["nm", "-D", _lib.CUDA_SO], capture_output=True, text=True, check=True).stdout
    assert "orc_" not in out

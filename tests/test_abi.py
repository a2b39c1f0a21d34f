"""CPU: the C-ABI library loads, exports every symbol the headers declare, and fails loudly
(no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from fugu_b200 import _native as nat
from fugu_b200 import dataset as dsm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fgh?_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol():
    L = nat.lib()
    names = _declared("fugu_gpu.h") + _declared("fugu_host.h")
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ but not exported by libfugu_gpu.so"
    assert sorted(nat.ABI_SYMBOLS) == _declared("fugu_gpu.h")
    assert sorted(dsm.HOST_SYMBOLS) == _declared("fugu_host.h")


def test_abi_struct_sizes():
    # layouts the reference-side binding must reproduce
    assert nat.LEAF_DT.itemsize == 12 and nat.CLAUSE_DT.itemsize == 12 and nat.QUERY_DT.itemsize == 12
    assert nat.HIT_DT.itemsize == 8
    assert C.sizeof(nat.FieldDesc) == 56 and C.sizeof(nat.IndexDesc) == 40 and C.sizeof(nat.QueryBatch) == 40


def test_no_device_is_a_hard_error():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(nat.FgError) as e:
        nat.Context(0)
    assert e.value.code == nat.FG_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)
    # the host layer plans without a device but refuses to search
    ds = dsm.Dataset(None)
    ds.upsert([dsm.ObjectRecord(id="a", text="hello world")], commit=False)
    assert ds.plan("hello").n_clauses == 1
    with pytest.raises(nat.FgError) as e:
        ds.commit()
    assert e.value.code == nat.FG_ERR_NO_DEVICE
    with pytest.raises(nat.FgError) as e:
        ds.search("hello")
    assert e.value.code == nat.FG_ERR_NO_DEVICE
    ds.close()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fugu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("the oracle", "").replace("not the oracle", "").replace("NOT part of the oracle", "").replace("# noqa", "") \
                    or f in ("synth.py", "_native.py", "synth.cpp"), f"{f} references the oracle"
    for f in ("_native.py", "synth.py", "dataset.py", "__init__.py"):
        src = open(os.path.join(pkg, f)).read()
        assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_product_never_loads_the_emulated_library():
    """tests/emu/libfugu_emu.so (the kernel sources under a SIMT emulation) is test infrastructure: no
    product module, the bench or the driver entry points may name it, and the library the product loads
    is the nvcc build next to the package."""
    for f in ("bench.py", "__graft_entry__.py", "fugu_b200/_native.py", "fugu_b200/dataset.py", "fugu_b200/synth.py", "fugu_b200/__init__.py"):
        src = open(os.path.join(ROOT, f)).read()
        assert "libfugu_emu" not in src and "tests/emu" not in src and "FG_EMULATE" not in src, f
    assert nat.LIB_PATH == os.path.join(ROOT, "fugu_b200", "libfugu_gpu.so")
    # the product build defines nothing of the emulation: FG_EMULATE comes from tests/emu/cuda_runtime.h alone
    mk = open(os.path.join(ROOT, "Makefile")).read()
    assert "FG_EMULATE" not in mk and "tests/emu" not in mk

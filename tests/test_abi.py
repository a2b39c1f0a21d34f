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
    L, H = nat.lib(), nat.host_lib()
    dev_names, host_names = _declared("fugu_gpu.h"), _declared("fugu_host.h")
    assert len(dev_names) + len(host_names) >= 30
    for n in dev_names:
        assert hasattr(L, n), f"{n} declared in include/fugu_gpu.h but not exported by libfugu_gpu.so"
    for n in host_names:
        assert hasattr(H, n), f"{n} declared in include/fugu_host.h but not exported by libfugu_host.so"
    assert sorted(nat.ABI_SYMBOLS) == dev_names
    assert sorted(dsm.HOST_SYMBOLS) == host_names


def test_host_library_maps_no_cuda_code():
    """libfugu_host.so (planner, tokenizer, dataset builder) neither contains nor links CUDA code, and exports
    nothing of the device ABI: a process that only plans (the bench's reference arm, a GPU-less Rust build)
    maps no CUDA library. Checked in a fresh interpreter: planning works and /proc/self/maps stays clean."""
    import subprocess
    import sys

    host = os.path.join(ROOT, "fugu_b200", "libfugu_host.so")
    needed = subprocess.run(["objdump", "-p", host], capture_output=True, text=True).stdout
    assert "libcudart" not in needed and "libcuda" not in needed and "libfugu_gpu" not in needed and "libnccl" not in needed
    syms = subprocess.run(["nm", "-D", "--defined-only", host], capture_output=True, text=True).stdout
    exported = sorted(ln.split()[-1] for ln in syms.splitlines() if " T " in ln)
    assert exported and all(n.startswith("fgh_") for n in exported), exported
    dev = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "fugu_b200", "libfugu_gpu.so")], capture_output=True, text=True).stdout
    assert " T fgh_" not in dev  # the device library carries no host-layer code
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from fugu_b200 import dataset as dsm\n"
        "ds = dsm.Dataset(None)\n"
        "ds.upsert([dsm.ObjectRecord(id='a', text='hello world')], commit=False)\n"
        "assert ds.plan('hello AND world').n_clauses == 2\n"
        "assert dsm.tokenize('Hello, World') == ['hello', 'world']\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'libfugu_host.so' in maps\n"
        "assert 'libfugu_gpu' not in maps and 'libcudart' not in maps and 'libcuda.so' not in maps, 'CUDA code mapped'\n"
        "print('clean')\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "clean" in r.stdout, r.stdout + r.stderr


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

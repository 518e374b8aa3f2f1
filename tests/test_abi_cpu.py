"""CPU-only checks of the drop-in boundary: the C-ABI library builds/loads, exports every symbol
include/dfgpu.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re
import subprocess

import pytest

from datafusion_archive_b200 import _abi as A
from datafusion_archive_b200 import engine

ROOT = A.repo_root()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dfgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dfgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for s in ["dfgpu_init", "dfgpu_batch_upload", "dfgpu_filter_project", "dfgpu_aggregate_create",
              "dfgpu_aggregate_update", "dfgpu_aggregate_finish", "dfgpu_result_copy_col", "dfgpu_comm_init"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    engine.build()
    L = ctypes.CDLL(engine.lib_path())
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert missing == []
    assert L.dfgpu_abi_version() == A.ABI_VERSION


def test_host_mirror_library_exports_its_header():
    # include/dfhost.h: the harness API of the C++ host mirror (not the drop-in boundary)
    from datafusion_archive_b200 import host
    host.build()
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "dfhost.h")).read(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(dfhost_[a-z0-9_]+)\s*\(", src)))
    assert len(syms) >= 25
    L = ctypes.CDLL(os.path.join(ROOT, "datafusion_archive_b200", "libdfhost.so"))
    assert [s for s in syms if not hasattr(L, s)] == []


def test_struct_layout_matches_header():
    # sizes the C side static-asserts implicitly through use; keep the ctypes mirror honest
    assert ctypes.sizeof(A.Insn) == 24
    assert ctypes.sizeof(A.Col) == 56
    assert ctypes.sizeof(A.Agg) == 24


def test_no_cpu_fallback_without_gpu():
    if engine.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(engine.DfGpuError) as e:
        engine.GpuContext(0)
    assert e.value.code == A.ERR_CUDA
    assert "no CPU fallback" in e.value.msg


def test_product_does_not_reference_oracle():
    # the oracle is test infrastructure: nothing under the product package may import/link it
    pkg = os.path.join(ROOT, "datafusion_archive_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if "oracle" in txt.lower() and f != "test_abi_cpu.py":
                    for line in txt.splitlines():
                        if re.search(r"(import|include|dlopen|CDLL|-l).*oracle", line):
                            bad.append((f, line.strip()))
    assert bad == []
    # and the shared libraries do not depend on it
    for so in ("libdfgpu.so", "libdfhost.so"):
        path = os.path.join(pkg, so)
        if os.path.exists(path):
            assert "oracle" not in subprocess.run(["ldd", path], capture_output=True, text=True).stdout



def test_rust_shim_declares_every_entry_point():
    # shim/src/execution/gpu/ffi.rs is the `extern "C"` block a maintainer adds to the reference:
    # it has to name every function of include/dfgpu.h, with the same number of parameters
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "dfgpu.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    ffi = open(os.path.join(root, "shim", "src", "execution", "gpu", "ffi.rs")).read()
    ffi = re.sub(r"//[^\n]*", "", ffi)
    cdecl = {m.group(1): m.group(2) for m in re.finditer(r"\b(dfgpu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)}
    rdecl = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (dfgpu_[a-z0-9_]+)\s*\((.*?)\)\s*->", ffi, flags=re.S)}
    assert len(cdecl) >= 39
    for name, params in cdecl.items():
        assert name in rdecl, "%s is not declared in shim/src/execution/gpu/ffi.rs" % name
        nc = 0 if params.strip() in ("", "void") else params.count(",") + 1
        rp = rdecl[name].strip().rstrip(",")
        nr = 0 if not rp else rp.count(",") + 1
        assert nc == nr, "%s: %d parameters in the header, %d in ffi.rs" % (name, nc, nr)

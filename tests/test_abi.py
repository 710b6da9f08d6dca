"""CPU test: the C-ABI library loads and exports every symbol include/horae_gpu.h declares (no compute calls)."""
import os
import re

from conftest import ROOT
from horaedb_b200 import _ffi


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "horae_gpu.h")).read()
    declared = set(re.findall(r"^(?:int|void\*?|uint32_t|const char\*)\s+(hg_[a-z_]+)\s*\(", hdr, flags=re.M))
    assert declared, "no declarations found"
    L = _ffi.lib()
    for sym in sorted(declared):
        assert hasattr(L, sym), f"{sym} declared in horae_gpu.h but not exported"
    assert set(_ffi.EXPORTS) == declared
    want = int(re.search(r"#define HG_ABI_VERSION (\d+)u", hdr).group(1))
    assert L.hg_abi_version() == want


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(_ffi.HgPredicate) == 48
    assert C.sizeof(_ffi.HgSstDesc) == 64
    assert C.sizeof(_ffi.HgSchemaDesc) == 32
    assert C.sizeof(_ffi.HgAggSpec) == 24
    assert C.sizeof(_ffi.HgScanStats) == 88
    assert C.sizeof(_ffi.ArrowArrayStream) == 40


def test_engine_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(_ffi.HgError) as ei:
        _ffi.Engine(device=0)
    assert "no CPU fallback" in str(ei.value)

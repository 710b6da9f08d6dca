"""TEST INFRASTRUCTURE: `pytest -p emu_plugin -m gpu ...` (with tests/emu on PYTHONPATH) runs the GPU parity tests against the EMULATED
build of the library (tests/emu/build_engine_emu.py: the product's kernels compiled for the CPU, threads as coroutines) instead of
libhorae_gpu.so.  Only tests/test_emu_engine.py uses it, in a subprocess; the product and the `-m gpu` run on a GPU box never see it.
HORAE_EMU_ORDER = 0 | 1 | n chooses the order threads run in between two barriers (cuda_emu.h).  After every test the scheduler's error word
is read: lanes meeting in different collectives, collectives naming exited lanes, barriers that cannot complete fail the test."""
import ctypes as C
import os

import pytest

_emu = None


def pytest_configure(config):
    global _emu
    import build_engine_emu
    from horaedb_b200 import _ffi
    path = build_engine_emu.build()
    _ffi.LIB_PATH = path
    _ffi._lib = None
    _emu = C.CDLL(path)
    _emu.emu_set_order(int(os.environ.get("HORAE_EMU_ORDER", "0")))


# tests that need the real device: torch tensors over device pointers, NCCL, and the 100 M-row property test
NEEDS_DEVICE = ("test_aggregate_device_result", "test_config2_full_size_properties", "test_gpu_nccl_combine", "test_transient_selective_load_matches_resident",
                "test_transient_gate_column_prunes_row_groups")   # (the last two pin host memory through torch)


def pytest_collection_modifyitems(config, items):
    keep, drop = [], []
    for it in items:
        (drop if any(n in it.nodeid for n in NEEDS_DEVICE) else keep).append(it)
    if drop:
        config.hook.pytest_deselected(items=drop)
        items[:] = keep


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_call(item):
    line = C.c_int(0)
    _emu.emu_take_error(C.byref(line))
    yield
    err = _emu.emu_take_error(C.byref(line))
    if err:
        pytest.fail(f"emulated scheduler error {err} at source line {line.value}", pytrace=False)

"""tests/c/c_abi_smoke.c: the header bound from plain C (gcc), linked against libhorae_gpu.so — what the Rust FFI does.
Without a GPU it checks the host-only entry points and that engine creation fails loudly; on a GPU box it runs a scan."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_and_run(tmp_path):
    from horaedb_b200 import sstgen
    data, n = sstgen.synth_sst(0, 64, 300, 1000, seq=5, compression="snappy")
    sst = tmp_path / "5.sst"
    sst.write_bytes(data)
    exe = str(tmp_path / "c_abi_smoke")
    libdir = os.path.join(ROOT, "horaedb_b200", "csrc")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "c_abi_smoke.c"),
                           "-o", exe, "-L", libdir, "-lhorae_gpu", f"-Wl,-rpath,{libdir}"])
    p = subprocess.run([exe, str(sst)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def test_c_binding_host_only(tmp_path):
    out = _build_and_run(tmp_path)
    assert "c_abi_smoke:" in out


@pytest.mark.gpu
def test_c_binding_on_gpu(tmp_path):
    out = _build_and_run(tmp_path)
    assert "c_abi_smoke: ok" in out

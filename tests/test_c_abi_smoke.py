"""tests/c_abi_smoke.c: the C ABI driven from plain C (no Python/ctypes between the caller and libtrixi_b200.so),
checked against the oracle with the same descriptor.  Compiling and linking is a CPU test; running needs a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "trixi.jl_b200")
ORADIR = os.path.join(ROOT, "oracle")


def _build(tmp_path, oracle_module):
    import __graft_entry__ as entry
    if not os.path.exists(os.path.join(LIBDIR, "libtrixi_b200.so")):
        entry.build()
    exe = str(tmp_path / "c_abi_smoke")
    cmd = ["gcc", "-O2", "-std=gnu11", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe,
           os.path.join(LIBDIR, "libtrixi_b200.so"), os.path.join(ORADIR, "libtrixi_oracle.so"),
           "-lm", f"-Wl,-rpath,{LIBDIR}", f"-Wl,-rpath,{ORADIR}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_c_abi_smoke_links(tmp_path, oracle_module):
    """Every symbol the C host uses resolves against the shared library; without a device the program reports the
    library's ENODEVICE (no CPU fallback) with exit code 77."""
    exe = _build(tmp_path, oracle_module)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert r.returncode == 77, r.stderr
        assert "no CUDA device" in r.stderr or "CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_abi_smoke_runs(tmp_path, oracle_module):
    exe = _build(tmp_path, oracle_module)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_abi_smoke: OK" in r.stdout

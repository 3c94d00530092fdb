"""The boundary is a plain C ABI: tests/c/abi_smoke.c (C99, no C++ / Python) compiles against include/sclane_b200.h with gcc,
links libsclane_b200.so, and -- on a GPU box -- runs the whole path and checks the records."""
import os
import subprocess

import pytest

import __graft_entry__ as ge

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_smoke.c")
EXE = os.path.join(ROOT, "tests", "c", "abi_smoke")


def _build():
    ge.build()
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
           "-L", ge.LIBDIR, "-lsclane_b200", "-lm", f"-Wl,-rpath,{ge.LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_header_compiles_as_c99_and_links():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    # on a CPU box the program stops at the device check: no CPU fallback
    assert r.returncode in (0, 3), r.stdout + r.stderr
    if r.returncode == 3:
        assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_c_caller_runs_the_path():
    _build()
    r = subprocess.run([EXE, "12", "6000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bad 0" in r.stdout

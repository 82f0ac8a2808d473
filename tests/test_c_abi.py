"""A C program (gcc -std=c99, no Python in between) that includes include/b2s.h, links libb2s.so and runs config 1 through the
boundary -- what a maintainer's shim does.  The CPU test compiles and links it (every symbol it uses resolves, the header is
valid C99); the -m gpu test runs it on the device."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "config1.c")
LIBDIR = os.path.join(ROOT, "open3d_slam_b200")


def build(out):
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-pedantic", "-I", os.path.join(ROOT, "include"), SRC, "-o", out, "-L", LIBDIR, "-lb2s",
           "-Wl,-rpath," + LIBDIR, "-lm"]
    subprocess.check_call(cmd)
    return out


def test_c_program_compiles_and_links(tmp_path):
    exe = build(str(tmp_path / "config1"))
    assert os.path.exists(exe)
    # without a GPU the program must refuse loudly (exit 77: no device), never fall back
    from open3d_slam_b200 import _lib
    if _lib.lib().b2s_device_count() == 0:
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 77 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_program_runs_config1(tmp_path):
    exe = build(str(tmp_path / "config1"))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASS" in r.stdout

"""The C++ host mirror (rust-lz-fear_b200/host/lz_fear.hpp) built with g++ against the C ABI and run the way the
reference's own unit tests use the crate: on the SIMT-emulated build here, on the shipped .so in the gpu tier."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_and_run(tmp_path, libdir, libname, env_extra=None):
    exe = str(tmp_path / "host_mirror_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "host_mirror_test.cpp"),
                           "-L" + libdir, "-l" + libname, "-Wl,-rpath," + libdir])
    env = dict(os.environ, **(env_extra or {}))
    p = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0 and "HOST-MIRROR-OK" in p.stdout, p.stdout + p.stderr


def test_cpp_host_mirror_on_simt_build(tmp_path, simt_lib_path):
    _build_and_run(tmp_path, os.path.dirname(simt_lib_path), "simt_lzfear")


@pytest.mark.gpu
def test_cpp_host_mirror_on_gpu(tmp_path, gpu):
    _build_and_run(tmp_path, os.path.join(ROOT, "rust-lz-fear_b200"), "lzfear_b200")

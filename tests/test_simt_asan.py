"""Memory-safety tier: the product kernels + C-ABI layer under AddressSanitizer on the CPU SIMT
emulator (tests/simt).  Any out-of-bounds access in the kernels aborts the subprocess."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kernels_and_abi_under_asan():
    libasan = subprocess.run(["/usr/bin/gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not libasan or not os.path.exists(libasan):
        pytest.skip("libasan not available")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "simt"), "libsimt_lzfear_asan.so"])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "simt", "asan_run.py")], env=env,
                       capture_output=True, text=True, timeout=1500)
    assert p.returncode == 0 and "ASAN-RUN-OK" in p.stdout, (p.stdout[-2000:], p.stderr[-6000:])

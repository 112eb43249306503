"""CPU: the host-side C++ of the library under sanitizers (no GPU, no CUDA): the symbolic / coupling analyses with
AddressSanitizer + UBSan + bounds-checked libstdc++, the threaded copy pool with ThreadSanitizer.  The harnesses live in
``tools/`` (``asan_host_analysis.cpp``, ``tsan_copy_pool.cpp``); a short batch of each runs here."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "parapint_b200", "csrc")


def _build_and_run(tmp_path, source, flags, args):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "harness")
    cmd = ["g++", "-std=c++17", "-O1", "-g", *flags, "-pthread", "-I", CSRC, os.path.join(ROOT, "tools", source), "-o", exe]
    built = subprocess.run(cmd, capture_output=True, text=True)
    if built.returncode != 0 and ("sanitize" in built.stderr or "cannot find" in built.stderr):
        pytest.skip("sanitizer runtime not available: " + built.stderr.strip().splitlines()[-1])
    assert built.returncode == 0, built.stderr
    run = subprocess.run([exe, *args], capture_output=True, text=True, timeout=600)
    if run.returncode != 0 and "FATAL: ThreadSanitizer: unexpected memory mapping" in run.stderr:
        pytest.skip("ThreadSanitizer cannot run in this kernel configuration")
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-4000:]
    assert run.stdout.strip().splitlines()[-1].startswith("ok:")


def test_host_analyses_under_asan_ubsan(tmp_path):
    _build_and_run(tmp_path, "asan_host_analysis.cpp",
                   ["-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-D_GLIBCXX_ASSERTIONS"], ["60", "4"])


def test_copy_pool_under_tsan(tmp_path):
    _build_and_run(tmp_path, "tsan_copy_pool.cpp", ["-fsanitize=thread"], ["100", "5"])

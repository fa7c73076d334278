"""The C++ host layer (mirror of the Rust crates + divsuftest harness) on a GPU box."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
HOST = os.path.join(ROOT, "stringsearch_b200", "host")


def test_reference_unit_tests_through_cpp_mirror():
    """crates/sacapart/src/lib.rs:105-165 and crates/divsufsort/src/lib.rs:84-91 restated in C++."""
    r = subprocess.run([os.path.join(HOST, "host_tests")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all passed" in r.stdout


def test_divsuftest_cli(tmp_path):
    from stringsearch_b200 import synth

    p = tmp_path / "input.bin"
    synth.acgt(4 << 20, 1).tofile(p)  # BASELINE config 0: 4 MiB ACGT, GPU verified against cdivsufsort
    ref_lib = os.path.join(ROOT, "oracle", "_ref", "libdivsufsort_ref.so")
    cmd = [os.path.join(HOST, "divsuftest"), "bench", str(p), "4m", "--partitions", "4"]
    if os.path.exists(ref_lib):
        cmd += ["--cpu-lib", ref_lib]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "gpu-divsufsort (device only)" in r.stdout and "gpu-sacapart (4 partitions)" in r.stdout
    if os.path.exists(ref_lib):
        assert "IDENTICAL" in r.stdout
    r = subprocess.run([os.path.join(HOST, "divsuftest"), "verify", str(p), "512k"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "verified" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([os.path.join(HOST, "divsuftest"), "run", str(p), "1m"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Done in" in r.stdout

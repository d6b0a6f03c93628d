"""Builds and runs the host check of the in-register DFT templates (csrc/sml_dft.cuh) with g++."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_register_dfts_match_naive_dft(tmp_path):
    exe = str(tmp_path / "dft_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "tensor-cuda-fft-_b200", "csrc"),
                    os.path.join(ROOT, "tests", "host", "dft_host_check.cpp"), "-o", exe], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    assert "WORST" in res.stdout

"""CPU check of the fused 8192-point convolution stages (csrc/conv8k.cuh): the header is compiled with g++ against a small CUDA
shim (tests/host/cuda_shim/common.cuh) and run thread by thread, stage by stage, against a direct float64 circular
convolution -- index math, twiddles, the pre-permuted filter spectrum (float64 fft8k_dif order) and the two-filter stash path
of the FFT-form zc_freq kernel.  No GPU."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_conv8k_stages_on_the_host(tmp_path):
    csrc = ROOT / "ofdm_sync_math_b200" / "csrc"
    for name in ("fft4096.cuh", "conv8k.cuh"):       # copied so that their `#include "common.cuh"` finds the shim, not csrc/common.cuh
        shutil.copy(csrc / name, tmp_path / name)
    exe = tmp_path / "conv8k_host_test"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", str(ROOT / "tests" / "host" / "cuda_shim"), "-I", str(tmp_path), "-o", str(exe),
                    str(ROOT / "tests" / "host" / "conv8k_host_test.cpp")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "PASS" in out.stdout, out.stdout + out.stderr

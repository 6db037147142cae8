"""Test configuration.

Markers:
  gpu  -- needs a CUDA device (the parity tests proper; run on the B200 box)
Everything else runs on the CPU: the oracle against golden vectors, the host logic, the C-ABI
export check, and the *real kernel sources* executed by the CPU emulator (tools/cudaemu).
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def port():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle.Port()


@pytest.fixture(scope="session")
def ref():
    from oracle import pyoracle
    pyoracle.build()
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return pyoracle.Ref()


EMU_CUDA = os.path.join(ROOT, "bch_fft_b200", "csrc", "libbchfft_emu.so")
EMU_HOST = os.path.join(ROOT, "host", "libbchfft_host_emu.so")


@pytest.fixture(scope="session")
def emu_libs():
    """(C-ABI ctypes lib, Host binding) of the CPU-emulator build of the kernel sources."""
    subprocess.run(["make", "-C", ROOT, "emu", "-j8"], check=True, stdout=subprocess.DEVNULL)
    from bch_fft_b200 import _lib, api
    cabi = _lib.bind(EMU_CUDA)
    assert cabi.bchfft_is_emulated() == 1
    return cabi, api.Host(EMU_HOST)


@pytest.fixture(scope="session")
def gpu_libs():
    """(C-ABI ctypes lib, Host binding) of the real CUDA build; fails loudly if it is missing."""
    from bch_fft_b200 import _lib, api
    cabi = _lib.lib()
    assert cabi.bchfft_is_emulated() == 0
    assert cabi.bchfft_device_count() > 0, cabi.bchfft_last_error()
    return cabi, api.Host()

"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` : oracle vs golden vectors / the real reference, host logic, C-ABI
                 loading -- runs anywhere (no CUDA calls that need a device).
`-m gpu`       : the parity tests proper; every one goes through the C ABI
                 (liblorab200.so) on cuda:0 and checks against oracle/.
"""
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def reference():
    """The real reference headers compiled in place (oracle/_ref). Skips if unbuilt."""
    from oracle import pyoracle
    if not pyoracle.have_reference():
        pytest.skip("oracle/_ref/libloraref.so not available (needs /root/reference at build time)")
    return pyoracle.Reference()


@pytest.fixture(scope="session")
def lib():
    """The product library, built in-tree (nvcc cross-compiles without a GPU)."""
    import lora_sdr_b200 as m
    m.build()
    m.load_library()
    return m


@pytest.fixture(scope="session")
def cuda_ok():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("test marked gpu but no CUDA device is visible (no CPU fallback exists)")
    return True

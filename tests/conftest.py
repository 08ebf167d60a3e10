import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "needs_cuda: gpu test that cannot run on the SIMT emulator (MPMB_SIMT=1)")


def _simt():
    """MPMB_SIMT=1: run the gpu-marked tests on the SIMT emulator build of the engine (tests/simt) — no GPU needed;
    tests that need a real device (torch CUDA tensors, NCCL, IPC, a linked C++ host) carry `needs_cuda` and are skipped."""
    return os.environ.get("MPMB_SIMT", "") == "1"


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _simt():
        from tests.simt import build_simt
        defines = [d for d in os.environ.get("MPMB_SIMT_DEFINES", "").split(",") if d]   # e.g. MPMB_EXP_P2G_IPLANE: build-time experiments
        os.environ["MPMB_LIB"] = build_simt.build(defines)
        skip = pytest.mark.skip(reason="needs a real CUDA device (not the SIMT emulator)")
        for item in items:
            if "needs_cuda" in item.keywords:
                item.add_marker(skip)
        return
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


# B2_EMU_RUN=1: run the gpu-marked parity tests against the CPU emulation of the kernels (tests/emu) instead of
# skipping them — a development aid for kernels that have not been on hardware yet; never a substitute for `-m gpu`
# on a B200 (the emulator knows nothing about memory ordering or speed). Tests that touch torch.cuda directly fail.
EMU_RUN = os.environ.get("B2_EMU_RUN", "0") == "1"


def pytest_collection_modifyitems(config, items):
    if _has_gpu() or EMU_RUN:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def plc():
    """The pylibcudf-named binding; building the library first if needed (build is the driver's job normally)."""
    import __graft_entry__ as g

    g.build()
    if EMU_RUN and not _has_gpu():
        from tests.emu.harness import install

        install()
    import cudf_b200.pylibcudf as plc

    return plc

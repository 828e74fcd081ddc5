import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def level_source(name):
    """(text, KitchenBatch level argument) of a reference level or of a custom one under tests/golden/levels/"""
    import gym_cooking_b200 as gcb
    if name in gcb.levels.LEVEL_NAMES:
        return gcb.levels.level_text(name), name
    path = os.path.join(GOLDEN, "levels", name + ".txt")
    return open(path).read(), path


TRACE_FILES = ("env_traces.npz", "env_traces_custom.npz")

import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of erroring out of
    Context(0)."""
    if has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device (the product path has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_library():
    """a fresh checkout has no libbamse_cuda.so yet (built artefacts are git-ignored): build it once
    (nvcc cross-compiles for sm_100a without a GPU) for the tests that load it."""
    lib = os.path.join(ROOT, "bamse_b200", "libbamse_cuda.so")
    if not os.path.isfile(lib):
        import shutil
        import subprocess
        if not (shutil.which("nvcc") and shutil.which("make")):
            pytest.skip("libbamse_cuda.so is not built and nvcc/make are unavailable")
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "bamse_b200", "csrc"), "-j4"],
                              stdout=subprocess.DEVNULL)
    return lib


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_scores():
    return load_golden("tree_scores.json")


@pytest.fixture(scope="session")
def golden_vectors():
    return load_golden("vectors.json")


@pytest.fixture(scope="session")
def golden_canon():
    return load_golden("canon.json")


@pytest.fixture(scope="session")
def golden_constants():
    return load_golden("constants.json")


@pytest.fixture(scope="session")
def golden_search():
    return load_golden("search.json")


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False

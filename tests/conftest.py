import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


collect_ignore_glob = ["_refsrc/*"]          # the staged reference files are inputs of the tests, not tests


def pytest_configure(config):
    # the oracle must be deterministic: multi-threaded torch CPU reductions / scatters change the
    # summation order from run to run (the reference itself runs tiny B = 1 tensors, never parallelised)
    import torch
    torch.set_num_threads(1)
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

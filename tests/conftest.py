"""Shared pytest plumbing: the `gpu` marker, repo on sys.path, golden-case loader."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: g[k] for k in g.files}
    d["nrow"], d["ncol"] = int(d["dim"][0]), int(d["dim"][1])
    return d


@pytest.fixture(params=golden_names())
def golden(request):
    d = load_golden(request.param)
    d["name"] = request.param
    return d

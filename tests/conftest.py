import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("PARADIS_REFERENCE", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_cases(prefix="adv_"):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    for k in ("field", "u", "v", "grad_out", "out", "grad_field", "grad_u", "grad_v"):
        if k in d:
            d[k] = torch.from_numpy(d[k])
    for k in ("H", "W", "B", "V"):
        if k in d:
            d[k] = int(d[k])
    if "poles" in d:
        d["poles"] = bool(d["poles"])
        d["interpolation"] = str(d["interpolation"])
        d["dt"] = float(d["dt"])
    return d


def relmax(a, b):
    """max-norm error relative to max |b|"""
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-300))


def bad_fraction(a, b, tol):
    scale = float(b.double().abs().max()) + 1e-300
    return float(((a.double() - b.double()).abs() > tol * scale).double().mean())

import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    """-> dict of torch tensors (0-d string arrays become python objects)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = {}
    for k in z.files:
        a = z[k]
        if a.dtype.kind in "US":
            out[k] = str(a)
        else:
            out[k] = torch.from_numpy(np.array(a))
    return out


def split_golden(g):
    """-> (inputs/others, state_dict, grads, after)"""
    sd = {k: v for k, v in g.items() if torch.is_tensor(v) and "." in k and not k.startswith(("grad.", "after.", "att.", "blk."))}
    grads = {k[5:]: v for k, v in g.items() if k.startswith("grad.")}
    after = {k[6:]: v for k, v in g.items() if k.startswith("after.")}
    return sd, grads, after


def relerr(a, b):
    """normwise relative error ||a-b|| / ||b||  (SURVEY App. B-9)."""
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    d = (a - b).norm().item()
    n = b.norm().item()
    return d / n if n > 0 else d


@pytest.fixture(scope="session")
def golden():
    return load_golden

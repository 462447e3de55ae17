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
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden:
    """A fixture written by tests/golden/make_golden.py: the reference's own outputs on a small mesh."""

    def __init__(self, name):
        from scipy.sparse import csc_matrix

        from femder_b200.mesh import Grid
        d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.d = d
        self.N = len(d["nos"])
        self.order = int(d["order"])
        self.grid = Grid(d["nos"], d["elem_vol"], d["elem_surf"], d["domain_index_surf"], order=self.order)
        self.freq = d["freq"]
        self.c0 = float(d["c0"])
        self.rho0 = float(d["rho0"])
        self.group_ids = d["group_ids"]
        self.mu = {int(k): d["mu"][i] for i, k in enumerate(self.group_ids)}
        N = self.N
        self.H = csc_matrix((d["H_data"], d["H_indices"], d["H_indptr"]), shape=(N, N))
        self.Q = csc_matrix((d["Q_data"], d["Q_indices"], d["Q_indptr"]), shape=(N, N))
        self.A = [csc_matrix((d[f"A{i}_data"], d[f"A{i}_indices"], d[f"A{i}_indptr"]), shape=(N, N))
                  for i in range(len(self.group_ids))]
        self.areas = d["areas"]
        self.q = d["q"]
        self.pN = d["pN"]
        self.pR = d["pR"]
        self.src = d["src"]
        self.src_q = d["src_q"]
        self.rcv = d["rcv"]
        self.v = {int(k): d["v"][i] for i, k in enumerate(d["v_ids"])} if "v_ids" in d else None   # velocity BCs
        self.rcv_off = d["rcv_off"] if "rcv_off" in d else None
        self.pR_off = d["pR_off"] if "pR_off" in d else None


@pytest.fixture(params=golden_names())
def golden(request):
    return Golden(request.param)


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA library on a usable GPU; GPU tests fail (not skip) if the native path is unavailable."""
    from femder_b200 import _lib
    lib = _lib.load()
    assert _lib.device_count() > 0, "no CUDA device: -m gpu tests must run on the B200 box"
    return lib

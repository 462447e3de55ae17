"""Pins the oracle against the LIVE reference when /root/reference is present (build container only)."""
import numpy as np
import pytest

from oracle import femder_oracle as O
from oracle.reference_shim import reference_available, run_reference_compute

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")


def test_live_reference_tet4_matches_oracle():
    from femder_b200.mesh import shoebox_tet4
    g = shoebox_tet4(5, 4, 3, jitter=0.1, seed=7)
    freq = np.array([40.0, 95.0])
    mu = {t: np.zeros(2) for t in range(2, 8)}
    mu[5] = np.ones(2) * 0.1 / (343 * 1.21)
    rc = g.nos[g.closest_node([2.0, 3.0, 1.0])]
    obj = run_reference_compute(g, freq, mu, [[1.0, 1.0, 1.2]], [1e-4], rc)
    H, Q = O.assemble_tet4(g.elem_vol, g.nos, 343.0, 1.21)
    assert np.array_equal(H.indices, obj.H.indices) and np.array_equal(H.indptr, obj.H.indptr)
    assert np.abs(H.data - obj.H.data).max() / np.abs(H.data).max() < 2.5e-7
    assert np.abs(Q.data - obj.Q.data).max() / np.abs(Q.data).max() < 2.5e-7
    q = O.source_vector(g.nos, [[1.0, 1.0, 1.2]], [1e-4])
    pN = O.solve_sweep(obj.H.astype(np.complex128), obj.Q.astype(np.complex128), obj.A, mu, np.sort([*mu]), freq, q)
    assert (np.linalg.norm(pN - obj.pN, axis=1) / np.linalg.norm(obj.pN, axis=1)).max() < 1e-9

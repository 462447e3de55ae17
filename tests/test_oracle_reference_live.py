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


def test_room_setup_stand_ins_carry_the_reference_attributes():
    """room_setup.py (the problem-setup holders of the benches and tests on machines without femder) against the reference's own
    ``controlsair.AirProperties / AlgControls`` and ``BoundaryConditions.BC``: every attribute FEM3D reads has the same value."""
    import room_setup as rs
    from oracle.reference_shim import load_reference
    ref = load_reference()
    for kw in (dict(freq_init=20.0, freq_end=200.0, freq_step=1), dict(freq_vec=[31.5, 63.0, 125.0, 250.0])):
        APr, APm = ref.controlsair.AirProperties(c0=340.0, rho0=1.2), rs.AirProperties(c0=340.0, rho0=1.2)
        assert float(APr.c0) == float(APm.c0) and float(APr.rho0) == float(APm.rho0)
        ACr, ACm = ref.controlsair.AlgControls(APr, **kw), rs.AlgControls(APm, **kw)
        assert np.array_equal(np.asarray(ACr.freq, dtype=float), np.asarray(ACm.freq, dtype=float))
        assert np.allclose(ACr.w, ACm.w, rtol=0, atol=0) and np.allclose(ACr.k0, ACm.k0, rtol=1e-15)
        BCr, BCm = ref.BoundaryConditions.BC(ACr, APr), rs.BC(ACm, APm)
        BCr.normalized_admittance(2, 0.02); BCm.normalized_admittance(2, 0.02)
        BCr.rigid(3); BCm.rigid(3)
        BCr.admittance(4, 1e-4 + 2e-5j); BCm.admittance(4, 1e-4 + 2e-5j)
        BCr.velocity(5, 1e-3); BCm.velocity(5, 1e-3)
        BCr.fluid(2, 180.0 - 35j, 2.1 - 0.9j); BCm.fluid(2, 180.0 - 35j, 2.1 - 0.9j)
        assert sorted(BCr.mu) == sorted(BCm.mu) and sorted(BCr.v) == sorted(BCm.v)
        for k in BCr.mu:
            assert np.allclose(np.asarray(BCr.mu[k]).ravel(), np.asarray(BCm.mu[k]).ravel(), rtol=1e-15, atol=0)
        for k in BCr.v:
            assert np.allclose(np.asarray(BCr.v[k]).ravel(), np.asarray(BCm.v[k]).ravel(), rtol=1e-15, atol=0)
        for k in BCr.rhoc:
            assert np.allclose(np.asarray(BCr.rhoc[k]).ravel(), np.asarray(BCm.rhoc[k]).ravel())
            assert np.allclose(np.asarray(BCr.cc[k]).ravel(), np.asarray(BCm.cc[k]).ravel())

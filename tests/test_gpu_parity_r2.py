"""-m gpu: parity at the settings the product ships with, and the behaviours round 1 left unchecked:
default tolerance / batch / preconditioner across a room resonance, multi-RHS against the oracle, BASELINE
config 3 on the reference's own control-room meshes with on- and off-node receivers, the velocity branch without
admittances on the driven faces, loud failure when a sweep does not converge or the surface matrices are missing,
re-assembly after in-place node edits, and FEM3D.eigenfrequency."""
import warnings

import numpy as np
import pytest
from scipy.sparse.linalg import eigsh, splu

import femder_b200 as fd

import room_setup as rs  # noqa: E402
from femder_b200._lib import FemderB200Error
from oracle import femder_oracle as O

pytestmark = pytest.mark.gpu
C0, RHO0 = 343.0, 1.21


def _problem(grid, freq, rigid=range(2, 8), adm=None, src=(1.0, 1.0, 1.2), rcv=(2.0, 3.0, 1.25), **kw):
    AP = rs.AirProperties(c0=C0, rho0=RHO0)
    AC = rs.AlgControls(AP, freq_vec=np.asarray(freq, dtype=np.float64))
    BC = rs.BC(AC, AP)
    for t in rigid:
        BC.rigid(t)
    for t, y in (adm or {}).items():
        BC.normalized_admittance(t, y)
    S = rs.Source(coord=grid.nos[grid.closest_node(src)], q=[1e-4])
    R = rs.Receiver(coord=grid.nos[grid.closest_node(rcv)])
    return fd.FEM3D(grid, S, R, AP, AC, BC, **kw), S, R, AC, BC


def test_config1_defaults_across_resonance(gpu_lib):
    """BASELINE config 1 (21 735-DOF rigid shoebox) through FEM3D.compute() with EVERY default (tol, batch, preconditioner,
    check interval) on 18 frequencies that straddle the first axial mode (42.875 Hz): true residuals of all of them against
    the oracle's FP64 matrices, and relative L2 <= 1e-6 / receiver SPL <= 0.01 dB against the oracle's direct solve at the
    three frequencies next to the resonance and at three others."""
    grid = fd.shoebox_tet4(26, 34, 22)
    freq = np.array([20.0, 31.5, 42.8, 42.9, 43.0, 50.0, 57.2, 63.0, 71.5, 80.0, 85.75, 100.0, 114.0, 125.0, 150.0, 171.5, 187.0, 200.0])
    obj, S, R, AC, BC = _problem(grid, freq)
    with warnings.catch_warnings():
        warnings.simplefilter("error")          # a non-converged frequency would warn
        obj.compute(timeit=False)
    assert obj.stats.precond_used == "modal" and obj.n_unconverged == 0
    assert obj.relres.max() <= obj.tol == 1e-8
    H, Q = O.assemble_tet4(grid.elem_vol, grid.nos, C0, RHO0)
    Hr, Qr = H.real.tocsr(), Q.real.tocsr()
    q = O.source_vector(grid.nos, S.coord, S.q.ravel())
    for i, f in enumerate(freq):
        w = 2 * np.pi * f
        b = -1j * w * q
        r = b - (Hr @ obj.pN[i] - w * w * (Qr @ obj.pN[i]))
        assert np.linalg.norm(r) / np.linalg.norm(b) <= 1.5e-8, (f, np.linalg.norm(r) / np.linalg.norm(b))
    rn = grid.closest_node(R.coord[0])
    pR = obj.evaluate(R)
    for i in (2, 3, 4, 9, 14, 17):
        w = 2 * np.pi * freq[i]
        ref = splu((Hr - w * w * Qr).tocsc().astype(np.complex128)).solve(-1j * w * q)
        rel = np.linalg.norm(obj.pN[i] - ref) / np.linalg.norm(ref)
        dspl = abs(O.p2SPL(pR[i, 0]) - O.p2SPL(ref[rn]))
        assert rel <= 1e-6 and dspl <= 0.01, (freq[i], rel, dspl)


def test_multi_rhs_against_the_oracle(gpu_lib):
    """Source candidates batched as right-hand sides sharing G(w) (FEM_3D.py:1633-1648): every (set, frequency) system against
    the oracle's direct solve of that set - not against a second GPU run."""
    from conftest import Golden
    g = Golden("tet4_cplx_room")
    s = fd.System(0)
    s.set_HQ(g.H, g.Q)
    s.set_surface_matrices(g.A)
    freq = np.linspace(30.0, 200.0, 7)
    w = 2 * np.pi * freq
    mu_by = {int(k): np.resize(g.mu[int(k)], len(freq)) for k in g.group_ids}
    mu = np.array([mu_by[int(k)] for k in g.group_ids])
    sets = [([5], [1e-4]), ([40, 7], [2e-4j, 1e-4]), ([100, 100], [9.0, 3e-4])]   # the last one: overwrite semantics
    pN, _, st = s.sweep(w, mu, None, None, tol=1e-10, batch=8, src_sets=sets)
    assert st.n_unconverged == 0 and pN.shape == (3, len(freq), g.N)
    H, Q = g.H.astype(np.complex128), g.Q.astype(np.complex128)
    for k, (nodes, qv) in enumerate(sets):
        q = np.zeros(g.N, dtype=np.complex128)
        for n, v in zip(nodes, qv):
            q[n] = v                                    # assignment: a later source on the same node overwrites
        ref = O.solve_sweep(H, Q, g.A, mu_by, g.group_ids, freq, q)
        rel = np.linalg.norm(pN[k] - ref, axis=1) / np.linalg.norm(ref, axis=1)
        assert rel.max() < 1e-6, (k, rel)


@pytest.mark.parametrize("name", ["tet4_caixa", "tet4_current_mesh"])
def test_config3_reference_meshes(gpu_lib, name):
    """BASELINE config 3 on the reference repository's unstructured control-room meshes, mixed and frequency-dependent
    admittances, on- and off-node receivers.  The solver is fed the reference's OWN H, Q, A (its complex64 Tet4
    matrices), so pN and the receiver SPLs must agree with the reference's verbatim output."""
    from conftest import Golden
    from femder_b200.fem3d import receiver_stencil
    g = Golden(name)
    s = fd.System(0)
    s.set_HQ(g.H, g.Q)
    s.set_surface_matrices(g.A)
    w = 2 * np.pi * g.freq
    mu = np.array([g.mu[int(k)] for k in g.group_ids])
    src_nodes = [O.closest_node(g.grid.nos, c) for c in g.src]
    for precond in ("block", "modal"):
        if precond == "modal":
            m = s.modal_compute(float(np.hypot(g.freq.max(), 300.0)), n_hint=0)
            assert m > 10
        pN, _, st = s.sweep(w, mu, src_nodes, g.src_q, tol=1e-9, precond=precond)
        assert st.n_unconverged == 0 and st.precond_used == precond
        rel = np.linalg.norm(pN - g.pN, axis=1) / np.linalg.norm(g.pN, axis=1)
        assert rel.max() < 1e-6, (precond, rel)
        on = np.array([pN[:, O.closest_node(g.grid.nos, c)] for c in g.rcv]).T
        assert np.abs(O.p2SPL(on) - O.p2SPL(g.pR)).max() < 0.01
        off = []
        for c in g.rcv_off:
            nodes, wts = receiver_stencil(g.grid.nos, g.grid.elem_vol, c)
            off.append(pN[:, nodes.astype(np.int64)] @ wts)
        off = np.array(off).T
        assert np.abs(O.p2SPL(off) - O.p2SPL(g.pR_off.reshape(off.shape))).max() < 0.01


def test_config3_fem3d_evaluate_off_node(gpu_lib):
    """FEM3D.compute() + evaluate() on caixa.msh with the batched GPU point location for off-node receivers, against the oracle
    (FP64 assembly) end to end."""
    from conftest import Golden
    g = Golden("tet4_caixa")
    AP = rs.AirProperties(c0=g.c0, rho0=g.rho0)
    AC = rs.AlgControls(AP, freq_vec=g.freq)
    BC = rs.BC(AC, AP)
    for k, v in g.mu.items():
        BC.mu[k] = v
    S = rs.Source(coord=g.src, q=g.src_q)
    R = rs.Receiver(coord=np.vstack([g.rcv, g.rcv_off]))
    obj = fd.FEM3D(g.grid, S, R, AP, AC, BC)
    obj.compute(timeit=False)
    pR = obj.evaluate(R)
    r = O.compute(g.grid, g.freq, g.mu, g.src, g.src_q, g.c0, g.rho0)
    rel = np.linalg.norm(obj.pN - r["pN"], axis=1) / np.linalg.norm(r["pN"], axis=1)
    assert rel.max() < 1e-6
    ref = O.evaluate(g.grid.nos, g.grid.elem_vol, r["pN"], R.coord)
    assert np.abs(O.p2SPL(pR) - O.p2SPL(ref)).max() < 0.01


def test_velocity_branch_without_admittance_on_driven_faces(gpu_lib):
    """The reference's usual velocity set-up: BC.velocity() on the driven faces fills only BC.v, admittance elsewhere, the other
    faces carry nothing (FEM_3D.py:1320-1347 walks np.sort([*self.mu]) only).  It must run, and match the oracle."""
    grid = fd.shoebox_tet4(6, 7, 5, jitter=0.2)
    freq = np.array([30.0, 75.0, 140.0])
    AP = rs.AirProperties(c0=C0, rho0=RHO0)
    AC = rs.AlgControls(AP, freq_vec=freq)
    BC = rs.BC(AC, AP)
    BC.velocity(2, 1e-3)
    BC.velocity(4, 5e-4)
    BC.normalized_admittance(7, 0.1)
    S = rs.Source(coord=grid.nos[grid.closest_node([1.0, 1.0, 1.2])], q=[1e-4])
    R = rs.Receiver(coord=grid.nos[grid.closest_node([2.0, 3.0, 1.25])])
    obj = fd.FEM3D(grid, S, R, AP, AC, BC, tol=1e-10)
    obj.compute(timeit=False)
    assert set(BC.mu) == {7} and set(BC.v) == {2, 4}
    r = O.compute(grid, freq, BC.mu, S.coord, S.q.ravel(), C0, RHO0, v_by_group=BC.v)
    rel = np.linalg.norm(obj.pN - r["pN"], axis=1) / np.linalg.norm(r["pN"], axis=1)
    assert rel.max() < 1e-6


def test_unconverged_sweep_is_reported(gpu_lib):
    grid = fd.shoebox_tet4(10, 12, 8)
    obj, *_ = _problem(grid, [150.0, 180.0, 200.0, 220.0], precond="jacobi", max_iter=5, check_every=1)
    with pytest.warns(RuntimeWarning, match="did not reach tol"):
        obj.compute(timeit=False)
    assert obj.n_unconverged == 4 and obj.stats.n_unconverged == 4 and (obj.relres > obj.tol).all()


def test_missing_surface_matrices_fail_loudly(gpu_lib):
    """Non-zero admittances with surface matrices that were never assembled used to solve the rigid problem silently."""
    grid = fd.shoebox_tet4(5, 6, 4)
    s = fd.System(0)
    s.set_volume_mesh(grid.nos, grid.elem_vol)
    s.assemble_volume(RHO0, C0)
    s.set_surface_mesh(grid.elem_surf, grid.domain_index_surf, np.arange(2, 8))     # ... but no assemble_surface()
    mu = np.full((6, 2), 0.1 / (RHO0 * C0), dtype=np.complex128)
    with pytest.raises(FemderB200Error, match="surface matrices are not assembled"):
        s.sweep(2 * np.pi * np.array([50.0, 60.0]), mu, [3], [1e-4], tol=1e-8, batch=2)
    # all-zero admittances are the rigid problem: fine without them
    pN, _, st = s.sweep(2 * np.pi * np.array([50.0, 60.0]), np.zeros_like(mu), [3], [1e-4], tol=1e-8, batch=2)
    assert st.n_unconverged == 0


def test_reassembly_sees_in_place_node_edits(gpu_lib):
    """`obj.H = None` is the reference's re-assembly hook (FEM_3D.py:1269); the reference then assembles from the CURRENT
    self.nos, so an optimizer loop that edits the node array in place must get the new H, Q."""
    grid = fd.shoebox_tet4(6, 7, 5, jitter=0.1)
    obj, *_ = _problem(grid, [40.0, 90.0], precond="block", batch=2)
    obj.compute(timeit=False)
    H0 = obj.H.copy()
    grid.nos *= 1.1                       # in place: same array object
    obj.H = None
    obj.compute(timeit=False)
    H1, _ = O.assemble_tet4(grid.elem_vol, grid.nos, C0, RHO0)
    assert np.abs(obj.H.data - H1.data.real).max() / np.abs(H1.data).max() < 1e-12
    assert np.abs(obj.H.data - H0.data).max() / np.abs(H0.data).max() > 1e-2


def test_eigenfrequency_matches_eigsh(gpu_lib):
    """FEM3D.eigenfrequency (FEM_3D.py:1699-1768) on the GPU against scipy's shift-invert Lanczos on the oracle's matrices, and
    the rigid shoebox's analytic first axial mode c / (2 L) = 42.875 Hz."""
    grid = fd.shoebox_tet4(24, 32, 20)
    obj, *_ = _problem(grid, [50.0])
    F_n = obj.eigenfrequency(12)
    assert F_n.shape == (12,) and obj.Vc.shape == (grid.NumNosC, 12)
    H, Q = O.assemble_tet4(grid.elem_vol, grid.nos, C0, RHO0)
    lam = np.sort(eigsh(H.real.tocsc(), k=12, M=Q.real.tocsc(), sigma=-(2 * np.pi * 30.0) ** 2, which="LM")[0])
    ref = np.sqrt(np.maximum(lam, 0)) / (2 * np.pi)
    assert abs(F_n[0]) < 1e-3 and np.abs(F_n[1:] / ref[1:] - 1).max() < 1e-5
    assert abs(F_n[1] - C0 / (2 * 4.0)) < 0.2       # mesh dispersion only
    G = obj.Vc.T @ (Q.real @ obj.Vc)
    assert np.abs(G - np.eye(12)).max() < 1e-8


def test_compute_candidates_against_the_oracle(gpu_lib):
    """The optimizer's source / receiver candidate loop (FEM_3D.py:1633-1648) as one multi-RHS sweep: every candidate's pR
    against the oracle's compute() + evaluate() of that candidate alone."""
    grid = fd.shoebox_tet4(10, 12, 8, jitter=0.2)
    freq = np.array([35.0, 80.0, 125.0, 160.0, 200.0])
    obj, S0, R0, AC, BC = _problem(grid, freq, rigid=range(2, 7), adm={7: 0.15})
    srcs = [rs.Source(coord=grid.nos[grid.closest_node(c)], q=[qv]) for c, qv in
            (([1.0, 1.0, 1.2], 1e-4), ([2.2, 0.7, 0.6], 2e-4), ([0.5, 3.2, 1.9], 1e-4))]
    rcvs = [rs.Receiver(coord=grid.nos[grid.closest_node([2.0, 3.0, 1.25])]),
            rs.Receiver(coord=[[1.93, 2.71, 1.33], list(grid.nos[grid.closest_node([0.4, 0.4, 0.4])])]),     # off-node + on-node
            rs.Receiver(coord=[1.11, 1.77, 0.83])]
    out = obj.compute_candidates(srcs, rcvs)
    assert len(out) == 3 and [o.shape for o in out] == [(5, 1), (5, 2), (5, 1)] and obj.n_unconverged == 0
    for S, R, pR in zip(srcs, rcvs, out):
        r = O.compute(grid, freq, BC.mu, S.coord, S.q.ravel(), C0, RHO0)
        ref = O.evaluate(grid.nos, grid.elem_vol, r["pN"], R.coord)
        assert np.abs(pR - ref).max() / np.abs(ref).max() < 1e-6
        assert np.abs(O.p2SPL(pR) - O.p2SPL(ref)).max() < 0.01
    assert obj.pR is out[-1] and obj.S is srcs[-1]


def test_config5_geometry_loop(gpu_lib):
    """BASELINE config 5 in the small: perturbed geometries on ONE connectivity, re-assembly + sweep each on the same device
    system (pattern, internal order and tile tables are reused; matrices, block inverses and room modes are rebuilt)."""
    from femder_b200.mesh import scale_grid
    base = fd.shoebox_tet4(12, 16, 10)
    freq = np.array([40.0, 63.0, 100.0, 160.0, 200.0])
    rng = np.random.default_rng(0)
    prev = None
    for fac in rng.uniform(0.95, 1.05, size=(3, 3)):
        grid = scale_grid(base, fac)
        obj, S, R, AC, BC = _problem(grid, freq, src=np.array([1.0, 1.0, 1.2]) * fac, rcv=np.array([2.0, 3.0, 1.25]) * fac)
        if prev is not None:
            obj._sys = prev._sys
        obj.compute(timeit=False, keep_pN=False)
        assert obj.stats.precond_used == "modal" and obj.n_unconverged == 0
        r = O.compute(grid, freq, BC.mu, S.coord, S.q.ravel(), C0, RHO0)
        ref = O.evaluate(grid.nos, grid.elem_vol, r["pN"], R.coord)
        assert np.abs(O.p2SPL(obj.pR) - O.p2SPL(ref)).max() < 0.01
        prev = obj


def test_config2_full_size_tet10(gpu_lib):
    """BASELINE config 2 at its size: ~21k-DOF Tet10 shoebox, normalized admittance 0.02 on one wall, 20-200 Hz at 1 Hz with all
    defaults: every frequency converged on the true residual; the residual re-checked against the oracle's FP64 Tet10 matrices
    and the solution against the oracle's direct solve at one frequency."""
    grid = fd.to_tet10(fd.shoebox_tet4(13, 17, 11, jitter=0.1))
    freq = np.arange(20.0, 201.0, 1.0)
    obj, S, R, AC, BC = _problem(grid, freq, rigid=range(2, 7), adm={7: 0.02})
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        obj.compute(timeit=False)
    assert obj.n_unconverged == 0 and obj.relres.max() <= 1e-8 and obj.pN.shape == (181, grid.NumNosC)
    H, Q = O.assemble_tet10(grid.elem_vol, grid.nos, C0, RHO0)
    A = O.assemble_surface(grid.elem_surf, grid.nos, grid.domain_index_surf, [7], 2)[0]
    q = O.source_vector(grid.nos, S.coord, S.q.ravel())
    Hr, Qr, Ar = H.tocsr(), Q.tocsr(), A.tocsr()
    for i in (0, 90, 180):
        w = 2 * np.pi * freq[i]
        x = obj.pN[i]
        r = -1j * w * q - (Hr @ x - w * w * (Qr @ x) + 1j * w * BC.mu[7][i] * (Ar @ x))
        assert np.linalg.norm(r) / np.linalg.norm(w * q) <= 2e-8
    w = 2 * np.pi * freq[120]
    G = (H - w * w * Q + 1j * w * BC.mu[7][120] * A).tocsc().astype(np.complex128)
    ref = splu(G).solve(-1j * w * q)
    assert np.linalg.norm(obj.pN[120] - ref) / np.linalg.norm(ref) < 1e-6


def test_equivalent_fluid_branch(gpu_lib):
    """The equivalent-fluid branch (FEM_3D.py:1367-1399, BC.fluid: complex c and rho in a volume domain) against the golden the
    unmodified reference produced (tests/golden/make_golden_fluid.py; the reference evaluates the mass shape functions in
    single precision, FEM_3D.py:800) and against the oracle's FP64 restatement."""
    import os
    from conftest import GOLDEN_DIR
    from femder_b200.mesh import Grid
    d = np.load(os.path.join(GOLDEN_DIR, "fluid", "tet4_fluid.npz"))
    g = Grid(d["nos"], d["elem_vol"], d["elem_surf"], d["domain_index_surf"], domain_index_vol=d["domain_index_vol"])
    freq = d["freq"]
    AP = rs.AirProperties(c0=float(d["c0"]), rho0=float(d["rho0"]))
    AC = rs.AlgControls(AP, freq_vec=freq)
    BC = rs.BC(AC, AP)
    for i, k in enumerate(d["group_ids"]):
        BC.mu[int(k)] = d["mu"][i]
    BC.fluid(int(d["fluid_domain"]), complex(d["cc"]), complex(d["rhoc"]))
    S = rs.Source(coord=d["src"], q=d["src_q"])
    R = rs.Receiver(coord=d["rcv"])
    obj = fd.FEM3D(g, S, R, AP, AC, BC, tol=1e-10)
    obj.compute(timeit=False)
    assert obj.n_unconverged == 0 and set(obj.rho) == {1, 2}
    nf = len(freq)
    c = {1: np.ones(nf) * float(d["c0"]), 2: np.ones(nf, dtype=complex) * complex(d["cc"])}
    rho = {1: np.ones(nf) * float(d["rho0"]), 2: np.ones(nf, dtype=complex) * complex(d["rhoc"])}
    r64 = O.compute_fluid(g, freq, BC.mu, c, rho, d["src"], d["src_q"], float32_shape=False)
    rel = np.linalg.norm(obj.pN - r64["pN"], axis=1) / np.linalg.norm(r64["pN"], axis=1)
    assert rel.max() < 1e-6, rel
    relg = np.linalg.norm(obj.pN - d["pN"], axis=1) / np.linalg.norm(d["pN"], axis=1)
    assert relg.max() < 2e-5, relg
    pR = obj.evaluate(R)
    assert np.abs(O.p2SPL(pR) - O.p2SPL(d["pR"].reshape(pR.shape))).max() < 0.01

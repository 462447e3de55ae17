"""-m gpu: K5 fused SpMV, K6 COCG sweep, K7 receivers through the C ABI against the oracle's direct solve."""
import numpy as np
import pytest

from oracle import femder_oracle as O

pytestmark = pytest.mark.gpu


def _rel_rows(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


def _solver_system(H, Q, A_list):
    """Feed the solver the given matrices (e.g. the reference's own arrays) - assembly is not involved."""
    from femder_b200 import System
    s = System(0)
    s.set_HQ(H, Q)
    s.set_surface_matrices(A_list)
    return s


@pytest.mark.parametrize("F", [1, 2, 4, 8, 16, 32, 64])
def test_spmv_matches_scipy(gpu_lib, golden, F):
    g = golden
    s = _solver_system(g.H, g.Q, g.A)
    rng = np.random.default_rng(F)
    x = rng.standard_normal((g.N, F)) + 1j * rng.standard_normal((g.N, F))
    freq = np.linspace(30.0, 400.0, F)
    w = 2 * np.pi * freq
    mu = np.array([np.resize(g.mu[int(k)], F) * (1 + 0.1 * np.arange(F)) for k in g.group_ids])
    y = s.spmv(w, mu, x)
    Hd, Qd = g.H.astype(np.complex128), g.Q.astype(np.complex128)
    for f in range(F):
        G = O.system_matrix(Hd, Qd, g.A, mu[:, f], w[f])
        ref = G @ x[:, f]
        assert np.linalg.norm(y[:, f] - ref) / np.linalg.norm(ref) < 1e-13


@pytest.mark.parametrize("tile", [64, 128])
def test_spmv_blob_tile_rows(gpu_lib, golden, monkeypatch, tile):
    """k_spmv_blob with 64-row tiles (two resident blocks per SM) and 128-row tiles (one): both against scipy, complex
    batch of 8 with the boundary overlay; then a rigid sweep in real arithmetic (batch 16) against the complex one."""
    monkeypatch.setenv("FDB_TILE_ROWS", str(tile))
    g = golden
    s = _solver_system(g.H, g.Q, g.A)
    F = 8
    rng = np.random.default_rng(tile)
    x = rng.standard_normal((g.N, F)) + 1j * rng.standard_normal((g.N, F))
    w = 2 * np.pi * np.linspace(30.0, 400.0, F)
    mu = np.array([np.resize(g.mu[int(k)], F) * (1 + 0.1 * np.arange(F)) for k in g.group_ids])
    y = s.spmv(w, mu, x)
    Hd, Qd = g.H.astype(np.complex128), g.Q.astype(np.complex128)
    for f in range(F):
        ref = O.system_matrix(Hd, Qd, g.A, mu[:, f], w[f]) @ x[:, f]
        assert np.linalg.norm(y[:, f] - ref) / np.linalg.norm(ref) < 1e-13
    w16 = 2 * np.pi * np.linspace(30.0, 200.0, 16)
    mu0 = np.zeros((len(g.group_ids), 16), dtype=complex)
    nodes, q = np.array([5], dtype=np.int32), np.array([1e-4 + 0j])
    pr, _, st_r = s.sweep(w16, mu0, nodes, q, tol=1e-11, batch=16)
    assert st_r.real_arithmetic and st_r.n_unconverged == 0
    pc, _, st_c = s.sweep(w16, mu0, nodes, q, tol=1e-11, batch=8, force_complex=True)
    assert not st_c.real_arithmetic
    assert _rel_rows(pr, pc).max() < 1e-7


def test_sweep_on_reference_matrices(gpu_lib, golden):
    """Solver parity on identical inputs: the reference's own H, Q, A arrays in, its pN out (north star: relative
    L2 error <= 1e-6, receiver SPL within 0.01 dB)."""
    g = golden
    s = _solver_system(g.H, g.Q, g.A)
    w = 2 * np.pi * g.freq
    mu = np.array([g.mu[int(k)] for k in g.group_ids])
    src = [O.closest_node(g.grid.nos, c) for c in g.src]
    if g.v:   # velocity-BC branch: b = -1j*w*sum_i v_i (V_i @ 1), point sources ignored (FEM_3D.py:1341)
        V = O.assemble_surface(g.grid.elem_surf, g.grid.nos, g.grid.domain_index_surf, np.sort([*g.v]), g.order)
        vecs = [np.asarray(Vm.sum(axis=1)).ravel() for Vm in V]
        coef = np.array([-1j * w * g.v[k] for k in np.sort([*g.v])])
        pN, _, st = s.sweep(w, mu, [], [], tol=1e-10, batch=4, max_iter=20000, rhs_vecs=vecs, rhs_coef=coef)
    else:
        pN, _, st = s.sweep(w, mu, src, g.src_q, tol=1e-10, batch=4, max_iter=20000)
    assert st.n_unconverged == 0, st
    assert _rel_rows(pN, g.pN).max() < 1e-6
    pR = O.evaluate(g.grid.nos, g.grid.elem_vol, pN, g.rcv)
    assert np.abs(O.p2SPL(pR) - O.p2SPL(g.pR)).max() < 0.01


@pytest.mark.parametrize("batch", [1, 2, 8, 16, 64])
def test_sweep_batches_agree(gpu_lib, batch):
    from conftest import Golden
    g = Golden("tet4_shoebox_jitter")
    s = _solver_system(g.H, g.Q, g.A)
    freq = np.arange(20.0, 31.0)  # 11 frequencies: ragged last batch for every width
    w = 2 * np.pi * freq
    mu = np.array([np.resize(g.mu[int(k)], len(freq)) for k in g.group_ids])
    src = [O.closest_node(g.grid.nos, c) for c in g.src]
    pN, _, st = s.sweep(w, mu, src, g.src_q, tol=1e-11, batch=batch)
    assert st.n_unconverged == 0
    muo = {int(k): mu[i] for i, k in enumerate(g.group_ids)}
    ref = O.solve_sweep(g.H.astype(np.complex128), g.Q.astype(np.complex128), g.A, muo, g.group_ids, freq,
                        O.source_vector(g.grid.nos, g.src, g.src_q))
    assert _rel_rows(pN, ref).max() < 1e-7


def test_sweep_empty_and_single(gpu_lib):
    from conftest import Golden
    g = Golden("tet4_cplx_room")
    s = _solver_system(g.H, g.Q, g.A)
    mu = np.array([g.mu[int(k)] for k in g.group_ids])
    pN, pR, st = s.sweep(np.array([]), mu[:, :0], [3], [1e-4])
    assert pN.shape == (0, g.N) and len(st.iters) == 0
    pN, _, st = s.sweep(2 * np.pi * g.freq[:1], mu[:, :1], [3], [1e-4], tol=1e-10)
    assert pN.shape == (1, g.N) and st.n_unconverged == 0
    # no source: zero right-hand side -> zero solution, zero iterations
    pN, _, st = s.sweep(2 * np.pi * g.freq[:2], mu[:, :2], [], [], tol=1e-10)
    assert not pN.any() and st.iters.max() == 0


def test_sources_overwrite_and_linearity(gpu_lib):
    from conftest import Golden
    g = Golden("tet4_cplx_room")
    s = _solver_system(g.H, g.Q, g.A)
    w = 2 * np.pi * g.freq[:2]
    mu = np.array([g.mu[int(k)][:2] for k in g.group_ids])
    a, _, _ = s.sweep(w, mu, [5], [1e-4], tol=1e-12)
    b, _, _ = s.sweep(w, mu, [40], [2e-4j], tol=1e-12)
    ab, _, _ = s.sweep(w, mu, [5, 40], [1e-4, 2e-4j], tol=1e-12)
    assert _rel_rows(ab, a + b).max() < 1e-8                       # linearity in q
    dup, _, _ = s.sweep(w, mu, [5, 5], [7.0, 1e-4], tol=1e-12)      # q[node] = S.q: the later source overwrites
    assert _rel_rows(dup, a).max() < 1e-8
    # reciprocity p(S->R) = p(R->S) for the symmetric operator
    assert abs(a[0, 40] - s.sweep(w, mu, [40], [1e-4], tol=1e-12)[0][0, 5]) / abs(a[0, 40]) < 1e-7


def test_receivers_only_mode_matches_pN(gpu_lib):
    from conftest import Golden
    from femder_b200 import receiver_stencil
    g = Golden("tet4_shoebox_jitter")
    s = _solver_system(g.H, g.Q, g.A)
    w = 2 * np.pi * g.freq
    mu = np.array([g.mu[int(k)] for k in g.group_ids])
    src = [O.closest_node(g.grid.nos, c) for c in g.src]
    coords = np.concatenate([g.rcv, g.rcv_off])
    st = [receiver_stencil(g.grid.nos, g.grid.elem_vol, c) for c in coords]
    nodes = np.array([t[0] for t in st]); wts = np.array([t[1] for t in st])
    pN, pR, stats = s.sweep(w, mu, src, g.src_q, nodes, wts, tol=1e-11, batch=2)
    ref = O.evaluate(g.grid.nos, g.grid.elem_vol, pN, coords)
    assert np.abs(pR - ref).max() / np.abs(ref).max() < 1e-13
    assert np.array_equal(pR[:, 0], pN[:, nodes[0, 0]])            # on-node receiver: exact copy
    _, pR2, _ = s.sweep(w, mu, src, g.src_q, nodes, wts, want_pN=False, tol=1e-11, batch=2)
    assert np.array_equal(pR, pR2)
    # against the reference's own evaluate() output, on its own pN
    gold = np.concatenate([g.pR, g.pR_off], axis=1)
    assert np.abs(O.p2SPL(pR) - O.p2SPL(gold)).max() < 0.01


def test_warm_start_and_determinism(gpu_lib):
    from conftest import Golden
    g = Golden("tet4_shoebox_jitter")
    s = _solver_system(g.H, g.Q, g.A)
    freq = np.arange(50.0, 66.0)
    w = 2 * np.pi * freq
    mu = np.array([np.resize(g.mu[int(k)], len(freq)) for k in g.group_ids])
    src = [O.closest_node(g.grid.nos, c) for c in g.src]
    cold, _, st0 = s.sweep(w, mu, src, g.src_q, tol=1e-11, batch=4, warm_start=False)
    cold2, _, _ = s.sweep(w, mu, src, g.src_q, tol=1e-11, batch=4, warm_start=False)
    assert np.array_equal(cold, cold2)                              # deterministic reductions
    warm, _, st1 = s.sweep(w, mu, src, g.src_q, tol=1e-11, batch=4, warm_start=True)
    assert st1.n_unconverged == 0 and _rel_rows(warm, cold).max() < 1e-8


def test_unconverged_is_reported(gpu_lib):
    from conftest import Golden
    g = Golden("tet4_shoebox_jitter")
    s = _solver_system(g.H, g.Q, g.A)
    mu = np.array([g.mu[int(k)] for k in g.group_ids])
    _, _, st = s.sweep(2 * np.pi * g.freq, mu, [3], [1e-4], tol=1e-12, max_iter=4, check_every=2)
    assert st.n_unconverged == len(g.freq) and st.iters.max() <= 6 and st.relres.min() > 1e-12


def test_config1_size_residual_and_oracle(gpu_lib):
    """Config 1 mesh (21 735 DOF, rigid): GPU assembly + sweep vs the oracle's direct solve on a frequency sample,
    and the size-independent property ||b - G x|| / ||b|| <= tol checked with scipy on every frequency solved."""
    from femder_b200 import System
    from femder_b200.mesh import shoebox_tet4
    grid = shoebox_tet4(26, 34, 22, jitter=0.2)
    c0, rho0 = 343.0, 1.21
    s = System(0)
    s.set_volume_mesh(grid.nos, grid.elem_vol)
    s.set_surface_mesh(grid.elem_surf, grid.domain_index_surf, np.arange(2, 8))
    s.assemble_volume(rho0, c0)
    s.assemble_surface()
    H, Q = s.get_HQ()
    A = s.get_A()
    freq = np.array([20.0, 63.0, 125.0, 200.0])
    w = 2 * np.pi * freq
    mu = np.zeros((6, len(freq)), dtype=complex)
    mu[5] = 0.02 / (c0 * rho0)
    src = grid.closest_node([1.0, 1.0, 1.2])
    pN, _, st = s.sweep(w, mu, [src], [1e-4], tol=1e-9, batch=4, max_iter=50000)
    assert st.n_unconverged == 0, st
    q = np.zeros(grid.NumNosC, dtype=complex); q[src] = 1e-4
    for n in range(len(freq)):
        G = O.system_matrix(H, Q, A, mu[:, n], w[n])
        b = -1j * w[n] * q
        assert np.linalg.norm(b - G @ pN[n]) / np.linalg.norm(b) < 1e-8
    ref = O.solve_sweep(H, Q, A, {k + 2: mu[k] for k in range(6)}, np.arange(2, 8), freq[[0, 3]], q)
    assert _rel_rows(pN[[0, 3]], ref).max() < 1e-6


def test_real_arithmetic_matches_complex(gpu_lib):
    """Rigid walls + real sources: the sweep runs in real arithmetic (p = 1j*u); same answer as the complex path."""
    from conftest import Golden
    g = Golden("tet4_shoebox_rigid")
    s = _solver_system(g.H, g.Q, g.A)
    freq = np.arange(30.0, 47.0)       # 17 frequencies: refill of the 16 slots, then tail narrowing
    w = 2 * np.pi * freq
    mu = np.zeros((len(g.group_ids), len(freq)), dtype=complex)
    src = [O.closest_node(g.grid.nos, c) for c in g.src]
    a, _, sa = s.sweep(w, mu, src, g.src_q, tol=1e-11, batch=16)
    b, _, sb = s.sweep(w, mu, src, g.src_q, tol=1e-11, batch=16, force_complex=True)
    assert sa.real_arithmetic and not sb.real_arithmetic and sa.n_unconverged == 0 and sb.n_unconverged == 0
    assert not a.real.any()                                # purely imaginary pressures, exactly
    assert _rel_rows(a, b).max() < 1e-8
    ref = O.solve_sweep(g.H.astype(np.complex128), g.Q.astype(np.complex128), [], {}, [], freq,
                        O.source_vector(g.grid.nos, g.src, g.src_q))
    assert _rel_rows(a, ref).max() < 1e-6
    # a complex source amplitude or a non-zero admittance switches the real path off
    _, _, sc = s.sweep(w[:2], mu[:, :2], src, [1e-4 + 1e-5j], tol=1e-9, batch=16)
    assert not sc.real_arithmetic
    mu2 = mu.copy(); mu2[0] = 1e-5
    _, _, sd = s.sweep(w[:2], mu2[:, :2], src, g.src_q, tol=1e-9, batch=16)
    assert not sd.real_arithmetic


@pytest.mark.parametrize("batch", [1, 2, 4, 8, 32])
def test_real_arithmetic_narrow_batches(gpu_lib, batch, monkeypatch):
    """Narrow real batches run the lane-per-row kernel k_spmv_rows (F = 2, 4, 8): same answer as the complex path and
    as the older kernels (FDB_NO_ROWS_KERNEL), residual checked with scipy on the same matrices."""
    from conftest import Golden
    g = Golden("tet4_shoebox_rigid")
    s = _solver_system(g.H, g.Q, g.A)
    freq = np.linspace(25.0, 210.0, 11)
    w = 2 * np.pi * freq
    mu = np.zeros((len(g.group_ids), len(freq)), dtype=complex)
    src = [O.closest_node(g.grid.nos, c) for c in g.src]
    a, _, sa = s.sweep(w, mu, src, g.src_q, tol=1e-11, batch=batch, check_every=16)
    b, _, sb = s.sweep(w, mu, src, g.src_q, tol=1e-11, batch=8, force_complex=True)
    assert sa.real_arithmetic and sa.n_unconverged == 0 and sb.n_unconverged == 0
    assert _rel_rows(a, b).max() < 1e-8
    q = O.source_vector(g.grid.nos, g.src, g.src_q)
    Hd, Qd = g.H.astype(np.complex128), g.Q.astype(np.complex128)
    for n in (0, 5, 10):
        G = O.system_matrix(Hd, Qd, [], np.zeros(0), w[n])
        rhs = -1j * w[n] * q
        assert np.linalg.norm(rhs - G @ a[n]) / np.linalg.norm(rhs) < 1e-9
    monkeypatch.setenv("FDB_NO_ROWS_KERNEL", "1")
    s2 = _solver_system(g.H, g.Q, g.A)
    c, _, sc = s2.sweep(w, mu, src, g.src_q, tol=1e-11, batch=batch, check_every=16)
    assert sc.real_arithmetic and _rel_rows(a, c).max() < 1e-8


@pytest.mark.parametrize("batch", [16, 4, 1])
def test_block_jacobi_preconditioner(gpu_lib, batch):
    """Block-Jacobi (FP32 inverses of the 128-row diagonal blocks of H, real sweeps): same solutions as point Jacobi, true
    residual checked with scipy, clearly fewer iterations, bit-equal repeats; the narrowing tail keeps the preconditioner."""
    from femder_b200 import System
    from femder_b200.mesh import shoebox_tet4
    grid = shoebox_tet4(26, 34, 22, jitter=0.2)
    s = System(0)
    s.set_volume_mesh(grid.nos, grid.elem_vol)
    s.assemble_volume(1.21, 343.0)
    H, Q = s.get_HQ()
    freq = np.linspace(30.0, 260.0, 21)
    w = 2 * np.pi * freq
    mu = np.zeros((0, len(freq)), dtype=complex)
    src = grid.closest_node([1.0, 1.0, 1.2])
    a, _, sa = s.sweep(w, mu, [src], [1e-4], tol=1e-10, batch=batch, check_every=16, precond="block", max_iter=100000)
    b, _, sb = s.sweep(w, mu, [src], [1e-4], tol=1e-10, batch=batch, check_every=16, precond="jacobi", max_iter=100000)
    assert sa.real_arithmetic and sb.real_arithmetic and sa.n_unconverged == 0 and sb.n_unconverged == 0
    assert _rel_rows(a, b).max() < 1e-7
    assert sa.iters.sum() < 0.75 * sb.iters.sum(), (sa.iters.sum(), sb.iters.sum())
    q = np.zeros(grid.NumNosC, dtype=complex); q[src] = 1e-4
    for n in (0, 10, 20):
        G = O.system_matrix(H, Q, [], [], w[n])
        rhs = -1j * w[n] * q
        assert np.linalg.norm(rhs - G @ a[n]) / np.linalg.norm(rhs) < 1e-9
    a2, _, sa2 = s.sweep(w, mu, [src], [1e-4], tol=1e-10, batch=batch, check_every=16, precond="block", max_iter=100000)
    assert np.array_equal(a, a2) and np.array_equal(sa.iters, sa2.iters)


@pytest.mark.parametrize("batch", [8, 2, 1])
def test_block_jacobi_complex_sweep(gpu_lib, batch):
    """Block-Jacobi on a COMPLEX sweep (admittance wall, complex source amplitude): a batch of F complex slots is 2 F real
    columns of the same tile product.  Same solutions as point Jacobi, residual checked with scipy, fewer iterations,
    bit-equal repeats."""
    from femder_b200 import System
    from femder_b200.mesh import shoebox_tet4
    grid = shoebox_tet4(26, 34, 22, jitter=0.2)
    c0, rho0 = 343.0, 1.21
    s = System(0)
    s.set_volume_mesh(grid.nos, grid.elem_vol)
    s.set_surface_mesh(grid.elem_surf, grid.domain_index_surf, np.arange(2, 8))
    s.assemble_volume(rho0, c0)
    s.assemble_surface()
    H, Q = s.get_HQ()
    A = s.get_A()
    freq = np.linspace(30.0, 150.0, 13)   # k h <= 0.32 on this mesh: the regime the frequency-independent blocks are built for
    w = 2 * np.pi * freq
    mu = np.zeros((6, len(freq)), dtype=complex)
    mu[5] = (0.02 + 0.005j) / (c0 * rho0)
    mu[2] = 0.1 / (c0 * rho0)
    src = grid.closest_node([1.0, 1.0, 1.2])
    qs = [1e-4 + 2e-5j]
    a, _, sa = s.sweep(w, mu, [src], qs, tol=1e-10, batch=batch, check_every=16, precond="block", max_iter=100000)
    b, _, sb = s.sweep(w, mu, [src], qs, tol=1e-10, batch=batch, check_every=16, precond="jacobi", max_iter=100000)
    assert not sa.real_arithmetic and sa.n_unconverged == 0 and sb.n_unconverged == 0
    assert _rel_rows(a, b).max() < 1e-7
    assert sa.iters.sum() < 0.8 * sb.iters.sum(), (sa.iters.sum(), sb.iters.sum())
    q = np.zeros(grid.NumNosC, dtype=complex); q[src] = qs[0]
    for n in (0, 6, 12):
        G = O.system_matrix(H, Q, A, mu[:, n], w[n])
        rhs = -1j * w[n] * q
        assert np.linalg.norm(rhs - G @ a[n]) / np.linalg.norm(rhs) < 1e-9
    a2, _, sa2 = s.sweep(w, mu, [src], qs, tol=1e-10, batch=batch, check_every=16, precond="block", max_iter=100000)
    assert np.array_equal(a, a2) and np.array_equal(sa.iters, sa2.iters)


@pytest.mark.parametrize("block_jacobi", [True, False])
def test_iteration_enqueue_runs_every_kernel(gpu_lib, block_jacobi):
    """fdb_iteration_enqueue (bench.py's per-kernel rooflines): prepare, then K5 / K6a / K6b on resident vectors; the
    algorithmic bytes follow DESIGN.md section 4."""
    from femder_b200 import System
    from femder_b200.mesh import shoebox_tet4
    grid = shoebox_tet4(12, 14, 10)
    s = System(0)
    s.set_volume_mesh(grid.nos, grid.elem_vol)
    s.assemble_volume(1.21, 343.0)
    N, F = grid.NumNosC, 16
    H, _ = s.get_HQ()
    n_tiles = (N + 127) // 128
    s.iteration_enqueue(F, 1, -1, real=True, block_jacobi=block_jacobi)
    nb = [s.iteration_enqueue(F, 3, which, real=True, block_jacobi=block_jacobi) for which in (0, 1, 2)]
    s.synchronize()
    vb = 8 * N * F
    assert nb[0] == 4 * (N + 1) + 20 * H.nnz + 2 * vb
    if block_jacobi:
        assert nb[1] == 3 * vb + 4 * N * F + n_tiles * 8256 * 4 and nb[2] == 4 * vb + 4 * N * F
    else:
        assert nb[1] == 3 * vb + 16 * N and nb[2] == 5 * vb + 16 * N
    with pytest.raises(Exception):
        s.iteration_enqueue(F, 3, 5)


def test_results_land_at_their_frequency_index(gpu_lib):
    """Slots are loaded in descending frequency order and refilled as they converge: outputs stay in input order."""
    from conftest import Golden
    g = Golden("tet4_cplx_room")
    s = _solver_system(g.H, g.Q, g.A)
    rng = np.random.default_rng(3)
    freq = rng.permutation(np.linspace(25.0, 240.0, 23))
    w = 2 * np.pi * freq
    mu = np.array([np.resize(g.mu[int(k)], len(freq)) * rng.uniform(0.5, 2.0, len(freq)) for k in g.group_ids])
    src = [O.closest_node(g.grid.nos, c) for c in g.src]
    pN, _, st = s.sweep(w, mu, src, g.src_q, tol=1e-11, batch=4, check_every=8)
    assert st.n_unconverged == 0
    muo = {int(k): mu[i] for i, k in enumerate(g.group_ids)}
    ref = O.solve_sweep(g.H.astype(np.complex128), g.Q.astype(np.complex128), g.A, muo, g.group_ids, freq,
                        O.source_vector(g.grid.nos, g.src, g.src_q))
    assert _rel_rows(pN, ref).max() < 1e-7


def test_config2_tet10_admittance_wall(gpu_lib):
    """Config 2 in the small: Tet10 shoebox (5 733 DOF) with a normalized-admittance wall, GPU assembly + sweep vs oracle."""
    import femder_b200 as fd
    import room_setup as rs  # noqa: E402
    grid = fd.to_tet10(fd.shoebox_tet4(8, 10, 6, jitter=0.15))
    AP = rs.AirProperties()
    AC = rs.AlgControls(AP, freq_vec=[20.0, 57.0, 101.0, 160.0, 200.0])
    BC = rs.BC(AC, AP)
    for t in range(2, 7):
        BC.rigid(t)
    BC.normalized_admittance(7, 0.02)
    S = rs.Source(coord=grid.nos[grid.closest_node([1.0, 1.0, 1.2])], q=[1e-4])
    R = rs.Receiver(coord=grid.nos[grid.closest_node([2.0, 3.0, 1.25])])
    obj = fd.FEM3D(grid, S, R, AP, AC, BC, tol=1e-10)
    obj.compute(timeit=False)
    obj.evaluate(R)
    ref = O.compute(grid, AC.freq, BC.mu, S.coord, S.q.ravel(), 343.0, 1.21)
    assert np.array_equal(obj.H.indices, ref["H"].indices)
    assert np.abs(obj.H.data - ref["H"].data).max() / np.abs(ref["H"].data).max() < 1e-12
    assert np.abs(obj.A[-1].data - ref["A"][-1].data).max() / np.abs(ref["A"][-1].data).max() < 1e-12
    assert _rel_rows(obj.pN, ref["pN"]).max() < 1e-6
    assert np.abs(obj.spl - O.p2SPL(O.evaluate(grid.nos, grid.elem_vol, ref["pN"], R.coord))).max() < 0.01


def test_multi_rhs_source_sets(gpu_lib):
    """Several source candidates per frequency in the same batches (FEM_3D.py:1633-1648): each
    (set, frequency) system equals the single-set solve."""
    from conftest import Golden
    g = Golden("tet4_cplx_room")
    s = _solver_system(g.H, g.Q, g.A)
    freq = np.linspace(30.0, 200.0, 7)
    w = 2 * np.pi * freq
    mu = np.array([np.resize(g.mu[int(k)], len(freq)) for k in g.group_ids])
    sets = [([5], [1e-4]), ([40, 7], [2e-4j, 1e-4]), ([100, 100], [9.0, 3e-4])]   # the last one: overwrite semantics
    rn = np.array([[3, 3, 3, 3], [50, 60, 70, 80]]); rw = np.array([[1.0, 0, 0, 0], [0.1, 0.2, 0.3, 0.4]])
    pN, pR, st = s.sweep(w, mu, None, None, rn, rw, tol=1e-11, batch=8, src_sets=sets)
    assert pN.shape == (3, len(freq), g.N) and pR.shape == (3, len(freq), 2) and st.iters.shape == (3, len(freq))
    assert st.n_unconverged == 0
    for k, (nodes, q) in enumerate(sets):
        one, oneR, _ = s.sweep(w, mu, nodes, q, rn, rw, tol=1e-11, batch=8)
        assert _rel_rows(pN[k], one).max() < 1e-8
        assert np.abs(pR[k] - oneR).max() / np.abs(oneR).max() < 1e-8


def test_config4_spmv_full_size(gpu_lib):
    """The fused multi-frequency SpMV at BASELINE config-4 size (1M DOF) with an admittance wall: every slot against
    scipy's G_f @ x built from the GPU-assembled H, Q, A, and symmetry x^T G y = y^T G x."""
    import femder_b200 as fd
    grid = fd.shoebox_tet4(99, 133, 74, jitter=0.2)
    s = fd.System(0)
    s.set_volume_mesh(grid.nos, grid.elem_vol)
    s.set_surface_mesh(grid.elem_surf, grid.domain_index_surf, np.arange(2, 8))
    s.assemble_volume(1.21, 343.0)
    s.assemble_surface()
    H, Q = s.get_HQ()
    A = s.get_A()
    N, F = grid.NumNosC, 8
    rng = np.random.default_rng(0)
    x = rng.standard_normal((N, F)) + 1j * rng.standard_normal((N, F))
    w = 2 * np.pi * np.linspace(40.0, 480.0, F)
    mu = np.zeros((6, F), dtype=complex)
    mu[5] = (0.05 + 0.01j) / (343 * 1.21)
    mu[1] = 0.2 / (343 * 1.21)
    y = s.spmv(w, mu, x)
    Hr, Qr = H.tocsr(), Q.tocsr()
    Ar = [a.tocsr() for a in A]
    for f in (0, 3, 7):
        ref = Hr @ x[:, f] - w[f] ** 2 * (Qr @ x[:, f]) + 1j * w[f] * sum(mu[g, f] * (Ar[g] @ x[:, f]) for g in range(6))
        assert np.linalg.norm(y[:, f] - ref) / np.linalg.norm(ref) < 1e-13
    z = rng.standard_normal((N, F)) + 1j * rng.standard_normal((N, F))
    yz = s.spmv(w, mu, z)
    lhs = np.einsum("nf,nf->f", z, y)      # z^T G x (unconjugated)
    rhs = np.einsum("nf,nf->f", x, yz)     # x^T G z
    assert np.abs(lhs - rhs).max() / np.abs(lhs).max() < 1e-11


def test_shared_queue_single_process(gpu_lib):
    """fdb_sweep_params.shared_next: systems are claimed from a caller-owned counter (the common queue of the ranks of a node).
    One process alone claims them all: same solutions as without the queue, every system marked solved, the counter past the end;
    a counter that starts beyond the list leaves nothing to solve."""
    import ctypes
    import femder_b200 as fd
    grid = fd.shoebox_tet4(6, 7, 5, jitter=0.2)
    s = fd.System(0)
    s.set_volume_mesh(grid.nos, grid.elem_vol)
    s.assemble_volume(1.21, 343.0)
    freq = np.linspace(40.0, 160.0, 11)
    w = 2 * np.pi * freq
    mu = np.zeros((0, len(w)))
    pN0, _, st0 = s.sweep(w, mu, [5], [1e-4], tol=1e-10, batch=4)
    counter = np.zeros(2, dtype=np.int64)
    addr = counter.ctypes.data
    pN1, _, st1 = s.sweep(w, mu, [5], [1e-4], tol=1e-10, batch=4, shared_next=addr)
    assert st0.solved.all() and st1.solved.all() and counter[0] >= len(w) and counter[1] == 0
    assert np.abs(pN1 - pN0).max() <= 1e-9 * np.abs(pN0).max()
    counter[0] = 7                                   # somebody else took the first seven systems of the dispatch order
    pN2, _, st2 = s.sweep(w, mu, [5], [1e-4], tol=1e-10, batch=4, shared_next=addr)
    assert st2.solved.sum() == len(w) - 7 and st2.n_unconverged == 0
    assert np.abs(pN2[st2.solved] - pN0[st2.solved]).max() <= 1e-9 * np.abs(pN0).max()
    counter[0] = 100
    _, _, st3 = s.sweep(w, mu, [5], [1e-4], tol=1e-10, batch=4, shared_next=addr)
    assert not st3.solved.any()

"""Room modes on the GPU (the reference's FEM3D.eigenfrequency, FEM_3D.py:1699-1768) and the modal coarse correction of the
sweep's preconditioner (tcgen05 kernels of csrc/modal.cu), against scipy / the oracle.  All calls go through the C ABI."""
import numpy as np
import pytest
from scipy.sparse.linalg import eigsh

import femder_b200 as fd
from oracle import femder_oracle as O

pytestmark = pytest.mark.gpu
C0, RHO0 = 343.0, 1.21


def bf16(a):
    a32 = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((a32 + 0x7FFF + ((a32 >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32).astype(np.float64)


@pytest.fixture(scope="module")
def room(gpu_lib):
    g = fd.shoebox_tet4(18, 24, 15, jitter=0.2)
    H, Q = O.assemble_tet4(g.elem_vol, g.nos, C0, RHO0)
    H, Q = H.real.astype(np.float64).tocsc(), Q.real.astype(np.float64).tocsc()
    k = 60
    lam, V = eigsh(H, k=k, M=Q, sigma=-(2 * np.pi * 30.0) ** 2, which="LM")
    o = np.argsort(lam)
    s = fd.System(0)
    s.set_volume_mesh(g.nos, g.elem_vol)
    s.assemble_volume(RHO0, C0)
    return dict(g=g, H=H, Q=Q, lam=lam[o], V=V[:, o].T.copy(), sys=s)


def test_modal_apply_matches_numpy(room):
    """z = V diag(1/(lam - w^2)) V^T r through the projection / coefficient / expansion kernels vs numpy with the same
    bf16-rounded V: what is left is the (hi, lo) bf16 split of r and e and FP32 accumulation."""
    s, lam, V = room["sys"], room["lam"], room["V"]
    s.modal_set(lam, V)
    assert s.modal_count() == len(lam)
    s.modal_compression(False)          # the full-resolution tiles: the products themselves are checked here
    N = V.shape[1]
    rng = np.random.default_rng(0)
    r = rng.standard_normal((N, 16))
    f = np.linspace(55.0, 190.0, 16)
    om = 2 * np.pi * f
    for n_use in (16, 48, 60):
        m = min(((n_use + 15) // 16) * 16, len(lam))
        Vb = bf16(V[:m])
        z = s.modal_apply(om, n_use, r)
        ref = np.empty_like(r)
        for c in range(16):
            ref[:, c] = Vb.T @ ((Vb @ r[:, c]) / (lam[:m] - om[c] ** 2))
        err = np.linalg.norm(z - ref, axis=0) / np.linalg.norm(ref, axis=0)
        assert err.max() < 2e-4, (n_use, err)
        # and the bf16 rounding of V itself is a small perturbation of the exact operator
        ex = np.empty_like(r)
        for c in range(16):
            ex[:, c] = V[:m].T @ ((V[:m] @ r[:, c]) / (lam[:m] - om[c] ** 2))
        assert (np.linalg.norm(z - ex, axis=0) / np.linalg.norm(ex, axis=0)).max() < 2e-2


def test_tile_compressed_modes_match_the_exact_operator(room):
    """The default form of the correction: per 128-node tile V_t ~ Phi_t C_t with 16 basis functions.  Applied to the smooth part of a
    residual (what the correction exists for) it reproduces V diag(1/(lam - w^2)) V^T r to the accuracy of the bf16 storage; on white
    noise, where the tile bases see only their own span of r, it stays a small perturbation of the exact operator."""
    s, lam, V, Q = room["sys"], room["lam"], room["V"], room["Q"]
    s.modal_set(lam, V)
    s.modal_compression("force")        # whatever the fit on this small mesh (the default falls back to the full tiles above 1.5 %)
    N = V.shape[1]
    rng = np.random.default_rng(1)
    f = np.linspace(55.0, 190.0, 16)
    om = 2 * np.pi * f
    m = 48
    r_smooth = Q @ (V[:m].T @ rng.standard_normal((m, 16)))       # residuals in the span of Q V
    r_noise = rng.standard_normal((N, 16))
    for r, tol in ((r_smooth, 2e-2), (r_noise, 0.25)):
        z = s.modal_apply(om, m, r)
        ex = np.empty_like(r)
        for c in range(16):
            ex[:, c] = V[:m].T @ ((V[:m] @ r[:, c]) / (lam[:m] - om[c] ** 2))
        err = np.linalg.norm(z - ex, axis=0) / np.linalg.norm(ex, axis=0)
        assert err.max() < tol, err
    s.modal_compression(False)
    z_fine = s.modal_apply(om, m, r_smooth)
    s.modal_compression("force")
    z_comp = s.modal_apply(om, m, r_smooth)
    assert (np.linalg.norm(z_comp - z_fine, axis=0) / np.linalg.norm(z_fine, axis=0)).max() < 2e-2


def test_modal_roundtrip(room):
    s, lam, V = room["sys"], room["lam"], room["V"]
    s.modal_set(lam, V)
    l2, V2 = s.modal_get()
    assert np.array_equal(l2, lam) and np.array_equal(V2, V)


def test_modal_compute_matches_eigsh(room):
    """Eigenvalues of the GPU subspace iteration vs scipy's shift-invert Lanczos on the oracle's matrices, Q-orthonormality and
    residuals of the returned vectors on the consistent pencil."""
    s, lam, H, Q = room["sys"], room["lam"], room["H"], room["Q"]
    f_cut = 200.0
    want = int(np.searchsorted(np.sqrt(np.maximum(lam, 0)) / (2 * np.pi), f_cut))
    assert want < len(lam) - 2
    m = s.modal_compute(f_cut, n_hint=want, tol=1e-3, max_outer=30)
    info = s.modal_info()
    print("modal_compute:", {k: v for k, v in info.items() if k != "residuals"})
    assert m == want
    l2, V2 = s.modal_get()
    QV = Q @ V2.T
    R = H @ V2.T - QV * l2[None, :]
    res = np.linalg.norm(R, axis=0) / np.linalg.norm(QV, axis=0)
    rel = res / np.maximum(l2, lam[1])
    # the residuals the library reports are the residuals of the consistent pencil (H, Q) ...
    assert np.allclose(info["residuals"], res, rtol=1e-3, atol=1e-9 * lam[1])
    assert rel.max() < 1e-2          # kh = 0.6 at 200 Hz on this mesh: the polynomial mass approximation stops here
    # ... the Ritz values are upper bounds of the eigenvalues, quadratically close in the residual ...
    err = l2[1:] / lam[1:m] - 1
    assert err.min() > -1e-12 and err.max() < 3 * rel.max() ** 2, (err.max(), rel.max())
    assert abs(l2[0]) < 1e-6 * lam[1]
    # ... and the vectors are Q-orthonormal
    G = V2 @ QV
    assert np.abs(G - np.eye(m)).max() < 1e-8
    # no estimate of the mode count: the block widens by itself; default tolerance (what the sweep needs)
    m2 = s.modal_compute(f_cut, n_hint=0)
    info2 = s.modal_info()
    print("modal_compute (no hint):", {k: v for k, v in info2.items() if k != "residuals"})
    assert m2 == want and info2["worst_residual"] <= 5e-2


@pytest.mark.parametrize("walls", ["rigid", "admittance"])
def test_modal_sweep_matches_direct_solve(room, walls):
    """precond="modal" vs the oracle's sparse LU on identical H, Q, A: relative L2 <= 1e-6 at tol 1e-9, true residuals <= tol,
    and far fewer iterations than block-Jacobi alone."""
    g, s = room["g"], room["sys"]
    f1 = np.sqrt(room["lam"][1]) / (2 * np.pi)     # first room mode (42.9 Hz): drive right next to it and on both sides
    f = np.array([f1 - 0.1, f1 - 0.001, f1 + 0.01, 57.5, 80.0, 101.0, 120.0, 133.0, 150.0, 171.5, 190.0, 200.0, 201.59, 201.6, 203.55])
    w = 2 * np.pi * f
    gid = np.arange(2, 8)
    s.set_surface_mesh(g.elem_surf, g.domain_index_surf, gid)
    s.assemble_surface()
    mu = np.zeros((6, len(f)), dtype=np.complex128)
    if walls == "admittance":
        mu[5] = 0.3 / (RHO0 * C0)
        mu[2] = (0.05 + 0.02j) / (RHO0 * C0)
    src = g.closest_node([1.0, 1.0, 1.2])
    s.modal_compute(np.sqrt(204.0 ** 2 + 300.0 ** 2), n_hint=70, f_acc=204.0)
    print("modal_compute:", {k: v for k, v in s.modal_info().items() if k != "residuals"})
    pN, _, st = s.sweep(w, mu, [src], [1e-4], tol=1e-9, precond="modal", check_every=8)
    print(walls, st, st.iters)
    pB, _, sb = s.sweep(w, mu, [src], [1e-4], tol=1e-9, precond="block", batch=16 if walls == "rigid" else 8, check_every=8)
    assert st.precond_used == "modal" and sb.precond_used == "block"
    print(sb, sb.iters)
    assert st.n_unconverged == 0 and st.relres.max() <= 1e-9
    H, Q = room["H"], room["Q"]
    A = O.assemble_surface(g.elem_surf, g.nos, g.domain_index_surf, gid, 1)
    q = np.zeros(len(g.nos), dtype=np.complex128)
    q[src] = 1e-4
    ref = O.solve_sweep(H, Q, A, {int(k): mu[i] for i, k in enumerate(gid)}, gid, f, q)
    rel = np.linalg.norm(pN - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert rel.max() < 1e-6, rel
    assert st.iters.mean() < 0.5 * sb.iters.mean(), (st.iters, sb.iters)

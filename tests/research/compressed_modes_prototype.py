"""Research note (CPU; uses the oracle, so it lives under tests/): can the mode tiles the coarse correction streams twice per
iteration (2 bytes x modes x nodes) be replaced by a low-order polynomial fit per 128-node tile,
V_t ~ Phi_t C_t (Phi_t: 128 x {10, 20} monomials of the centred coordinates, C_t: {10, 20} x modes)?  The fitted space is
re-diagonalised (Rayleigh-Ritz on the consistent H, Q), so the correction stays V'' diag(1 / (theta'' - w^2)) V''^T.

    python tests/research/compressed_modes_prototype.py [nx ny nz] [fmax_modes] [f1,f2,...] [delta]
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
from scipy.linalg import eigh
from scipy.sparse.linalg import eigsh

sys.path.insert(0, "/root/repo")
sys.argv = sys.argv[:1] + (sys.argv[1:] if len(sys.argv) > 1 else [])
from femder_b200.mesh import shoebox_tet4  # noqa: E402
from oracle import femder_oracle as O  # noqa: E402

nx, ny, nz = [int(t) for t in sys.argv[1:4]] if len(sys.argv) >= 4 else (34, 44, 29)
FMODES = float(sys.argv[4]) if len(sys.argv) >= 5 else 230.0
FREQS = [float(t) for t in (sys.argv[5].split(",") if len(sys.argv) >= 6 else "60,100,140,170".split(","))]
DELTA = float(sys.argv[6]) if len(sys.argv) >= 7 else 100.0
TILE = 128
g = shoebox_tet4(nx, ny, nz, jitter=0.2)
c0, rho0 = 343.0, 1.21
H, Q = O.assemble_tet4(g.elem_vol, g.nos, c0, rho0)
H = H.tocsr().real.astype(np.float64)
Q = Q.tocsr().real.astype(np.float64)
N = g.NumNosC
src = g.closest_node([1.0, 1.0, 1.2])


def mode_count(f):
    x = f / c0
    return int(4 * np.pi / 3 * 30.0 * x ** 3 + np.pi / 4 * 59.0 * x ** 2 + 38.0 / 8 * x) + 1


k = min(N - 2, int(mode_count(FMODES) * 1.12) + 8)
t0 = time.time()
lam, V = eigsh(H, k=k, M=Q, sigma=-(2 * np.pi * 30.0) ** 2, which="LM")
o = np.argsort(lam); lam, V = lam[o], V[:, o]
print(f"N={N}, h={3.0 / nx:.3f} m: {k} modes up to {np.sqrt(lam[-1]) / 2 / np.pi:.0f} Hz in {time.time() - t0:.0f} s", flush=True)

exec(open("/root/repo/tests/research/modal_deflation_variants.py").read().split("def blob_order")[1].join(["def blob_order", ""]).split("\n\n\norder = ")[0])
order = blob_order(H.indptr, H.indices, N, TILE)  # noqa: F821
Pm = sp.csr_matrix((np.ones(N), (np.arange(N), order)), shape=(N, N))
Hp = (Pm @ H @ Pm.T).tocsr(); Qp = (Pm @ Q @ Pm.T).tocsr()
Vp = V[order]
xyz = np.asarray(g.nos)[order]
srcp = int(np.where(order == src)[0][0])
n_tiles = (N + TILE - 1) // TILE
Ms = (Hp + (2 * np.pi * 50.0) ** 2 * Qp).tocsr()
Bs = []
diam = []
for t in range(n_tiles):
    a, b = t * TILE, min(N, (t + 1) * TILE)
    Bs.append(np.linalg.inv(Ms[a:b, a:b].toarray()).astype(np.float32).astype(np.float64))
    diam.append(np.linalg.norm(xyz[a:b].max(axis=0) - xyz[a:b].min(axis=0)))
print(f"tile bounding-box diagonal: mean {np.mean(diam):.2f} m (wavelength at the highest mode {c0 / (np.sqrt(lam[-1]) / 2 / np.pi):.2f} m)")


def bj(r):
    z = np.empty_like(r)
    for t, B in enumerate(Bs):
        a, b = t * TILE, min(N, (t + 1) * TILE)
        z[a:b] = B @ r[a:b]
    return z


def monomials(x, deg):
    cols = [np.ones(len(x))]
    for d in range(1, deg + 1):
        for i in range(d + 1):
            for j in range(d - i + 1):
                cols.append(x[:, 0] ** i * x[:, 1] ** j * x[:, 2] ** (d - i - j))
    return np.stack(cols, axis=1)


def compress(Vm, deg, weight=None):
    """Per tile: least-squares fit of every mode by polynomials of degree `deg` in the centred coordinates."""
    out = np.empty_like(Vm)
    nf = 0
    for t in range(n_tiles):
        a, b = t * TILE, min(N, (t + 1) * TILE)
        x = xyz[a:b] - xyz[a:b].mean(axis=0)
        x = x / (np.abs(x).max() + 1e-30)
        Phi, _ = np.linalg.qr(monomials(x, deg))
        nf = Phi.shape[1]
        out[a:b] = Phi @ (Phi.T @ Vm[a:b])
    return out, nf


def compress_pca(Vm, rank, lam_w=None):
    """Per tile: the best rank-`rank` basis of the tile's rows of the modes themselves (left singular vectors)."""
    out = np.empty_like(Vm)
    for t in range(n_tiles):
        a, b = t * TILE, min(N, (t + 1) * TILE)
        U, s_, _ = np.linalg.svd(Vm[a:b], full_matrices=False)
        Phi = U[:, :rank]
        out[a:b] = Phi @ (Phi.T @ Vm[a:b])
    return out


def ritz(Vc):
    Hs, Qs = Vc.T @ (Hp @ Vc), Vc.T @ (Qp @ Vc)
    th, Y = eigh(Hs, Qs)
    return th, Vc @ Y


def cocg(G, b, prec, tol=1e-8, maxit=5000):
    x = np.zeros_like(b); r = b.copy(); z = prec(r); p = z.copy(); rho = r @ z
    bb = b @ b
    for it in range(1, maxit + 1):
        q = G @ p
        alpha = rho / (p @ q)
        x += alpha * p; r -= alpha * q
        if r @ r <= tol * tol * bb:
            return it
        z = prec(r); rho_new = r @ z
        p = z + (rho_new / rho) * p; rho = rho_new
    return maxit


variants = {"exact": (lam, Vp, 0)}
for deg in (1, 2, 3):
    t0 = time.time()
    Vc, nf = compress(Vp, deg)
    err = np.linalg.norm(Vc - Vp, axis=0) / np.linalg.norm(Vp, axis=0)
    th, Vr = ritz(Vc)
    res = np.linalg.norm(Hp @ Vr - (Qp @ Vr) * th, axis=0) / (np.maximum(np.abs(th), (2 * np.pi * 10) ** 2) * np.linalg.norm(Qp @ Vr, axis=0))
    variants[f"poly{deg}+RR"] = (th, Vr, nf)
    variants[f"poly{deg}"] = (lam, Vc, nf)
    print(f"degree {deg}: {nf} functions per tile ({nf / TILE:.3f} of the tile's bytes); fit error of the modes: median {np.median(err):.2e}, "
          f"highest mode {err[-1]:.2e}; Ritz frequency shift max {np.abs(np.sqrt(np.abs(th)) - np.sqrt(np.abs(lam))).max() / 2 / np.pi:.3f} Hz; "
          f"eigen-residual median {np.median(res):.2e} max {res.max():.2e}  ({time.time() - t0:.0f} s)", flush=True)

for rank in (8, 12, 16, 20):
    Vc = compress_pca(Vp, rank)
    err = np.linalg.norm(Vc - Vp, axis=0) / np.linalg.norm(Vp, axis=0)
    variants[f"pca{rank}"] = (lam, Vc, rank)
    print(f"pca rank {rank}: fit error of the modes: median {np.median(err):.2e}, highest mode {err[-1]:.2e}", flush=True)
for name in ("poly1+RR", "poly1", "poly2+RR", "poly3+RR"):
    variants.pop(name)

print(f"{'f':>6} {'modes':>6} {'bj':>6} " + " ".join(f"{n:>10}" for n in variants))
for f in FREQS:
    w = 2 * np.pi * f
    G = (Hp - w * w * Qp).tocsr()
    b = np.zeros(N); b[srcp] = -w * 1e-4
    fc = np.hypot(f, DELTA)
    row = [cocg(G, b, bj)]
    m = 0
    for name, (th, Vm, nf) in variants.items():
        m = int(np.searchsorted(np.sqrt(np.maximum(th, 0)) / 2 / np.pi, fc))
        Vu, d = Vm[:, :m], 1.0 / (th[:m] - w * w)
        row.append(cocg(G, b, lambda r: bj(r) + Vu @ (d * (Vu.T @ r))))
    print(f"{f:6.1f} {m:6d} " + " ".join(f"{v:>{6 if i == 0 else 10}d}" for i, v in enumerate(row)), flush=True)

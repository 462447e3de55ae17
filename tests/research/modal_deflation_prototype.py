"""Research note (CPU, numpy/scipy; uses the oracle, so it lives under tests/): low room modes as an additive coarse
correction next to block-Jacobi in the COCG sweep (the modal path as a deflation basis,
femder/FEM_3D.py:1699-1768).

    M^-1 = B + V diag(1 / (lambda_i - w^2)) V^T        V^T Q V = I, V^T H V = diag(lambda)

B = block-Jacobi over the 128-row tiles of the blob order.  Iteration counts against plain block-Jacobi for the modes
below gamma * f, exact or rounded to bfloat16.

    python tests/research/modal_deflation_prototype.py [nx ny nz] [fmax_modes]
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import eigsh

sys.path.insert(0, "/root/repo")
from femder_b200.mesh import shoebox_tet4  # noqa: E402
from oracle import femder_oracle as O  # noqa: E402

nx, ny, nz = [int(t) for t in sys.argv[1:4]] if len(sys.argv) >= 4 else (26, 34, 22)
FMODES = float(sys.argv[4]) if len(sys.argv) >= 5 else 400.0
TILE = 128
g = shoebox_tet4(nx, ny, nz, jitter=0.0)
c0, rho0 = 343.0, 1.21
H, Q = O.assemble_tet4(g.elem_vol, g.nos, c0, rho0)
H = H.tocsr().real.astype(np.float64)
Q = Q.tocsr().real.astype(np.float64)
N = g.NumNosC
h = 3.0 / nx
src = g.closest_node([1.0, 1.0, 1.2])
print(f"N={N} h={h:.4f}")


def mode_count(f):
    V, S, L = 30.0, 59.0, 38.0
    x = f / c0
    return int(4 * np.pi / 3 * V * x ** 3 + np.pi / 4 * S * x ** 2 + L / 8 * x) + 1


k = min(N - 2, int(mode_count(FMODES) * 1.15) + 8)
t0 = time.time()
sigma = -(2 * np.pi * 30.0) ** 2
lam, V = eigsh(H, k=k, M=Q, sigma=sigma, which="LM")
o = np.argsort(lam)
lam, V = lam[o], V[:, o]
fm = np.sqrt(np.maximum(lam, 0)) / (2 * np.pi)
print(f"{k} modes in {time.time() - t0:.1f}s; f = {fm[1]:.2f} .. {fm[-1]:.1f} Hz")


def bf16(a):
    a32 = a.astype(np.float32).view(np.uint32)
    r = ((a32 + 0x7FFF + ((a32 >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32).astype(np.float64)


def blob_order(indptr, indices, n, tile):
    order = np.empty(n, dtype=np.int64)
    mark = -np.ones(n, dtype=np.int64)
    done = np.zeros(n, dtype=bool)
    frontier, fpos, next_seed, cnt, blob = [], 0, 0, 0, 0
    while cnt < n:
        seed = -1
        while fpos < len(frontier):
            c = frontier[fpos]; fpos += 1
            if not done[c]:
                seed = c; break
        if seed < 0:
            while done[next_seed]:
                next_seed += 1
            seed = next_seed
        q = [seed]; mark[seed] = blob; qh = 0; got = 0
        want = tile - cnt % tile
        while got < want and qh < len(q):
            u = q[qh]; qh += 1
            order[cnt] = u; cnt += 1; done[u] = True; got += 1
            for v in indices[indptr[u]:indptr[u + 1]]:
                if not done[v] and mark[v] != blob:
                    mark[v] = blob; q.append(v)
        frontier.extend(q[qh:])
        blob += 1
    return order


order = blob_order(H.indptr, H.indices, N, TILE)
Pm = sp.csr_matrix((np.ones(N), (np.arange(N), order)), shape=(N, N))
Hp = (Pm @ H @ Pm.T).tocsr()
Qp = (Pm @ Q @ Pm.T).tocsr()
Vp = V[order]
srcp = int(np.where(order == src)[0][0])
n_tiles = (N + TILE - 1) // TILE
shift = (2 * np.pi * 50.0) ** 2
Ms = (Hp + shift * Qp).tocsr()
Bs = []
for t in range(n_tiles):
    a, b = t * TILE, min(N, (t + 1) * TILE)
    Bs.append(np.linalg.inv(Ms[a:b, a:b].toarray()).astype(np.float32))


def apply_blocks(r):
    z = np.empty_like(r)
    for t, B in enumerate(Bs):
        a, b = t * TILE, min(N, (t + 1) * TILE)
        z[a:b] = (B @ r[a:b].astype(np.float32)).astype(np.float64)
    return z


def cocg(G, b, prec, tol=1e-8, maxit=40000):
    x = np.zeros_like(b); r = b.copy(); z = prec(r); p = z.copy(); rho = r @ z
    bb = b @ b
    for it in range(1, maxit + 1):
        q = G @ p
        alpha = rho / (p @ q)
        x += alpha * p; r -= alpha * q
        if r @ r <= tol * tol * bb:
            return x, it
        z = prec(r); rho_new = r @ z
        p = z + (rho_new / rho) * p; rho = rho_new
    return x, maxit


Vb = bf16(Vp)
freqs = [float(t) for t in (sys.argv[5].split(",") if len(sys.argv) >= 6 else "60,100,150,200,250,300".split(","))]
gammas = [1.0, 1.15, 1.3, 1.6]
print(f"{'f':>6} {'kh':>6} {'modes<f':>8} {'blockJ':>8} " + " ".join(f"{'g=%.2f' % gm:>14}" for gm in gammas) + f" {'bf16 g=1.3':>11} {'true relres':>12}")
for f in freqs:
    w = 2 * np.pi * f
    G = (Hp - w * w * Qp).tocsr()
    b = np.zeros(N); b[srcp] = -w * 1e-4
    _, it_b = cocg(G, b, apply_blocks)
    row = []
    for gm in gammas:
        m = int(np.searchsorted(fm, gm * f))
        if m >= k:
            row.append("   (need more)")
            continue
        d = 1.0 / (lam[:m] - w * w)
        Vm = Vp[:, :m]
        x, it = cocg(G, b, lambda r: apply_blocks(r) + Vm @ (d * (Vm.T @ r)))
        row.append(f"{it:8d}/{m:5d}")
    m = int(np.searchsorted(fm, 1.3 * f))
    if m < k:
        d = 1.0 / (lam[:m] - w * w)
        Vm = Vb[:, :m]
        x, it16 = cocg(G, b, lambda r: apply_blocks(r) + Vm @ (d * (Vm.T @ bf16(r) + Vm.T @ bf16(r - bf16(r)))))
        rel = np.linalg.norm(b - G @ x) / np.linalg.norm(b)
    else:
        it16, rel = -1, 0.0
    print(f"{f:6.0f} {w / c0 * h:6.3f} {int(np.searchsorted(fm, f)):8d} {it_b:8d} " + " ".join(f"{r:>14}" for r in row) + f" {it16:11d} {rel:12.2e}", flush=True)

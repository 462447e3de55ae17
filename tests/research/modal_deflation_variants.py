"""Research note (CPU; uses the oracle, so it lives under tests/): variants of the modal coarse correction of
modal_deflation_prototype.py that decide what the CUDA path may do cheaply.

  * modes of the LUMPED-mass problem (H, diag) followed by Rayleigh-Ritz with the consistent (H, Q) - what a Chebyshev-filtered
    subspace iteration without mass solves would deliver - against the exact modes of (H, Q);
  * block-Jacobi inverses rounded to bfloat16 (tensor-core operand) against FP32;
  * cut-off rule f_cut = sqrt(f^2 + D^2);
  * one admittance wall (complex sweep): d_i = 1 / (lambda_i - w^2 + 1j w sum_g mu_g v_i^T A_g v_i).

    python tests/research/modal_deflation_variants.py [nx ny nz] [fmax_modes]
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
from scipy.linalg import eigh
from scipy.sparse.linalg import eigsh

sys.path.insert(0, "/root/repo")
from femder_b200.mesh import shoebox_tet4  # noqa: E402
from oracle import femder_oracle as O  # noqa: E402

nx, ny, nz = [int(t) for t in sys.argv[1:4]] if len(sys.argv) >= 4 else (26, 34, 22)
FMODES = float(sys.argv[4]) if len(sys.argv) >= 5 else 420.0
TILE = 128
g = shoebox_tet4(nx, ny, nz, jitter=0.2)
c0, rho0 = 343.0, 1.21
H, Q = O.assemble_tet4(g.elem_vol, g.nos, c0, rho0)
H = H.tocsr().real.astype(np.float64)
Q = Q.tocsr().real.astype(np.float64)
N = g.NumNosC
h = 3.0 / nx
src = g.closest_node([1.0, 1.0, 1.2])
A7 = O.assemble_surface(g.elem_surf, g.nos, g.domain_index_surf, [7], 1)[0].tocsr().real.astype(np.float64)
print(f"N={N} h={h:.4f} (jittered mesh)")


def mode_count(f):
    V, S, L = 30.0, 59.0, 38.0
    x = f / c0
    return int(4 * np.pi / 3 * V * x ** 3 + np.pi / 4 * S * x ** 2 + L / 8 * x) + 1


k = min(N - 2, int(mode_count(FMODES) * 1.12) + 8)
sigma = -(2 * np.pi * 30.0) ** 2
t0 = time.time()
lam, V = eigsh(H, k=k, M=Q, sigma=sigma, which="LM")
o = np.argsort(lam); lam, V = lam[o], V[:, o]
print(f"{k} exact modes in {time.time() - t0:.0f}s up to {np.sqrt(lam[-1]) / 2 / np.pi:.0f} Hz")
t0 = time.time()
D = sp.diags(np.asarray(Q.sum(axis=1)).ravel())
lamL, VL = eigsh(H, k=k, M=D.tocsc(), sigma=sigma, which="LM")
o = np.argsort(lamL); VL = VL[:, o]
# Rayleigh-Ritz of the consistent problem in the lumped-mass subspace
Hs, Qs = VL.T @ (H @ VL), VL.T @ (Q @ VL)
lamR, Y = eigh(Hs, Qs)
VR = VL @ Y
print(f"lumped modes + RR in {time.time() - t0:.0f}s; max rel eigenvalue error of RR vs exact: {np.abs(lamR[1:] / lam[1:] - 1).max():.2e}")


def bf16(a):
    a32 = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
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
Hp = (Pm @ H @ Pm.T).tocsr(); Qp = (Pm @ Q @ Pm.T).tocsr(); Ap = (Pm @ A7 @ Pm.T).tocsr()
Vp, VRp = V[order], VR[order]
srcp = int(np.where(order == src)[0][0])
n_tiles = (N + TILE - 1) // TILE
shift = (2 * np.pi * 50.0) ** 2
Ms = (Hp + shift * Qp).tocsr()
Bs32, Bs16 = [], []
for t in range(n_tiles):
    a, b = t * TILE, min(N, (t + 1) * TILE)
    Bi = np.linalg.inv(Ms[a:b, a:b].toarray())
    Bs32.append(Bi.astype(np.float32).astype(np.float64))
    Bs16.append(bf16(Bi))


def blocks(Bs):
    def f(r):
        z = np.empty_like(r)
        for t, B in enumerate(Bs):
            a, b = t * TILE, min(N, (t + 1) * TILE)
            z[a:b] = B @ r[a:b]
        return z
    return f


def cocg(G, b, prec, tol=1e-8, maxit=20000):
    x = np.zeros_like(b); r = b.copy(); z = prec(r); p = z.copy(); rho = r @ z
    bb = np.vdot(b, b).real
    for it in range(1, maxit + 1):
        q = G @ p
        alpha = rho / (p @ q)
        x += alpha * p; r -= alpha * q
        if np.vdot(r, r).real <= tol * tol * bb:
            return x, it
        z = prec(r); rho_new = r @ z
        p = z + (rho_new / rho) * p; rho = rho_new
    return x, maxit


def modal(Vm, d):
    return lambda r: Vm @ (d * (Vm.T @ r))


fm = np.sqrt(np.maximum(lam, 0)) / (2 * np.pi)
fmR = np.sqrt(np.maximum(lamR, 0)) / (2 * np.pi)
b32, b16 = blocks(Bs32), blocks(Bs16)
aR = np.einsum("ij,ij->j", VRp, Ap @ VRp)     # v_i^T A v_i
print(f"{'f':>5} {'bj32':>6} {'bj16':>6} | D=200: {'exact':>6} {'lumpRR':>7} {'lRR+bj16':>9} | D=300: {'exact':>6} {'lumpRR':>7} | wall Y=0.02: {'bj32':>6} {'+modes':>7} | Y=0.3: {'bj32':>6} {'+modes':>7}")
for f in [float(t) for t in (sys.argv[5].split(",") if len(sys.argv) >= 6 else "60,100,150,200,250,300,350".split(","))]:
    w = 2 * np.pi * f
    G = (Hp - w * w * Qp).tocsr()
    b = np.zeros(N); b[srcp] = -w * 1e-4
    _, i32 = cocg(G, b, b32)
    _, i16 = cocg(G, b, b16)
    out = []
    for Dl in (200.0, 300.0):
        fc = np.sqrt(f * f + Dl * Dl)
        m = int(np.searchsorted(fm, fc)); mR = int(np.searchsorted(fmR, fc))
        if m >= k - 2:
            out += [-1, -1, -1][:3 if Dl == 200.0 else 2]
            continue
        pe = modal(Vp[:, :m], 1.0 / (lam[:m] - w * w))
        pr = modal(VRp[:, :mR], 1.0 / (lamR[:mR] - w * w))
        _, ie = cocg(G, b, lambda r: b32(r) + pe(r))
        _, ir = cocg(G, b, lambda r: b32(r) + pr(r))
        out += [ie, ir]
        if Dl == 200.0:
            prb = modal(bf16(VRp[:, :mR]), 1.0 / (lamR[:mR] - w * w))
            _, irb = cocg(G, b, lambda r: b16(r) + prb(r))
            out.append(irb)
    cp = []
    for Y in (0.02, 0.3):
        mu = Y / (rho0 * c0)
        Gc = (Hp - w * w * Qp + 1j * w * mu * Ap).tocsr()
        bc = b.astype(np.complex128) * 1j
        _, ic0 = cocg(Gc, bc, lambda r: b32(r.real) + 1j * b32(r.imag))
        fc = np.sqrt(f * f + 300.0 ** 2)
        mR = int(np.searchsorted(fmR, fc))
        if mR >= k - 2:
            cp += [ic0, -1]
            continue
        d = 1.0 / (lamR[:mR] - w * w + 1j * w * mu * aR[:mR])
        Vm = VRp[:, :mR]
        x, ic1 = cocg(Gc, bc, lambda r: b32(r.real) + 1j * b32(r.imag) + Vm @ (d * (Vm.T @ r)))
        cp += [ic0, ic1]
    print(f"{f:5.0f} {i32:6d} {i16:6d} |        {out[0]:6d} {out[1]:7d} {out[2]:9d} |        {out[3]:6d} {out[4]:7d} |              {cp[0]:6d} {cp[1]:7d} |         {cp[2]:6d} {cp[3]:7d}", flush=True)

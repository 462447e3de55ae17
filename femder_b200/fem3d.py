"""``FEM3D`` - drop-in for the hot path of the reference's ``femder.FEM_3D.FEM3D`` (``FEM_3D.py:1136-1432``,
``2042-2090``): same constructor, same ``compute()`` / ``evaluate()`` calls, same attributes left behind
(``H, Q`` scipy csc, ``A`` list of csc ordered by sorted ``mu`` keys, ``q`` csc (N,1), ``areas``, ``pN``
(n_freq, N) complex128, ``pR``, ``t``), with the assembly and the per-frequency solve running on the B200
through ``libfemder_b200.so``.  There is no CPU fallback.

Differences from the reference, all deliberate:
* Tet4 ``H, Q`` are float64 csc; the reference rounds them to complex64 (``fem_numerical.py:307-312``).
  ``compat_complex64=True`` reproduces that rounding (values rounded to float32 before the solve).
* every frequency is solved by preconditioned COCG to ``tol`` (true residual) instead of a PARDISO factorisation.  The
  preconditioner is block-Jacobi plus a coarse correction over the low room modes (``precond="modal"``, the default
  whenever it pays: the modes are what ``eigenfrequency`` computes, found once per assembly on the GPU), plain
  block-Jacobi (``"block"``) or point Jacobi (``"jacobi"``).
* a surface group without a boundary condition raises ``KeyError`` as in the reference (``FEM_3D.py:1310``);
  ``mu`` keys are paired with their own group instead of by position.
* the velocity-BC branch (``FEM_3D.py:1320-1347``) is supported: ``b = -1j*w*sum_i v_i V_i 1`` (the reference's
  cumulative ``V[indx] = v`` assignment reduces to this because ``V_i`` only has columns on group i's nodes); as in
  the reference, point sources are ignored in that branch.
* the equivalent-fluid branch (``FEM_3D.py:1367-1399``: complex ``rho(w), c(w)`` per volume domain, ``BC.fluid``) does not
  re-assemble per frequency as the reference does: ``H, Q`` are assembled once for air, the fluid domains' own stiffness and
  mass (``H_d, Q_d``) once more, and ``G(w) = H - w^2 Q + sum_d [(rho0/rho_d - 1) H_d - w^2 (rho0 c0^2 / (rho_d c_d^2) - 1) Q_d]
  + 1j w sum_g mu_g A_g`` - the bracket rides in the SpMV's overlay with per-frequency complex coefficients.  ``obj.H, obj.Q``
  are the air matrices (the reference leaves the last frequency's).
"""
from __future__ import annotations

import time

import numpy as np
from scipy.sparse import csc_matrix

from . import sharding
from .system import System


def p2SPL(p):
    """``FEM_3D.py:240-242``."""
    SPL = 10 * np.log10(0.5 * p * np.conj(p) / (2e-5) ** 2)
    return np.real(SPL)


def closest_node(nodes, node):
    """``FEM_3D.py:220-224``."""
    nodes = np.asarray(nodes)
    deltas = nodes - node
    dist_2 = np.einsum("ij,ij->i", deltas, deltas)
    return int(np.argmin(dist_2))


def receiver_stencil(nos, elem_vol, coord, interpolation_tolerance=1e-3):
    """(nodes[4], weights[4]) with p(coord) = sum_j w_j p[node_j], following ``evaluate`` (``FEM_3D.py:2077-2086``):
    the closest node if it lies within ``interpolation_tolerance``, otherwise linear Tet4 interpolation in the
    tetrahedron found the way ``coord_interpolation`` / ``which_tetra`` do (``FEM_3D.py:173-218``): candidates are
    the elements containing the closest node, the last candidate passing the barycentric test wins, and when
    none passes the reference's ``-1`` sentinel selects the last candidate."""
    nos = np.asarray(nos, dtype=np.float64)
    coord = np.asarray(coord, dtype=np.float64)
    n = closest_node(nos, coord)
    if np.linalg.norm(coord - nos[n]) < interpolation_tolerance:
        return np.array([n, n, n, n], dtype=np.int32), np.array([1.0, 0.0, 0.0, 0.0])
    ev = np.asarray(elem_vol)
    cand = np.where((ev == n).any(axis=1))[0]
    if len(cand) == 0:
        raise ValueError("receiver's closest node belongs to no volume element")
    corners = ev[cand][:, :4].astype(np.int64)
    hit = -1
    lam_hit = None
    for t in range(len(cand)):
        o = nos[corners[t, 0]]
        M = np.stack([nos[corners[t, 1]] - o, nos[corners[t, 2]] - o, nos[corners[t, 3]] - o], axis=1)
        try:
            lam = np.linalg.solve(M, coord - o)
        except np.linalg.LinAlgError:
            continue
        if np.all(lam >= 0) and np.all(lam <= 1) and lam.sum() <= 1:
            hit, lam_hit = t, lam
    if hit < 0:
        hit = len(cand) - 1
        o = nos[corners[hit, 0]]
        M = np.stack([nos[corners[hit, 1]] - o, nos[corners[hit, 2]] - o, nos[corners[hit, 3]] - o], axis=1)
        lam_hit = np.linalg.solve(M, coord - o)
    w = np.array([1.0 - lam_hit.sum(), lam_hit[0], lam_hit[1], lam_hit[2]])
    return corners[hit].astype(np.int32), w


class FEM3D:
    def __init__(self, Grid, S, R, AP, AC, BC=None, device=None, stream=None, tol=1e-8, max_iter=200000, batch="auto",
                 check_every=16, warm_start=False, compat_complex64=False, shard=True, precond="auto", modal_delta=300.0):
        self.BC = BC
        if BC is not None:
            self.mu = BC.mu
            self.v = BC.v
        else:
            self.mu = {}
            self.v = {}
        self.freq = AC.freq
        self.w = AC.w
        self.AC = AC
        self.AP = AP
        self.c0 = AP.c0
        self.rho0 = AP.rho0
        self.S = S
        self.R = R
        if Grid is not None:
            self.grid = Grid
            self.nos = Grid.nos
            self.elem_surf = Grid.elem_surf
            self.elem_vol = Grid.elem_vol
            self.domain_index_surf = Grid.domain_index_surf
            self.domain_index_vol = Grid.domain_index_vol
            self.number_ID_faces = Grid.number_ID_faces
            self.number_ID_vol = Grid.number_ID_vol
            self.NumNosC = Grid.NumNosC
            self.NumElemC = Grid.NumElemC
            self.order = Grid.order
            self.path_to_geo = getattr(Grid, "path_to_geo", None)
            self.path_to_geo_unrolled = getattr(Grid, "path_to_geo_unrolled", None)
        self.npg = 4
        self.pR = None
        self.pN = None
        self.F_n = None
        self.Vc = None
        # H, Q, A live on the GPU after compute(); the scipy objects the reference leaves behind are built on first access
        self._H = self._Q = self._A = None
        self._assembled = False
        self._A_ready = False
        # equivalent-fluid volume domains (FEM_3D.py:1203-1218): complex rho(w), c(w) per domain in BC.rhoc / BC.cc, air elsewhere
        self.rho = {}
        self.c = {}
        if BC is not None and (len(getattr(BC, "rhoc", {})) > 0 or len(getattr(BC, "cc", {})) > 0):
            if set(BC.rhoc) != set(BC.cc):
                raise ValueError("BC.rhoc and BC.cc must name the same volume domains (BC.fluid sets both)")
            for d in np.asarray(self.number_ID_vol).tolist():
                if d not in BC.rhoc:
                    self.rho[d] = np.ones_like(np.asarray(self.freq, dtype=np.float64)) * self.rho0
                    self.c[d] = np.ones_like(np.asarray(self.freq, dtype=np.float64)) * self.c0
            self.rho.update(BC.rhoc)
            self.c.update(BC.cc)
        # solver controls (no counterpart in the reference, which uses a direct solver)
        self.tol = tol
        self.max_iter = max_iter
        self.batch = batch
        self.check_every = check_every
        self.precond = precond          # "auto" | "modal" | "block" | "jacobi"
        self.modal_delta = modal_delta  # Hz: a frequency f is solved with the modes below sqrt(f^2 + modal_delta^2)
        self._modes_key = None
        self.pN_out = None              # optional caller-owned (n_freq, N) complex128 destination for pN (e.g. pinned host memory)
        self.warm_start = warm_start
        self.compat_complex64 = compat_complex64
        self.shard = shard  # under torch.distributed: split the frequencies over the ranks and gather the rows at the end
        self.stats = None
        self._device = device
        self._stream = stream
        self._sys = None
        self._sys_key = None

    # ---- H, Q, A: same attributes as the reference, fetched from the device on demand ------------------
    def _fetch_HQ(self):
        H, Q = self._sys.get_HQ()
        if self.compat_complex64:
            H, Q = H.astype(np.complex64), Q.astype(np.complex64)
        self._H, self._Q = H, Q

    @property
    def H(self):
        if self._H is None and self._assembled:
            self._fetch_HQ()
        return self._H

    @H.setter
    def H(self, value):
        self._H = value
        if value is None:
            self._assembled = False   # "H is None" is the reference's cache test (FEM_3D.py:1269): forces re-assembly

    @property
    def Q(self):
        if self._Q is None and self._assembled:
            self._fetch_HQ()
        return self._Q

    @Q.setter
    def Q(self, value):
        self._Q = value

    @property
    def A(self):
        if self._A is None and self._A_ready:
            self._A = self._sys.get_A(drop_zeros=(self.order == 2))
        return self._A

    @A.setter
    def A(self, value):
        self._A = value
        if value is None:
            self._A_ready = False

    # ---- reference properties (FEM_3D.py:1220-1238) ------------------------------------------
    @property
    def spl(self):
        return p2SPL(self.pR)

    @property
    def avg_spl(self):
        return np.mean(self.spl, axis=1)

    @property
    def spl_S(self):
        return p2SPL(self.pR.T / (1j * self.w * self.rho0)).T

    @property
    def avg_spl_S(self):
        return np.mean(self.spl_S, axis=1)

    # ---- device system -----------------------------------------------------------------------
    def _system(self):
        rank, world, local_rank = sharding.dist_info()
        dev = self._device if self._device is not None else (local_rank if world > 1 else 0)  # one process per GPU
        if self._sys is None:
            self._sys = System(dev, self._stream)
        return self._sys

    def _ensure_volume(self, sys):
        key = (id(self.nos), id(self.elem_vol), self.NumNosC, self.NumElemC)
        if self._sys_key != key:
            sys.set_volume_mesh(self.nos, self.elem_vol)
            self._sys_key = key
            return True
        return False

    # ---- compute (FEM_3D.py:1242-1432) -------------------------------------------------------
    def compute(self, timeit=True, printless=True, keep_pN=True):
        """Assemble (once; ``H`` is the cache test as in the reference) and sweep every frequency.

        ``keep_pN=False`` is an addition for meshes whose ``(n_freq, N)`` pressure array is not wanted on the
        host (1M DOF x 481 frequencies = 7.7 GB): only the receiver pressures ``pR`` are returned."""
        then = time.time()
        sys = self._system()
        N = self.NumNosC
        q = np.zeros([N, 1], dtype=np.complex128)
        group_ids, mu, rhs_vecs, rhs_coef = self._assemble(sys)

        src_nodes, src_q = [], []
        if self.S is not None:
            for ii in range(len(self.S.coord)):
                n = closest_node(self.nos, self.S.coord[ii, :])
                q[n] = np.asarray(self.S.q[ii]).ravel()[0]
                src_nodes.append(n)
                src_q.append(complex(np.asarray(self.S.q[ii]).ravel()[0]))
        self.q = csc_matrix(q)
        if rhs_vecs is not None:
            src_nodes, src_q = [], []   # FEM_3D.py:1341: b = -1j*w*Vn only
        if len(self.rho) > 0:
            if rhs_vecs is not None:
                raise NotImplementedError("velocity boundary conditions together with equivalent-fluid domains (the reference's "
                                          "fluid branch, FEM_3D.py:1367-1399, has no velocity term either)")
            mu = self._fluid_overlay(sys, group_ids, mu)

        nf = len(self.freq)
        w = np.asarray(self.w, dtype=np.float64)

        rcv_nodes = rcv_w = None
        if self.R is not None and not keep_pN:
            rcv_nodes, rcv_w = sys.locate_points(np.atleast_2d(self.R.coord), 1e-3)   # batched on the GPU

        self._sweep(sys, w, mu, src_nodes, src_q, rcv_nodes, rcv_w, keep_pN, rhs_vecs, rhs_coef)
        self.t = time.time() - then
        if timeit:
            if self.t <= 60:
                print(f"Time taken: {self.t} s")
            else:
                print(f"Time taken: {self.t / 60} min")

    def compute_candidates(self, sources, receivers, timeit=False):
        """The source / receiver candidate loop of the reference's optimizer (``FEM_3D.py:1633-1648``: per candidate pair
        ``self.S, self.R = sC[i], rC[i]; self.compute(); self.pR = self.evaluate(self.R)``) as ONE sweep: every candidate's source
        is a right-hand side of the same ``G(w)``, so the (candidate, frequency) systems fill the solver's batches together and the
        matrices, block inverses and room modes are built once.  ``sources`` / ``receivers``: equally long lists of ``Source`` /
        ``Receiver`` objects.  Returns the list of ``pR`` arrays, ``(n_freq, n_receivers_i)`` each - what the loop appends to
        ``pOptim`` - and leaves them in ``self.pR_candidates``; ``self.S``, ``self.R``, ``self.pR`` end as after the loop's last pass."""
        then = time.time()
        if len(sources) != len(receivers) or len(sources) == 0:
            raise ValueError("sources and receivers must be equally long, non-empty lists")
        if self.BC is not None and len(self.v) > 0:
            raise NotImplementedError("candidate batching covers the point-source branches (FEM_3D.py:1304-1319, 1348-1365)")
        sys = self._system()
        group_ids, mu, _, _ = self._assemble(sys)
        w = np.asarray(self.w, dtype=np.float64)
        nf = len(w)
        sets, counts, coords = [], [], []
        for S, R in zip(sources, receivers):
            nodes = [closest_node(self.nos, c) for c in np.atleast_2d(S.coord)]
            sets.append((nodes, [complex(np.asarray(v).ravel()[0]) for v in S.q]))
            rc = np.atleast_2d(R.coord)
            counts.append(len(rc))
            coords.append(rc)
        rcv_nodes, rcv_w = sys.locate_points(np.vstack(coords), 1e-3)
        precond = self._prepare_preconditioner(sys, nf)
        _, pR, stats = sys.sweep(w, mu, None, None, rcv_nodes, rcv_w, want_pN=False, tol=self.tol, max_iter=self.max_iter,
                                 batch=self.batch if precond != "modal" else "auto", check_every=self.check_every, precond=precond,
                                 src_sets=sets, modal_delta=self.modal_delta)
        if precond == "modal" and stats.n_unconverged:
            _, pB, sb = sys.sweep(w, mu, None, None, rcv_nodes, rcv_w, want_pN=False, tol=self.tol, max_iter=self.max_iter, batch=self.batch,
                                  check_every=max(self.check_every, 32), precond="block", src_sets=sets)
            bad = ~(stats.relres <= self.tol)
            pR[bad] = pB[bad]
            stats.relres[bad] = sb.relres[bad]
            stats.iters[bad] += sb.iters[bad]
            stats.n_unconverged = int((~(stats.relres <= self.tol)).sum())
        self.stats = stats
        self.iters, self.relres = stats.iters, stats.relres
        self.n_unconverged = int((~(stats.relres <= self.tol)).sum())
        if self.n_unconverged:
            import warnings
            warnings.warn(f"femder_b200: {self.n_unconverged} of {stats.relres.size} (candidate, frequency) systems did not reach "
                          f"tol={self.tol:g} (worst relative residual {np.nanmax(stats.relres):.2e})", RuntimeWarning, stacklevel=2)
        out, k = [], 0
        for c in counts:
            out.append(pR[len(out), :, k:k + c].copy())
            k += c
        self.pR_candidates = out
        self.S, self.R, self.pR = sources[-1], receivers[-1], out[-1]
        self.pN = None
        self.t = time.time() - then
        if timeit:
            print(f"Time taken: {self.t} s")
        return out

    # ---- equivalent-fluid domains (FEM_3D.py:1367-1399) --------------------------------------------------------------
    def _fluid_overlay(self, sys, group_ids, mu):
        """Add the fluid domains' correction to the overlay the SpMV reads and return the admittance table extended by its
        per-frequency coefficients.  The element matrices are those the volume assembly just left on the device (assembled with
        rho0, c0): ``H_d, Q_d`` = their sum over domain d's elements."""
        from scipy.sparse import csc_matrix as _csc
        N = self.NumNosC
        w = np.asarray(self.w, dtype=np.float64)
        nf = len(w)
        He, Qe = sys.element_matrices()                      # (nper, nper, E) as assembled for air
        ev = np.asarray(self.elem_vol).astype(np.int64)
        dom = np.asarray(self.domain_index_vol)
        mats = list(sys.get_A(drop_zeros=(self.order == 2))) if len(group_ids) else []
        rows_mu = [mu[i] for i in range(len(group_ids))]
        rho0, c0 = complex(np.real(self.rho0)), complex(np.real(self.c0))
        for d in sorted(self.BC.rhoc):
            sel = np.flatnonzero(dom == d)
            if len(sel) == 0:
                continue
            r = np.repeat(ev[sel][:, :, None], ev.shape[1], axis=2).ravel()
            c = np.repeat(ev[sel][:, None, :], ev.shape[1], axis=1).ravel()
            Hd = _csc((np.transpose(He[:, :, sel], (2, 0, 1)).ravel(), (r, c)), shape=(N, N))
            Qd = _csc((np.transpose(Qe[:, :, sel], (2, 0, 1)).ravel(), (r, c)), shape=(N, N))
            rho_d = np.asarray(self.rho[d], dtype=np.complex128).ravel()[:nf]
            c_d = np.asarray(self.c[d], dtype=np.complex128).ravel()[:nf]
            a_d = rho0 / rho_d - 1.0                               # multiplies H_d
            b_d = rho0 * c0 ** 2 / (rho_d * c_d ** 2) - 1.0        # multiplies -w^2 Q_d
            mats += [Hd, Qd]
            # the overlay applies 1j * w * mu_g to group g
            rows_mu += [a_d / (1j * w), 1j * w * b_d]
        sys.set_surface_matrices(mats)
        return np.array(rows_mu, dtype=np.complex128).reshape(len(mats), nf)

    # ---- set-up shared by compute() and compute_candidates(): mesh, matrices, boundary conditions --------------
    def _assemble(self, sys):
        new_mesh = self._ensure_volume(sys)

        group_ids = np.sort([*self.mu]) if self.BC is not None else np.array([], dtype=np.int64)
        if self.BC is not None and len(self.v) == 0:
            for bl in self.number_ID_faces:
                if bl not in self.mu:
                    raise KeyError(bl)  # FEM_3D.py:1310 indexes self.mu[bl] for every surface group (admittance branch only:
                                        # the velocity branch, 1330, walks np.sort([*self.mu]) and never touches the other groups)
        sys.set_surface_mesh(self.elem_surf, self.domain_index_surf, group_ids)
        self.areas = sys.heron_areas().astype(np.complex128)  # compute_areas returns complex128 (FEM_3D.py:743)

        if not self._assembled or new_mesh:
            if not new_mesh:
                sys.set_volume_mesh(self.nos, self.elem_vol)   # "H = None" re-assembles from the CURRENT nodes (in-place edits included)
            sys.assemble_volume(self.rho0, self.c0)
            self._modes_key = None
            if self.compat_complex64:
                H, Q = sys.get_HQ_values()
                sys.set_HQ_values(H.astype(np.float32).astype(np.float64), Q.astype(np.float32).astype(np.float64))
            self._H = self._Q = None
            self._assembled = True
        if self.BC is not None:
            sys.assemble_surface()
            self._A = None
            self._A_ready = True

        # velocity boundary conditions (FEM_3D.py:1286-1287, 1320-1347): V_i are the surface matrices of the groups in
        # BC.v; the right-hand side of frequency n is -1j*w[n] * sum_i v_i[n] * (V_i @ 1)
        rhs_vecs = rhs_coef = None
        if self.BC is not None and len(self.v) > 0:
            from . import fem_numerical as fn
            vkeys = np.sort([*self.v])
            dev = sys.device
            if self.order == 1:
                self.V = fn.assemble_surface_matrices(self.elem_surf, self.nos, self.areas, self.domain_index_surf, vkeys, 1, device=dev)
            else:
                self.V = fn.assemble_A10_3_FAST(self.domain_index_surf, vkeys, self.NumElemC, self.NumNosC, self.elem_surf, self.nos,
                                                self.c0, self.rho0, device=dev)
            rhs_vecs = [np.asarray(Vm.sum(axis=1)).ravel() for Vm in self.V]
            wv = np.asarray(self.w, dtype=np.float64)
            rhs_coef = np.array([-1j * wv * np.asarray(self.v[k]).ravel()[:len(wv)] for k in vkeys])

        nf = len(self.freq)
        mu = np.zeros((len(group_ids), nf), dtype=np.complex128)
        for i, g in enumerate(group_ids):
            mu[i] = np.asarray(self.mu[g]).ravel()[:nf]
        return group_ids, mu, rhs_vecs, rhs_coef

    # ---- the sweep of compute(): preconditioner, frequency sharding, safety net, gather ---------------------------
    def _sweep(self, sys, w, mu, src_nodes, src_q, rcv_nodes, rcv_w, keep_pN, rhs_vecs, rhs_coef):
        nf = len(w)
        precond = self._prepare_preconditioner(sys, nf)
        rank, world, _ = sharding.dist_info() if self.shard else (0, 1, 0)
        shared = None
        if world > 1 and not keep_pN and sharding.same_node():
            # one common queue for the ranks of the node (a counter in shared memory, claimed with atomic adds inside the sweep):
            # a rank that runs out of work takes the next frequency instead of idling.  (With keep_pN every rank would need the
            # whole (n_freq, N) array: static dealing then.)
            shared = sharding.counters()
            owned = np.arange(nf, dtype=np.int64)
        elif world > 1 and precond == "modal":
            # iterations are flat in frequency with the modal correction, except next to a resonance (up to several times the usual
            # count): deal those first so every rank gets its share of them, then the rest by frequency
            lam, _ = sys.modal_get(vectors=False)
            w2 = w * w
            k = np.clip(np.searchsorted(lam, w2), 1, len(lam) - 1)
            near = np.minimum(np.abs(lam[k] - w2), np.abs(lam[k - 1] - w2)) / w2
            owned = sharding.deal_by_cost(np.where(near < 1e-3, 10.0, 1.0) + w / w.max(), rank, world)
        else:
            owned = sharding.owned_frequencies(nf, 1, rank, world)
        pN_out = self.pN_out if (keep_pN and self.pN_out is not None and world == 1) else None
        pN, pR, stats = sys.sweep(w[owned], mu[:, owned], src_nodes, src_q, rcv_nodes, rcv_w, want_pN=keep_pN, pN_out=pN_out,
                                  tol=self.tol, max_iter=self.max_iter, batch=self.batch if precond != "modal" else "auto",
                                  check_every=self.check_every, warm_start=self.warm_start, precond=precond, rhs_vecs=rhs_vecs,
                                  rhs_coef=None if rhs_coef is None else rhs_coef[:, owned], modal_delta=self.modal_delta,
                                  shared_next=None if shared is None else shared.address(shared.fresh()))
        if shared is not None:
            owned = np.flatnonzero(stats.solved)
            pR = None if pR is None else pR[owned]
            stats.iters, stats.relres = stats.iters[owned], stats.relres[owned]
        self.stats = stats
        if precond == "modal" and stats.n_unconverged:
            # Safety net: what the modal path handed back above tol (a mesh too coarse for its modes near the top of the sweep, an
            # unlucky breakdown) is solved again with plain block-Jacobi, from scratch.
            bad = np.flatnonzero(~(stats.relres <= self.tol))
            ob = owned[bad]
            pN2, pR2, st2 = sys.sweep(w[ob], mu[:, ob], src_nodes, src_q, rcv_nodes, rcv_w, want_pN=keep_pN, tol=self.tol,
                                      max_iter=self.max_iter, batch=self.batch, check_every=max(self.check_every, 32),
                                      warm_start=False, precond="block", rhs_vecs=rhs_vecs, rhs_coef=None if rhs_coef is None else rhs_coef[:, ob])
            if keep_pN:
                pN[bad] = pN2
            if pR is not None:
                pR[bad] = pR2
            stats.iters[bad] += st2.iters
            stats.relres[bad] = st2.relres
            stats.n_unconverged = int(st2.n_unconverged)
            stats.n_resolved_block = int(len(bad))
            stats.seconds += st2.seconds
            stats.n_kernels += st2.n_kernels
        if world > 1:
            pN = sharding.gather_rows(pN, owned, nf) if keep_pN else None
            pR = sharding.gather_rows(pR, owned, nf) if pR is not None else None
            it = sharding.gather_rows(stats.iters[:, None], owned, nf)
            rr = sharding.gather_rows(stats.relres[:, None], owned, nf)
            self.iters, self.relres = it[:, 0], rr[:, 0]
        else:
            self.iters, self.relres = stats.iters, stats.relres
        self.pN = pN
        if pR is not None:
            self.pR = pR
        # The reference's direct solver cannot "not converge"; an iterative one can (max_iter, breakdown): never hand back such
        # a sweep silently.  relres is the TRUE residual ||b - G p|| / ||b|| of every returned system.
        bad = np.flatnonzero(~(np.asarray(self.relres) <= self.tol))
        self.n_unconverged = int(len(bad))
        if len(bad):
            import warnings
            fr = np.asarray(self.freq)[bad]
            warnings.warn(f"femder_b200: {len(bad)} of {nf} frequencies did not reach tol={self.tol:g} "
                          f"(worst relative residual {np.nanmax(np.asarray(self.relres)[bad]):.2e} at "
                          f"{fr[np.nanargmax(np.asarray(self.relres)[bad])]:g} Hz; first: {fr[:8]})", RuntimeWarning, stacklevel=2)

    # ---- room modes (FEM_3D.py:1699-1768) and the preconditioner that uses them ----------------
    def _prepare_preconditioner(self, sys, nf):
        """Resolve ``precond="auto"`` and make sure the modes the modal correction needs are on the device."""
        precond = self.precond
        if precond == "auto":
            # the modes cost a few sweeps' worth of SpMVs once per assembly: worth it for anything but tiny problems
            precond = "modal" if (self.NumNosC >= 2000 and nf >= 4) else "block"
        if precond != "modal":
            return precond
        f_hi = float(np.max(np.asarray(self.freq, dtype=np.float64)))
        delta = float(self.modal_delta)
        f_cut = float(np.hypot(f_hi, delta))
        # at most ~1000 modes are kept: shrink the margin above the sweep, then the cut itself, until the estimate fits
        # (Weyl's estimate from the room's volume and boundary area, computed by the library from the assembled mass)
        while sys.modal_estimate(f_cut) > 960 and f_cut > 1.02 * f_hi:
            f_cut = max(1.02 * f_hi, 0.97 * f_cut)
        while sys.modal_estimate(f_cut) > 960:
            f_cut *= 0.97
        key = (f_cut,)
        if self._modes_key is None or self._modes_key[0] < f_cut:
            if self.shard and sharding.dist_info()[1] > 1:
                sharding.share_modal_setup(sys)      # the ranks split the set-up's columns (every rank computes the same key)
            sys.modal_compute(f_cut, f_acc=min(f_hi, f_cut))
            self._modes_key = key
        return precond

    def eigenfrequency(self, neigs=12, near_freq=None, optimize=False):
        """The reference's modal analysis (``FEM_3D.py:1699-1768``): the ``neigs`` lowest room modes.  Leaves ``F_n`` (Hz,
        ascending) and ``Vc`` (N, neigs) like the reference, returns ``F_n``.  Computed on the GPU by Chebyshev-filtered subspace
        iteration + Rayleigh-Ritz on (H, Q); the reference runs ``spsolve(Q, H)`` + ARPACK."""
        sys = self._system()
        if not self._assembled:
            self._ensure_volume(sys)
            sys.assemble_volume(self.rho0, self.c0)
            self._assembled = True
            self._modes_key = None
        c0 = float(np.real(self.c0))
        # frequency below which Weyl's estimate predicts a few more than neigs modes
        f = 20.0
        for _ in range(200):
            if sys.modal_estimate(f) >= neigs + 4:
                break
            f *= 1.1
        for _ in range(8):
            sys.modal_compute(f, tol=1e-4, max_outer=30)
            self._modes_key = (f,)
            if sys.modal_count() >= neigs:
                break
            f *= 1.15
        lam, V = sys.modal_get()
        self.F_n = np.sqrt(np.maximum(lam[:neigs], 0.0)) / (2.0 * np.pi)
        self.Vc = V[:neigs].T.copy()
        return self.F_n

    # ---- evaluate (FEM_3D.py:2042-2090) ------------------------------------------------------
    def evaluate(self, R, interpolation_tolerance=0.001, plot=False):
        self.R = R
        if self.pN is None:
            raise RuntimeError("evaluate() needs pN: call compute() (with keep_pN=True) first")
        cols = []
        coords = np.atleast_2d(R.coord)
        if self._sys is not None and self._sys_key is not None:
            all_nodes, all_w = self._sys.locate_points(coords, interpolation_tolerance)   # batched on the GPU
        else:
            st = [receiver_stencil(self.nos, self.elem_vol, c, interpolation_tolerance) for c in coords]
            all_nodes, all_w = np.array([t[0] for t in st]), np.array([t[1] for t in st])
        for i in range(len(coords)):
            nodes, wts = all_nodes[i], all_w[i]
            if wts[0] == 1.0 and not wts[1:].any():
                cols.append(self.pN[:, nodes[0]])
            else:
                cols.append(self.pN[:, nodes.astype(np.int64)] @ wts)
        self.pR = np.asarray(cols).T.reshape(self.pN.shape[0], len(R.coord))
        return self.pR

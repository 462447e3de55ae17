"""``System``: one mesh on one B200, a thin object layer over the C ABI handle ``fdb_system``.

Everything numerical happens in ``libfemder_b200.so``; this file only shapes numpy / scipy objects into
the plain pointers the ABI takes and back into the types the reference leaves behind
(scipy ``csc_matrix`` with int32 indices, ``complex128`` ndarrays).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
from scipy.sparse import csc_matrix

from . import _lib
from ._lib import SweepParams, SweepResult, check, ptr

VALID_BATCH = (1, 2, 4, 8, 16, 32, 64)
PRECOND_CODES = {"jacobi": 0, "block": 1, "modal": 2}
PRECOND_NAMES = {v: k for k, v in PRECOND_CODES.items()}
SMALL_MESH_NODES = 200_000  # below this the kernels are launch-bound: use wider batches


def _i32(a):
    a = np.asarray(a)
    if a.dtype != np.int32:
        if a.size and (a.max() > np.iinfo(np.int32).max or a.min() < 0):
            raise ValueError("index does not fit int32")
        a = a.astype(np.int32)
    return np.ascontiguousarray(a)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


class SweepStats:
    def __init__(self, iters, relres, seconds, n_spmv, n_kernels, n_unconverged, n_resumed=0, precond_used=0):
        self.iters = iters
        self.relres = relres
        self.seconds = seconds
        self.n_spmv = n_spmv
        self.n_kernels = n_kernels
        self.n_unconverged = n_unconverged
        self.n_resumed = n_resumed          # restarts from the true residual (a slot the recurrence stopped above tol)
        self.precond_used = PRECOND_NAMES.get(precond_used, str(precond_used))

    def __repr__(self):
        it = self.iters
        return (f"SweepStats(n_freq={len(it)}, iters[min/mean/max]={it.min() if len(it) else 0}/"
                f"{it.mean() if len(it) else 0:.0f}/{it.max() if len(it) else 0}, max_relres="
                f"{self.relres.max() if len(it) else 0:.2e}, seconds={self.seconds:.3f}, "
                f"n_unconverged={self.n_unconverged}, n_resumed={self.n_resumed}, precond={self.precond_used})")


class System:
    def __init__(self, device=0, stream=None):
        """``stream``: a ``cudaStream_t`` as int (e.g. ``torch.cuda.Stream().cuda_stream``); it must not be the
        legacy default stream because the sweep captures CUDA graphs on it.  ``None`` -> library-owned stream."""
        self._lib = _lib.load()
        self.comm_world = 1
        self._h = C.c_void_p()
        check(self._lib.fdb_system_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.device = int(device)
        self.n_nodes = 0
        self.n_elem = 0
        self.nper = 0
        self.nnz = 0
        self.n_groups = 0
        self.n_tri = 0
        self.nper_s = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.fdb_system_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(self._lib.fdb_synchronize(self._h))

    # ---- volume -------------------------------------------------------------------------------
    def set_volume_mesh(self, nos, elem_vol):
        nos = _f64(nos)
        conn = _i32(elem_vol)
        if nos.ndim != 2 or nos.shape[1] != 3:
            raise ValueError("nos must be (N, 3)")
        if conn.ndim != 2 or conn.shape[1] not in (4, 10):
            raise ValueError("elem_vol must be (E, 4) or (E, 10)")
        check(self._lib.fdb_set_volume_mesh(self._h, nos.shape[0], ptr(nos), conn.shape[0], conn.shape[1], ptr(conn)))
        self.n_nodes, self.n_elem, self.nper = nos.shape[0], conn.shape[0], conn.shape[1]
        nnz = C.c_int64()
        check(self._lib.fdb_pattern_nnz(self._h, C.byref(nnz)))
        self.nnz = nnz.value

    def set_pattern(self, indptr, indices):
        indptr, indices = _i32(indptr), _i32(indices)
        check(self._lib.fdb_set_pattern(self._h, len(indptr) - 1, len(indices), ptr(indptr), ptr(indices)))
        self.n_nodes, self.nnz, self.n_elem, self.nper = len(indptr) - 1, len(indices), 0, 0

    def pattern(self):
        indptr = np.empty(self.n_nodes + 1, dtype=np.int32)
        indices = np.empty(self.nnz, dtype=np.int32)
        check(self._lib.fdb_pattern_get(self._h, ptr(indptr), ptr(indices)))
        return indptr, indices

    def elem2nnz(self):
        m = np.empty((self.n_elem, self.nper, self.nper), dtype=np.int32)
        check(self._lib.fdb_elem2nnz_get(self._h, ptr(m)))
        return m

    def assemble_volume(self, rho0, c0):
        check(self._lib.fdb_assemble_volume(self._h, float(np.real(rho0)), float(np.real(c0))))

    def get_HQ_values(self):
        H = np.empty(self.nnz, dtype=np.float64)
        Q = np.empty(self.nnz, dtype=np.float64)
        check(self._lib.fdb_get_HQ(self._h, ptr(H), ptr(Q)))
        return H, Q

    def get_HQ(self):
        """scipy csc (N, N) float64, canonical int32 indices: the types ``compute()`` leaves in ``self.H/Q``."""
        indptr, indices = self.pattern()
        H, Q = self.get_HQ_values()
        N = self.n_nodes
        # the pattern is symmetric and the values are those of a symmetric operator stored by rows:
        # CSR(row r, col c) of the operator is CSC(row c, col r), both equal by symmetry of the entries' positions
        return (csc_matrix((H, indices, indptr), shape=(N, N)), csc_matrix((Q, indices.copy(), indptr.copy()), shape=(N, N)))

    def set_HQ_values(self, H, Q):
        H, Q = _f64(H), _f64(Q)
        if len(H) != self.nnz or len(Q) != self.nnz:
            raise ValueError("H/Q value arrays must have nnz entries")
        check(self._lib.fdb_set_HQ(self._h, ptr(H), ptr(Q)))

    def set_HQ(self, H, Q):
        """Solver-only use: take H and Q as scipy sparse matrices (e.g. the reference's own csc arrays, complex64
        with zero imaginary part for Tet4).  They are converted to CSR explicitly: the reference's float32
        accumulation leaves H(r,c) != H(c,r) in the last bits, so its CSC arrays are not its CSR arrays."""
        from scipy.sparse import csr_matrix
        mats = []
        for M in (H, Q):
            M = csr_matrix(M)
            if np.iscomplexobj(M.data):
                if M.data.imag.any():
                    raise ValueError("H and Q must be real (air-only path)")
                M = M.real
            M = M.astype(np.float64)
            M.sort_indices()
            mats.append(M)
        Hr, Qr = mats
        if not (np.array_equal(Hr.indptr, Qr.indptr) and np.array_equal(Hr.indices, Qr.indices)):
            raise ValueError("H and Q must share one sparsity pattern")
        self.set_pattern(Hr.indptr, Hr.indices)
        self.set_HQ_values(Hr.data, Qr.data)

    def set_nodes(self, nos):
        """Node coordinates only: enough for the surface routines (areas, boundary matrices) without a volume mesh."""
        nos = _f64(nos)
        if nos.ndim != 2 or nos.shape[1] != 3:
            raise ValueError("nos must be (N, 3)")
        check(self._lib.fdb_set_nodes(self._h, nos.shape[0], ptr(nos)))
        self.n_nodes = int(nos.shape[0])
        self.n_elem = 0

    def element_matrices(self):
        n2 = self.nper * self.nper
        He = np.empty((n2, self.n_elem), dtype=np.float64)
        Qe = np.empty((n2, self.n_elem), dtype=np.float64)
        check(self._lib.fdb_get_element_matrices(self._h, ptr(He), ptr(Qe)))
        return He.reshape(self.nper, self.nper, self.n_elem), Qe.reshape(self.nper, self.nper, self.n_elem)

    # ---- surface ------------------------------------------------------------------------------
    def set_surface_mesh(self, elem_surf, domain_index_surf, group_ids):
        """``group_ids``: the sorted boundary-condition keys (``np.sort([*mu])``, FEM_3D.py:1283)."""
        conn = _i32(elem_surf)
        if conn.ndim != 2 or conn.shape[1] not in (3, 6):
            raise ValueError("elem_surf must be (T, 3) or (T, 6)")
        group_ids = np.asarray(group_ids)
        dom = np.asarray(domain_index_surf)
        grp = np.full(len(conn), -1, dtype=np.int32)
        for i, g in enumerate(group_ids):
            grp[dom == g] = i
        check(self._lib.fdb_set_surface_mesh(self._h, conn.shape[0], conn.shape[1], ptr(conn), ptr(grp), len(group_ids)))
        self.n_tri, self.nper_s, self.n_groups = conn.shape[0], conn.shape[1], len(group_ids)

    def heron_areas(self):
        a = np.empty(self.n_tri, dtype=np.float64)
        check(self._lib.fdb_heron_areas(self._h, ptr(a)))
        return a

    def assemble_surface(self):
        check(self._lib.fdb_assemble_surface(self._h))

    def get_A(self, drop_zeros=False):
        """list of csc (N, N) float64, one per group, each with its own pattern (entries its triangles touch)."""
        nnz_b = C.c_int64()
        check(self._lib.fdb_surface_nnz(self._h, C.byref(nnz_b)))
        Zb = nnz_b.value
        N = self.n_nodes
        indptr = np.empty(N + 1, dtype=np.int32)
        indices = np.empty(Zb, dtype=np.int32)
        check(self._lib.fdb_surface_pattern_get(self._h, ptr(indptr), ptr(indices)))
        vals = np.empty((self.n_groups, Zb), dtype=np.float64)
        touched = np.empty((self.n_groups, Zb), dtype=np.uint8)
        check(self._lib.fdb_surface_get(self._h, ptr(vals), ptr(touched)))
        rows = np.repeat(np.arange(N, dtype=np.int32), np.diff(indptr))
        out = []
        for g in range(self.n_groups):
            keep = touched[g].astype(bool)
            if drop_zeros:
                keep &= vals[g] != 0.0
            cnt = np.bincount(rows[keep], minlength=N)
            ip = np.zeros(N + 1, dtype=np.int32)
            np.cumsum(cnt, out=ip[1:])
            out.append(csc_matrix((vals[g][keep], indices[keep], ip), shape=(N, N)))
        return out

    def set_surface_matrices(self, A_list):
        """Solver-only use: give A_g as scipy matrices (any pattern); they are merged on one union pattern."""
        N = self.n_nodes
        ng = len(A_list)
        if ng == 0:
            ip = np.zeros(N + 1, dtype=np.int32)
            check(self._lib.fdb_set_surface(self._h, 0, 0, ptr(ip), None, None))
            self.n_groups = 0
            return
        mats = [csc_matrix(A).astype(np.float64).T.tocsc() for A in A_list]  # CSC of A^T == CSR of A
        for m in mats:
            m.sort_indices()
        union = None
        for m in mats:
            p = csc_matrix((np.ones(m.nnz), m.indices, m.indptr), shape=(N, N))
            union = p if union is None else union + p
        union = union.tocsc()
        union.sort_indices()
        ip, ix = union.indptr.astype(np.int32), union.indices.astype(np.int32)
        vals = np.zeros((ng, len(ix)), dtype=np.float64)
        cols = np.repeat(np.arange(N), np.diff(ip))
        key_u = cols.astype(np.int64) * N + ix
        for g, m in enumerate(mats):
            cm = np.repeat(np.arange(N), np.diff(m.indptr))
            key = cm.astype(np.int64) * N + m.indices
            pos = np.searchsorted(key_u, key)
            vals[g, pos] = m.data
        check(self._lib.fdb_set_surface(self._h, ng, len(ix), ptr(ip), ptr(ix), ptr(vals)))
        self.n_groups = ng

    # ---- room modes ---------------------------------------------------------------------------
    def modal_compute(self, f_cut, n_hint=0, tol=0.0, max_outer=12, f_acc=0.0):
        """All modes ``H v = lambda Q v`` below ``f_cut`` Hz, on the GPU (``FEM3D.eigenfrequency``'s job,
        ``FEM_3D.py:1699-1768``).  ``n_hint``: expected number of modes (sizes the block); ``tol``: relative residual of the
        modes below ``f_acc`` (default: all; the rest may be 10 x coarser).  Returns the number kept."""
        n = C.c_int(0)
        check(self._lib.fdb_modal_compute(self._h, float(f_cut), float(f_acc), int(n_hint), float(tol), int(max_outer), C.byref(n)))
        return n.value

    def modal_estimate(self, f):
        """Weyl's estimate of the number of room modes below ``f`` Hz (0 if the room's size is unknown)."""
        n = C.c_int(0)
        check(self._lib.fdb_modal_estimate(self._h, float(f), C.byref(n)))
        return n.value

    def modal_compression(self, on=True):
        """Stream the modes tile-compressed (default: 16 basis functions per 128-node tile) or at full resolution."""
        check(self._lib.fdb_modal_compression(self._h, 2 if on == "force" else (1 if on else 0)))

    @staticmethod
    def comm_unique_id():
        """128 bytes identifying a new group of ranks (one rank creates it, all ranks pass it to :meth:`comm_init`)."""
        buf = C.create_string_buffer(128)
        check(_lib.load().fdb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id, rank, world):
        """Join the group of ranks that share this room's modal set-up (collective; NCCL).  ``modal_compute`` is then a collective call."""
        if len(unique_id) != 128:
            raise ValueError("unique_id must be the 128 bytes of comm_unique_id()")
        check(self._lib.fdb_comm_init(self._h, C.create_string_buffer(bytes(unique_id), 128), int(rank), int(world)))
        self.comm_world = int(world)

    def comm_free(self):
        check(self._lib.fdb_comm_free(self._h))
        self.comm_world = 1

    def modal_info(self):
        """dict(outer, block, mass_order, worst_residual, residuals) of the stored modes."""
        o, b, t, wr = C.c_int(0), C.c_int(0), C.c_int(0), C.c_double(0.0)
        res = np.empty(self.modal_count(), dtype=np.float64)
        check(self._lib.fdb_modal_info(self._h, C.byref(o), C.byref(b), C.byref(t), C.byref(wr), ptr(res)))
        return dict(outer=o.value, block=b.value, mass_order=t.value, worst_residual=wr.value, residuals=res)

    def modal_set(self, lam, V):
        """Modes computed elsewhere: ``lam`` (m,) ascending, ``V`` (m, N) Q-orthonormal rows."""
        lam, V = _f64(lam), _f64(V)
        if V.shape != (len(lam), self.n_nodes):
            raise ValueError("V must be (n_modes, n_nodes)")
        check(self._lib.fdb_modal_set(self._h, len(lam), ptr(lam), ptr(V)))

    def modal_count(self):
        n = C.c_int(0)
        check(self._lib.fdb_modal_count(self._h, C.byref(n)))
        return n.value

    def modal_get(self, vectors=True):
        m = self.modal_count()
        lam = np.empty(m, dtype=np.float64)
        V = np.empty((m, self.n_nodes), dtype=np.float64) if vectors else None
        check(self._lib.fdb_modal_get(self._h, ptr(lam), ptr(V)))
        return lam, V

    def modal_apply(self, omega, n_use, r):
        """``z = V diag(1 / (lam - omega_f^2)) V^T r`` through the solver's tensor-core kernels; r is (N, 16) real."""
        omega, r = _f64(omega), _f64(r)
        if r.shape != (self.n_nodes, 16) or len(omega) != 16:
            raise ValueError("r must be (N, 16) and omega (16,)")
        z = np.empty_like(r)
        check(self._lib.fdb_modal_apply(self._h, ptr(omega), int(n_use), ptr(r), ptr(z)))
        return z

    # ---- receivers ----------------------------------------------------------------------------
    def locate_points(self, coords, tol=1e-3):
        """(nodes (P, 4) int32, weights (P, 4)) with p(coord) = sum_j w_j p[node_j] for every point, on the GPU."""
        pts = _f64(np.atleast_2d(coords))
        nodes = np.empty((len(pts), 4), dtype=np.int32)
        w = np.empty((len(pts), 4), dtype=np.float64)
        check(self._lib.fdb_locate_points(self._h, len(pts), ptr(pts), float(tol), ptr(nodes), ptr(w)))
        return nodes, w

    # ---- sweep --------------------------------------------------------------------------------
    def sweep(self, omega, mu, src_nodes, src_q, rcv_nodes=None, rcv_w=None, want_pN=True, tol=1e-8,
              max_iter=200000, batch=8, check_every=32, warm_start=False, pN_out=None, force_complex=False, rhs_vecs=None,
              rhs_coef=None, src_sets=None, precond="block", modal_delta=0.0, shared_next=None):
        """Solve ``(H - w^2 Q + 1j w sum_g mu_g A_g) p = -1j w q`` for every w in ``omega``.

        mu: (n_groups, n_freq) complex.  ``rhs_vecs`` (list of real length-N vectors, dense or scipy sparse) with
        ``rhs_coef`` (n_vec, n_freq) complex add ``sum_k rhs_coef[k, n] * vec_k`` to the right-hand side of frequency n
        (velocity boundary conditions).  ``src_sets`` = list of ``(nodes, q)`` pairs solves several right-hand sides
        per frequency in the same batches (source candidate loops); outputs then carry a leading set axis.
        ``precond``: "block" = block-Jacobi over the 128-row tiles of the sweep's internal order where implemented
        (batches of up to 16 real or 8 complex slots), point Jacobi elsewhere; "jacobi" = point Jacobi everywhere;
        "modal" = block-Jacobi plus the coarse correction over the room modes (``modal_compute`` / ``modal_set`` first; the
        batch is then 16 real or 8 complex slots; a batch applies the modes below sqrt(f_max^2 + modal_delta^2), default 300 Hz).
        Returns (pN or None, pR or None, SweepStats)."""
        omega = _f64(np.atleast_1d(omega))
        nf = len(omega)
        mu = _c128(np.asarray(mu).reshape(self.n_groups, nf)) if self.n_groups else None
        n_sets = 1
        set_ptr = None
        if src_sets is not None:
            n_sets = len(src_sets)
            set_ptr = np.zeros(n_sets + 1, dtype=np.int32)
            set_ptr[1:] = np.cumsum([len(np.atleast_1d(n)) for n, _ in src_sets])
            src_nodes = np.concatenate([np.atleast_1d(n) for n, _ in src_sets]) if n_sets else []
            src_q = np.concatenate([np.atleast_1d(q).ravel() for _, q in src_sets]) if n_sets else []
        src_nodes = _i32(np.atleast_1d(src_nodes))
        src_q = _c128(np.atleast_1d(src_q).ravel())
        n_vec = 0 if rhs_vecs is None else len(rhs_vecs)
        if batch in (0, None, "auto"):
            # a rigid-wall sweep with real sources runs in real arithmetic (half the bytes per frequency), where 16
            # frequencies fill the 128-byte gather line that 8 complex ones do
            real = (mu is None or not mu.any()) and not src_q.imag.any() and not force_complex and n_vec == 0
            batch = 16 if real else 8
            if precond == "modal":
                pass    # the tensor-core kernels of the modal correction fix the batch: 16 real or 8 complex slots
            elif self.n_nodes <= SMALL_MESH_NODES:
                # small meshes: the kernels are latency-bound, wide batches amortise the fixed cost per launch.  (The
                # block-Jacobi kernels cover batches up to 16 only; on the coarse config-1 mesh, k h up to 1, 64 slots with point
                # Jacobi sweep 20-500 Hz in 1.39 s against 2.18 s for 16 slots with block-Jacobi.)
                batch *= 4
        if batch not in VALID_BATCH:
            raise ValueError(f"batch must be one of {VALID_BATCH} or 'auto'")
        if len(src_q) != len(src_nodes):
            raise ValueError("src_nodes and src_q differ in length")
        prm = SweepParams()
        prm.n_freq = nf
        prm.omega = ptr(omega)
        prm.mu = ptr(mu) if mu is not None else None
        prm.n_src = len(src_nodes)
        prm.src_nodes = ptr(src_nodes)
        prm.src_q = ptr(src_q)
        prm.n_src_sets = n_sets if src_sets is not None else 0
        prm.src_set_ptr = ptr(set_ptr) if src_sets is not None else None
        prm.tol = float(tol)
        prm.max_iter = int(max_iter)
        prm.batch = int(batch)
        prm.check_every = int(check_every)
        prm.warm_start = int(bool(warm_start))
        if precond not in PRECOND_CODES:
            raise ValueError("precond must be 'modal', 'block' or 'jacobi'")
        prm.precond = PRECOND_CODES[precond]
        prm.modal_delta_hz = float(modal_delta)
        prm.arithmetic = 1 if force_complex else 0
        prm.n_rhs_vec = n_vec
        if n_vec:
            from scipy.sparse import issparse
            ptr_, nodes_, vals_ = [0], [], []
            for v in rhs_vecs:
                if issparse(v):
                    v = np.asarray(v.todense())
                v = np.asarray(v, dtype=np.float64).ravel()
                if len(v) != self.n_nodes:
                    raise ValueError("right-hand-side vectors must have one entry per node")
                nz = np.flatnonzero(v)
                nodes_.append(nz.astype(np.int32)); vals_.append(v[nz]); ptr_.append(ptr_[-1] + len(nz))
            rhs_ptr = np.asarray(ptr_, dtype=np.int32)
            rhs_nodes = _i32(np.concatenate(nodes_)) if ptr_[-1] else np.zeros(0, dtype=np.int32)
            rhs_vals = _f64(np.concatenate(vals_)) if ptr_[-1] else np.zeros(0)
            rhs_coef = _c128(np.asarray(rhs_coef).reshape(n_vec, nf))
            prm.rhs_ptr, prm.rhs_nodes, prm.rhs_vals, prm.rhs_coef = ptr(rhs_ptr), ptr(rhs_nodes), ptr(rhs_vals), ptr(rhs_coef)
        n_rcv = 0
        if rcv_nodes is not None:
            rcv_nodes = _i32(np.asarray(rcv_nodes).reshape(-1, 4))
            rcv_w = _f64(np.asarray(rcv_w).reshape(-1, 4))
            n_rcv = len(rcv_nodes)
        prm.n_rcv = n_rcv
        prm.rcv_nodes = ptr(rcv_nodes) if n_rcv else None
        prm.rcv_w = ptr(rcv_w) if n_rcv else None
        res = SweepResult()
        pN = None
        if pN_out is not None:
            pN = pN_out
        elif want_pN:
            pN = np.empty((n_sets * nf, self.n_nodes), dtype=np.complex128)
        pR = np.empty((n_sets * nf, n_rcv), dtype=np.complex128) if n_rcv else None
        iters = np.zeros(n_sets * nf, dtype=np.int32)
        relres = np.zeros(n_sets * nf, dtype=np.float64)
        res.pN = ptr(pN)
        res.pR = ptr(pR)
        res.iters = ptr(iters)
        res.relres = ptr(relres)
        # shared_next: address of an int64 counter shared by the processes of this node that were given the same list
        # (sharding.SharedCounters): systems are claimed from one common queue; stats.solved marks the ones solved here
        solved = np.zeros(n_sets * nf, dtype=np.uint8)
        res.solved = ptr(solved)
        prm.shared_next = C.c_void_p(int(shared_next)) if shared_next else None
        check(self._lib.fdb_sweep(self._h, C.byref(prm), C.byref(res)))
        if src_sets is not None:
            if pN is not None and pN_out is None:
                pN = pN.reshape(n_sets, nf, self.n_nodes)
            if pR is not None:
                pR = pR.reshape(n_sets, nf, n_rcv)
            iters, relres = iters.reshape(n_sets, nf), relres.reshape(n_sets, nf)
        st = SweepStats(iters, relres, res.seconds, res.n_spmv, res.n_kernels, res.n_unconverged, res.n_resumed, res.precond_used)
        st.real_arithmetic = bool(res.real_arithmetic)
        st.solved = solved.astype(bool).reshape(iters.shape)
        return pN, pR, st

    def spmv(self, omega, mu, x):
        """y = G_f x for a batch; x is (N, F) complex128."""
        x = _c128(x)
        F = x.shape[1]
        omega = _f64(np.atleast_1d(omega))
        mu = _c128(np.asarray(mu).reshape(self.n_groups, F)) if self.n_groups else None
        y = np.empty_like(x)
        check(self._lib.fdb_spmv(self._h, F, ptr(omega), ptr(mu) if mu is not None else None, ptr(x), ptr(y)))
        return y

    def spmv_bench(self, batch, reps, with_boundary=True, real=False):
        ms = C.c_double()
        nbytes = C.c_int64()
        flags = 2 if real else int(bool(with_boundary))
        check(self._lib.fdb_spmv_bench(self._h, int(batch), int(reps), flags, C.byref(ms), C.byref(nbytes)))
        return ms.value, nbytes.value

    def iteration_enqueue(self, batch, reps, which, real=True, block_jacobi=True, modal=False):
        """Enqueue ``reps`` launches of one kernel of a solver iteration (which: -1 prepare, 0 K5, 1 K6a, 2 K6b, with ``modal``
        3 projection, 4 expansion, 5 coefficients) on the system's stream without synchronising; returns that kernel's algorithmic
        bytes per launch."""
        nbytes = C.c_int64()
        flags = (2 if real else 0) | (4 if block_jacobi else 0) | (8 if modal else 0)
        check(self._lib.fdb_iteration_enqueue(self._h, int(batch), int(reps), flags, int(which), C.byref(nbytes)))
        return nbytes.value

    def spmv_enqueue(self, batch, reps, with_boundary=True, real=False):
        """Enqueue ``reps`` SpMV launches on the system's stream without synchronising; returns algorithmic bytes."""
        nbytes = C.c_int64()
        flags = 2 if real else int(bool(with_boundary))
        check(self._lib.fdb_spmv_enqueue(self._h, int(batch), int(reps), flags, C.byref(nbytes)))
        return nbytes.value

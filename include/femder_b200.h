/*
 * femder_b200 - C ABI of the B200 (sm_100a) implementation of femder's FEM3D.compute() hot path.
 *
 * The reference (gutoalvim/femder) is pure Python and has no FFI layer; its boundary for this path is the
 * set of Python call sites inside FEM3D.compute() (femder/FEM_3D.py:1242-1432).  Each entry point below
 * names the reference call site(s) it replaces.  A maintainer binds them with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; fdb_last_error() then describes it;
 *   - plain pointers and sizes only; arrays are C-contiguous; indices are int32 (scipy's index dtype for
 *     these sizes), values are IEEE double; complex numbers are interleaved (re, im) doubles;
 *   - every data pointer may be a host pointer OR a device pointer of the system's device (copies use
 *     cudaMemcpyDefault), so PyTorch-owned device buffers can be passed without staging;
 *   - a system handle owns its device memory; the library keeps no pointer to caller memory past a call;
 *   - there is no CPU fallback: without a usable CUDA device every call fails.
 */
#ifndef FEMDER_B200_H
#define FEMDER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fdb_system fdb_system;

/* ---- library / device ------------------------------------------------------------------------- */
const char* fdb_last_error(void);
int fdb_abi_version(void);   /* 2 */
int fdb_device_count(int* count);

/* One system = one mesh on one GPU.  `stream` is a cudaStream_t (0 = library-owned non-blocking stream);
 * pass torch.cuda.current_stream().cuda_stream to have torch events bracket the library's work. */
int fdb_system_create(int device, void* stream, fdb_system** out);
int fdb_system_destroy(fdb_system* sys);
int fdb_synchronize(fdb_system* sys);

/* ---- volume mesh, CSR pattern, element->nnz map (K4) ------------------------------------------ *
 * Replaces the COO->CSC conversion inside csc_matrix(...) at femder/fem_numerical.py:337-343 and
 * coo_matrix(...).tocsc() at femder/FEM_3D.py:413-416: the pattern is the canonical (sorted, duplicate-
 * free) pattern scipy produces, symmetric, so CSR == CSC.  nper = 4 (Tet4) or 10 (Tet10, gmsh node order). */
int fdb_set_volume_mesh(fdb_system* sys, int64_t n_nodes, const double* nos /* [n_nodes][3] */,
                        int64_t n_elem, int nper, const int32_t* conn /* [n_elem][nper] */);
/* Node coordinates only, for the surface routines on their own (fdb_set_surface_mesh, fdb_heron_areas, fdb_assemble_surface:
 * the reference's compute_areas / assemble_surface_matrices take vertices and faces, no volume mesh - femder/FEM_3D.py:722-744,
 * 475-523).  Drops any volume pattern, matrices and modes the system held. */
int fdb_set_nodes(fdb_system* sys, int64_t n_nodes, const double* nos /* [n_nodes][3] */);
int fdb_pattern_nnz(fdb_system* sys, int64_t* nnz);
int fdb_pattern_get(fdb_system* sys, int32_t* indptr /* [n_nodes+1] */, int32_t* indices /* [nnz] */);
/* elem2nnz[e][a][b] = position in `indices` of entry (row conn[e][a], col conn[e][b]). */
int fdb_elem2nnz_get(fdb_system* sys, int32_t* elem2nnz /* [n_elem][nper][nper] */);
/* Use an externally built pattern instead (solver-only use: the reference's own H/Q/A arrays). */
int fdb_set_pattern(fdb_system* sys, int64_t n_nodes, int64_t nnz, const int32_t* indptr, const int32_t* indices);

/* ---- volume assembly (K1 Tet4, K2 Tet10) ------------------------------------------------------- *
 * Replaces fem_numerical.pre_compute_volume_assemble_vars + assemble_volume_matrices
 * (femder/fem_numerical.py:129-150, 267-345; call site FEM_3D.py:1274-1275) for nper = 4 and
 * assemble_Q_H_4_FAST_2order + int_tetra10_4gauss (FEM_3D.py:398-417, 888-919; call site 1277) for
 * nper = 10.  FP64 throughout; values land in the pattern's nnz order by a deterministic segmented sum. */
int fdb_assemble_volume(fdb_system* sys, double rho0, double c0);
int fdb_get_HQ(fdb_system* sys, double* H_vals /* [nnz] */, double* Q_vals /* [nnz] */);
int fdb_set_HQ(fdb_system* sys, const double* H_vals, const double* Q_vals);
/* Element matrices of the last fdb_assemble_volume, [nper*nper][n_elem] (debug / parity). */
int fdb_get_element_matrices(fdb_system* sys, double* He, double* Qe);

/* ---- surface mesh, Heron areas, admittance matrices (K3) -------------------------------------- *
 * Replaces compute_areas/area_normal_ph (FEM_3D.py:722-771; call site 1267), assemble_surface_matrices
 * (FEM_3D.py:475-523; call site 1283) for nper = 3 and assemble_A10_3_FAST + int_tri10_3gauss
 * (FEM_3D.py:541-554, 1053-1092; call site 1290) for nper = 6.  group_of_tri[t] is an index into the
 * caller's sorted group list (0..n_groups-1) or -1 for triangles with no boundary condition. */
int fdb_set_surface_mesh(fdb_system* sys, int64_t n_tri, int nper, const int32_t* conn /* [n_tri][nper] */,
                         const int32_t* group_of_tri /* [n_tri] */, int n_groups);
int fdb_heron_areas(fdb_system* sys, double* areas /* [n_tri] */);
int fdb_assemble_surface(fdb_system* sys);
int fdb_surface_nnz(fdb_system* sys, int64_t* nnz_b);
int fdb_surface_pattern_get(fdb_system* sys, int32_t* indptr /* [n_nodes+1] */, int32_t* indices /* [nnz_b] */);
/* A_vals[g][k] and touched[g][k] (1 where group g has at least one triangle contributing to entry k). */
int fdb_surface_get(fdb_system* sys, double* A_vals /* [n_groups][nnz_b] */, uint8_t* touched /* same */);
/* Solver-only use: give the per-group matrices directly, all on one pattern. */
int fdb_set_surface(fdb_system* sys, int n_groups, int64_t nnz_b, const int32_t* indptr, const int32_t* indices,
                    const double* A_vals /* [n_groups][nnz_b] */);

/* ---- receivers (K7, set-up half) ------------------------------------------------------------------ *
 * Batched receiver location: for every query point the 4 nodes and weights with p(point) = sum_j w_j p[node_j].
 * Replaces, for many points at once, the nearest-node rule of evaluate() (femder/FEM_3D.py:2077-2086) and
 * closest_node + prob_elem + which_tetra + coord_interpolation (FEM_3D.py:173-224): the closest node if it lies
 * within `tol`, otherwise linear interpolation in the LAST element around that node that passes the barycentric
 * test (the last candidate if none does).  The result feeds fdb_sweep_params.rcv_nodes / rcv_w. */
int fdb_locate_points(fdb_system* sys, int64_t n_pts, const double* pts /* [n_pts][3] */, double tol,
                      int32_t* nodes /* [n_pts][4] */, double* weights /* [n_pts][4] */);

/* ---- frequency sweep (K5 fused multi-frequency SpMV, K6 COCG, K7 receiver gather) -------------- *
 * Replaces the hot loop FEM_3D.py:1304-1319: for every frequency n
 *     G = H + 1j*w[n]*sum_g mu[g][n]*A[g] - w[n]**2*Q ;  b = -1j*w[n]*q ;  pN[n] = solve(G, b)
 * where the reference factorises G from scratch with PARDISO (pyMKL, mtype 13, phases 12 and 33). */
typedef struct fdb_sweep_params {
    int64_t n_freq;
    const double* omega;      /* [n_freq] angular frequencies w */
    const double* mu;         /* [n_groups][n_freq] complex admittances, may be NULL when n_groups == 0 */
    int64_t n_src;
    const int32_t* src_nodes; /* [n_src]  node of each source (later duplicates overwrite, as the reference) */
    const double* src_q;      /* [n_src]  complex volume velocities q */
    /* multi-RHS: the sources are split into n_src_sets right-hand sides sharing G(w) (source/receiver candidate loops,
     * FEM_3D.py:1633-1648).  0 = one set.  Outputs are then [n_src_sets][n_freq][...]. */
    int64_t n_src_sets;
    const int32_t* src_set_ptr; /* [n_src_sets+1] set k = sources src_set_ptr[k] .. src_set_ptr[k+1] */
    double tol;               /* relative residual to reach: every returned solution satisfies ||b - G x|| / ||b|| <= tol on the TRUE
                               * residual (a slot the recurrence stopped early is restarted from its true residual), or is
                               * counted in n_unconverged */
    int32_t max_iter;
    int32_t batch;            /* solver slots = frequencies per fused SpMV: 1, 2, 4, 8, 16, 32 or 64 */
    int32_t check_every;      /* iterations between host convergence checks (one CUDA graph launch) */
    int32_t warm_start;       /* 1: start each batch from the previous batch's last solution */
    int32_t precond;          /* 0: point Jacobi; 1: block-Jacobi (FP32 inverses of the 128-row diagonal blocks of H + s Q, s = (2 pi 50 Hz)^2)
                               * where it is implemented - batches of up to 16 real or 8 complex slots - and point Jacobi elsewhere;
                               * 2: block-Jacobi plus the modal coarse correction V diag(d) V^T over the low room modes
                               * (fdb_modal_compute / fdb_modal_set must have been called; the batch is then 16 real or 8 complex slots) */
    int32_t arithmetic;       /* 0: automatic (real arithmetic for rigid-wall sweeps with real sources), 1: always complex */
    /* extra right-hand-side terms (velocity boundary conditions, FEM_3D.py:1320-1347): b_n += sum_k rhs_coef[k][n] * vec_k */
    int64_t n_rhs_vec;
    const int32_t* rhs_ptr;   /* [n_rhs_vec+1] vec_k = entries rhs_ptr[k] .. rhs_ptr[k+1] */
    const int32_t* rhs_nodes; /* node of each entry */
    const double* rhs_vals;   /* real value of each entry */
    const double* rhs_coef;   /* [n_rhs_vec][n_freq] complex coefficients (the caller folds -1j*w in) */
    int64_t n_rcv;            /* receivers: pR[n][i] = sum_j rcv_w[i][j] * p[n][rcv_nodes[i][j]] */
    const int32_t* rcv_nodes; /* [n_rcv][4] (unused slots: weight 0, any valid node) */
    const double* rcv_w;      /* [n_rcv][4] */
    double modal_delta_hz;    /* precond 2: a batch applies the modes below sqrt(f_max^2 + modal_delta_hz^2); 0 = default (300 Hz) */
    int64_t* shared_next;     /* NULL: this call solves every system it is given.  Otherwise a HOST counter shared by the processes
                               * that were all given the SAME list (one per GPU of a node; e.g. in POSIX shared memory, zero before
                               * the first of them starts): systems are claimed one at a time, in the sweep's own dispatch order,
                               * with an atomic fetch-and-add, so a process that runs out of work takes the next system of the
                               * common queue instead of idling; fdb_sweep_result.solved says which ones this call solved */
} fdb_sweep_params;
/* Multi-GPU: frequencies are independent, so each rank creates its own system on its own GPU and either passes
 * only the frequencies it owns or shares one queue with the other ranks (shared_next); there is no collective on this path. */

typedef struct fdb_sweep_result {
    double* pR;       /* [n_sets][n_freq][n_rcv] complex; may be NULL */
    double* pN;       /* [n_sets][n_freq][n_nodes] complex; may be NULL */
    int32_t* iters;   /* [n_sets][n_freq] iterations used; may be NULL */
    double* relres;   /* [n_sets][n_freq] true relative residual ||b - G x|| / ||b|| of what is returned; may be NULL */
    double seconds;   /* device time of the sweep (CUDA events on the system's stream) */
    int64_t n_spmv;   /* fused SpMV launches */
    int64_t n_kernels;/* all kernel launches of the sweep */
    int32_t n_unconverged;    /* systems returned with ||b - G x|| / ||b|| > tol (max_iter reached, breakdown, NaN) */
    int32_t real_arithmetic;  /* 1 if the sweep ran in real arithmetic (G real, b imaginary: p = 1j*u) */
    int32_t n_resumed;        /* restarts from the true residual (see tol) */
    int32_t precond_used;     /* the preconditioner the sweep actually ran: 0, 1 or 2 as in fdb_sweep_params.precond */
    uint8_t* solved;  /* [n_sets][n_freq] 1 where this call solved the system (all of them without shared_next); may be NULL */
} fdb_sweep_result;

int fdb_sweep(fdb_system* sys, const fdb_sweep_params* prm, fdb_sweep_result* res);
/* sizeof(fdb_sweep_params) / sizeof(fdb_sweep_result) as compiled, so a binding can verify its struct layout. */
int fdb_sizeof_sweep_params(void);
int fdb_sizeof_sweep_result(void);

/* ---- room modes (modal analysis; coarse correction of the sweep) --------------------------------------------------- *
 * The eigenpairs H v = lambda Q v (lambda = (2 pi f)^2) that the reference computes in FEM3D.eigenfrequency
 * (femder/FEM_3D.py:1699-1768: spsolve(Q, H) + scipy eigs, F_n = sqrt(lambda) / 2 pi, Vc = eigenvectors) and uses for modal
 * superposition (1842-1934).  fdb_modal_compute finds all modes below f_cut_hz on the GPU (Chebyshev-filtered subspace
 * iteration with a polynomial approximation of the inverse mass matrix, Rayleigh-Ritz with the consistent H, Q; csrc/modal.cu);
 * fdb_modal_set takes modes computed elsewhere.  Either makes precond = 2 available to fdb_sweep.  V is [n_modes][n_nodes],
 * Q-orthonormal, eigenvalues ascending.  At most 1024 modes are kept.
 *   f_acc_hz   the modes below it must meet `tol` (0 = all of them); those between f_acc_hz and f_cut_hz may be 10 x coarser
 *   tol        relative residual ||H v - lambda Q v|| / (lambda ||Q v||) to aim for (0 = default 4e-2: what the sweep's coarse
 *              correction needs; eigenvalues are then good to ~tol^2).  The polynomial mass approximation bounds what can be
 *              reached on coarse meshes; fdb_modal_info reports what was.
 *   n_modes_hint  expected number of modes below f_cut_hz (sizes the block; 0 = find out by widening) */
int fdb_modal_compute(fdb_system* sys, double f_cut_hz, double f_acc_hz, int n_modes_hint, double tol, int max_outer, int* n_modes_found);
/* The form in which fdb_sweep streams the modes.  on = 1 (the default after fdb_modal_compute / fdb_modal_set): tile-compressed -
 * per 128-node tile V_t ~ Phi_t C_t with 16 basis functions, ~0.55 bytes per node and mode and iteration instead of 4, same iteration
 * counts (csrc/modal.cu "Tile-compressed modes") - unless the mesh is so coarse that 16 functions per tile miss a mode by more than
 * 1.5 % (the tiles are then not small against the shortest wavelength), in which case the full-resolution tiles stay in use.
 * on = 2: compressed regardless of the fit.  on = 0: the full-resolution bf16 tiles. */
int fdb_modal_compression(fdb_system* sys, int on);
/* Several GPUs, one process each, sweeping shares of the same frequency list over the same room (the reference's frequency loop,
 * femder/FEM_3D.py:1304-1319, has independent iterations): they all need the same modes, so they can share the work of finding them.
 * One rank creates an id (fdb_comm_unique_id, 128 bytes) and hands it to the others by whatever means the host program has; every
 * rank then calls fdb_comm_init on its own system (same mesh and matrices on all of them) and, collectively, fdb_modal_compute with
 * the same arguments: the block's columns are dealt to the ranks and exchanged over NCCL (NVLink), every rank ends with all the
 * modes.  NCCL (libnccl.so.2) is loaded on first use.  fdb_sweep itself never communicates. */
int fdb_comm_unique_id(char* out128);
int fdb_comm_init(fdb_system* sys, const char* id128, int rank, int world);
int fdb_comm_free(fdb_system* sys);
/* Weyl's estimate of the number of room modes below f_hz from the assembled room's volume (total mass) and boundary area; 0 when the
 * matrices were given from outside.  fdb_modal_compute uses it when n_modes_hint is 0. */
int fdb_modal_estimate(fdb_system* sys, double f_hz, int* n_modes);
/* Diagnostics of the stored modes: outer iterations / block width / order of the mass approximation of the last fdb_modal_compute,
 * its worst relative residual, and per mode ||H v - lambda Q v|| / ||Q v|| (what fdb_sweep bounds 1 / (lambda - w^2) with).  Any pointer may be NULL. */
int fdb_modal_info(fdb_system* sys, int* outer_iterations, int* block_columns, int* mass_order, double* worst_residual, double* residuals /* [n_modes] */);
int fdb_modal_set(fdb_system* sys, int n_modes, const double* lam /* [n_modes] */, const double* V /* [n_modes][n_nodes] */);
int fdb_modal_count(fdb_system* sys, int* n_modes);
int fdb_modal_get(fdb_system* sys, double* lam /* [n_modes] or NULL */, double* V /* [n_modes][n_nodes] or NULL */);
/* The modal correction alone on caller data, for parity tests: z = V diag(1 / (lam - omega[f]^2)) V^T r in the kernels'
 * arithmetic (bf16 V, (hi, lo) bf16 operands, FP32 accumulation); r, z are [n_nodes][16] real. */
int fdb_modal_apply(fdb_system* sys, const double* omega /* [16] */, int n_modes_use, const double* r, double* z);

/* y = G_f x for a batch of frequencies on caller data: x, y are [n_nodes][batch] complex. (K5 alone; parity) */
int fdb_spmv(fdb_system* sys, int batch, const double* omega, const double* mu /* [n_groups][batch] complex */,
             const double* x, double* y);
/* Time `reps` launches of K5 on resident synthetic vectors; returns mean milliseconds per launch.
 * flags: bit 0 = include the boundary (admittance) overlay; bit 1 = real-arithmetic variant (what a rigid-wall
 * sweep with real sources launches; excludes bit 0). */
#define FDB_SPMV_BOUNDARY 1
#define FDB_SPMV_REAL 2
int fdb_spmv_bench(fdb_system* sys, int batch, int reps, int flags, double* ms_per_launch, int64_t* algorithmic_bytes);
/* Enqueue `reps` launches of K5 (with its dot-product epilogue, as the solver launches it) on the system's
 * stream and return without synchronising, so the caller can bracket them with its own CUDA events.
 * Call fdb_spmv_bench once before to fill the resident vectors. */
int fdb_spmv_enqueue(fdb_system* sys, int batch, int reps, int flags, int64_t* algorithmic_bytes);
/* The three kernels of one solver iteration, one at a time, for per-kernel rooflines (bench.py): enqueue `reps` launches
 * of kernel `which` on resident synthetic vectors without synchronising.  which = -1 prepares the vectors (call it first),
 * 0 = K5 (fused SpMV with its dot-product epilogue), 1 = K6a (residual update + preconditioner + reductions),
 * 2 = K6b (solution and search-direction update); with FDB_ITER_MODAL also 3 = modal projection c = V^T r, 4 = modal expansion
 * z += V e (+ rho', ||r||^2), 5 = modal coefficients.  flags: FDB_SPMV_REAL, FDB_ITER_BLOCK_JACOBI, FDB_ITER_MODAL (all stored
 * modes applied; 16 real slots; rigid-wall systems only).  algorithmic_bytes = what one launch of that kernel has to move. */
#define FDB_ITER_BLOCK_JACOBI 4
#define FDB_ITER_MODAL 8
int fdb_iteration_enqueue(fdb_system* sys, int batch, int reps, int flags, int which, int64_t* algorithmic_bytes);

#ifdef __cplusplus
}
#endif
#endif /* FEMDER_B200_H */

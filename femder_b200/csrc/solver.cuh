// Declarations shared by sweep.cu (K5/K6 kernels, sweep driver) and modal.cu (modal coarse correction, modal set-up).
#pragma once
#include <cstdlib>
#include "common.cuh"
#include "tc.cuh"
#include "../../include/femder_b200.h"
#include <type_traits>

namespace fdb {

constexpr int kMaxF = 64;   // widest batch: small meshes are launch-bound, wide batches amortise the fixed costs
constexpr int kBlock = 256;

struct __align__(16) Scalars {
    double2 rho[kMaxF];    // r^T z (unconjugated)
    double2 pq[kMaxF];     // p^T G p
    double2 beta[kMaxF];
    double2 alpha[kMaxF];  // step length of the iteration whose x update is still pending (applied by k_pupdate)
    double rr[kMaxF];      // ||r||^2
    double bb[kMaxF];      // ||b||^2
    double w2[kMaxF];      // omega^2
    double2 rhs_scale[kMaxF];  // -1j*omega
    double rt[kMaxF];      // ||b - G x||^2 of the returned solution (true residual check)
    int active[kMaxF];
    int iters[kMaxF];
    int reset[kMaxF];      // slots being (re)loaded with a new frequency: set-up kernels touch only these
    int xpend[kMaxF];      // the slot took a step in the last k_update: k_pupdate owes it x += alpha p
    int src_set[kMaxF];    // which source set (right-hand side) the slot solves: multi-RHS sweeps share G(w)
    int keepx[kMaxF];      // the slot is RESUMED, not loaded: x and the iteration count are kept, r = b - G x (run_sweep)
    int bb_from_r0;        // ||b||^2 is taken from the start residual (general right-hand sides, zero starting guess)
    int modes_use;         // modal coarse correction: modes applied by the current launches (a multiple of 16)
    double tol2;
    unsigned int ticket[4];
};

// Chebyshev-filter epilogue of k_spmv_blob (modal set-up): y[row] = c1 (G x)[row] d[row] + c2 x[row] + c3 xold[row]
struct ChebEpilogue {
    const double* d = nullptr;      // [n_rows] 1 / (lumped mass); nullptr = plain SpMV
    const double* xold = nullptr;   // [n_rows][16] previous iterate (read by the owner of the row only; y may alias it)
    double c1 = 0.0, c2 = 0.0, c3 = 0.0;
    bool f32 = false;               // the vectors hold 32 single-precision columns per row (same 128 bytes)
};
struct SweepWork {
    int F = 0;
    int64_t N = 0;
    int n_groups = 0;
    int grid = 0;
    DevBuf<double2> x, r, p, q, dinv, tmp;  // [N][F]
    DevBuf<double2> coef;              // [n_groups][F]  1j*omega*mu
    DevBuf<double2> rcoef;             // [n_rhs_vec][F] coefficients of the extra right-hand-side vectors
    DevBuf<int32_t> src_set_ptr;       // [n_sets+1] partition of the (deduplicated) source arrays into right-hand sides
    int n_rn = 0;                      // nodes touched by extra right-hand-side vectors (complex sweeps only)
    DevBuf<int32_t> rn_nodes, rn_ptr, rn_vec;   // per listed node: its (vector, value) terms
    DevBuf<double> rn_val;
    DevBuf<double2> partial;           // [grid][F][2]
    DevBuf<Scalars> sc;
    DevBuf<double2> stage;             // [F][N] pN export staging
    DevBuf<double2> rcv_out;           // [F][n_rcv]
    cudaGraphExec_t graph = nullptr;
    int graph_iters = 0;
    int graph_F = 0;
    bool graph_overlay = false;
    int spmv_cfg = 0;             // FDB_SPMV_CFG (tuning aid): 0 default, 1..3 other bulk-copy pipeline shapes, >= 8 plain kernel only
    int tma_stage_entries[3] = {0, 0, 0};
    int spmv_bps = 0;             // FDB_SPMV_BPS: cap on resident blocks per SM (tuning aid)
    int blob_lpr = 2;             // FDB_BLOB_LPR: lanes per row of k_spmv_blob (tuning aid)
    int precond = 0;              // 0 Jacobi, 1 block-Jacobi (16 real columns per row at most, 128-row tiles), 2 block-Jacobi + modal
    int graph_precond = 0;
    DevBuf<float> c_part;         // modal: per-block partial projections [grid][kModalMaxModes][16]
    DevBuf<double> tcv;           // modal, tile-compressed: t = Phi^T r [16 n_tiles (padded to 128)][16]
    DevBuf<float> uc;             // modal, tile-compressed: u = C e  [16 n_tiles (padded to 128)][16]
    DevBuf<uint16_t> e_img;       // modal: coefficients e = D V^T r as the bf16 (hi, lo) operand image of the expansion
    ChebEpilogue cheb;            // modal set-up: the SpMV writes a Chebyshev filter step (see k_spmv_blob)
    bool fused_precond = true;    // modal: K6a + projection as one tensor-core kernel (k_precond_fused); false = k_update_bj + k_modal_project
    DevBuf<float> z32;            // [N][F] preconditioned residual of the block-Jacobi path
    bool rows_kernel = true;      // FDB_NO_ROWS_KERNEL: narrow real batches fall back to the older kernels (A/B aid)
    int blob_dbg = 0;             // FDB_BLOB_DBG: timing experiments (1 no compute, 2 no x fill, 4 no matrix stream) - results are wrong
    bool real = false;            // the current sweep runs in real arithmetic (rigid walls, real sources)
    bool graph_real = false;
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cdiv(double2 a, double2 b) {
    const double d = b.x * b.x + b.y * b.y;
    return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

// ---- value type of the Krylov vectors: double2 (complex sweep) or double (rigid-wall sweep, see run_sweep) ----
// A rigid-wall system G = H - w^2 Q is real and b = -1j*w*q is purely imaginary for real q, so p = 1j*u with
// G u = -w q: the whole solve runs in real arithmetic on half the bytes.  Scalars (alpha, beta, rho) stay double2.
__device__ __forceinline__ double2 s_mul(double2 s, double2 v) { return cmul(s, v); }
__device__ __forceinline__ double s_mul(double2 s, double v) { return s.x * v; }
__device__ __forceinline__ double2 r_mul(double g, double2 v) { return make_double2(g * v.x, g * v.y); }
__device__ __forceinline__ double r_mul(double g, double v) { return g * v; }
__device__ __forceinline__ double2 v_mul(double2 a, double2 b) { return cmul(a, b); }
__device__ __forceinline__ double v_mul(double a, double b) { return a * b; }
__device__ __forceinline__ double2 v_dot(double2 a, double2 b) { return cmul(a, b); }   // unconjugated
__device__ __forceinline__ double2 v_dot(double a, double b) { return make_double2(a * b, 0.0); }
__device__ __forceinline__ double v_n2(double2 a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ double v_n2(double a) { return a * a; }
__device__ __forceinline__ double2 v_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double v_add(double a, double b) { return a + b; }
__device__ __forceinline__ double2 v_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double v_sub(double a, double b) { return a - b; }
template <typename T> __device__ __forceinline__ T v_zero();
template <> __device__ __forceinline__ double2 v_zero<double2>() { return make_double2(0.0, 0.0); }
template <> __device__ __forceinline__ double v_zero<double>() { return 0.0; }
// what leaves the solver is always complex: p = u (complex sweep) or p = 1j*u (real sweep)
__device__ __forceinline__ double2 v_out(double2 a) { return a; }
__device__ __forceinline__ double2 v_out(double a) { return make_double2(0.0, a); }

// Sum `v` over the lanes that share lane % F, result valid in lanes 0..F-1.
template <int F>
__device__ __forceinline__ double warp_sum_by_f(double v) {
#pragma unroll
    for (int o = 16; o >= F; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// The last block to arrive adds the per-block partials partial[b][f][slot] of all blocks in a fixed order
// (deterministic) and calls `fin(f, totals)` from one thread per f.  Call with the block's partials already
// written by threads that then reach this point (all threads of the block must call it).
template <int F, int NV, int BLOCK, typename Fin>
__device__ __forceinline__ void finalize_last_block(double* __restrict__ partial, unsigned int* __restrict__ ticket, Fin fin) {
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // each f gets BLOCK/F threads striding over the blocks, then a fixed-order sum over those threads
    constexpr int TPF = BLOCK / F;
    const int f = threadIdx.x % F, j = threadIdx.x / F;
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0;
    if (j < TPF) {
        // U independent loads in flight per round, added in block order: the chain of gridDim.x / TPF dependent L2 round
        // trips was the tail of every reducing kernel (30 us of k_update's 48 us on a 21 735-DOF mesh at 64 slots)
        constexpr int U = 8;
        for (int b = j; b < (int)gridDim.x; b += TPF * U) {
            double t[U][NV];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int bb = b + u * TPF;
#pragma unroll
                for (int i = 0; i < NV; ++i) t[u][i] = (bb < (int)gridDim.x) ? __ldcg(&partial[((int64_t)bb * F + f) * NV + i]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int i = 0; i < NV; ++i) acc[i] += t[u][i];
            }
        }
    }
    __shared__ double s_fin[BLOCK][NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) s_fin[threadIdx.x][i] = acc[i];
    __syncthreads();
    if (threadIdx.x < F) {
        double tot[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) tot[i] = 0;
        for (int k = 0; k < TPF; ++k) {
#pragma unroll
            for (int i = 0; i < NV; ++i) tot[i] += s_fin[k * F + threadIdx.x][i];
        }
        fin(threadIdx.x, tot);
        if (threadIdx.x == 0) *ticket = 0;
    }
}

// lanes = (.., f): per-thread values of frequency f = threadIdx.x % F -> block partial -> finalize
template <int F, int NV, int BLOCK, typename Fin>
__device__ __forceinline__ void block_reduce_finalize(const double (&v)[NV], double* __restrict__ partial,
                                                      unsigned int* __restrict__ ticket, Fin fin) {
    __shared__ double s_red[BLOCK / 32][32][NV];
    constexpr int SW = F < 32 ? F : 32;   // distinct frequencies inside one warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double w[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) w[i] = warp_sum_by_f<F>(v[i]);
    if (lane < SW) {
#pragma unroll
        for (int i = 0; i < NV; ++i) s_red[warp][lane][i] = w[i];
    }
    __syncthreads();
    if (threadIdx.x < F) {
        const int l = threadIdx.x & 31;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0;
            for (int k = 0; k < BLOCK / 32; ++k)
                if ((k * 32 + l) % F == (int)threadIdx.x) s += s_red[k][l][i];   // warps whose lane l carries this frequency
            partial[((int64_t)blockIdx.x * F + threadIdx.x) * NV + i] = s;
        }
    }
    finalize_last_block<F, NV, BLOCK>(partial, ticket, fin);
}

// last-block finalisation of K6a: beta = rho'/rho, rho = rho', convergence flag, iteration count, alpha kept for K6b
__device__ __forceinline__ void fin_update(Scalars* __restrict__ sc, int ff, const double* tot) {
    if (sc->active[ff]) {
        const double2 rho_old = sc->rho[ff];
        const double2 rho_new = make_double2(tot[0], tot[1]);
        const double2 pq = sc->pq[ff];
        sc->alpha[ff] = (pq.x != 0.0 || pq.y != 0.0) ? cdiv(rho_old, pq) : make_double2(0.0, 0.0);   // what every block used
        sc->xpend[ff] = 1;
        double2 beta = make_double2(0.0, 0.0);
        if (rho_old.x != 0.0 || rho_old.y != 0.0) beta = cdiv(rho_new, rho_old);
        sc->rho[ff] = rho_new;
        sc->rr[ff] = tot[2];
        sc->iters[ff] += 1;
        const bool conv = !(tot[2] > sc->tol2 * sc->bb[ff]);  // also stops on NaN
        if (conv) { sc->active[ff] = 0; beta = make_double2(0.0, 0.0); }
        sc->beta[ff] = beta;
    } else {
        sc->beta[ff] = make_double2(0.0, 0.0);
        sc->xpend[ff] = 0;
    }
}
// ... of the slot (re)load: rho = r^T z, activate the slot if its residual is above tolerance
__device__ __forceinline__ void fin_start(Scalars* __restrict__ sc, int ff, const double* tot) {
    if (!sc->reset[ff]) return;
    sc->rho[ff] = make_double2(tot[0], tot[1]);
    sc->rr[ff] = tot[2];
    sc->beta[ff] = make_double2(0.0, 0.0);
    sc->pq[ff] = make_double2(0.0, 0.0);
    if (!sc->keepx[ff]) {               // a resumed slot keeps its count and its ||b||
        sc->iters[ff] = 0;
        if (sc->bb_from_r0) sc->bb[ff] = tot[2];
    }
    sc->active[ff] = (tot[2] > sc->tol2 * sc->bb[ff]) ? 1 : 0;
    sc->reset[ff] = 0;
}


// ---- modal.cu: the modal coarse correction inside a solver iteration ----
constexpr int kModalCols = 16;   // real columns per row the tensor-core kernels are built for (16 real or 8 complex slots)
void modal_prepare(fdb_system* sys, SweepWork* w);
void pupdate_tiles(fdb_system* sys, SweepWork* w);   // K6b on 128-byte vector rows (16 real / 8 complex slots), bulk copies both ways
void modal_set_compression(fdb_system* sys, bool on, bool force = false);   // force: even where the fit is poor (tests)
int modal_modes_for(fdb_system* sys, double w2_max, double delta_hz);
// parts: 1 c = V^T r ; 2 e = d c ; 4 z += V e, rho', ||r||^2, finalisation.  Tile-compressed modes: 1 c = C^T t (t from
// modal_precond_fused) ; 2 e = d c ; 4 u = C e ; 8 z += Phi u, rho', ||r||^2, finalisation
void modal_apply(fdb_system* sys, SweepWork* w, bool overlay, bool start, int parts = 15);
int64_t modal_bytes(fdb_system* sys, SweepWork* w, int modes_use, int part);
void modal_precond_fused(fdb_system* sys, SweepWork* w, bool start);   // r -= alpha q ; z = B r ; c = V^T r  (one tcgen05 kernel)
int64_t modal_fused_bytes(fdb_system* sys, SweepWork* w, int modes_use);
void modal_build_adiag(fdb_system* sys);
void modal_set(fdb_system* sys, int n_modes, const double* lam, const double* V);
void modal_get(fdb_system* sys, double* lam, double* V);
void modal_finish(fdb_system* sys, int n_modes);
int modal_estimate(fdb_system* sys, double f_hz);
void modal_compute(fdb_system* sys, double f_cut_hz, double f_acc_hz, int n_hint, double tol, int max_outer, int* n_found);
// ---- sweep.cu: the fused SpMV on 16 real columns for the modal set-up (y = (H - w2 Q) x, one w2 for all columns) ----
SweepWork* modal_spmv_begin(fdb_system* sys);
void modal_spmv_w2(fdb_system* sys, SweepWork* w, double w2);
void modal_spmv(fdb_system* sys, SweepWork* w, const double* x, double* y);
bool modal_spmv_cheb(fdb_system* sys, SweepWork* w, const double* x, const double* xold, double* out, const double* dinv, double c1, double c2,
                     double c3, bool f32 = false);
void run_modal_apply(fdb_system* sys, const double* omega, int n_use, const double* r, double* z);

// A/B knobs of the solver and of the modal set-up are environment variables that exist only in builds made with -DFDB_TUNING
// (`FDB_BUILD_TUNING=1 python -m femder_b200.build` -> femder_b200/_ab/libfemder_b200_tuning.so, loaded with FDB_LIB_PATH): a
// stray variable cannot change what the shipped library computes.
#ifdef FDB_TUNING
inline const char* tune_env(const char* name) { return getenv(name); }
#else
inline const char* tune_env(const char*) { return nullptr; }
#endif

// ---- comm.cu: the ranks that share one modal set-up (NCCL, opened lazily) ----
struct RankComm;
void comm_unique_id(char* out128);
RankComm* comm_create(const char* id128, int rank, int world);
void comm_destroy(RankComm* c);
int comm_rank(const RankComm* c);
int comm_world(const RankComm* c);
void comm_allgatherv(RankComm* c, double* buf, const std::vector<size_t>& offs, cudaStream_t s);   // in place; rank r holds [offs[r], offs[r+1])
void comm_bcast(RankComm* c, double* buf, size_t n, int root, cudaStream_t s);
void comm_allreduce_sum(RankComm* c, double* buf, size_t n, cudaStream_t s);

}  // namespace fdb

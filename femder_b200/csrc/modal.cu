// Low room modes as a coarse correction of the sweep's preconditioner, on the tensor cores.
//
// The reference computes room modes in FEM3D.eigenfrequency (femder/FEM_3D.py:1699-1768: spsolve(Q, H) + eigs) and uses them for
// modal superposition (1842-1934).  Here the same eigenpairs  H v_i = lambda_i Q v_i,  V^T Q V = I  make the Krylov solver
// of FEM3D.compute() (sweep.cu) independent of how many room modes lie below the drive frequency:
//
//     M^-1 = B + V diag(d) V^T ,   d_i = 1 / (lambda_i - w^2 + 1j w sum_g mu_g v_i^T A_g v_i)
//
// B = block-Jacobi (sweep.cu).  G = H - w^2 Q has one negative eigenvalue per mode below the drive frequency and a
// near-zero one at every resonance - that, not the mesh, is what makes COCG need thousands of iterations; V diag(d) V^T
// inverts G exactly on span(V), so with all modes below sqrt(f^2 + D^2) in V the preconditioned operator is positive
// with its spectrum bounded away from zero (tests/research/modal_deflation_prototype.py: 1879 -> 37 iterations at 300 Hz on
// the config-1 mesh; bfloat16 V and modes of the lumped-mass problem do as well as exact FP64 modes).
//
// The products with V are the dense contractions of the hot path, and they run on tcgen05:
//   projection  c = V^T r     (k_modal_project)   D[128 modes x 32] += A[128 modes x 16 nodes] B[16 nodes x 32]
//   expansion   z += V e      (k_modal_expand)    D[128 nodes x 32] += A[128 nodes x 16 modes] B[16 modes x 32]
// Both read the SAME bytes of V: a tile of 128 nodes x 128 modes is stored as 8 x 8 bf16 core matrices, the expansion reads
// it K-major, the projection MN-major (tc.cuh).  The 32 operand columns are the batch's 16 real columns (16 real slots or
// 8 complex ones) split into bf16 (hi, lo) pairs: 16 mantissa bits of the FP32 value, accumulated in FP32 in TMEM.
// By default V is not streamed at all: per 128-node tile V_t ~ Phi_t C_t with 16 basis functions ("Tile-compressed modes"
// below), t = Phi^T r is one more MMA group inside the fused K6a, the two kernels above work on the coefficient tiles C and a
// vector of 16 numbers per tile, and z += Phi u is k_coarse_finish.  Everything here is bound by its HBM stream: the tensor pipe
// is ~5 % busy.  This file also holds K6b (k_pupdate_tiles), the modal set-up (Chebyshev-filtered subspace iteration) and the
// tile-basis kernel.
#include <cublas_v2.h>
#include <cusolverDn.h>
#include "solver.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <ctime>

namespace fdb {

constexpr int kMC = 128;                       // modes per chunk = rows of one projection accumulator
constexpr int kModalMaxChunks = 8;             // 1024 modes: 8 x 32 TMEM columns in the projection, 64 KB coefficient image
constexpr int kModalMaxModes = kMC * kModalMaxChunks;
constexpr int kChunkBytes = 128 * kMC * 2;     // one tile of V for one chunk of modes: 32 KB
constexpr int kStages = 4;                     // ring of V chunks in shared memory
constexpr int kModalThreads = 192;             // warp 0 producer, warp 1 MMA issue + TMEM owner, warps 2..5 operand / epilogue
constexpr int kFusedThreads = 64 + 2 * 128;    // k_precond_fused: two groups of operand / epilogue warps
constexpr uint32_t kSm = 2048, kSn = 128;      // byte strides of a chunk: mode block (8 modes), node block (8 nodes)
constexpr int kCF = 16;                        // functions of a tile basis (tile-compressed modes)
constexpr int kPhiBytes = 128 * kCF * 2;       // one tile basis in the chunk layout: the first two "mode blocks", 4 KB
constexpr double kMaxFitError = 1.5e-2;        // worst relative fit error of a mode above which the tile-compressed form is not used
// k_modal_coef keeps |lambda - w^2| >= res^2 / lambda (res = the stored vector's residual): below that the inexact vector cannot
// tell which side of the resonance the drive frequency is on.  Nothing coarser: a bound proportional to res itself was tried
// and makes V diag(d) V^T stop being the inverse of G on span(V), which the indefinite CG recurrence does not survive
// (measured: occasional stagnation with the bound, none without, on modes with 5e-3 relative residual).

void modal_reset(fdb_system* sys) {
    sys->have_modes = false;
    sys->have_adiag = false;
    sys->n_modes = 0;
}

// ------------------------------------------------------------------------------------------------
// storage: FP64 modes in batches [b][row][16] (internal order) -> bf16 tiles
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_modal_pack(const double* __restrict__ X, int64_t n_rows, int64_t n_tiles, int n_modes, int n_chunks,
                                                    uint16_t* __restrict__ V) {
    // one thread per (row of the padded tiles, block of 8 modes); rows fastest: 8 consecutive threads write one 128-byte core matrix
    const int64_t rows_pad = n_tiles * 128;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows_pad * n_chunks * 16) return;
    const int64_t row = idx % rows_pad;
    const int mbg = (int)(idx / rows_pad);          // global mode block
    const int j = mbg >> 4, mb = mbg & 15;
    const int64_t t = row >> 7;
    const int rin = (int)(row & 127), nb = rin >> 3, ni = rin & 7;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (row < n_rows) {
        const int m0 = mbg * 8;
        const double* src = X + ((int64_t)(m0 >> 4) * n_rows + row) * 16 + (m0 & 15);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t h = (m0 + k < n_modes) ? bf16_rn((float)src[k]) : 0u;
            w[k >> 1] |= h << (16 * (k & 1));
        }
    }
    uint16_t* dst = V + (((int64_t)t * n_chunks + j) * (kChunkBytes / 2)) + (mb * kSm + nb * kSn + ni * 16) / 2;
    *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
}

// X[i / 16][perm[n]][i % 16] = v[n]  (one mode from the caller's numbering into the batch layout)
__global__ void k_modal_scatter(const double* __restrict__ v, const int32_t* __restrict__ perm, int64_t n_rows, int mode, double* __restrict__ X) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < n_rows) X[((int64_t)(mode >> 4) * n_rows + perm[n]) * 16 + (mode & 15)] = v[n];
}
__global__ void k_modal_gather(const double* __restrict__ X, const int32_t* __restrict__ perm, int64_t n_rows, int mode, double* __restrict__ v) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < n_rows) v[n] = X[((int64_t)(mode >> 4) * n_rows + perm[n]) * 16 + (mode & 15)];
}

// adiag[g][i] = v_i^T A_g v_i from the boundary overlay (rows in the internal order): one block per mode, fixed-order sums
__global__ void __launch_bounds__(256) k_modal_adiag(const double* __restrict__ X, const int32_t* __restrict__ ov_ptr, const int2* __restrict__ ov_cg,
                                                     const double* __restrict__ ov_val, int64_t n_rows, int n_groups, int n_modes,
                                                     double* __restrict__ adiag) {
    const int i = blockIdx.x;
    const double* x = X + (int64_t)(i >> 4) * n_rows * 16 + (i & 15);
    extern __shared__ double s_acc[];   // [256][n_groups]
    for (int g = 0; g < n_groups; ++g) s_acc[threadIdx.x * n_groups + g] = 0.0;
    for (int64_t r = threadIdx.x; r < n_rows; r += 256) {
        const int ob = ov_ptr[r], oe = ov_ptr[r + 1];
        if (ob == oe) continue;
        const double xr = x[r * 16];
        for (int o = ob; o < oe; ++o) {
            const int2 cg = ov_cg[o];
            s_acc[threadIdx.x * n_groups + cg.y] += xr * ov_val[o] * x[(int64_t)cg.x * 16];
        }
    }
    __syncthreads();
    if (threadIdx.x < n_groups) {
        double s = 0.0;
        for (int k = 0; k < 256; ++k) s += s_acc[k * n_groups + threadIdx.x];
        adiag[(int64_t)threadIdx.x * n_modes + i] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// projection  c = V^T r :  per block, partial sums over its tiles in TMEM; c_part[block][mode][16]
// ------------------------------------------------------------------------------------------------
struct ModalSmem {
    uint64_t a_full[kStages], a_empty[kStages];
    uint64_t b_full[2], b_empty[2];
    uint64_t e_full, done;
    uint32_t tmem_base;
};

// chunks and the bytes / K steps of the last one for `modes_use` modes (a multiple of 16)
__device__ __forceinline__ void modal_shape(int modes_use, int& n_use, uint32_t& bytes_last, int& ksteps_last) {
    n_use = (modes_use + kMC - 1) / kMC;
    const int last = modes_use - (n_use - 1) * kMC;          // modes in the last chunk
    bytes_last = (uint32_t)((last + 7) / 8) * kSm;
    ksteps_last = last / 16;
}

// the V chunks of this block's tiles, in order, into the ring (one thread)
__device__ __forceinline__ void modal_produce(const uint16_t* __restrict__ V, int n_chunks, int64_t n_tiles, int n_use, uint32_t bytes_last,
                                              unsigned char* ring, ModalSmem* sm) {
    uint32_t it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int j = 0; j < n_use; ++j, ++it) {
            const int st = it % kStages;
            mbar_wait_bounded(&sm->a_empty[st], ((it / kStages) & 1u) ^ 1u);
            const uint32_t bytes = (j == n_use - 1) ? bytes_last : (uint32_t)kChunkBytes;
            mbar_expect_tx(&sm->a_full[st], bytes);
            bulk_g2s(ring + (size_t)st * kChunkBytes, V + ((int64_t)t * n_chunks + j) * (kChunkBytes / 2), bytes, &sm->a_full[st]);
        }
    }
}

__global__ void __launch_bounds__(kModalThreads, 1)
k_modal_project(const double* __restrict__ r, const uint16_t* __restrict__ V, int n_chunks, int64_t n_rows, int64_t n_tiles,
                const Scalars* __restrict__ sc, float* __restrict__ c_part) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* ring = smem_raw;                                      // [kStages][32 KB]
    unsigned char* bimg = smem_raw + (size_t)kStages * kChunkBytes;      // [2][8 KB]: (hi, lo) image of a residual tile
    ModalSmem* sm = reinterpret_cast<ModalSmem*>(bimg + 2 * 8192);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int n_use, ksteps_last;
    uint32_t bytes_last;
    modal_shape(sc->modes_use, n_use, bytes_last, ksteps_last);
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&sm->a_full[i], 1); mbar_init(&sm->a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&sm->b_full[i], 128); mbar_init(&sm->b_empty[i], 1); }
        mbar_init(&sm->done, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&sm->tmem_base, 32 * kModalMaxChunks);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm->tmem_base;

    if (warp == 0) {
        if (lane == 0) modal_produce(V, n_chunks, n_tiles, n_use, bytes_last, ring, sm);
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc_idesc_bf16(128, 32, 1, 1);   // A = V^T chunk, MN-major; B = residual image, MN-major
            uint32_t it = 0, ti = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
                const int bb = ti & 1;
                mbar_wait_bounded(&sm->b_full[bb], (ti >> 1) & 1u);
                tc_fence_after();
                const uint32_t b_addr = smem_u32(bimg + bb * 8192);
                for (int j = 0; j < n_use; ++j, ++it) {
                    const int st = it % kStages;
                    mbar_wait_bounded(&sm->a_full[st], (it / kStages) & 1u);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(ring + (size_t)st * kChunkBytes);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)   // K = 16 nodes per step: two node blocks
                        tc_mma_bf16(tmem + 32 * j, tc_desc(a_addr + ks * 2 * kSn, kSn, kSm), tc_desc(b_addr + ks * 256, 128, 2048), idesc,
                                    ti > 0 || ks > 0);
                    tc_commit(&sm->a_empty[st]);
                }
                tc_commit(&sm->b_empty[bb]);
            }
            tc_commit(&sm->done);
        }
    } else {
        // ---- residual tile -> bf16 (hi, lo) operand image: thread = node row ----
        const int i = tid - 64;
        uint32_t ti = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
            const int bb = ti & 1;
            const int64_t row = t * 128 + i;
            double v[16];
            if (row < n_rows) {
                const double4* src = reinterpret_cast<const double4*>(r + row * 16);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double4 d = src[k];
                    v[4 * k] = d.x; v[4 * k + 1] = d.y; v[4 * k + 2] = d.z; v[4 * k + 3] = d.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = 0.0;
            }
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t h0, l0, h1, l1;
                bf16_split((float)v[2 * k], h0, l0);
                bf16_split((float)v[2 * k + 1], h1, l1);
                hi[k] = h0 | (h1 << 16);
                lo[k] = l0 | (l1 << 16);
            }
            mbar_wait_bounded(&sm->b_empty[bb], ((ti >> 1) & 1u) ^ 1u);
            unsigned char* img = bimg + bb * 8192 + i * 16;   // core (node block, column block) at column block * 2048 + node block * 128
            *reinterpret_cast<uint4*>(img) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(img + 2048) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            *reinterpret_cast<uint4*>(img + 4096) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4*>(img + 6144) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            fence_proxy_async();
            mbar_arrive(&sm->b_full[bb]);
        }
        // ---- epilogue: this block's partial sums, thread = mode row of a chunk ----
        mbar_wait_bounded(&sm->done, 0);
        tc_fence_after();
        const int q = warp & 3;   // the TMEM lanes a warp may read: 32 (warp % 4) .. + 31
        float* out = c_part + (int64_t)blockIdx.x * kModalMaxModes * 16;
        for (int j = 0; j < n_use; ++j) {
            float acc[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 32 * j, acc);
            float4* o = reinterpret_cast<float4*>(out + (int64_t)(j * kMC + q * 32 + lane) * 16);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                o[k] = make_float4(acc[4 * k] + acc[16 + 4 * k], acc[4 * k + 1] + acc[17 + 4 * k], acc[4 * k + 2] + acc[18 + 4 * k],
                                   acc[4 * k + 3] + acc[19 + 4 * k]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 32 * kModalMaxChunks);
}

// ------------------------------------------------------------------------------------------------
// K6a of the modal path, fused with the projection: per 128-row tile
//     r <- r - alpha q (FP64, stored)          z_bj = B_t r  (tcgen05: B_t = bf16 inverse of the tile's diagonal block)
//     c += V_t^T r      (tcgen05, partial sums of the block's tiles in TMEM as in k_modal_project)
// One pass over r, q and the tile's block inverse instead of k_update_bj's CUDA-core product (the dominant, issue-bound kernel
// of round 1) plus a second read of r by the projection.  The residual tile's bf16 (hi, lo) image is the B operand of BOTH
// products.  Everything arrives by bulk copies - [r | q] into two 32 KB buffers, B_t and the V chunks into a ring of 32 KB slots - and
// the updated residual leaves by a bulk store out of the buffer it arrived in.
// START: slot load - r as it stands (no q), only the columns flagged in sc.reset carry data.
// ------------------------------------------------------------------------------------------------
struct FusedSmem {
    uint64_t s_full[kStages], s_empty[kStages];   // ring of B_t and V chunks: filled by the producer, consumed IN ORDER by the MMA thread
    uint64_t rq_full[4], rq_empty[4];             // [r | q] tiles: consumed by the converter warps.  (A consumer that skipped phases of a
                                                  // shared ring could not tell them apart by parity: each consumer has its own barriers.)
    uint64_t b_full[2], b_empty[2];
    uint64_t z_full[2], z_empty[2];
    uint64_t phi_full[2], phi_empty[2];           // COARSE: the tile bases
    uint64_t done;
    uint32_t tmem_base;
    double2 alpha[16];
    int act[16];
};

// COARSE (tile-compressed modes, see "tile bases" below): instead of the V chunks the tile's basis Phi_t (128 nodes x 16 functions,
// 4 KB) is the second A operand: t = Phi_t^T r, 16 numbers per column, written to tc[tile][16][16] for the coarse projection.
// Phi_t sits at the start of a 32 KB operand extent whose remaining rows are whatever shared memory holds: they only feed
// accumulator rows nobody reads.
template <bool START, bool CPX, bool COARSE = false>
__global__ void __launch_bounds__(kFusedThreads, 1)
k_precond_fused(double* __restrict__ r, const double* __restrict__ q, const uint16_t* __restrict__ binv16, float* __restrict__ z32,
                const uint16_t* __restrict__ V, int n_chunks, int64_t n_rows, int64_t n_tiles, const Scalars* __restrict__ sc,
                float* __restrict__ c_part, const uint16_t* __restrict__ phi16, double* __restrict__ tc) {
    constexpr int F = CPX ? 8 : 16;
    constexpr uint32_t kZCol = 32 * kModalMaxChunks;   // TMEM columns: [0, 256) projection accumulators, [256, 320) two z_bj accumulators
    constexpr uint32_t kTCol = kZCol + 64;             // COARSE: [320, 384) two t accumulators
    // shared memory: six 32 KB buffers.  With V chunks in the ring it takes four of them and [r | q] is double-buffered; the COARSE
    // ring carries one item per tile (B_t), so two slots do and [r | q] gets four: three tiles of vectors in flight ahead of the
    // converter warps, whose per-tile work is one long dependent chain
    constexpr int RS = COARSE ? 2 : kStages;           // ring slots
    constexpr int RQ = COARSE ? 4 : 2;                 // [r | q] buffers
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* phis = smem_raw;                                      // COARSE: [2][4 KB] tile bases (reads run on into the ring)
    unsigned char* ring = smem_raw + (COARSE ? 2 * kPhiBytes : 0);       // [kStages][32 KB]
    unsigned char* rqbuf = ring + (size_t)RS * kChunkBytes;              // [RQ][32 KB]: r (16 KB) | q (16 KB)
    unsigned char* bimg = rqbuf + RQ * (size_t)kChunkBytes;              // [2][8 KB]
    FusedSmem* sm = reinterpret_cast<FusedSmem*>(bimg + 2 * 8192);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int n_use, ksteps_last;
    uint32_t bytes_last;
    modal_shape(sc->modes_use, n_use, bytes_last, ksteps_last);
    if (COARSE) n_use = 0;
    const uint32_t items = 1u + (uint32_t)n_use;   // ring items per tile: B_t, then the V chunks
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&sm->s_full[i], 1); mbar_init(&sm->s_empty[i], 1); }
        for (int i = 0; i < RQ; ++i) { mbar_init(&sm->rq_full[i], 1); mbar_init(&sm->rq_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm->b_full[i], 128); mbar_init(&sm->b_empty[i], 1);
            mbar_init(&sm->z_full[i], 1); mbar_init(&sm->z_empty[i], 128);
            mbar_init(&sm->phi_full[i], 1); mbar_init(&sm->phi_empty[i], 1);
        }
        mbar_init(&sm->done, 1);
        mbar_fence_init();
    }
    if (tid < F) {
        const bool a = START ? (sc->reset[tid] != 0) : (sc->active[tid] != 0);
        double2 al = make_double2(0.0, 0.0);
        if (!START && a) {
            const double2 pq = sc->pq[tid];
            if (pq.x != 0.0 || pq.y != 0.0) al = cdiv(sc->rho[tid], pq);
        }
        sm->act[tid] = a ? 1 : 0;
        sm->alpha[tid] = al;
    }
    if (warp == 1) tmem_alloc(&sm->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0, ti = 0;
            auto slot_for = [&](uint32_t i) { mbar_wait_bounded(&sm->s_empty[i % RS], ((i / RS) & 1u) ^ 1u); return (int)(i % RS); };
            auto load_rq = [&](int64_t t, uint32_t k) {   // [r | q] of tile t (the block's k-th) into buffer k % RQ
                const int bb = k % RQ;
                const uint32_t rows = (uint32_t)min((int64_t)128, n_rows - t * 128);
                mbar_wait_bounded(&sm->rq_empty[bb], ((k / RQ) & 1u) ^ 1u);
                unsigned char* dst = rqbuf + (size_t)bb * kChunkBytes;
                mbar_expect_tx(&sm->rq_full[bb], rows * 128u * (START ? 1u : 2u));
                bulk_g2s(dst, r + t * 128 * 16, rows * 128u, &sm->rq_full[bb]);
                if (!START) bulk_g2s(dst + 16384, q + t * 128 * 16, rows * 128u, &sm->rq_full[bb]);
            };
            int64_t t_pref = blockIdx.x;
            uint32_t k_pref = 0;
            auto prefetch_upto = [&](uint32_t k_limit) {   // the vectors of the block's tiles 0 .. k_limit
                while (k_pref <= k_limit && t_pref < n_tiles) { load_rq(t_pref, k_pref); t_pref += gridDim.x; ++k_pref; }
            };
            prefetch_upto(RQ - 2);
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
                prefetch_upto(ti + RQ - 1);   // the next tiles' vectors ahead of this tile's block inverse and mode chunks
                {   // B_t
                    const int st = slot_for(it++);
                    mbar_expect_tx(&sm->s_full[st], (uint32_t)kChunkBytes);
                    bulk_g2s(ring + (size_t)st * kChunkBytes, binv16 + t * (kChunkBytes / 2), (uint32_t)kChunkBytes, &sm->s_full[st]);
                }
                if constexpr (COARSE) {
                    const int pb = ti & 1;
                    mbar_wait_bounded(&sm->phi_empty[pb], ((ti >> 1) & 1u) ^ 1u);
                    mbar_expect_tx(&sm->phi_full[pb], (uint32_t)kPhiBytes);
                    bulk_g2s(phis + (size_t)pb * kPhiBytes, phi16 + t * (kPhiBytes / 2), (uint32_t)kPhiBytes, &sm->phi_full[pb]);
                }
                for (int j = 0; j < n_use; ++j) {
                    const int st = slot_for(it++);
                    const uint32_t bytes = (j == n_use - 1) ? bytes_last : (uint32_t)kChunkBytes;
                    mbar_expect_tx(&sm->s_full[st], bytes);
                    bulk_g2s(ring + (size_t)st * kChunkBytes, V + ((int64_t)t * n_chunks + j) * (kChunkBytes / 2), bytes, &sm->s_full[st]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_p = tc_idesc_bf16(128, 32, 1, 1);   // projection: A = V^T chunk (MN-major), B = residual image (MN-major)
            constexpr uint32_t idesc_z = tc_idesc_bf16(128, 32, 0, 1);   // block product: A = B_t (K-major), same B
            uint32_t ti = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
                const int bb = ti & 1;
                const uint32_t it0 = ti * items;
                mbar_wait_bounded(&sm->b_full[bb], (ti >> 1) & 1u);
                const uint32_t b_addr = smem_u32(bimg + bb * 8192);
                {   // z_bj = B_t r
                    const uint32_t it = it0;
                    const int st = it % RS;
                    mbar_wait_bounded(&sm->s_full[st], (it / RS) & 1u);
                    mbar_wait_bounded(&sm->z_empty[bb], ((ti >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(ring + (size_t)st * kChunkBytes);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        tc_mma_bf16(tmem + kZCol + 32 * bb, tc_desc(a_addr + ks * 2 * kSm, kSm, kSn), tc_desc(b_addr + ks * 256, 128, 2048), idesc_z, ks > 0);
                    tc_commit(&sm->s_empty[st]);
                    if constexpr (COARSE) {   // t = Phi_t^T r: rows 0..15 of the accumulator
                        mbar_wait_bounded(&sm->phi_full[bb], (ti >> 1) & 1u);
                        tc_fence_after();
                        const uint32_t p_addr = smem_u32(phis + (size_t)bb * kPhiBytes);
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
                            tc_mma_bf16(tmem + kTCol + 32 * bb, tc_desc(p_addr + ks * 2 * kSn, kSn, kSm), tc_desc(b_addr + ks * 256, 128, 2048), idesc_p, ks > 0);
                        tc_commit(&sm->phi_empty[bb]);
                    }
                    tc_commit(&sm->z_full[bb]);   // both accumulators of the tile
                }
                for (int j = 0; j < n_use; ++j) {
                    const uint32_t it = it0 + 1 + j;
                    const int st = it % RS;
                    mbar_wait_bounded(&sm->s_full[st], (it / RS) & 1u);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(ring + (size_t)st * kChunkBytes);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        tc_mma_bf16(tmem + 32 * j, tc_desc(a_addr + ks * 2 * kSn, kSn, kSm), tc_desc(b_addr + ks * 256, 128, 2048), idesc_p, ti > 0 || ks > 0);
                    tc_commit(&sm->s_empty[st]);
                }
                tc_commit(&sm->b_empty[bb]);
            }
            tc_commit(&sm->done);
        }
    } else {
        // Two groups of four converter / epilogue warps take alternate tiles (group = parity of the block's tile counter = the
        // operand image and accumulator pair the tile uses): a tile's conversion is one long dependent chain, and with a single
        // warp per scheduler every instruction's latency was exposed (the kernel ran at 0.78 of the HBM rate whatever was in flight).
        const int grp = (tid - 64) >> 7;
        const int i = (tid - 64) & 127;    // node row of the tile this thread converts
        const int qd = warp & 3;           // TMEM lanes this warp may read
        const int zrow = qd * 32 + lane;   // ... = the row whose z_bj it stores
        auto epilogue = [&](int64_t t, uint32_t ti) {
            const int bb = ti & 1;
            mbar_wait_bounded(&sm->z_full[bb], (ti >> 1) & 1u);
            tc_fence_after();
            float acc[32];
            tmem_ld32(tmem + ((uint32_t)(qd * 32) << 16) + kZCol + 32 * bb, acc);
            if constexpr (COARSE) {
                if (qd == 0) {   // lanes 0..15 of the accumulator = the 16 functions
                    float ta[32];
                    tmem_ld32(tmem + kTCol + 32 * bb, ta);
                    if (lane < kCF) {
                        double4* o = reinterpret_cast<double4*>(tc + (t * kCF + lane) * 16);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            o[k] = make_double4((double)ta[4 * k] + (double)ta[16 + 4 * k], (double)ta[4 * k + 1] + (double)ta[17 + 4 * k],
                                                (double)ta[4 * k + 2] + (double)ta[18 + 4 * k], (double)ta[4 * k + 3] + (double)ta[19 + 4 * k]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&sm->z_empty[bb]);
            const int64_t row = t * 128 + zrow;
            if (row < n_rows) {
                float4* zd = reinterpret_cast<float4*>(z32 + row * 16);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    zd[k] = make_float4(acc[4 * k] + acc[16 + 4 * k], acc[4 * k + 1] + acc[17 + 4 * k], acc[4 * k + 2] + acc[18 + 4 * k],
                                        acc[4 * k + 3] + acc[19 + 4 * k]);
            }
        };
        uint32_t ti = 0, ti_prev = 0;
        int64_t t_prev = -1;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
            const int bb = ti & 1;
            if (bb != grp) continue;
            const int64_t row = t * 128 + i;
            const bool valid = row < n_rows;
            const int rb = ti % RQ;
            mbar_wait_bounded(&sm->rq_full[rb], (ti / RQ) & 1u);
            double* rs = reinterpret_cast<double*>(rqbuf + (size_t)rb * kChunkBytes) + i * 16;
            const double* qs = rs + 2048;
            double v[16];
            // 16-byte chunks of the row, rotated by the row index: the 8 threads of a quarter-warp phase hit 8 different bank groups
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int ch = (jj + i) & 7;
                double2 rv = make_double2(0.0, 0.0), qv = make_double2(0.0, 0.0);
                if (valid) {
                    rv = *reinterpret_cast<const double2*>(rs + 2 * ch);
                    if (!START) qv = *reinterpret_cast<const double2*>(qs + 2 * ch);
                }
                double2 rn;
                if constexpr (CPX) {
                    const double2 al = sm->alpha[ch];
                    rn = make_double2(rv.x - (al.x * qv.x - al.y * qv.y), rv.y - (al.x * qv.y + al.y * qv.x));
                } else {
                    rn = make_double2(rv.x - sm->alpha[2 * ch].x * qv.x, rv.y - sm->alpha[2 * ch + 1].x * qv.y);
                }
                if (!START && valid) *reinterpret_cast<double2*>(rs + 2 * ch) = rn;
                const bool a0 = sm->act[CPX ? ch : 2 * ch] != 0, a1 = sm->act[CPX ? ch : 2 * ch + 1] != 0;
                // registers are indexed statically: select the chunk's place with predicated moves
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c == ch) { v[2 * c] = a0 ? rn.x : 0.0; v[2 * c + 1] = a1 ? rn.y : 0.0; }
            }
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t h0, l0, h1, l1;
                bf16_split((float)v[2 * k], h0, l0);
                bf16_split((float)v[2 * k + 1], h1, l1);
                hi[k] = h0 | (h1 << 16);
                lo[k] = l0 | (l1 << 16);
            }
            mbar_wait_bounded(&sm->b_empty[bb], ((ti >> 1) & 1u) ^ 1u);
            unsigned char* img = bimg + bb * 8192 + i * 16;
            *reinterpret_cast<uint4*>(img) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(img + 2048) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            *reinterpret_cast<uint4*>(img + 4096) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4*>(img + 6144) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            fence_proxy_async();
            mbar_arrive(&sm->b_full[bb]);
            // the updated residual leaves from the slot it arrived in; the slot goes back to the producer when the store has read it
            if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
            if (i == 0) {
                if (!START) {
                    const uint32_t rows = (uint32_t)min((int64_t)128, n_rows - t * 128);
                    bulk_s2g(r + t * 128 * 16, rqbuf + (size_t)rb * kChunkBytes, rows * 128u);
                    bulk_commit();
                    bulk_wait_read<0>();
                }
                mbar_arrive(&sm->rq_empty[rb]);
            }
            if (t_prev >= 0) epilogue(t_prev, ti_prev);
            t_prev = t;
            ti_prev = ti;
        }
        if (t_prev >= 0) epilogue(t_prev, ti_prev);
        if (i == 0 && !START) bulk_wait<0>();
        // ---- this block's partial projections ----
        mbar_wait_bounded(&sm->done, 0);
        tc_fence_after();
        float* out = c_part + (int64_t)blockIdx.x * kModalMaxModes * 16;
        for (int j = 0; j < (grp == 0 ? n_use : 0); ++j) {   // (none when COARSE)
            float acc[32];
            tmem_ld32(tmem + ((uint32_t)(qd * 32) << 16) + 32 * j, acc);
            float4* o = reinterpret_cast<float4*>(out + (int64_t)(j * kMC + qd * 32 + lane) * 16);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                o[k] = make_float4(acc[4 * k] + acc[16 + 4 * k], acc[4 * k + 1] + acc[17 + 4 * k], acc[4 * k + 2] + acc[18 + 4 * k],
                                   acc[4 * k + 3] + acc[19 + 4 * k]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// coefficients  e = d * c :  fixed-order sum of the block partials, d from lambda, w^2 and the modes' boundary damping;
// written as the bf16 (hi, lo) MN-major operand image of the expansion: chunk j, column block cb, mode k at
// j * 8192 + cb * 2048 + k * 16 (+ 2 bytes per column)
// ------------------------------------------------------------------------------------------------
template <bool CPX>
__global__ void __launch_bounds__(256) k_modal_coef(const float* __restrict__ c_part, int n_blocks, const double* __restrict__ lam,
                                                    const double* __restrict__ res, int n_modes, const double* __restrict__ adiag, const double2* __restrict__ coef, int n_groups, int F,
                                                    const Scalars* __restrict__ sc, uint16_t* __restrict__ e_img) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int i = idx >> 4, c = idx & 15;          // mode, real column
    if (i >= sc->modes_use) return;                // uniform per 16 threads
    double s = 0.0;
    for (int b0 = 0; b0 < n_blocks; b0 += 8) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = (b0 + u < n_blocks) ? __ldcg(&c_part[((int64_t)(b0 + u) * kModalMaxModes + i) * 16 + c]) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) s += (double)t[u];
    }
    double e = 0.0;
    double so = 0.0;
    if constexpr (CPX) so = __shfl_xor_sync(0xffffffffu, s, 1);            // the slot's other component (all lanes take part)
    if (i < n_modes) {
        const double l = lam[i];
        const double floor_den = res[i] * res[i] / fmax(fabs(l), 1e-300) + 1e-13 * fabs(l);
        if constexpr (CPX) {
            const int f = c >> 1;
            double2 den = make_double2(l - sc->w2[f], 0.0);
            for (int g = 0; g < n_groups; ++g) {
                const double2 cf = coef[g * F + f];
                const double a = adiag[(int64_t)g * n_modes + i];
                den.x += cf.x * a;
                den.y += cf.y * a;
            }
            double dn = den.x * den.x + den.y * den.y;
            if (dn < floor_den * floor_den) {
                if (dn > 0.0) { const double g = floor_den / sqrt(dn); den.x *= g; den.y *= g; } else { den = make_double2(floor_den, 0.0); }
                dn = floor_den * floor_den;
            }
            if (f < F && dn > 0.0) {
                const double2 cv = (c & 1) ? make_double2(so, s) : make_double2(s, so);
                const double2 ev = cdiv(cv, den);
                e = (c & 1) ? ev.y : ev.x;
            }
        } else {
            double den = l - sc->w2[c];
            if (fabs(den) < floor_den) den = den < 0.0 ? -floor_den : floor_den;
            if (c < F && den != 0.0) e = s / den;
        }
    }
    uint32_t hi, lo;
    bf16_split((float)e, hi, lo);
    const int j = i >> 7, k = i & 127;
    unsigned char* base = reinterpret_cast<unsigned char*>(e_img) + (size_t)j * 8192 + (size_t)k * 16 + (c & 7) * 2;
    *reinterpret_cast<uint16_t*>(base + (c >> 3) * 2048) = (uint16_t)hi;
    *reinterpret_cast<uint16_t*>(base + (2 + (c >> 3)) * 2048) = (uint16_t)lo;
}

// ------------------------------------------------------------------------------------------------
// expansion  z += V e , then with the complete z:  rho' = r^T z, ||r||^2 and the finalisation of the iteration
// (fin_update) or of the slot load (START: p = z, fin_start).  T = double (real sweep, 16 slots) or double2 (8 complex).
// ------------------------------------------------------------------------------------------------
// PLAIN: z32 = V e and nothing else (the coarse stage of the tile-compressed correction: V = the coefficient tiles, z32 = u).
template <bool CPX, bool START, bool PLAIN = false>
__global__ void __launch_bounds__(kModalThreads, 1)
k_modal_expand(const double* __restrict__ r, float* __restrict__ z32, double* __restrict__ p_start, const uint16_t* __restrict__ V, int n_chunks,
               const uint16_t* __restrict__ e_img, int64_t n_rows, int64_t n_tiles, Scalars* __restrict__ sc, double* __restrict__ partial) {
    constexpr int F = CPX ? 8 : 16;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* ring = smem_raw;                                      // [kStages][32 KB]
    unsigned char* eimg = smem_raw + (size_t)kStages * kChunkBytes;      // [kModalMaxChunks][8 KB]
    ModalSmem* sm = reinterpret_cast<ModalSmem*>(eimg + (size_t)kModalMaxChunks * 8192);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int n_use, ksteps_last;
    uint32_t bytes_last;
    modal_shape(sc->modes_use, n_use, bytes_last, ksteps_last);
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&sm->a_full[i], 1); mbar_init(&sm->a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&sm->b_full[i], 1); mbar_init(&sm->b_empty[i], 128); }   // here: accumulator full / drained
        mbar_init(&sm->e_full, 1);
        mbar_init(&sm->done, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&sm->tmem_base, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm->tmem_base;
    double acc_rho[16], acc_rr[F];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc_rho[k] = 0.0;
#pragma unroll
    for (int k = 0; k < F; ++k) acc_rr[k] = 0.0;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t eb = (uint32_t)(n_use - 1) * 8192u + 8192u;
            mbar_expect_tx(&sm->e_full, eb);
            bulk_g2s(eimg, e_img, eb, &sm->e_full);
            modal_produce(V, n_chunks, n_tiles, n_use, bytes_last, ring, sm);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc_idesc_bf16(128, 32, 0, 1);   // A = V tile, K-major; B = coefficient image, MN-major
            mbar_wait_bounded(&sm->e_full, 0);
            tc_fence_after();
            uint32_t it = 0, ti = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
                const int ab = ti & 1;
                mbar_wait_bounded(&sm->b_empty[ab], ((ti >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator
                tc_fence_after();
                for (int j = 0; j < n_use; ++j, ++it) {
                    const int st = it % kStages;
                    mbar_wait_bounded(&sm->a_full[st], (it / kStages) & 1u);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(ring + (size_t)st * kChunkBytes);
                    const uint32_t b_addr = smem_u32(eimg + (size_t)j * 8192);
                    const int ksn = (j == n_use - 1) ? ksteps_last : 8;
                    for (int ks = 0; ks < ksn; ++ks)   // K = 16 modes per step: two mode blocks
                        tc_mma_bf16(tmem + 32 * ab, tc_desc(a_addr + ks * 2 * kSm, kSm, kSn), tc_desc(b_addr + ks * 256, 128, 2048), idesc,
                                    j > 0 || ks > 0);
                    tc_commit(&sm->a_empty[st]);
                }
                tc_commit(&sm->b_full[ab]);
            }
        }
    } else {
        // ---- epilogue: thread = node row of the tile ----
        const int q = warp & 3;
        const int rin = q * 32 + lane;
        uint32_t ti = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
            const int ab = ti & 1;
            const int64_t row = t * 128 + rin;
            const bool valid = row < n_rows;
            // the row's residual and block-Jacobi part while the tensor core works
            double rv[16];
            float zv[16];
            if (valid && !PLAIN) {
                const double4* src = reinterpret_cast<const double4*>(r + row * 16);
                const float4* zs = reinterpret_cast<const float4*>(z32 + row * 16);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double4 d = src[k];
                    rv[4 * k] = d.x; rv[4 * k + 1] = d.y; rv[4 * k + 2] = d.z; rv[4 * k + 3] = d.w;
                    const float4 zf = zs[k];
                    zv[4 * k] = zf.x; zv[4 * k + 1] = zf.y; zv[4 * k + 2] = zf.z; zv[4 * k + 3] = zf.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) { rv[k] = 0.0; zv[k] = 0.f; }
            }
            mbar_wait_bounded(&sm->b_full[ab], (ti >> 1) & 1u);
            tc_fence_after();
            float acc[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 32 * ab, acc);
            tc_fence_before();
            mbar_arrive(&sm->b_empty[ab]);
#pragma unroll
            for (int k = 0; k < 16; ++k) zv[k] += acc[k] + acc[16 + k];
            if (valid) {
                float4* zd = reinterpret_cast<float4*>(z32 + row * 16);
#pragma unroll
                for (int k = 0; k < 4; ++k) zd[k] = make_float4(zv[4 * k], zv[4 * k + 1], zv[4 * k + 2], zv[4 * k + 3]);
                if constexpr (PLAIN) continue;
                if constexpr (START) {
                    // first search direction of the slots being loaded
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (sc->reset[CPX ? (k >> 1) : k]) p_start[row * 16 + k] = (double)zv[k];
                }
                if constexpr (CPX) {
#pragma unroll
                    for (int f = 0; f < 8; ++f) {
                        const double zr = (double)zv[2 * f], zi = (double)zv[2 * f + 1];
                        acc_rho[2 * f] += rv[2 * f] * zr - rv[2 * f + 1] * zi;
                        acc_rho[2 * f + 1] += rv[2 * f] * zi + rv[2 * f + 1] * zr;
                        acc_rr[f] += rv[2 * f] * rv[2 * f] + rv[2 * f + 1] * rv[2 * f + 1];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        acc_rho[k] += rv[k] * (double)zv[k];
                        acc_rr[k] += rv[k] * rv[k];
                    }
                }
            }
        }
    }
    // ---- block sums in a fixed order (the ring is free: every MMA that reads it has been waited for), then the last block ----
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 64);
    if constexpr (PLAIN) return;
    double* s_red = reinterpret_cast<double*>(ring);   // [128][3 F]
    if (tid >= 64) {
        const int e = tid - 64;
#pragma unroll
        for (int f = 0; f < F; ++f) {
            s_red[e * 3 * F + 3 * f] = CPX ? acc_rho[2 * f] : acc_rho[f];
            s_red[e * 3 * F + 3 * f + 1] = CPX ? acc_rho[2 * f + 1] : 0.0;
            s_red[e * 3 * F + 3 * f + 2] = acc_rr[f];
        }
    }
    __syncthreads();
    if (tid < 3 * F) {
        double s = 0.0;
        for (int e = 0; e < 128; ++e) s += s_red[e * 3 * F + tid];
        partial[(int64_t)blockIdx.x * 3 * F + tid] = s;
    }
    finalize_last_block<F, 3, kModalThreads>(partial, &sc->ticket[START ? 2 : 1], [&](int ff, const double* tot) {
        if constexpr (START) fin_start(sc, ff, tot); else fin_update(sc, ff, tot);
    });
}

// ================================================================================================
// Tile-compressed modes.  The mode tiles are what an iteration streams (2 bytes per node and mode, twice), and inside one
// 128-node tile the modes below the cut are smooth: a tile is a fraction of the shortest wavelength across, so its rows of V
// are, to the accuracy bf16 stores them with, combinations of kCF = 16 functions (measured: tests/research/
// compressed_modes_prototype.py - rank 16 reproduces the exact modes' iteration counts, rank 12 and a cubic polynomial fit
// nearly).  So per tile
//       V_t ~ Phi_t C_t ,   Phi_t: 128 x 16 (the tile's dominant left singular vectors, bf16), C_t: 16 x modes (bf16)
// and the correction  V diag(d) V^T  becomes  Phi (C diag(d) C^T) Phi^T:  t = Phi^T r inside K6a (one more MMA group per tile),
// the two products with C on the coarse vector (16 numbers per tile: 1/8 of the nodes) by the same tensor-core kernels that
// handled V, and z += Phi u with the scalars of the iteration in k_coarse_finish.  The stream per iteration drops from
// 4 bytes to ~0.55 bytes per node and mode.
// ================================================================================================

// One block per tile: Gram matrix of the tile's rows of the modes, orthogonal iteration for its 16 dominant eigenvectors, the
// bf16 basis in the operand layout, and the tile's rows of the coarse modes  Xc = (Phi_b^T Phi_b)^-1 Phi_b^T X  (the least-squares
// coefficients for the ROUNDED basis).  FP64 throughout: the spectrum of the Gram matrix falls by ~1e5 over the 16 kept directions.
constexpr int kTbThreads = 256;
constexpr int kTbGs = 129, kTbPs = 17;   // padded strides (bank conflicts)
static size_t tile_basis_smem() { return (size_t)(128 * kTbGs + 3 * 128 * kTbPs + 16 * 16 + 128) * sizeof(double); }

__device__ __forceinline__ double tb_block_sum(double v, double* s_w) {
    // sum over the block's 256 threads, the same value to every thread (fixed order)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kTbThreads / 32; ++k) t += s_w[k];
    return t;
}

// modified Gram-Schmidt on the 16 columns of A [128][kTbPs]; a column without norm (fewer than 16 independent rows) becomes zero
__device__ void tb_orthonormalise(double* A, double* s_w, double* s_d) {
    const int tid = threadIdx.x, i = tid & 127;
    const bool own = tid < 128;
    for (int k = 0; k < kCF; ++k) {
        const double a = own ? A[i * kTbPs + k] : 0.0;
        const double n2 = tb_block_sum(a * a, s_w);
        const double inv = n2 > 1e-280 ? rsqrt(n2) : 0.0;
        if (own) A[i * kTbPs + k] = a * inv;
        // dots of the normalised column with the later ones: 16 threads per later column would idle most of the block, so every
        // row thread forms its 15 products and the warps reduce them
        double d[kCF];
#pragma unroll
        for (int j = 0; j < kCF; ++j) d[j] = (own && j > k) ? a * inv * A[i * kTbPs + j] : 0.0;
#pragma unroll
        for (int j = 0; j < kCF; ++j)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) d[j] += __shfl_xor_sync(0xffffffffu, d[j], o);
        __syncthreads();
        if ((tid & 31) == 0 && own)
#pragma unroll
            for (int j = 0; j < kCF; ++j) s_d[(tid >> 5) * kCF + j] = d[j];
        __syncthreads();
        if (own) {
#pragma unroll
            for (int j = 0; j < kCF; ++j)
                if (j > k) {
                    const double dj = s_d[j] + s_d[kCF + j] + s_d[2 * kCF + j] + s_d[3 * kCF + j];
                    A[i * kTbPs + j] -= dj * a * inv;
                }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kTbThreads, 1)
k_tile_basis(const double* __restrict__ X, int64_t n_rows, int n_modes, int n_iter, uint16_t* __restrict__ phi16, double* __restrict__ Xc,
             int64_t nc_rows, double* __restrict__ fit /* [n_modes][2]: squared fit error and squared norm of every mode, summed over the tiles */) {
    extern __shared__ __align__(16) unsigned char tb_raw[];
    double* G = reinterpret_cast<double*>(tb_raw);   // [128][kTbGs]
    double* blk = G + 128 * kTbGs;                   // [128][kTbPs]  one panel's rows of the tile
    double* P = blk + 128 * kTbPs;                   // [128][kTbPs]
    double* Y = P + 128 * kTbPs;                     // [128][kTbPs]
    double* M = Y + 128 * kTbPs;                     // [16][16]
    double* s_w = M + 256;                           // [8] + [4][16]
    double* s_d = s_w + 8;
    const int tid = threadIdx.x, i = tid & 127, half = tid >> 7;
    const int64_t t = blockIdx.x;
    const int nbat = (n_modes + 15) / 16;
    auto load_panel = [&](int bt) {
        for (int e = tid; e < 128 * 16; e += kTbThreads) {
            const int node = e >> 4, k = e & 15;
            const int64_t row = t * 128 + node;
            blk[node * kTbPs + k] = (row < n_rows && 16 * bt + k < n_modes) ? X[((int64_t)bt * n_rows + row) * 16 + k] : 0.0;
        }
    };
    // ---- Gram matrix: thread (i, half) owns G[i][64 half .. 64 half + 63] ----
    double acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.0;
    for (int bt = 0; bt < nbat; ++bt) {
        __syncthreads();
        load_panel(bt);
        __syncthreads();
#pragma unroll 2
        for (int k = 0; k < 16; ++k) {
            const double a = blk[i * kTbPs + k];
#pragma unroll
            for (int j = 0; j < 64; ++j) acc[j] = fma(a, blk[(64 * half + j) * kTbPs + k], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 64; ++j) G[i * kTbGs + 64 * half + j] = acc[j];
    __syncthreads();
    // ---- orthogonal iteration, started from 16 spread columns of G (vectors in the range of the tile's modes) ----
    if (tid < 128)
        for (int k = 0; k < kCF; ++k) P[i * kTbPs + k] = G[i * kTbGs + 8 * k + 3];
    __syncthreads();
    tb_orthonormalise(P, s_w, s_d);
    for (int itn = 0; itn < n_iter; ++itn) {
        double y[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = 0.0;
        for (int j = 0; j < 128; ++j) {
            const double g = G[i * kTbGs + j];
#pragma unroll
            for (int k = 0; k < 8; ++k) y[k] = fma(g, P[j * kTbPs + 8 * half + k], y[k]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) Y[i * kTbPs + 8 * half + k] = y[k];
        __syncthreads();
        tb_orthonormalise(Y, s_w, s_d);
        for (int e = tid; e < 128 * kCF; e += kTbThreads) P[(e >> 4) * kTbPs + (e & 15)] = Y[(e >> 4) * kTbPs + (e & 15)];
        __syncthreads();
    }
    // ---- the basis as stored (bf16), M = Phi_b^T Phi_b, W = Phi_b M^-1 ----
    for (int e = tid; e < 128 * kCF; e += kTbThreads) {
        const int node = e >> 4, k = e & 15;
        P[node * kTbPs + k] = (double)bf16_to_float(bf16_rn((float)P[node * kTbPs + k]));
    }
    __syncthreads();
    {
        const int a = tid >> 4, b = tid & 15;
        double m = 0.0;
        for (int node = 0; node < 128; ++node) m = fma(P[node * kTbPs + a], P[node * kTbPs + b], m);
        M[a * 16 + b] = m;
    }
    __syncthreads();
    if (tid == 0) {
        // Gauss-Jordan on [M | I] in place (M is the identity to bf16 rounding, or has zero rows / columns where the basis has
        // zero columns: those stay out)
        double inv[16][16];
        for (int a = 0; a < 16; ++a)
            for (int b = 0; b < 16; ++b) inv[a][b] = (a == b) ? 1.0 : 0.0;
        for (int c = 0; c < 16; ++c) {
            const double piv = M[c * 16 + c];
            if (!(fabs(piv) > 1e-6)) {
                for (int b = 0; b < 16; ++b) { inv[c][b] = 0.0; M[c * 16 + b] = 0.0; }
                for (int a = 0; a < 16; ++a) M[a * 16 + c] = 0.0;
                continue;
            }
            const double ip = 1.0 / piv;
            for (int b = 0; b < 16; ++b) { M[c * 16 + b] *= ip; inv[c][b] *= ip; }
            for (int a = 0; a < 16; ++a) {
                if (a == c) continue;
                const double f = M[a * 16 + c];
                if (f == 0.0) continue;
                for (int b = 0; b < 16; ++b) { M[a * 16 + b] -= f * M[c * 16 + b]; inv[a][b] -= f * inv[c][b]; }
            }
        }
        for (int a = 0; a < 16; ++a)
            for (int b = 0; b < 16; ++b) M[a * 16 + b] = inv[a][b];
    }
    __syncthreads();
    for (int e = tid; e < 128 * kCF; e += kTbThreads) {
        const int node = e >> 4, k = e & 15;
        double wv = 0.0;
#pragma unroll
        for (int a = 0; a < 16; ++a) wv = fma(P[node * kTbPs + a], M[a * 16 + k], wv);
        Y[node * kTbPs + k] = wv;
    }
    // ---- the basis in the operand layout: function block fb, node block nb, node ni -> fb * 2048 + nb * 128 + ni * 16 (8 functions) ----
    if (tid < 256) {
        const int node = tid & 127, fb = tid >> 7;
        uint32_t w4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t lo = bf16_rn((float)P[node * kTbPs + 8 * fb + 2 * k]), hi = bf16_rn((float)P[node * kTbPs + 8 * fb + 2 * k + 1]);
            w4[k] = lo | (hi << 16);
        }
        unsigned char* dst = reinterpret_cast<unsigned char*>(phi16) + (size_t)t * kPhiBytes + fb * kSm + (node >> 3) * kSn + (node & 7) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
    // ---- coarse modes: Xc[bt][16 t + k][c] = sum_node W[node][k] X[bt][128 t + node][c] ----
    const int kk = tid >> 4, cc = tid & 15;
    for (int bt = 0; bt < nbat; ++bt) {
        __syncthreads();
        load_panel(bt);
        __syncthreads();
        double sacc = 0.0;
        for (int node = 0; node < 128; ++node) sacc = fma(Y[node * kTbPs + kk], blk[node * kTbPs + cc], sacc);
        Xc[((int64_t)bt * nc_rows + t * kCF + kk) * 16 + cc] = sacc;
        // what the compressed form loses of this panel's modes on this tile: || V_t - Phi_b C_t ||^2 per mode (the host decides from
        // the sums over all tiles whether the tiles are small enough against the shortest wavelength for 16 functions)
        M[kk * 16 + cc] = sacc;
        __syncthreads();
        double e2[8], n2[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double v = blk[i * kTbPs + 8 * half + c];
            double rec = 0.0;
#pragma unroll
            for (int k = 0; k < kCF; ++k) rec = fma(P[i * kTbPs + k], M[k * 16 + 8 * half + c], rec);
            e2[c] = (v - rec) * (v - rec);
            n2[c] = v * v;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                e2[c] += __shfl_xor_sync(0xffffffffu, e2[c], o);
                n2[c] += __shfl_xor_sync(0xffffffffu, n2[c], o);
            }
        if ((tid & 31) == 0) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int mode = 16 * bt + 8 * half + c;
                if (mode < n_modes) { atomicAdd(&fit[2 * mode], e2[c]); atomicAdd(&fit[2 * mode + 1], n2[c]); }
            }
        }
    }
}

// z += Phi u per tile (CUDA cores: 16 functions x 16 columns per node row), then with the complete z what k_modal_expand does:
// rho' = r^T z, ||r||^2, the iteration's scalars (fin_update) or the slot load's (START: p = z, fin_start).
// A streaming kernel with little arithmetic: one block per SM, a producer thread keeps kCfStages tiles (r | z | Phi | u, 29 KB
// each) in flight with bulk copies, two groups of eight consumer warps take alternate tiles, two threads per node row (8 columns
// each).  (With one thread per row and 4 to 12 warps per SM the kernel ran at 3 - 3.6 TB/s whatever was in flight: every
// instruction's latency was exposed.)
constexpr int kCfGroups = 2;
constexpr int kCfThreads = 32 + kCfGroups * 256;
constexpr int kCfStages = 6;
constexpr int kCfOffZ = 16384, kCfOffPhi = kCfOffZ + 8192, kCfOffU = kCfOffPhi + kPhiBytes, kCfStageBytes = kCfOffU + kCF * 16 * 4;
struct CfSmem {
    uint64_t full[kCfStages], empty[kCfStages];
};
static size_t finish_smem() { return (size_t)kCfStages * kCfStageBytes + sizeof(CfSmem) + 64; }

template <bool CPX, bool START>
__global__ void __launch_bounds__(kCfThreads, 1)
k_coarse_finish(const double* __restrict__ r, float* z32, double* __restrict__ p_start, const uint16_t* __restrict__ phi16,
                const float* __restrict__ uc, int64_t n_rows, int64_t n_tiles, Scalars* __restrict__ sc, double* __restrict__ partial) {
    constexpr int F = CPX ? 8 : 16;
    constexpr int FL = F / 2;   // slots a thread carries: columns 8 h .. 8 h + 7 of its row
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    CfSmem* sm = reinterpret_cast<CfSmem*>(smem_raw + (size_t)kCfStages * kCfStageBytes);
    __shared__ double s_part[kCfGroups * 8][2][3 * FL];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < kCfStages; ++i) { mbar_init(&sm->full[i], 1); mbar_init(&sm->empty[i], 8); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int st = it % kCfStages;
                mbar_wait_bounded(&sm->empty[st], ((it / kCfStages) & 1u) ^ 1u);
                const uint32_t rows = (uint32_t)min((int64_t)128, n_rows - t * 128);
                unsigned char* dst = smem_raw + (size_t)st * kCfStageBytes;
                mbar_expect_tx(&sm->full[st], rows * 128u + rows * 64u + (uint32_t)kPhiBytes + (uint32_t)(kCF * 64));
                bulk_g2s(dst, r + t * 128 * 16, rows * 128u, &sm->full[st]);
                bulk_g2s(dst + kCfOffZ, z32 + t * 128 * 16, rows * 64u, &sm->full[st]);
                bulk_g2s(dst + kCfOffPhi, phi16 + t * (kPhiBytes / 2), (uint32_t)kPhiBytes, &sm->full[st]);
                bulk_g2s(dst + kCfOffU, uc + t * (kCF * 16), (uint32_t)(kCF * 64), &sm->full[st]);
            }
        }
    } else {
        const int c = tid - 32, grp = c >> 8, j = c & 255;
        const int i = j >> 1, h = j & 1;   // node row of the tile, half of its columns
        double a0[FL], a1[FL], a2[FL];     // rho' (re, im) and ||r||^2 of the thread's slots
#pragma unroll
        for (int k = 0; k < FL; ++k) a0[k] = a1[k] = a2[k] = 0.0;
        uint32_t it = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            if ((int)(it & 1u) != grp) continue;
            const int st = it % kCfStages;
            const int64_t row = t * 128 + i;
            const bool valid = row < n_rows;
            mbar_wait_bounded(&sm->full[st], (it / kCfStages) & 1u);
            const unsigned char* stg = smem_raw + (size_t)st * kCfStageBytes;
            double rv[8];
            float zv[8], ph[kCF];
            // 16-byte chunks rotated by the row: the 8 lanes of a quarter-warp phase (4 rows x 2 halves) hit 8 different bank groups
            const double* rs = reinterpret_cast<const double*>(stg) + i * 16 + 8 * h;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int ch = (jj + i) & 3;
                const double2 v = valid ? *reinterpret_cast<const double2*>(rs + 2 * ch) : make_double2(0.0, 0.0);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    if (cc == ch) { rv[2 * cc] = v.x; rv[2 * cc + 1] = v.y; }
            }
            const float* zs = reinterpret_cast<const float*>(stg + kCfOffZ) + i * 16 + 8 * h;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int ch = (jj + (i >> 1)) & 1;
                const float4 v = valid ? *reinterpret_cast<const float4*>(zs + 4 * ch) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int cc = 0; cc < 2; ++cc)
                    if (cc == ch) { zv[4 * cc] = v.x; zv[4 * cc + 1] = v.y; zv[4 * cc + 2] = v.z; zv[4 * cc + 3] = v.w; }
            }
            const unsigned char* pb = stg + kCfOffPhi + (i >> 3) * kSn + (i & 7) * 16;
#pragma unroll
            for (int fb = 0; fb < 2; ++fb) {
                const uint4 wv = *reinterpret_cast<const uint4*>(pb + fb * kSm);
                const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    ph[8 * fb + 2 * k] = valid ? bf16_to_float(ww[k] & 0xffffu) : 0.f;
                    ph[8 * fb + 2 * k + 1] = valid ? bf16_to_float(ww[k] >> 16) : 0.f;
                }
            }
            const float4* us = reinterpret_cast<const float4*>(stg + kCfOffU) + 2 * h;
#pragma unroll
            for (int k = 0; k < kCF; ++k) {
                const float4 u0 = us[k * 4], u1 = us[k * 4 + 1];
                zv[0] = fmaf(ph[k], u0.x, zv[0]); zv[1] = fmaf(ph[k], u0.y, zv[1]); zv[2] = fmaf(ph[k], u0.z, zv[2]); zv[3] = fmaf(ph[k], u0.w, zv[3]);
                zv[4] = fmaf(ph[k], u1.x, zv[4]); zv[5] = fmaf(ph[k], u1.y, zv[5]); zv[6] = fmaf(ph[k], u1.z, zv[6]); zv[7] = fmaf(ph[k], u1.w, zv[7]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm->empty[st]);   // this warp has read the stage
            if (valid) {
                float4* zd = reinterpret_cast<float4*>(z32 + row * 16 + 8 * h);
                zd[0] = make_float4(zv[0], zv[1], zv[2], zv[3]);
                zd[1] = make_float4(zv[4], zv[5], zv[6], zv[7]);
                if constexpr (START) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (sc->reset[CPX ? ((8 * h + k) >> 1) : (8 * h + k)]) p_start[row * 16 + 8 * h + k] = (double)zv[k];
                }
                if constexpr (CPX) {
#pragma unroll
                    for (int f = 0; f < FL; ++f) {
                        const double zr = (double)zv[2 * f], zi = (double)zv[2 * f + 1];
                        a0[f] += rv[2 * f] * zr - rv[2 * f + 1] * zi;
                        a1[f] += rv[2 * f] * zi + rv[2 * f + 1] * zr;
                        a2[f] += rv[2 * f] * rv[2 * f] + rv[2 * f + 1] * rv[2 * f + 1];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < FL; ++k) {
                        a0[k] += rv[k] * (double)zv[k];
                        a2[k] += rv[k] * rv[k];
                    }
                }
            }
        }
        // ---- block sums in a fixed order: the 16 lanes of a warp that carry the same half by shuffle, then the warps ----
#pragma unroll
        for (int k = 0; k < FL; ++k) {
            double v0 = a0[k], v1 = a1[k], v2 = a2[k];
#pragma unroll
            for (int o = 16; o > 1; o >>= 1) {
                v0 += __shfl_xor_sync(0xffffffffu, v0, o);
                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                v2 += __shfl_xor_sync(0xffffffffu, v2, o);
            }
            if (lane < 2) { s_part[warp - 1][lane][3 * k] = v0; s_part[warp - 1][lane][3 * k + 1] = v1; s_part[warp - 1][lane][3 * k + 2] = v2; }
        }
    }
    __syncthreads();
    if (tid < 3 * F) {
        const int f = tid / 3, i3 = tid % 3, hh = f / FL, k = f % FL;
        double sacc = 0.0;
#pragma unroll
        for (int wq = 0; wq < kCfGroups * 8; ++wq) sacc += s_part[wq][hh][3 * k + i3];
        partial[(int64_t)blockIdx.x * 3 * F + tid] = sacc;
    }
    finalize_last_block<F, 3, kCfThreads>(partial, &sc->ticket[START ? 2 : 1], [&](int ff, const double* tot) {
        if constexpr (START) fin_start(sc, ff, tot); else fin_update(sc, ff, tot);
    });
}

// ------------------------------------------------------------------------------------------------
// K6b on 128-byte vector rows (16 real or 8 complex slots): x += alpha p ; p = z + beta p.  Same shape as k_coarse_finish: a
// producer thread streams tiles (x | p | z, 40 KB) into a ring with bulk copies, eight warps update them in place (two threads
// per row) and the tiles leave by bulk stores out of the buffers they arrived in.  (The grid-stride version with 8-byte accesses
// ran at 0.74 of the HBM rate.  A second consumer group on alternate tiles bought nothing here - there is next to no arithmetic
// to overlap - and with an odd number of stages, where a stage alternates between the groups, it corrupted tiles now and then:
// in the kernels that do use two groups every buffer always returns to the same group.)
// ------------------------------------------------------------------------------------------------
constexpr int kPuThreads = 32 + 256;
constexpr int kPuStages = 5;
constexpr int kPuOffP = 16384, kPuOffZ = 32768, kPuStageBytes = 32768 + 8192;
struct PuSmem {
    uint64_t full[kPuStages], empty[kPuStages];
    double2 alpha[16], beta[16];
    int xpend[16], act[16];
};
static size_t pupdate_smem() { return (size_t)kPuStages * kPuStageBytes + sizeof(PuSmem) + 64; }

template <bool CPX>
__global__ void __launch_bounds__(kPuThreads, 1)
k_pupdate_tiles(double* x, double* p, const float* __restrict__ z32, int64_t n_rows, int64_t n_tiles, const Scalars* __restrict__ sc) {
    constexpr int F = CPX ? 8 : 16;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    PuSmem* sm = reinterpret_cast<PuSmem*>(smem_raw + (size_t)kPuStages * kPuStageBytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < kPuStages; ++i) { mbar_init(&sm->full[i], 1); mbar_init(&sm->empty[i], 1); }
        mbar_fence_init();
    }
    if (tid < F) {
        sm->alpha[tid] = sc->alpha[tid];
        sm->beta[tid] = sc->beta[tid];
        sm->xpend[tid] = sc->xpend[tid];
        sm->act[tid] = sc->active[tid];
    }
    __syncthreads();
    bool any = false;
    for (int f = 0; f < F; ++f) any |= sm->xpend[f] != 0;
    if (!any) return;   // nothing pending in any slot (uniform)

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int st = it % kPuStages;
                mbar_wait_bounded(&sm->empty[st], ((it / kPuStages) & 1u) ^ 1u);
                const uint32_t rows = (uint32_t)min((int64_t)128, n_rows - t * 128);
                unsigned char* dst = smem_raw + (size_t)st * kPuStageBytes;
                mbar_expect_tx(&sm->full[st], rows * 128u * 2u + rows * 64u);
                bulk_g2s(dst, x + t * 128 * 16, rows * 128u, &sm->full[st]);
                bulk_g2s(dst + kPuOffP, p + t * 128 * 16, rows * 128u, &sm->full[st]);
                bulk_g2s(dst + kPuOffZ, z32 + t * 128 * 16, rows * 64u, &sm->full[st]);
            }
        }
    } else {
        const int j = tid - 32;
        const int i = j >> 1, h = j & 1;   // node row of the tile, half of its columns
        uint32_t it = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int st = it % kPuStages;
            const uint32_t rows = (uint32_t)min((int64_t)128, n_rows - t * 128);
            mbar_wait_bounded(&sm->full[st], (it / kPuStages) & 1u);
            unsigned char* stg = smem_raw + (size_t)st * kPuStageBytes;
            if (i < (int)rows) {
                double* xs = reinterpret_cast<double*>(stg) + i * 16 + 8 * h;
                double* ps = reinterpret_cast<double*>(stg + kPuOffP) + i * 16 + 8 * h;
                const float* zs = reinterpret_cast<const float*>(stg + kPuOffZ) + i * 16 + 8 * h;
                // 16-byte chunks rotated by the row (bank groups), each chunk = 2 real slots or one complex slot
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int ch = (jj + i) & 3;
                    const int col = 8 * h + 2 * ch;             // first real column of the chunk
                    const int f0 = CPX ? (col >> 1) : col, f1 = CPX ? f0 : col + 1;
                    if (!sm->xpend[f0] && !sm->xpend[f1]) continue;
                    double2 xv = *reinterpret_cast<const double2*>(xs + 2 * ch);
                    double2 pv = *reinterpret_cast<const double2*>(ps + 2 * ch);
                    const float2 zf = *reinterpret_cast<const float2*>(zs + 2 * ch);
                    if constexpr (CPX) {
                        const double2 al = sm->alpha[f0], be = sm->beta[f0];
                        xv = make_double2(xv.x + al.x * pv.x - al.y * pv.y, xv.y + al.x * pv.y + al.y * pv.x);
                        if (sm->act[f0]) pv = make_double2((double)zf.x + be.x * pv.x - be.y * pv.y, (double)zf.y + be.x * pv.y + be.y * pv.x);
                    } else {
                        if (sm->xpend[f0]) {
                            xv.x += sm->alpha[f0].x * pv.x;
                            if (sm->act[f0]) pv.x = (double)zf.x + sm->beta[f0].x * pv.x;
                        }
                        if (sm->xpend[f1]) {
                            xv.y += sm->alpha[f1].x * pv.y;
                            if (sm->act[f1]) pv.y = (double)zf.y + sm->beta[f1].x * pv.y;
                        }
                    }
                    *reinterpret_cast<double2*>(xs + 2 * ch) = xv;
                    *reinterpret_cast<double2*>(ps + 2 * ch) = pv;
                }
            }
            fence_proxy_async();   // the generic-proxy writes above before the bulk stores read the tile
            __syncwarp();
            asm volatile("barrier.sync 1, 256;" ::: "memory");
            if (j == 0) {
                bulk_s2g(x + t * 128 * 16, stg, rows * 128u);
                bulk_s2g(p + t * 128 * 16, stg + kPuOffP, rows * 128u);
                bulk_commit();
                bulk_wait_read<0>();
                mbar_arrive(&sm->empty[st]);
            }
        }
        if (j == 0) bulk_wait<0>();
    }
}

void pupdate_tiles(fdb_system* sys, SweepWork* w) {
    static thread_local bool done[64] = {};
    if (!done[sys->device & 63]) {
        FDB_CUDA(cudaFuncSetAttribute(k_pupdate_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pupdate_smem()));
        FDB_CUDA(cudaFuncSetAttribute(k_pupdate_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pupdate_smem()));
        done[sys->device & 63] = true;
    }
    const int64_t n_tiles = (w->N + 127) / 128;
    const int G = (int)std::max<int64_t>(1, std::min<int64_t>(n_tiles, sys->num_sms));
    if (w->real) k_pupdate_tiles<false><<<G, kPuThreads, pupdate_smem(), sys->stream>>>((double*)w->x.p, (double*)w->p.p, w->z32.p, w->N, n_tiles, w->sc.p);
    else k_pupdate_tiles<true><<<G, kPuThreads, pupdate_smem(), sys->stream>>>((double*)w->x.p, (double*)w->p.p, w->z32.p, w->N, n_tiles, w->sc.p);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static size_t fused_smem(bool coarse = false) { return (size_t)(kStages + 2) * kChunkBytes + 2 * 8192 + sizeof(FusedSmem) + 64 + (coarse ? 2 * kPhiBytes : 0); }
static size_t project_smem() { return (size_t)kStages * kChunkBytes + 2 * 8192 + sizeof(ModalSmem) + 64; }
static size_t expand_smem() { return (size_t)kStages * kChunkBytes + (size_t)kModalMaxChunks * 8192 + sizeof(ModalSmem) + 64; }

int modal_grid(fdb_system* sys) { return (int)std::max<int64_t>(1, std::min<int64_t>(sys->n_tiles, sys->num_sms)); }
static int coarse_grid(fdb_system* sys) { return (int)std::max<int64_t>(1, std::min<int64_t>(sys->n_ctiles, sys->num_sms)); }
static int finish_grid(fdb_system* sys) { return modal_grid(sys); }   // one persistent block per SM

// the full-resolution bf16 tiles of V (the uncompressed path; with compression only fdb_modal_apply's reference product asks for them)
static void modal_ensure_fine(fdb_system* sys) {
    if (sys->have_fine_v) return;
    FDB_REQUIRE(sys->modal_x.p != nullptr, "the FP64 modes are gone");
    sys->modal_v.alloc((size_t)sys->n_tiles * sys->mode_chunks * (kChunkBytes / 2));
    const int64_t work = sys->n_tiles * 128 * sys->mode_chunks * 16;
    k_modal_pack<<<grid_for(work, 256), 256, 0, sys->stream>>>(sys->modal_x.p, sys->n_nodes, sys->n_tiles, sys->n_modes, sys->mode_chunks, sys->modal_v.p);
    FDB_CUDA(cudaGetLastError());
    sys->have_fine_v = true;
}

void modal_prepare(fdb_system* sys, SweepWork* w) {
    FDB_REQUIRE(sys->have_modes && sys->n_modes > 0, "no modes: call fdb_modal_compute or fdb_modal_set first");
    FDB_REQUIRE(sys->tile_rows == 128 && sys->n_tiles > 0, "the modal correction needs the 128-row tiles of the internal order");
    w->c_part.alloc((size_t)modal_grid(sys) * kModalMaxModes * 16);
    w->e_img.alloc((size_t)kModalMaxChunks * 4096);
    if (sys->modal_compress) {
        w->fused_precond = true;   // t = Phi^T r exists only inside the fused kernel
        w->tcv.alloc((size_t)sys->n_ctiles * 128 * 16);
        w->uc.alloc((size_t)sys->n_ctiles * 128 * 16);
        w->tcv.zero(sys->stream);  // the rows that pad the last coarse tile
    } else {
        modal_ensure_fine(sys);
    }
    static thread_local bool done[64] = {};
    FDB_REQUIRE(sys->device >= 0 && sys->device < 64, "device index out of range");
    if (!done[sys->device]) {
        FDB_CUDA(cudaFuncSetAttribute(k_modal_project, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)project_smem()));
        FDB_CUDA(cudaFuncSetAttribute(k_precond_fused<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem()));
        FDB_CUDA(cudaFuncSetAttribute(k_precond_fused<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem()));
        FDB_CUDA(cudaFuncSetAttribute(k_precond_fused<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem()));
        FDB_CUDA(cudaFuncSetAttribute(k_precond_fused<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem()));
        FDB_CUDA(cudaFuncSetAttribute(k_modal_expand<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)expand_smem()));
        FDB_CUDA(cudaFuncSetAttribute(k_modal_expand<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)expand_smem()));
        FDB_CUDA(cudaFuncSetAttribute(k_modal_expand<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)expand_smem()));
        FDB_CUDA(cudaFuncSetAttribute(k_modal_expand<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)expand_smem()));
        FDB_CUDA(cudaFuncSetAttribute(k_modal_expand<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)expand_smem()));
        FDB_CUDA(cudaFuncSetAttribute(k_precond_fused<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem(true)));
        FDB_CUDA(cudaFuncSetAttribute(k_precond_fused<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem(true)));
        FDB_CUDA(cudaFuncSetAttribute(k_precond_fused<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem(true)));
        FDB_CUDA(cudaFuncSetAttribute(k_precond_fused<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem(true)));
        FDB_CUDA(cudaFuncSetAttribute(k_coarse_finish<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)finish_smem()));
        FDB_CUDA(cudaFuncSetAttribute(k_coarse_finish<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)finish_smem()));
        FDB_CUDA(cudaFuncSetAttribute(k_coarse_finish<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)finish_smem()));
        FDB_CUDA(cudaFuncSetAttribute(k_coarse_finish<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)finish_smem()));
        done[sys->device] = true;
    }
}

// modes the current batch applies: every mode below sqrt(f_max^2 + delta^2), a multiple of 16, at most what is stored
// Tile-compressed modes: ALL stored modes, always - the two products with the coefficient tiles are ~15 % of an iteration at the full
// count, and the extra deflation (a 100 Hz system then sees the modes up to the sweep's cut, not only those below 316 Hz) is worth
// far more: 109 -> 79 iterations per frequency at config 4.  The cut only ever limited what full-resolution tiles had to stream.
int modal_modes_for(fdb_system* sys, double w2_max, double delta_hz) {
    if (sys->modal_compress) return std::min(((sys->n_modes + 15) / 16) * 16, sys->mode_chunks * kMC);
    const double two_pi = 6.283185307179586;
    const double lam_cut = w2_max + (two_pi * delta_hz) * (two_pi * delta_hz);
    const auto& l = sys->modal_lam_host;
    int m = (int)(std::upper_bound(l.begin(), l.end(), lam_cut) - l.begin());
    m = std::max(m, std::min<int>(16, (int)l.size()));
    m = ((m + 15) / 16) * 16;
    return std::min(m, sys->mode_chunks * kMC);
}

// the three launches between K6a and K6b: c = V^T r, e = d c, z += V e (+ rho', ||r||^2, finalisation)
// K6a of the modal path: residual update + block-Jacobi product + projection in one tensor-core kernel
void modal_precond_fused(fdb_system* sys, SweepWork* w, bool start) {
    FDB_REQUIRE(sys->blk_inv16.p != nullptr, "bf16 block inverses missing");
    const int G = modal_grid(sys);
    const bool coarse = sys->modal_compress;
    auto go = [&](auto kern) {
        kern<<<G, kFusedThreads, fused_smem(coarse), sys->stream>>>((double*)w->r.p, (const double*)w->q.p, sys->blk_inv16.p, w->z32.p, sys->modal_v.p,
                                                                   sys->mode_chunks, w->N, sys->n_tiles, w->sc.p, w->c_part.p, sys->modal_phi.p, w->tcv.p);
    };
    if (coarse) {
        if (w->real) { if (start) go(k_precond_fused<true, false, true>); else go(k_precond_fused<false, false, true>); }
        else { if (start) go(k_precond_fused<true, true, true>); else go(k_precond_fused<false, true, true>); }
        return;
    }
    if (w->real) { if (start) go(k_precond_fused<true, false>); else go(k_precond_fused<false, false>); }
    else { if (start) go(k_precond_fused<true, true>); else go(k_precond_fused<false, true>); }
}
int64_t modal_fused_bytes(fdb_system* sys, SweepWork* w, int modes_use) {
    const int64_t vec = (int64_t)w->N * kModalCols;
    const int64_t rest = 2 * 8 * vec /* q in, r out */ + 4 * vec /* z32 out */ + sys->n_tiles * (int64_t)kChunkBytes /* B_t */;
    if (sys->modal_compress) return 8 * vec /* r in */ + rest + sys->n_tiles * (int64_t)kPhiBytes + sys->nc_rows * 16 * 8 /* t out */;
    return modal_bytes(sys, w, modes_use, 1) + rest;
}

// parts: 1 = projection, 2 = coefficients, 4 = expansion (bench.py times them one at a time)
void modal_apply(fdb_system* sys, SweepWork* w, bool overlay, bool start, int parts) {
    const bool coarse = sys->modal_compress;
    const int G = coarse ? coarse_grid(sys) : modal_grid(sys);
    const double* r = (const double*)w->r.p;
    if ((parts & 1) && coarse)
        k_modal_project<<<G, kModalThreads, project_smem(), sys->stream>>>(w->tcv.p, sys->modal_vc.p, sys->mode_chunks, sys->nc_rows, sys->n_ctiles, w->sc.p,
                                                                          w->c_part.p);
    else if (parts & 1)
        k_modal_project<<<G, kModalThreads, project_smem(), sys->stream>>>(r, sys->modal_v.p, sys->mode_chunks, w->N, sys->n_tiles, w->sc.p, w->c_part.p);
    const int cg = grid_for((int64_t)sys->mode_chunks * kMC * 16, 256);
    if (!(parts & 2)) {
    } else if (w->real)
        k_modal_coef<false><<<cg, 256, 0, sys->stream>>>(w->c_part.p, G, sys->modal_lam.p, sys->modal_res.p, sys->n_modes, nullptr, w->coef.p, 0, w->F,
                                                        w->sc.p, w->e_img.p);
    else
        k_modal_coef<true><<<cg, 256, 0, sys->stream>>>(w->c_part.p, G, sys->modal_lam.p, sys->modal_res.p, sys->n_modes, sys->modal_adiag.p, w->coef.p,
                                                       overlay ? sys->n_groups : 0, w->F, w->sc.p, w->e_img.p);
    double* part = (double*)w->partial.p;
    if (coarse) {
        if (parts & 4)
            k_modal_expand<false, false, true><<<G, kModalThreads, expand_smem(), sys->stream>>>(nullptr, w->uc.p, nullptr, sys->modal_vc.p, sys->mode_chunks,
                                                                                                w->e_img.p, sys->nc_rows, sys->n_ctiles, w->sc.p, part);
        if (parts & 8) {
            auto fin = [&](auto kern) {
                kern<<<finish_grid(sys), kCfThreads, finish_smem(), sys->stream>>>(r, w->z32.p, (double*)w->p.p, sys->modal_phi.p, w->uc.p, w->N, sys->n_tiles,
                                                                                 w->sc.p, part);
            };
            if (w->real) { if (start) fin(k_coarse_finish<false, true>); else fin(k_coarse_finish<false, false>); }
            else { if (start) fin(k_coarse_finish<true, true>); else fin(k_coarse_finish<true, false>); }
        }
        return;
    }
    if (!(parts & 4)) return;
    auto go = [&](auto kern) {
        kern<<<G, kModalThreads, expand_smem(), sys->stream>>>(r, w->z32.p, (double*)w->p.p, sys->modal_v.p, sys->mode_chunks, w->e_img.p, w->N,
                                                              sys->n_tiles, w->sc.p, part);
    };
    if (w->real) { if (start) go(k_modal_expand<false, true>); else go(k_modal_expand<false, false>); }
    else { if (start) go(k_modal_expand<true, true>); else go(k_modal_expand<true, false>); }
}

// Algorithmic bytes of one launch at `modes_use` modes: 2 bytes per node and mode (the bf16 tile stream, whole 8-mode blocks) plus the
// FP64 residual tile (projection) or the residual, the FP32 z tile read and written and the coefficient image (expansion).
int64_t modal_bytes(fdb_system* sys, SweepWork* w, int modes_use, int part) {
    const int64_t vec = (int64_t)w->N * kModalCols;
    if (sys->modal_compress) {
        // the coefficient tiles C (2 bytes per coarse unknown and mode), the coarse vectors (16 numbers per tile and column), the tile bases
        const int64_t vc = sys->n_ctiles * 128 * (int64_t)(((modes_use + 7) / 8) * 8) * 2;
        const int64_t cvec = sys->nc_rows * 16;
        if (part == 1) return vc + 8 * cvec;
        if (part == 4) return vc + 4 * cvec + (int64_t)modes_use * 64;
        if (part == 8) return 8 * vec + 2 * 4 * vec + sys->n_tiles * (int64_t)kPhiBytes + 4 * cvec;
        return (int64_t)coarse_grid(sys) * modes_use * 64 + (int64_t)modes_use * 64;
    }
    const int64_t rows_pad = sys->n_tiles * 128;
    const int64_t v = rows_pad * (int64_t)(((modes_use + 7) / 8) * 8) * 2;
    if (part == 1) return v + 8 * vec;
    if (part == 4) return v + 8 * vec + 2 * 4 * vec + (int64_t)modes_use * 64;
    return (int64_t)modal_grid(sys) * modes_use * 64 + (int64_t)modes_use * 64;   // coefficients: the block partials in, the image out
}

// boundary damping of the modes, once per surface assembly
void modal_build_adiag(fdb_system* sys) {
    if (sys->have_adiag) return;
    FDB_REQUIRE(sys->have_modes && sys->modal_x.p, "FP64 modes are needed for the boundary damping terms");
    const int ng = std::max(1, sys->n_groups);
    sys->modal_adiag.alloc((size_t)ng * sys->n_modes);
    sys->modal_adiag.zero(sys->stream);
    if (sys->n_groups > 0 && sys->have_overlay && sys->n_ov > 0)
        k_modal_adiag<<<sys->n_modes, 256, (size_t)256 * ng * sizeof(double), sys->stream>>>(sys->modal_x.p, sys->ov_ptr.p, sys->ov_cg.p, sys->ov_val.p,
                                                                                          sys->n_nodes, sys->n_groups, sys->n_modes, sys->modal_adiag.p);
    FDB_CUDA(cudaGetLastError());
    sys->have_adiag = true;
}

// ================================================================================================
// Modal set-up on the GPU: all modes of H v = lambda Q v below f_cut (FEM3D.eigenfrequency's job, femder/FEM_3D.py:1699-1768,
// which runs spsolve(Q, H) + ARPACK on the CPU).
//
// Chebyshev-filtered subspace iteration.  The filter is a polynomial in Qt^-1 H where Qt^-1 is a fixed, cheap approximation
// of the inverse mass matrix (no mass solves, no factorisation):
//     Qt^-1 = sum_{j <= t} (I - D^-1 Q)^j D^-1 ,   D = diag(Q) scaled to the total mass,
// the order-t Taylor expansion of 1/x around the point of the mass spectrum where smooth functions sit (D^-1 Q u ~ u).
// t = 0 is mass lumping: its low invariant subspace differs from that of the consistent pencil by O((k h)^2), too much for
// a mode that a drive frequency comes close to (the coupling its residual leaves is amplified by 1 / (lambda - w^2));
// t = 1, 2 cost one or two more SpMVs per filter step and leave O((k h)^4), O((k h)^6).  The first outer iteration runs
// lumped (random start), the following ones raise t until the residuals of the CONSISTENT pencil meet the tolerance.
// Every outer iteration ends with a Rayleigh-Ritz step on (X^T H X, X^T Q X), so eigenvalues, Q-orthonormality and the
// residuals that decide convergence are those of (H, Q).  H X and Q X come from the sweep's own fused SpMV (16 columns per
// launch; Q x = (H + s Q) x / s with s so large that H x / s is below rounding); the dense FP64 Gram products and the small
// symmetric-definite eigenproblem are cuBLAS / cuSOLVER calls (plain library GEMMs: set-up, not the solver iteration).
// ================================================================================================
#define FDB_CUBLAS(expr)                                                                              \
    do {                                                                                              \
        cublasStatus_t _s = (expr);                                                                   \
        if (_s != CUBLAS_STATUS_SUCCESS) throw Error(std::string("femder_b200: cuBLAS call failed: ") + #expr + " (" + std::to_string((int)_s) + ")"); \
    } while (0)
#define FDB_CUSOLVER(expr)                                                                            \
    do {                                                                                              \
        cusolverStatus_t _s = (expr);                                                                 \
        if (_s != CUSOLVER_STATUS_SUCCESS) throw Error(std::string("femder_b200: cuSOLVER call failed: ") + #expr + " (" + std::to_string((int)_s) + ")"); \
    } while (0)

void modal_free_handles(fdb_system* sys) {
    if (sys->cublas) cublasDestroy((cublasHandle_t)sys->cublas);
    if (sys->cusolver) cusolverDnDestroy((cusolverDnHandle_t)sys->cusolver);
    sys->cublas = sys->cusolver = nullptr;
}

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// batch layout [b][row][16]: column c of batch b is mode 16 b + c
__global__ void k_modal_random(double* __restrict__ X, int64_t n, uint32_t seed) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        X[i] = (double)hash32((uint32_t)i * 2654435761u + seed + (uint32_t)(i >> 32)) * (2.0 / 4294967296.0) - 1.0;
}
// row sums of Q and Gershgorin bound: qsum[row] = sum_k Q_k, gmax = max_row sum_k |H_k| / Q_rr   (SELL stream, internal order)
__global__ void __launch_bounds__(256) k_modal_rows(const int32_t* __restrict__ sell_off, const double2* __restrict__ sell_hq,
                                                    const double2* __restrict__ diag_hq, int64_t n_rows, double* __restrict__ qsum_rows,
                                                    unsigned long long* __restrict__ gmax_bits) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double g = 0.0;
    if (row < n_rows) {
        const int64_t slice = row >> 3;
        const int off = sell_off[slice], width = sell_off[slice + 1] - off;
        double ah = 0.0, sq = 0.0;
        for (int k = 0; k < width; ++k) {
            const double2 m = sell_hq[((int64_t)off + k) * 8 + (row & 7)];
            ah += fabs(m.x);
            sq += m.y;
        }
        qsum_rows[row] = sq;
        const double d = diag_hq[row].y;
        g = d > 0.0 ? ah / d : 0.0;
    }
    __shared__ double s_g[256];
    s_g[threadIdx.x] = g;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_g[threadIdx.x] = fmax(s_g[threadIdx.x], s_g[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicMax(gmax_bits, (unsigned long long)__double_as_longlong(s_g[0]));   // non-negative doubles order like integers
}
// u = g / d  (and v = u: start of the Taylor sum), d = mass_scale * Q_rr
__global__ void __launch_bounds__(256) k_mass_start(const double* __restrict__ g, const double2* __restrict__ diag_hq, double mass_scale, int64_t n_elems,
                                                    double* __restrict__ u, double* __restrict__ v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += (int64_t)gridDim.x * blockDim.x) {
        const double t = g[i] / (diag_hq[i >> 4].y * mass_scale);
        u[i] = t;
        if (v) v[i] = t;
    }
}
// v <- v - (qv * qscale) / d ; u += v     (qv = s Q v from the SpMV, qscale = 1 / s)
__global__ void __launch_bounds__(256) k_mass_term(const double* __restrict__ qv, double qscale, const double2* __restrict__ diag_hq, double mass_scale,
                                                   int64_t n_elems, double* __restrict__ v, double* __restrict__ u) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += (int64_t)gridDim.x * blockDim.x) {
        const double t = v[i] - qv[i] * qscale / (diag_hq[i >> 4].y * mass_scale);
        v[i] = t;
        u[i] += t;
    }
}
// one Chebyshev step on a batch: Ynew = c1 * AY + c2 * Y + c3 * Xold   (Ynew may alias Xold)
__global__ void __launch_bounds__(256) k_cheb_step(const double* __restrict__ AY, const double* __restrict__ Y, const double* Xold, double c1,
                                                   double c2, double c3, int64_t n_elems, double* Ynew) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += (int64_t)gridDim.x * blockDim.x) {
        double v = c1 * AY[i] + c2 * Y[i];
        if (c3 != 0.0) v += c3 * Xold[i];
        Ynew[i] = v;
    }
}
// the same with A Y = (H Y) / d formed on the fly (mass order 0: lumped): one pass less per filter step
__global__ void __launch_bounds__(256) k_cheb_step_lumped(const double* __restrict__ HY, const double* __restrict__ Y, const double* Xold,
                                                          const double2* __restrict__ diag_hq, double mass_scale, double c1, double c2, double c3,
                                                          int64_t n_elems, double* Ynew) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += (int64_t)gridDim.x * blockDim.x) {
        double v = c1 * (HY[i] / (diag_hq[i >> 4].y * mass_scale)) + c2 * Y[i];
        if (c3 != 0.0) v += c3 * Xold[i];
        Ynew[i] = v;
    }
}
__global__ void __launch_bounds__(256) k_modal_dinv(const double2* __restrict__ diag_hq, double mass_scale, int64_t n, double* __restrict__ dinv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dinv[i] = 1.0 / (diag_hq[i].y * mass_scale);
}
__global__ void __launch_bounds__(256) k_modal_scale(double* __restrict__ x, double a, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] *= a;
}
// batch [row][16] <-> columns [16 b, 16 b + 16) of the big row-major matrix [row][nb]
// two 16-column double panels <-> one 32-column single panel (rows stay 128 bytes: the filter's SpMV stages them unchanged)
__global__ void __launch_bounds__(256) k_pair_to_f32(const double* __restrict__ xa, const double* __restrict__ xb, int64_t n_rows,
                                                     float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows * 32; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i >> 5;
        const int c = (int)(i & 31);
        out[i] = (float)(c < 16 ? xa[row * 16 + c] : xb[row * 16 + c - 16]);
    }
}
__global__ void __launch_bounds__(256) k_pair_from_f32(const float* __restrict__ in, int64_t n_rows, double* __restrict__ xa,
                                                       double* __restrict__ xb) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows * 32; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i >> 5;
        const int c = (int)(i & 31);
        if (c < 16) xa[row * 16 + c] = (double)in[i]; else xb[row * 16 + c - 16] = (double)in[i];
    }
}

__global__ void __launch_bounds__(256) k_modal_to_big(const double* __restrict__ Xb, int64_t n_rows, int nb, int b, double* __restrict__ big) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows * 16; i += (int64_t)gridDim.x * blockDim.x)
        big[(i >> 4) * nb + 16 * b + (i & 15)] = Xb[i];
}
__global__ void __launch_bounds__(256) k_modal_from_big(const double* __restrict__ big, int64_t n_rows, int nb, int b, double* __restrict__ Xb) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows * 16; i += (int64_t)gridDim.x * blockDim.x)
        Xb[i] = big[(i >> 4) * nb + 16 * b + (i & 15)];
}
// per column of a batch: partial[block][c] = sum over the block's rows of (HX - theta_c QX)^2 and of QX^2 (fixed order)
__global__ void __launch_bounds__(256) k_modal_resid(const double* __restrict__ HX, const double* __restrict__ QX, const double* __restrict__ theta,
                                                     int64_t n_rows, double* __restrict__ partial /* [grid][32] */) {
    const int c = threadIdx.x & 15;
    const double th = theta[c];
    double a = 0.0, q = 0.0;
    const int64_t rpb = (n_rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * rpb, r1 = min(n_rows, r0 + rpb);
    for (int64_t r = r0 + (threadIdx.x >> 4); r < r1; r += 16) {
        const double h = HX[r * 16 + c], qq = QX[r * 16 + c];
        const double d = h - th * qq;
        a += d * d;
        q += qq * qq;
    }
    __shared__ double s_a[256], s_q[256];
    s_a[threadIdx.x] = a; s_q[threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.x < 16) {
        double ta = 0.0, tq = 0.0;
        for (int k = 0; k < 16; ++k) { ta += s_a[k * 16 + threadIdx.x]; tq += s_q[k * 16 + threadIdx.x]; }
        partial[(int64_t)blockIdx.x * 32 + threadIdx.x] = ta;
        partial[(int64_t)blockIdx.x * 32 + 16 + threadIdx.x] = tq;
    }
}
__global__ void k_modal_symmetrize(double* __restrict__ A, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i < n && j < i) {
        const double v = 0.5 * (A[(int64_t)j * n + i] + A[(int64_t)i * n + j]);
        A[(int64_t)j * n + i] = v;
        A[(int64_t)i * n + j] = v;
    }
}

// scale s of (H + s Q): s |Q| >> 1e16 |H|, so ((H + s Q) x) / s is Q x to the last bit that matters
static double modal_qscale(double b_up) { return b_up * 1e22; }

// res[i] = ||H x_i - theta_i Q x_i||_2 / ||Q x_i||_2 (eigenvalue units) of the first n columns of the batches X
// [bt0, bt1): the panels this rank measures (the others' entries of res stay 0 for the caller to sum over the ranks)
static void modal_residuals(fdb_system* sys, SweepWork* w, const double* X, const std::vector<double>& theta, int n, double s_q,
                            DevBuf<double>& HX, DevBuf<double>& QX, std::vector<double>& res, int bt0 = 0, int bt1 = 1 << 30) {
    const int64_t N = sys->n_nodes;
    cudaStream_t s = sys->stream;
    const size_t bsz = (size_t)N * 16;
    const int rgrid = (int)std::max<int64_t>(1, std::min<int64_t>(sys->num_sms * 2, (N + 255) / 256));
    const int fgrid = (int)std::min<int64_t>(grid_for(N * 16, 256), (int64_t)sys->num_sms * 8);
    DevBuf<double> dtheta, rpart;
    dtheta.alloc(16);
    rpart.alloc((size_t)rgrid * 32);
    HX.alloc(bsz); QX.alloc(bsz);
    std::vector<double> hp((size_t)rgrid * 32), th16(16);
    res.assign(n, 0.0);
    for (int bt = bt0; bt < bt1 && 16 * bt < n; ++bt) {
        const double* xb = X + (size_t)bt * bsz;
        modal_spmv_w2(sys, w, 0.0);
        modal_spmv(sys, w, xb, HX.p);
        modal_spmv_w2(sys, w, -s_q);
        modal_spmv(sys, w, xb, QX.p);
        k_modal_scale<<<fgrid, 256, 0, s>>>(QX.p, 1.0 / s_q, (int64_t)bsz);
        for (int c = 0; c < 16; ++c) th16[c] = (16 * bt + c < (int)theta.size()) ? theta[16 * bt + c] : 0.0;
        FDB_CUDA(cudaMemcpyAsync(dtheta.p, th16.data(), 16 * sizeof(double), cudaMemcpyHostToDevice, s));
        k_modal_resid<<<rgrid, 256, 0, s>>>(HX.p, QX.p, dtheta.p, N, rpart.p);
        FDB_CUDA(cudaMemcpyAsync(hp.data(), rpart.p, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
        FDB_CUDA(cudaStreamSynchronize(s));
        for (int c = 0; c < 16 && 16 * bt + c < n; ++c) {
            double ra = 0.0, rq = 0.0;
            for (int g = 0; g < rgrid; ++g) { ra += hp[(size_t)g * 32 + c]; rq += hp[(size_t)g * 32 + 16 + c]; }
            res[16 * bt + c] = std::sqrt(ra) / (std::sqrt(rq) + 1e-300);
        }
    }
    FDB_CUDA(cudaGetLastError());
}

// upper bound of the spectrum of D^-1 H (Gershgorin) and the total-mass scale of D = mass_scale * diag(Q)
static void modal_bounds(fdb_system* sys, double& mass_scale, double& b_up, double* total_mass = nullptr) {
    const int64_t N = sys->n_nodes;
    cudaStream_t s = sys->stream;
    DevBuf<double> qrow;
    DevBuf<unsigned long long> gbits;
    qrow.alloc(N);
    gbits.alloc(1);
    gbits.zero(s);
    k_modal_rows<<<grid_for(N, 256), 256, 0, s>>>(sys->sell_off.p, sys->sell_hq.p, sys->diag_hq.p, N, qrow.p, gbits.p);
    FDB_CUDA(cudaGetLastError());
    std::vector<double> h_qrow(N);
    std::vector<double2> h_diag(N);
    unsigned long long h_g = 0;
    FDB_CUDA(cudaMemcpyAsync(h_qrow.data(), qrow.p, N * sizeof(double), cudaMemcpyDeviceToHost, s));
    FDB_CUDA(cudaMemcpyAsync(h_diag.data(), sys->diag_hq.p, N * sizeof(double2), cudaMemcpyDeviceToHost, s));
    FDB_CUDA(cudaMemcpyAsync(&h_g, gbits.p, sizeof(h_g), cudaMemcpyDeviceToHost, s));
    FDB_CUDA(cudaStreamSynchronize(s));
    double mass = 0.0, trq = 0.0;
    for (int64_t i = 0; i < N; ++i) { mass += h_qrow[i]; trq += h_diag[i].y; }
    FDB_REQUIRE(mass > 0.0 && trq > 0.0, "the mass matrix has no positive total mass");
    mass_scale = mass / trq;
    if (total_mass) *total_mass = mass;
    double gersh;
    memcpy(&gersh, &h_g, sizeof(double));
    b_up = 1.01 * gersh / mass_scale;
    FDB_REQUIRE(b_up > 0.0 && std::isfinite(b_up), "could not bound the spectrum of D^-1 H");
}

// Weyl's estimate of the number of room modes below f: N(f) = 4 pi / 3 V (f / c)^3 + pi / 4 S (f / c)^2 + L / 8 (f / c), with the
// room's volume from the total mass (1^T Q 1 = V / (rho c^2)), its surface from the boundary triangles' areas when they are on the
// device (6 V^(2/3) otherwise) and the edge term from a cube of that volume.  0 when the matrices were given from outside (no c).
int modal_estimate(fdb_system* sys, double f_hz) {
    if (!(sys->c0 > 0.0) || !(sys->rho0 > 0.0) || !sys->have_hq) return 0;
    double mass_scale, b_up, mass = 0.0;
    modal_bounds(sys, mass_scale, b_up, &mass);
    const double vol = mass * sys->rho0 * sys->c0 * sys->c0;
    double area = 6.0 * std::pow(vol, 2.0 / 3.0);
    if (sys->n_tri > 0 && sys->areas.p && sys->areas.n == (size_t)sys->n_tri) {
        std::vector<double> a(sys->n_tri);
        FDB_CUDA(cudaMemcpy(a.data(), sys->areas.p, a.size() * sizeof(double), cudaMemcpyDeviceToHost));
        double sa = 0.0;
        for (double v : a) sa += v;
        if (sa > 0.0) area = sa;
    }
    const double x = f_hz / sys->c0;
    return (int)(4.0 * 3.14159265358979323846 / 3.0 * vol * x * x * x + 3.14159265358979323846 / 4.0 * area * x * x + 1.5 * std::cbrt(vol) * x) + 1;
}

// modal_x (FP64 batches) + modal_lam_host -> bf16 tiles, device eigenvalues, per-mode residuals of the consistent pencil
// which form of the modes the sweep streams: the tile-compressed one (default) or the full-resolution tiles
void modal_set_compression(fdb_system* sys, bool on, bool force) {
    FDB_REQUIRE(sys->n_modes > 0 && sys->modal_x.p != nullptr, "no modes");
    const int n_modes = sys->n_modes;
    if (on && !sys->have_coarse_v) {
        // tile bases and coefficient tiles (see "Tile-compressed modes")
        sys->nc_rows = sys->n_tiles * kCF;
        sys->n_ctiles = (sys->nc_rows + 127) / 128;
        sys->modal_phi.alloc((size_t)sys->n_tiles * (kPhiBytes / 2));
        const int nbat = (n_modes + 15) / 16;
        DevBuf<double> Xc;
        Xc.alloc((size_t)nbat * sys->nc_rows * 16);
        static thread_local bool attr_done[64] = {};
        if (!attr_done[sys->device & 63]) {
            FDB_CUDA(cudaFuncSetAttribute(k_tile_basis, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_basis_smem()));
            attr_done[sys->device & 63] = true;
        }
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        const bool verbose = getenv("FDB_MODAL_VERBOSE") != nullptr;
        if (verbose) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, sys->stream); }
        int n_iter = 6;
        if (const char* e = tune_env("FDB_TILE_ITERS")) n_iter = std::max(0, atoi(e));
        DevBuf<double> fit;
        fit.alloc((size_t)2 * n_modes);
        fit.zero(sys->stream);
        k_tile_basis<<<(unsigned)sys->n_tiles, kTbThreads, tile_basis_smem(), sys->stream>>>(sys->modal_x.p, sys->n_nodes, n_modes, n_iter, sys->modal_phi.p,
                                                                                            Xc.p, sys->nc_rows, fit.p);
        std::vector<double> hfit((size_t)2 * n_modes);
        FDB_CUDA(cudaMemcpyAsync(hfit.data(), fit.p, hfit.size() * sizeof(double), cudaMemcpyDeviceToHost, sys->stream));
        if (verbose) {
            cudaEventRecord(e1, sys->stream);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            fprintf(stderr, "[fdb modal] tile bases: %lld tiles x %d modes in %.1f ms\n", (long long)sys->n_tiles, n_modes, ms);
            cudaEventDestroy(e0); cudaEventDestroy(e1);
        }
        sys->modal_vc.alloc((size_t)sys->n_ctiles * sys->mode_chunks * (kChunkBytes / 2));
        const int64_t work = sys->n_ctiles * 128 * sys->mode_chunks * 16;
        k_modal_pack<<<grid_for(work, 256), 256, 0, sys->stream>>>(Xc.p, sys->nc_rows, sys->n_ctiles, n_modes, sys->mode_chunks, sys->modal_vc.p);
        FDB_CUDA(cudaGetLastError());
        FDB_CUDA(cudaStreamSynchronize(sys->stream));   // Xc goes out of scope
        sys->have_coarse_v = true;
        double worst = 0.0;
        for (int i2 = 0; i2 < n_modes; ++i2)
            if (hfit[2 * i2 + 1] > 0.0) worst = std::max(worst, std::sqrt(hfit[2 * i2] / hfit[2 * i2 + 1]));
        sys->modal_fit_worst = worst;
        if (getenv("FDB_MODAL_VERBOSE")) fprintf(stderr, "[fdb modal] tile-compressed modes: worst relative fit error %.3e\n", worst);
    }
    // 16 functions per tile describe the modes only while a tile is small against the shortest wavelength (fit error 3e-3 on the 1M-DOF
    // room's highest mode; at 1e-2 the sweep starts to pay in iterations, tests/research/compressed_modes_prototype.py): a mesh too
    // coarse for that keeps the full-resolution tiles
    if (on && !force && sys->modal_fit_worst > kMaxFitError) on = false;
    if (!on) modal_ensure_fine(sys);
    sys->modal_compress = on;
    invalidate_graph(sys);
}

void modal_finish(fdb_system* sys, int n_modes) {
    FDB_REQUIRE(sys->have_order && sys->tile_rows == 128, "the modal correction needs the 128-row tiles of the internal order");
    n_modes = std::min(n_modes, kModalMaxModes);
    sys->n_modes = n_modes;
    sys->modal_lam_host.resize(n_modes);
    sys->mode_chunks = (n_modes + kMC - 1) / kMC;
    sys->modal_lam.alloc(n_modes);
    FDB_CUDA(cudaMemcpyAsync(sys->modal_lam.p, sys->modal_lam_host.data(), n_modes * sizeof(double), cudaMemcpyHostToDevice, sys->stream));
    sys->have_fine_v = false;
    sys->have_coarse_v = false;
    {
        const char* e = tune_env("FDB_MODAL_COMPRESS");
        modal_set_compression(sys, !(e && atoi(e) == 0), e && atoi(e) == 2);
    }
    // How far each stored vector is from an eigenvector: k_modal_coef bounds 1 / (lambda - w^2) by it, so an inaccurate mode that a
    // drive frequency comes close to is deflated only as far as it can be trusted.
    double mass_scale, b_up;
    modal_bounds(sys, mass_scale, b_up);
    SweepWork* w = modal_spmv_begin(sys);
    DevBuf<double> HX, QX;
    std::vector<double> res;
    modal_residuals(sys, w, sys->modal_x.p, sys->modal_lam_host, n_modes, modal_qscale(b_up), HX, QX, res);
    if (const char* e = tune_env("FDB_MODAL_FLOOR")) {   // A/B: scale of the bound k_modal_coef keeps |lambda - w^2| above (res^2 / lambda)
        const double f = std::sqrt(std::max(0.0, atof(e)));
        for (auto& r : res) r *= f;
    }
    sys->modal_res_host = res;
    sys->modal_res.alloc(n_modes);
    FDB_CUDA(cudaMemcpyAsync(sys->modal_res.p, res.data(), n_modes * sizeof(double), cudaMemcpyHostToDevice, sys->stream));
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
    sys->have_modes = true;
    sys->have_adiag = false;
    invalidate_graph(sys);
}

void modal_set(fdb_system* sys, int n_modes, const double* lam, const double* V) {
    FDB_REQUIRE(sys->have_order && sys->have_hq, "set the pattern and H, Q before the modes");
    FDB_REQUIRE(n_modes > 0 && lam && V, "bad arguments");
    const int64_t N = sys->n_nodes;
    const int m = std::min(n_modes, kModalMaxModes);
    sys->modal_lam_host.resize(m);
    FDB_CUDA(cudaMemcpy(sys->modal_lam_host.data(), lam, m * sizeof(double), cudaMemcpyDefault));
    for (int i = 1; i < m; ++i) FDB_REQUIRE(sys->modal_lam_host[i] >= sys->modal_lam_host[i - 1], "eigenvalues must ascend");
    const int nb = (m + 15) / 16;
    sys->modal_x.alloc((size_t)nb * N * 16);
    sys->modal_x.zero(sys->stream);
    DevBuf<double> tmp;
    tmp.alloc(N);
    for (int i = 0; i < m; ++i) {
        FDB_CUDA(cudaMemcpyAsync(tmp.p, V + (int64_t)i * N, N * sizeof(double), cudaMemcpyDefault, sys->stream));
        k_modal_scatter<<<grid_for(N, 256), 256, 0, sys->stream>>>(tmp.p, sys->perm.p, N, i, sys->modal_x.p);
    }
    FDB_CUDA(cudaGetLastError());
    modal_finish(sys, m);
}

void modal_get(fdb_system* sys, double* lam, double* V) {
    FDB_REQUIRE(sys->have_modes, "no modes");
    const int64_t N = sys->n_nodes;
    if (lam) FDB_CUDA(cudaMemcpy(lam, sys->modal_lam_host.data(), sys->n_modes * sizeof(double), cudaMemcpyDefault));
    if (V) {
        FDB_REQUIRE(sys->modal_x.p, "the FP64 modes were released");
        DevBuf<double> tmp;
        tmp.alloc(N);
        for (int i = 0; i < sys->n_modes; ++i) {
            k_modal_gather<<<grid_for(N, 256), 256, 0, sys->stream>>>(sys->modal_x.p, sys->perm.p, N, i, tmp.p);
            FDB_CUDA(cudaMemcpyAsync(V + (int64_t)i * N, tmp.p, N * sizeof(double), cudaMemcpyDefault, sys->stream));
        }
        FDB_CUDA(cudaStreamSynchronize(sys->stream));
    }
}

// f_cut_hz: every mode below it is kept.  f_acc_hz (<= f_cut_hz): the modes below it are the ones a drive frequency can come
// close to - they must meet `tol` (relative residual ||H x - theta Q x|| / (theta ||Q x||) of the consistent pencil); the modes
// between f_acc and f_cut only ever see |lambda - w^2| >= lambda_cut - lambda_acc and may stay an order of magnitude coarser.
void modal_compute(fdb_system* sys, double f_cut_hz, double f_acc_hz, int n_hint, double tol, int max_outer, int* n_found) {
    FDB_REQUIRE(sys->have_order && sys->have_hq, "set the pattern and H, Q before the modes");
    FDB_REQUIRE(sys->tile_rows == 128 && sys->n_tiles > 0, "the modal set-up needs the 128-row tiles of the internal order");
    FDB_REQUIRE(f_cut_hz > 0.0, "f_cut_hz must be positive");
    if (tol <= 0.0) tol = 4e-2;   // what the coarse correction needs (tests/research/modal_deflation_variants.py); modal analysis asks for less
    if (max_outer <= 0) max_outer = 12;
    if (f_acc_hz <= 0.0 || f_acc_hz > f_cut_hz) f_acc_hz = f_cut_hz;
    const int64_t N = sys->n_nodes;
    cudaStream_t s = sys->stream;
    const double two_pi = 6.283185307179586;
    const double lam_cut = (two_pi * f_cut_hz) * (two_pi * f_cut_hz);
    const double lam_acc = (two_pi * f_acc_hz) * (two_pi * f_acc_hz);
    if (!sys->cublas) {
        cublasHandle_t h;
        FDB_CUBLAS(cublasCreate(&h));
        sys->cublas = h;
    }
    if (!sys->cusolver) {
        cusolverDnHandle_t h;
        FDB_CUSOLVER(cusolverDnCreate(&h));
        sys->cusolver = h;
    }
    cublasHandle_t cb = (cublasHandle_t)sys->cublas;
    cusolverDnHandle_t cs = (cusolverDnHandle_t)sys->cusolver;
    FDB_CUBLAS(cublasSetStream(cb, s));
    FDB_CUSOLVER(cusolverDnSetStream(cs, s));

    double mass_scale, b_up;
    modal_bounds(sys, mass_scale, b_up);
    const double s_q = modal_qscale(b_up);

    // block size: the caller's estimate of the mode count plus a guard band, in batches of 16 columns
    const int cap_cols = (int)std::min<int64_t>(kModalMaxModes + 128, std::max<int64_t>(16, (N / 3 / 16) * 16));
    if (n_hint <= 0) n_hint = modal_estimate(sys, f_cut_hz);   // 0 if the room's size is unknown: start small and widen
    int nb = n_hint > 0 ? ((int)(1.12 * n_hint) + 24 + 15) / 16 * 16 : 64;
    nb = std::max(16, std::min(nb, cap_cols));

    // several ranks sharing the set-up (fdb_comm_init): the block's panels are dealt to them in pairs (the single-precision filter
    // takes two panels at once); a rank filters, multiplies, rotates and measures its own panels, the block itself is exchanged
    RankComm* comm = (RankComm*)sys->comm;
    const int c_rank = comm_rank(comm), c_world = comm_world(comm);
    int own0 = 0, own1 = 0;                 // this rank's panels [own0, own1) of the current block
    std::vector<size_t> own_offs;           // panel offsets of all ranks, in doubles of X
    DevBuf<double> dres;
    // panels [lo, hi) dealt to the ranks in pairs
    auto deal_panels = [&](int lo, int hi) {
        const int pairs = (hi - lo + 1) / 2;
        own_offs.assign(c_world + 1, 0);
        for (int r = 0; r <= c_world; ++r) own_offs[r] = (size_t)(lo + std::min(hi - lo, 2 * (int)((int64_t)pairs * r / c_world))) * (size_t)N * 16;
        own0 = (int)(own_offs[c_rank] / ((size_t)N * 16));
        own1 = (int)(own_offs[c_rank + 1] / ((size_t)N * 16));
    };
    // Locking (A/B knob, off): the leading Ritz vectors whose residuals are well below the tolerance are not filtered again (they
    // still take part in every Rayleigh-Ritz step).  It saves filter passes and costs the sweep more than it saves: on the 1M-DOF room,
    // locking the 128 lowest modes in the last outer iteration alone (residuals < 5e-3) took the sweep from 2.05 s to 2.80 s - the
    // solver's iteration count keeps falling with the accuracy of the modes well below the tolerance the set-up stops at.
    // lock_bat = locked panels (a multiple of 2: the single-precision filter takes pairs).
    int lock_bat = 0;
    SweepWork* w = modal_spmv_begin(sys);
    const bool verbose = getenv("FDB_MODAL_VERBOSE") != nullptr && c_rank == 0;   // progress lines on stderr; changes nothing that is computed
    std::vector<double> theta, res;
    int wanted = 0, wanted_prev = -1, outer = 0, order_t = 0;
    const int max_order = 4;
    double worst = 0.0, worst_prev = 1e300;
    DevBuf<double> X, Y, U, V1, T, big_x, big_hx, big_qx, big_new, Hs, Qs, W;
    DevBuf<int> dinfo;
    dinfo.alloc(1);
    const int fgrid = (int)std::min<int64_t>(grid_for(N * 16, 256), (int64_t)sys->num_sms * 8);
    bool have_ritz = false, fused_cheb = true;
    // The filter only has to ENRICH the block in the wanted directions - what comes out is measured by the Rayleigh-Ritz step and the
    // residuals, both in double precision - so it starts on single-precision copies of the columns, 32 to the SpMV's 128-byte row
    // instead of 16: half the passes.  Once the worst residual is below f32_until the filter runs in double precision: measured on the
    // 1M-DOF room, modes finished by a single-precision filter meet the same residual test but cost the sweep a third more
    // iterations (95 -> 129 mean); single precision for the first outer iterations only costs none.
    auto knob = [](const char* name, double dflt) { const char* e = tune_env(name); return e ? atof(e) : dflt; };
    bool use_f32 = knob("FDB_MODAL_F64", 0.0) == 0.0;
    const double f32_until = knob("FDB_MODAL_F32_UNTIL", 1.0);
    double worst_last = 1e300;   // the previous outer iteration's worst residual
    const bool locking = knob("FDB_MODAL_LOCK", 0.0) != 0.0;   // off: measured below
    const double lock_tol = knob("FDB_MODAL_LOCK_TOL", tol / 8.0);
    const bool f32_active = knob("FDB_MODAL_F32_ACTIVE", 0.0) != 0.0;   // single precision for unlocked columns throughout
    const double first_mult = knob("FDB_MODAL_DEG1", 2.0);
    const double late_mult = knob("FDB_MODAL_DEG_LATE", 1.0);   // degree of the double-precision passes
    const double growth = knob("FDB_MODAL_GROWTH", 0.05);
    DevBuf<double> dinv_mass;   // 1 / (mass_scale * Q_rr)
    dinv_mass.alloc(N);
    k_modal_dinv<<<grid_for(N, 256), 256, 0, s>>>(sys->diag_hq.p, mass_scale, N, dinv_mass.p);
    double a_low = lam_cut * 1.21;   // lower edge of the interval the filter damps (1.1 f_cut until the block's own upper edge is known)
    const size_t bsz = (size_t)N * 16;
    // u = Qt^-1 (H y) for one batch: order_t extra SpMVs with Q
    auto apply_a = [&](const double* y, double* u) {
        modal_spmv_w2(sys, w, 0.0);
        modal_spmv(sys, w, y, T.p);
        k_mass_start<<<fgrid, 256, 0, s>>>(T.p, sys->diag_hq.p, mass_scale, (int64_t)bsz, u, order_t > 0 ? V1.p : nullptr);
        if (order_t > 0) modal_spmv_w2(sys, w, -s_q);
        for (int j = 0; j < order_t; ++j) {
            modal_spmv(sys, w, V1.p, T.p);
            k_mass_term<<<fgrid, 256, 0, s>>>(T.p, 1.0 / s_q, sys->diag_hq.p, mass_scale, (int64_t)bsz, V1.p, u);
        }
    };
    for (;;) {
        const int nbat = nb / 16;
        if (!have_ritz) {
            X.alloc((size_t)nbat * bsz);
            k_modal_random<<<fgrid, 256, 0, s>>>(X.p, (int64_t)nbat * bsz, 12345u);
        }
        Y.alloc(bsz); U.alloc(bsz); V1.alloc(bsz); T.alloc(bsz);
        big_x.alloc((size_t)N * nb); big_hx.alloc((size_t)N * nb); big_qx.alloc((size_t)N * nb); big_new.alloc((size_t)N * nb);
        Hs.alloc((size_t)nb * nb); Qs.alloc((size_t)nb * nb); W.alloc(nb);
        int lwork = 0;
        FDB_CUSOLVER(cusolverDnDsygvd_bufferSize(cs, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, nb, Hs.p, nb, Qs.p, nb,
                                                 W.p, &lwork));
        DevBuf<double> work;
        work.alloc(std::max(1, lwork));
        bool grown = false, converged = false;
        for (; outer < max_outer && !converged;) {
            double t_ph[5] = {0, 0, 0, 0, 0};
            auto tick = [&]() {   // phase timing for the progress lines only (synchronises: not in normal runs)
                if (!verbose) return 0.0;
                cudaStreamSynchronize(s);
                timespec ts;
                clock_gettime(CLOCK_MONOTONIC, &ts);
                return ts.tv_sec + 1e-9 * ts.tv_nsec;
            };
            double t_a = tick();
            // ---- filter: X <- p(Qt^-1 H) X, degree from the ratio of the interval to damp and the interval to keep ----
            // sum_{j <= t} (1 - x)^j <= t + 1 on the mass spectrum x in (0, 1]: a safe upper end for the spectrum of Qt^-1 H (components
            // above the filter's interval would be amplified, not damped)
            const double a = a_low, b = b_up * (1.0 + order_t);
            // the first pass starts from noise: twice the degree there replaces the first two passes and the Rayleigh-Ritz step between
            // them (the amplification ratio between the lowest modes and those next to the cut stays inside single precision)
            const double deg_mult = (outer == 0 && !have_ritz) ? first_mult : ((use_f32 && worst_last > f32_until) ? 1.0 : late_mult);
            const int deg = (int)std::max(12.0, std::min(160.0, std::ceil(3.0 * deg_mult * std::sqrt(b / a))));
            const double e = 0.5 * (b - a), c = 0.5 * (b + a);
            // two panels at once in single precision (see use_f32): T and Y hold the 32-column single-precision iterates
            auto filter_pair_f32 = [&](int bt) -> bool {
                double* xa = X.p + (size_t)bt * bsz;
                double* xb2 = xa + bsz;
                k_pair_to_f32<<<fgrid, 256, 0, s>>>(xa, xb2, N, (float*)T.p);
                double sigma = e / (0.0 - c);
                const double sigma1 = sigma;
                modal_spmv_w2(sys, w, 0.0);
                if (!modal_spmv_cheb(sys, w, T.p, nullptr, Y.p, dinv_mass.p, sigma1 / e, -c * sigma1 / e, 0.0, true)) return false;
                double *px = T.p, *py = Y.p;
                for (int i = 2; i <= deg; ++i) {
                    const double sigma2 = 1.0 / (2.0 / sigma1 - sigma);
                    modal_spmv_cheb(sys, w, py, px, px, dinv_mass.p, 2.0 * sigma2 / e, -2.0 * sigma2 * c / e, -sigma * sigma2, true);
                    std::swap(px, py);
                    sigma = sigma2;
                }
                k_pair_from_f32<<<fgrid, 256, 0, s>>>((const float*)py, N, xa, xb2);
                return true;
            };
            deal_panels(lock_bat, nbat);   // the columns to filter
            for (int bt = own0; bt < own1; ++bt) {
                if (use_f32 && (worst_last > f32_until || f32_active) && order_t == 0 && fused_cheb && bt + 1 < own1) {
                    if (filter_pair_f32(bt)) { ++bt; continue; }
                    use_f32 = false;
                }
                double* xb = X.p + (size_t)bt * bsz;
                double sigma = e / (0.0 - c);
                const double sigma1 = sigma;
                // one filter step: Ynew = c1 A Y + c2 Y + c3 Xold
                auto step = [&](const double* y, const double* xold, double c1, double c2, double c3, double* out) {
                    if (order_t == 0) {
                        // lumped mass: the whole step inside the SpMV's epilogue where the mesh's tiles allow it
                        if (fused_cheb && modal_spmv_cheb(sys, w, y, xold, out, dinv_mass.p, c1, c2, c3)) return;
                        fused_cheb = false;
                        modal_spmv(sys, w, y, T.p);
                        k_cheb_step_lumped<<<fgrid, 256, 0, s>>>(T.p, y, xold, sys->diag_hq.p, mass_scale, c1, c2, c3, (int64_t)bsz, out);
                    } else {
                        apply_a(y, U.p);
                        k_cheb_step<<<fgrid, 256, 0, s>>>(U.p, y, xold, c1, c2, c3, (int64_t)bsz, out);
                    }
                };
                if (order_t == 0) modal_spmv_w2(sys, w, 0.0);
                step(xb, nullptr, sigma1 / e, -c * sigma1 / e, 0.0, Y.p);   // Y = (sigma1 / e)(A X - c X)
                double *px = xb, *py = Y.p;   // px = previous iterate, py = current one
                for (int i = 2; i <= deg; ++i) {
                    const double sigma2 = 1.0 / (2.0 / sigma1 - sigma);
                    // Ynew = (2 sigma2 / e)(A Y - c Y) - sigma sigma2 X, written over X (its last use)
                    step(py, px, 2.0 * sigma2 / e, -2.0 * sigma2 * c / e, -sigma * sigma2, px);
                    std::swap(px, py);
                    sigma = sigma2;
                }
                if (py != xb) FDB_CUDA(cudaMemcpyAsync(xb, py, bsz * sizeof(double), cudaMemcpyDeviceToDevice, s));
            }
            FDB_CUDA(cudaGetLastError());
            { const double t_b = tick(); t_ph[0] = t_b - t_a; t_a = t_b; }
            // ---- Rayleigh-Ritz on (H, Q) ----
            comm_allgatherv(comm, X.p, own_offs, s);         // every rank needs the whole filtered block on the left of the Gram products
            deal_panels(0, nbat);                            // the columns of the Gram products, the rotation and the residuals: all of them
            for (int bt = 0; bt < nbat; ++bt) {
                double* xb = X.p + (size_t)bt * bsz;
                k_modal_to_big<<<fgrid, 256, 0, s>>>(xb, N, nb, bt, big_x.p);
                if (bt < own0 || bt >= own1) continue;
                modal_spmv_w2(sys, w, 0.0);
                modal_spmv(sys, w, xb, Y.p);                 // H X
                modal_spmv_w2(sys, w, -s_q);
                modal_spmv(sys, w, xb, T.p);                 // s Q X
                k_modal_scale<<<fgrid, 256, 0, s>>>(T.p, 1.0 / s_q, (int64_t)bsz);
                k_modal_to_big<<<fgrid, 256, 0, s>>>(Y.p, N, nb, bt, big_hx.p);
                k_modal_to_big<<<fgrid, 256, 0, s>>>(T.p, N, nb, bt, big_qx.p);
            }
            FDB_CUDA(cudaGetLastError());
            const double one = 1.0, zero = 0.0;
            // row-major [N][nb] = column-major (nb x N): Hs = X^T (H X), Qs = X^T (Q X) - this rank's columns of them
            const int j0 = 16 * own0, jn = 16 * (own1 - own0);
            if (jn > 0) {
                FDB_CUBLAS(cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_T, nb, jn, (int)N, &one, big_x.p, nb, big_hx.p + j0, nb, &zero, Hs.p + (size_t)j0 * nb, nb));
                FDB_CUBLAS(cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_T, nb, jn, (int)N, &one, big_x.p, nb, big_qx.p + j0, nb, &zero, Qs.p + (size_t)j0 * nb, nb));
            }
            if (c_world > 1) {
                std::vector<size_t> so(c_world + 1);
                for (int r = 0; r <= c_world; ++r) so[r] = own_offs[r] / ((size_t)N) * (size_t)nb;   // 16 columns per panel x nb rows
                comm_allgatherv(comm, Hs.p, so, s);
                comm_allgatherv(comm, Qs.p, so, s);
            }
            { const double t_b = tick(); t_ph[1] = t_b - t_a; t_a = t_b; }
            k_modal_symmetrize<<<dim3((nb + 127) / 128, nb), 128, 0, s>>>(Hs.p, nb);
            k_modal_symmetrize<<<dim3((nb + 127) / 128, nb), 128, 0, s>>>(Qs.p, nb);
            FDB_CUSOLVER(cusolverDnDsygvd(cs, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, nb, Hs.p, nb, Qs.p, nb, W.p,
                                          work.p, lwork, dinfo.p));
            int info = 0;
            theta.resize(nb);
            // every rank solved the same small eigenproblem; rank 0's answer is the one all of them use (identical decisions below)
            comm_bcast(comm, Hs.p, (size_t)nb * nb, 0, s);
            comm_bcast(comm, W.p, (size_t)nb, 0, s);
            FDB_CUDA(cudaMemcpyAsync(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost, s));
            FDB_CUDA(cudaMemcpyAsync(theta.data(), W.p, nb * sizeof(double), cudaMemcpyDeviceToHost, s));
            FDB_CUDA(cudaStreamSynchronize(s));
            FDB_REQUIRE(info == 0, "the Rayleigh-Ritz eigenproblem of the modal set-up failed (the filtered block lost rank)");
            { const double t_b = tick(); t_ph[2] = t_b - t_a; t_a = t_b; }
            // X <- X Y (Ritz vectors, Q-orthonormal, ascending): (X Y)^T = Y^T X^T in the column-major view
            if (jn > 0)
                FDB_CUBLAS(cublasDgemm(cb, CUBLAS_OP_T, CUBLAS_OP_N, jn, (int)N, nb, &one, Hs.p + (size_t)j0 * nb, nb, big_x.p, nb, &zero, big_new.p + j0, nb));
            for (int bt = own0; bt < own1; ++bt) k_modal_from_big<<<fgrid, 256, 0, s>>>(big_new.p, N, nb, bt, X.p + (size_t)bt * bsz);
            comm_allgatherv(comm, X.p, own_offs, s);         // the Ritz vectors: every rank holds all of them again
            have_ritz = true;
            { const double t_b = tick(); t_ph[3] = t_b - t_a; t_a = t_b; }
            ++outer;
            wanted = (int)(std::upper_bound(theta.begin(), theta.end(), lam_cut) - theta.begin());
            if (wanted >= nb - 8 && nb < cap_cols) { grown = true; break; }   // the block holds no guard band: widen it
            // ---- residuals of the wanted pairs on the consistent pencil ----
            modal_residuals(sys, w, X.p, theta, wanted, s_q, Y, T, res, own0, own1);
            if (c_world > 1 && wanted > 0) {
                dres.alloc(wanted);
                FDB_CUDA(cudaMemcpyAsync(dres.p, res.data(), wanted * sizeof(double), cudaMemcpyHostToDevice, s));
                comm_allreduce_sum(comm, dres.p, wanted, s);
                FDB_CUDA(cudaMemcpyAsync(res.data(), dres.p, wanted * sizeof(double), cudaMemcpyDeviceToHost, s));
                FDB_CUDA(cudaStreamSynchronize(s));
            }
            const double th_ref = (two_pi * 10.0) * (two_pi * 10.0);
            worst = 0.0;
            double worst_hi = 0.0;
            for (int i = 0; i < wanted; ++i) {
                const double rel = res[i] / std::max(std::fabs(theta[i]), th_ref);
                if (theta[i] <= lam_acc) worst = std::max(worst, rel); else worst_hi = std::max(worst_hi, rel);
            }
            // next filter: damp everything above the guard band.  (The block's own upper Ritz value is the textbook choice, but it
            // starts far above the cut and comes down slowly; the columns between the cut and it need no protection.)
            a_low = lam_cut * 1.1;
            if (verbose) {
                t_ph[4] = tick() - t_a;
                fprintf(stderr, "[fdb modal] outer %d: block %d (%d columns locked), wanted %d, mass order %d, %s filter of degree %d, worst rel residual %.3e (below f_acc) %.3e (above), top Ritz f %.1f Hz"
                                " | seconds: filter %.3f, H X / Q X + Gram %.3f, eigenproblem %.3f, rotation %.3f, residuals %.3f\n",
                        outer, nb, 16 * lock_bat, wanted, order_t, (use_f32 && (worst_last > f32_until || f32_active)) ? "single-precision" : "double-precision", deg, worst, worst_hi, std::sqrt(std::max(theta[nb - 1], 0.0)) / two_pi, t_ph[0], t_ph[1], t_ph[2],
                        t_ph[3], t_ph[4]);
            }
            worst_last = worst;
            {   // leading modes good enough to lock
                int nl = 0;
                while (nl < wanted && res[nl] / std::max(std::fabs(theta[nl]), th_ref) < lock_tol) ++nl;
                lock_bat = std::max(lock_bat, (nl / 32) * 2);
                if (!locking) lock_bat = 0;
            }
            const bool stable = wanted == wanted_prev;   // the set of modes below the cut has stopped changing
            const int prev_wanted = wanted_prev;
            wanted_prev = wanted;
            // The modes below f_acc decide (a drive frequency can come close to them, and every one of them must be in the block or G
            // keeps a negative direction the correction knows nothing about).  The ones between f_acc and f_cut only ever see
            // lambda - w^2 > 0: whatever the block holds of them helps, none of them is needed.
            // ... but the block should have stopped collecting them: fewer modes between f_acc and f_cut cost the sweep iterations).
            if (outer >= 2 && worst <= tol && wanted - prev_wanted <= std::max(1, (int)(wanted * growth))) { converged = true; break; }
            // Once the subspace is in its asymptotic regime the mass approximation of the filter limits what it can reach: raise its
            // order when progress stalls, stop when the highest order stalls too (the residuals are reported; the sweep does not need
            // them small, see k_modal_coef).
            if (stable && worst < 0.2 && worst > 0.5 * worst_prev) {
                if (use_f32) use_f32 = false;
                else if (order_t < max_order) ++order_t;
                else { converged = true; break; }
            }
            worst_prev = worst;
        }
        if (!grown) break;
        // widen: keep the Ritz vectors, append random columns
        const int nb2 = std::min(cap_cols, std::max(nb + 64, (int)(nb * 1.6) / 16 * 16));
        DevBuf<double> X2;
        X2.alloc((size_t)(nb2 / 16) * bsz);
        k_modal_random<<<fgrid, 256, 0, s>>>(X2.p, (int64_t)(nb2 / 16) * bsz, 777u + (uint32_t)nb2);
        FDB_CUDA(cudaMemcpyAsync(X2.p, X.p, (size_t)nbat * bsz * sizeof(double), cudaMemcpyDeviceToDevice, s));
        FDB_CUDA(cudaStreamSynchronize(s));
        std::swap(X.p, X2.p); std::swap(X.n, X2.n);
        nb = nb2;
        a_low = lam_cut * 1.21;
        order_t = 0;
        worst_prev = 1e300;
        if (outer >= max_outer) break;
    }
    FDB_REQUIRE(wanted > 0, "no mode below f_cut");
    const int m = std::min(wanted, kModalMaxModes);
    sys->modal_lam_host.assign(theta.begin(), theta.begin() + m);
    const int keep_bat = (m + 15) / 16;
    sys->modal_x.alloc((size_t)keep_bat * N * 16);
    FDB_CUDA(cudaMemcpyAsync(sys->modal_x.p, X.p, (size_t)keep_bat * N * 16 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    FDB_CUDA(cudaStreamSynchronize(s));
    modal_finish(sys, m);
    sys->modal_outer = outer;
    sys->modal_resid = worst;
    sys->modal_block = nb;
    sys->modal_order = order_t;
    if (n_found) *n_found = m;
}

}  // namespace fdb

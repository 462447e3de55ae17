// K5 fused multi-frequency SpMV, K6 COCG vector kernels, K7 receiver gather, and the sweep driver.
//
// Replaces the hot loop of FEM3D.compute() (femder/FEM_3D.py:1304-1319): per frequency
//     G = H + 1j*w*sum_g mu_g A_g - w^2 Q ;  b = -1j*w*q ;  p = solve(G, b)
// The reference factorises G from scratch with PARDISO per frequency.  Here a batch of F frequencies is
// solved together by preconditioned COCG (conjugate orthogonal CG for complex symmetric G); the batch
// shares one pass over the column indices and the (H, Q) values per SpMV.
//
// HBM layout
//   hq[nnz]        double2 (H, Q) in CSR order, indices[nnz] int32, indptr[N+1] int32
//   overlay        ov_ptr[N+1], ov_cg[n_ov] (col, group), ov_val[n_ov]  boundary admittance entries
//   vectors        [N][F] complex128 (double2), frequency fastest: a gather of x[col][0..F) is one
//                  contiguous 16*F-byte segment (a full 128-byte line at F = 8)
#include "solver.cuh"
#include <algorithm>
#include <array>
#include <map>
#include <cstdlib>
#include <type_traits>

namespace fdb {
void free_sweep_work(SweepWork* w) {
    if (!w) return;
    if (w->graph) cudaGraphExecDestroy(w->graph);
    delete w;
}

void invalidate_graph(fdb_system* sys) {
    SweepWork* w = sys->work;
    if (w && w->graph) {
        cudaGraphExecDestroy(w->graph);
        w->graph = nullptr;
    }
}


// ------------------------------------------------------------------------------------------------
// K5: y[row][f] = sum_k (H_k - w2_f Q_k) x[col_k][f] + sum_ov (coef[g][f] * A_ov) x[col_ov][f]
//
// lanes = (row in warp, f): 32/F rows per warp.  The matrix is streamed from the SELL-8 copy, so the F-lane
// groups of a warp read consecutive entries (coalesced, each byte fetched once).  The x gather of one
// (row, k) is one contiguous 16*F-byte segment.  Every block owns one contiguous range of rows, so the x
// lines of neighbouring mesh lines / planes are re-used out of L1 instead of being re-fetched from L2.
// DOT: also accumulate x[row]^T y[row] (unconjugated) -> sc.pq
// ------------------------------------------------------------------------------------------------
// launch shapes (threads per block, resident blocks per SM); FDB_SPMV_CFG picks one at run time (tuning aid)
struct SpmvCfg { int block, per_sm; };
constexpr SpmvCfg kSpmvCfgs[] = {{512, 2}};   // launch shape of the plain kernel (measured best of 5)

__device__ __forceinline__ double2 ld_stream(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

template <typename T, int F, bool DOT, bool OVERLAY, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_spmv(const int32_t* __restrict__ sell_off, const int32_t* __restrict__ sell_idx, const double2* __restrict__ sell_hq,
       const int32_t* __restrict__ ov_ptr, const int2* __restrict__ ov_cg, const double* __restrict__ ov_val,
       const double2* __restrict__ coef, const T* __restrict__ x, T* __restrict__ y, int64_t n_rows,
       Scalars* __restrict__ sc, double* __restrict__ partial) {
    static_assert(!(OVERLAY && std::is_same<T, double>::value), "a boundary term makes the system complex");
    constexpr int RPW = 32 / F;                 // rows per warp
    constexpr int NWARP = BLOCK / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = lane % F;
    const int rl = lane / F;
    const double w2 = sc->w2[f];
    const int64_t n_groups = (n_rows + RPW - 1) / RPW;
    const int64_t gpb = (n_groups + gridDim.x - 1) / gridDim.x;   // contiguous groups per block
    const int64_t g0 = (int64_t)blockIdx.x * gpb;
    const int64_t g1 = min(g0 + gpb, n_groups);
    double dot[2] = {0.0, 0.0};
    for (int64_t g = g0 + warp; g < g1; g += NWARP) {
        const int64_t row = g * RPW + rl;
        if (row < n_rows) {
            const int64_t slice = row >> 3;
            const int off = sell_off[slice];
            const int width = sell_off[slice + 1] - off;
            const int32_t* ip = sell_idx + ((int64_t)off * 8 + (row & 7));
            const double2* mp = sell_hq + ((int64_t)off * 8 + (row & 7));
            T acc = v_zero<T>();
            int k = 0;
            for (; k + 4 <= width; k += 4) {
                int c[4];
                double2 m[4];
                T xv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) c[u] = __ldg(ip + (k + u) * 8);
#pragma unroll
                for (int u = 0; u < 4; ++u) xv[u] = __ldg(&x[(int64_t)c[u] * F + f]);
#pragma unroll
                for (int u = 0; u < 4; ++u) m[u] = ld_stream(mp + (k + u) * 8);
#pragma unroll
                for (int u = 0; u < 4; ++u) acc = v_add(acc, r_mul(m[u].x - w2 * m[u].y, xv[u]));
            }
            for (; k < width; ++k) {
                const int c = __ldg(ip + k * 8);
                const double2 m = ld_stream(mp + k * 8);
                acc = v_add(acc, r_mul(m.x - w2 * m.y, __ldg(&x[(int64_t)c * F + f])));
            }
            if constexpr (OVERLAY) {
                const int ob = ov_ptr[row], oe = ov_ptr[row + 1];
                for (int o = ob; o < oe; ++o) {
                    const int2 cg = ov_cg[o];
                    const double a = ov_val[o];
                    const double2 cf = coef[cg.y * F + f];
                    const double2 xv = __ldg(&x[(int64_t)cg.x * F + f]);
                    acc = v_add(acc, cmul(make_double2(cf.x * a, cf.y * a), xv));
                }
            }
            y[row * F + f] = acc;
            if (DOT) {
                const double2 d = v_dot(__ldg(&x[row * F + f]), acc);
                dot[0] += d.x;
                dot[1] += d.y;
            }
        }
    }
    if (DOT) {
        block_reduce_finalize<F, 2, BLOCK>(dot, partial, &sc->ticket[0], [&](int ff, const double* tot) {
            sc->pq[ff] = make_double2(tot[0], tot[1]);
        });
    }
}

// ------------------------------------------------------------------------------------------------
// K5 for NARROW real batches (F = 2, 4, 8: the tail of a sweep, when the last systems run in a narrowing batch):
// one lane per ROW, the row's F values in registers.  The plain kernel above gives every (row, f) its own lane, so F
// lanes fetch the same matrix entry and a warp's matrix loads cover 32/F rows only - at F = 2..4 it is bound by load
// instructions, not by HBM (197 us per iteration at F = 2 against 88 us at F = 1).  Here the matrix stream is read
// exactly as at F = 1 (fully coalesced, each lane its own row of a SELL-8 slice) and one 16/32-byte load gathers
// x[col][0..F).
// ------------------------------------------------------------------------------------------------
struct __align__(32) d4x { double v[4]; };
template <int FPL> struct RowVec { double v[FPL]; };
template <int FPL>
__device__ __forceinline__ RowVec<FPL> ld_row(const double* p) {
    RowVec<FPL> r;
    if constexpr (FPL == 2) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(p));
        r.v[0] = t.x; r.v[1] = t.y;
    } else {
#pragma unroll
        for (int h = 0; h < FPL / 4; ++h)
            asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
                         : "=d"(r.v[4 * h]), "=d"(r.v[4 * h + 1]), "=d"(r.v[4 * h + 2]), "=d"(r.v[4 * h + 3]) : "l"(p + 4 * h));
    }
    return r;
}
template <int FPL>
__device__ __forceinline__ void st_row(double* p, const RowVec<FPL>& r) {
    if constexpr (FPL == 2) {
        *reinterpret_cast<double2*>(p) = make_double2(r.v[0], r.v[1]);
    } else {
#pragma unroll
        for (int h = 0; h < FPL / 4; ++h) {
            d4x t;
#pragma unroll
            for (int j = 0; j < 4; ++j) t.v[j] = r.v[4 * h + j];
            *reinterpret_cast<d4x*>(p + 4 * h) = t;
        }
    }
}

template <int FPL, bool DOT, int BLOCK, int MINB, int UNR>
__global__ void __launch_bounds__(BLOCK, MINB)
k_spmv_rows(const int32_t* __restrict__ sell_off, const int32_t* __restrict__ sell_idx, const double2* __restrict__ sell_hq,
            const double* __restrict__ x, double* __restrict__ y, int64_t n_rows, Scalars* __restrict__ sc, double* __restrict__ partial) {
    constexpr int NWARP = BLOCK / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double w2[FPL];
#pragma unroll
    for (int j = 0; j < FPL; ++j) w2[j] = sc->w2[j];
    const int64_t n_groups = (n_rows + 31) / 32;
    const int64_t gpb = (n_groups + gridDim.x - 1) / gridDim.x;   // contiguous groups of 32 rows per block
    const int64_t g0 = (int64_t)blockIdx.x * gpb;
    const int64_t g1 = min(g0 + gpb, n_groups);
    double dot[FPL];
#pragma unroll
    for (int j = 0; j < FPL; ++j) dot[j] = 0.0;
    for (int64_t g = g0 + warp; g < g1; g += NWARP) {
        const int64_t row = g * 32 + lane;
        if (row < n_rows) {
            const int64_t slice = row >> 3;
            const int off = sell_off[slice];
            const int width = sell_off[slice + 1] - off;
            const int32_t* ip = sell_idx + ((int64_t)off * 8 + (row & 7));
            const double2* mp = sell_hq + ((int64_t)off * 8 + (row & 7));
            double acc[FPL];
#pragma unroll
            for (int j = 0; j < FPL; ++j) acc[j] = 0.0;
            int k = 0;
            for (; k + UNR <= width; k += UNR) {
                int c[UNR];
                double2 m[UNR];
                RowVec<FPL> xv[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) c[u] = __ldg(ip + (k + u) * 8);
#pragma unroll
                for (int u = 0; u < UNR; ++u) xv[u] = ld_row<FPL>(x + (int64_t)c[u] * FPL);
#pragma unroll
                for (int u = 0; u < UNR; ++u) m[u] = ld_stream(mp + (k + u) * 8);
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
#pragma unroll
                    for (int j = 0; j < FPL; ++j) acc[j] += (m[u].x - w2[j] * m[u].y) * xv[u].v[j];
                }
            }
            for (; k < width; ++k) {
                const int c = __ldg(ip + k * 8);
                const double2 m = ld_stream(mp + k * 8);
                const RowVec<FPL> xv = ld_row<FPL>(x + (int64_t)c * FPL);
#pragma unroll
                for (int j = 0; j < FPL; ++j) acc[j] += (m.x - w2[j] * m.y) * xv.v[j];
            }
            RowVec<FPL> out;
#pragma unroll
            for (int j = 0; j < FPL; ++j) out.v[j] = acc[j];
            st_row<FPL>(y + row * FPL, out);
            if (DOT) {
                const RowVec<FPL> xo = ld_row<FPL>(x + row * FPL);
#pragma unroll
                for (int j = 0; j < FPL; ++j) dot[j] += xo.v[j] * acc[j];
            }
        }
    }
    if (DOT) {
        // every lane holds all FPL frequencies: lanes -> warp (shuffles, fixed order) -> block (fixed order) -> last block
        __shared__ double s_w[NWARP][FPL];
#pragma unroll
        for (int j = 0; j < FPL; ++j) {
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) dot[j] += __shfl_xor_sync(0xffffffffu, dot[j], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < FPL; ++j) s_w[warp][j] = dot[j];
        }
        __syncthreads();
        if (threadIdx.x < FPL) {
            double sum = 0.0;
            for (int k = 0; k < NWARP; ++k) sum += s_w[k][threadIdx.x];
            partial[((int64_t)blockIdx.x * FPL + threadIdx.x) * 2 + 0] = sum;
            partial[((int64_t)blockIdx.x * FPL + threadIdx.x) * 2 + 1] = 0.0;
        }
        finalize_last_block<FPL, 2, BLOCK>(partial, &sc->ticket[0], [&](int ff, const double* tot) {
            sc->pq[ff] = make_double2(tot[0], tot[1]);
        });
    }
}

// ------------------------------------------------------------------------------------------------
struct __align__(32) d4 { double v[4]; };   // one 256-bit access (LDG/STG.E.256 on sm_100): 2 complex or 4 real values
__device__ __forceinline__ d4 ld256(const double* p) {
    d4 r;
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(double* p, const d4& r) { *reinterpret_cast<d4*>(p) = r; }

// lanes = (row in warp, group of frequencies): every lane carries 32 bytes of a row's vector entry - two
// frequencies of a complex sweep (CPX) or four of a real one - so one 256-bit load per lane gathers
// x[col][0..F) of a row as one contiguous segment (128 bytes at F = 8 complex / F = 16 real).
template <int F, bool CPX, bool DOT, bool OVERLAY, int NWARP, int TS, int NST, int KB>
__global__ void __launch_bounds__((NWARP + 1) * 32)
k_spmv_tma(const int32_t* __restrict__ sell_off, const int32_t* __restrict__ sell_idx, const double2* __restrict__ sell_hq,
           const int32_t* __restrict__ ov_ptr, const int2* __restrict__ ov_cg, const double* __restrict__ ov_val,
           const double2* __restrict__ coef, const double* __restrict__ x, double* __restrict__ y, int64_t n_rows,
           int64_t n_slices, int stage_entries, Scalars* __restrict__ sc, double* __restrict__ partial) {
    constexpr int BLOCK = (NWARP + 1) * 32;
    constexpr int EPL = CPX ? 2 : 4;    // frequencies per lane
    constexpr int LPR = F / EPL;        // lanes per row
    constexpr int RPW = 32 / LPR;       // rows per warp
    constexpr int DPR = CPX ? 2 * F : F;  // doubles per row of a vector
    constexpr int NPASS = (TS * 8 + NWARP * RPW - 1) / (NWARP * RPW);
    static_assert(LPR >= 1 && LPR <= 32, "batch too small for this lane mapping");
    static_assert(!(OVERLAY && !CPX), "a boundary term makes the system complex");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* s_hq = reinterpret_cast<double2*>(smem_raw);                                                    // [NST][stage_entries]
    int32_t* s_idx = reinterpret_cast<int32_t*>(smem_raw + (size_t)NST * stage_entries * sizeof(double2));   // [NST][stage_entries]
    __shared__ uint64_t full_bar[NST], empty_bar[NST];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], NWARP);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t n_tiles = (n_slices + TS - 1) / TS;
    const int64_t tpb = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * tpb;
    const int64_t t1 = min(t0 + tpb, n_tiles);
    double dot[4] = {0.0, 0.0, 0.0, 0.0};

    if (warp == NWARP) {
        // ===== producer warp: one elected lane streams the tiles of this block into the ring =====
        if (lane == 0) {
            int i = 0;
            for (int64_t t = t0; t < t1; ++t, ++i) {
                const int st = i % NST;
                const uint32_t ph = (uint32_t)((i / NST) & 1);
                mbar_wait(&empty_bar[st], ph ^ 1u);
                const int64_t sl0 = t * TS, sl1 = min(sl0 + TS, n_slices);
                const int64_t e0 = (int64_t)sell_off[sl0] * 8, e1 = (int64_t)sell_off[sl1] * 8;
                const uint32_t n = (uint32_t)(e1 - e0);
                if (n) {
                    mbar_expect_tx(&full_bar[st], n * 20u);
                    bulk_g2s(s_hq + (size_t)st * stage_entries, sell_hq + e0, n * 16u, &full_bar[st]);
                    bulk_g2s(s_idx + (size_t)st * stage_entries, sell_idx + e0, n * 4u, &full_bar[st]);
                } else {
                    mbar_arrive(&full_bar[st]);
                }
            }
        }
    } else {
        // ===== consumer warps =====
        const int fq = lane % LPR;
        const int rl = lane / LPR;
        double w2v[4];   // omega^2 of the frequency each of the lane's four doubles belongs to
#pragma unroll
        for (int j = 0; j < 4; ++j) w2v[j] = sc->w2[CPX ? 2 * fq + (j >> 1) : 4 * fq + j];
        const double* xq = x + 4 * fq;
        int i = 0;
        for (int64_t t = t0; t < t1; ++t, ++i) {
            const int st = i % NST;
            const uint32_t ph = (uint32_t)((i / NST) & 1);
            const int64_t row0 = t * TS * 8;
            const int tile_off = sell_off[t * TS];
            int off_r[NPASS], wid_r[NPASS], ob_r[NPASS], oe_r[NPASS];
#pragma unroll
            for (int np = 0; np < NPASS; ++np) {
                const int64_t row = row0 + (np * NWARP + warp) * RPW + rl;
                off_r[np] = 0; wid_r[np] = 0; ob_r[np] = 0; oe_r[np] = 0;
                if ((np * NWARP + warp) * RPW + rl < TS * 8 && row < n_rows) {
                    const int64_t slice = row >> 3;
                    const int o = sell_off[slice];
                    off_r[np] = o - tile_off;
                    wid_r[np] = sell_off[slice + 1] - o;
                    if (OVERLAY) { ob_r[np] = ov_ptr[row]; oe_r[np] = ov_ptr[row + 1]; }
                }
            }
            mbar_wait(&full_bar[st], ph);
            const double2* thq = s_hq + (size_t)st * stage_entries;
            const int32_t* tix = s_idx + (size_t)st * stage_entries;
#pragma unroll
            for (int np = 0; np < NPASS; ++np) {
                const int rin = (np * NWARP + warp) * RPW + rl;
                const int64_t row = row0 + rin;
                if (rin >= TS * 8 || row >= n_rows) continue;
                const int base = off_r[np] * 8 + (int)(row & 7);
                const int width = wid_r[np];
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
                // software pipeline of depth KB: the gather of entry k+KB is issued as entry k is consumed, so
                // KB-1 256-bit gathers per lane stay in flight whatever the compiler's schedule
                d4 xv[KB];
#pragma unroll
                for (int u = 0; u < KB; ++u) {
                    const int c = (u < width) ? tix[base + u * 8] : (int)row;
                    xv[u] = ld256(xq + (int64_t)c * DPR);
                }
                for (int k0 = 0; k0 < width; k0 += KB) {
#pragma unroll
                    for (int u = 0; u < KB; ++u) {
                        const int k = k0 + u;
                        if (k < width) {
                            const double2 m = thq[base + k * 8];
                            if constexpr (CPX) {
                                const double ga = m.x - w2v[0] * m.y, gb = m.x - w2v[2] * m.y;
                                acc[0] += ga * xv[u].v[0]; acc[1] += ga * xv[u].v[1];
                                acc[2] += gb * xv[u].v[2]; acc[3] += gb * xv[u].v[3];
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j) acc[j] += (m.x - w2v[j] * m.y) * xv[u].v[j];
                            }
                        }
                        if (k + KB < width) {
                            const int c = tix[base + (k + KB) * 8];
                            xv[u] = ld256(xq + (int64_t)c * DPR);
                        }
                    }
                }
                if constexpr (OVERLAY) {
                    for (int o = ob_r[np]; o < oe_r[np]; ++o) {
                        const int2 cg = ov_cg[o];
                        const double a = ov_val[o];
                        const double2 ca = coef[cg.y * F + 2 * fq], cb = coef[cg.y * F + 2 * fq + 1];
                        const d4 xo = ld256(xq + (int64_t)cg.x * DPR);
                        const double2 ta = cmul(make_double2(ca.x * a, ca.y * a), make_double2(xo.v[0], xo.v[1]));
                        const double2 tb = cmul(make_double2(cb.x * a, cb.y * a), make_double2(xo.v[2], xo.v[3]));
                        acc[0] += ta.x; acc[1] += ta.y; acc[2] += tb.x; acc[3] += tb.y;
                    }
                }
                d4 out;
#pragma unroll
                for (int j = 0; j < 4; ++j) out.v[j] = acc[j];
                st256(y + row * DPR + 4 * fq, out);
                if (DOT) {
                    const d4 xo = ld256(xq + row * DPR);
                    if constexpr (CPX) {
                        dot[0] += xo.v[0] * acc[0] - xo.v[1] * acc[1]; dot[1] += xo.v[0] * acc[1] + xo.v[1] * acc[0];
                        dot[2] += xo.v[2] * acc[2] - xo.v[3] * acc[3]; dot[3] += xo.v[2] * acc[3] + xo.v[3] * acc[2];
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) dot[j] += xo.v[j] * acc[j];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
        }
    }
    if (DOT) {
        // lanes sharing fq hold the same frequencies
        __shared__ double s_red[NWARP][kMaxF][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int o = 16; o >= LPR; o >>= 1) dot[i] += __shfl_xor_sync(0xffffffffu, dot[i], o);
        }
        if (warp < NWARP && lane < LPR) {
            if constexpr (CPX) {
                s_red[warp][2 * lane][0] = dot[0]; s_red[warp][2 * lane][1] = dot[1];
                s_red[warp][2 * lane + 1][0] = dot[2]; s_red[warp][2 * lane + 1][1] = dot[3];
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { s_red[warp][4 * lane + j][0] = dot[j]; s_red[warp][4 * lane + j][1] = 0.0; }
            }
        }
        __syncthreads();
        if (threadIdx.x < F) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                double sum = 0;
                for (int k = 0; k < NWARP; ++k) sum += s_red[k][threadIdx.x][i];
                partial[((int64_t)blockIdx.x * F + threadIdx.x) * 2 + i] = sum;
            }
        }
        finalize_last_block<F, 2, BLOCK>(partial, &sc->ticket[0], [&](int ff, const double* tot) {
            sc->pq[ff] = make_double2(tot[0], tot[1]);
        });
    }
}

// ------------------------------------------------------------------------------------------------
// K5 for vectors whose rows are 128 bytes (8 complex or 16 real frequencies): the x rows a 128-row tile touches are
// staged in SHARED MEMORY once and the ~15 gathers per row are served from there.
//
// The L1 global-load path returns one 128-byte line every ~2 cycles per SM; with 15 gathers of a full line per
// row that pipe, not HBM, bounded k_spmv_tma.  Shared memory returns 128 bytes per cycle, and
// in the sweep's blob order (reorder.cu) a tile's distinct columns are only ~2.5 x its 128 rows, so staging costs a
// fraction of the gathers it replaces.
//   One persistent block per SM, tiles dealt round-robin (the tiles in flight are neighbours in the blob order and
//   share their halo rows in L2).  Two stage buffers; producer warps fill one while consumer warps work on the other.
//   producers  thread 0 streams the tile's (H, Q) values, 16-bit local column positions and 80-byte header with bulk
//              copies (cp.async.bulk, complete_tx on the stage's `full` mbarrier); all producer threads copy the tile's
//              x rows with 16-byte cp.async (L2 -> smem, no registers, no L1 allocation) and hand their completion to
//              the same mbarrier (cp.async.mbarrier.arrive.noinc).  The column list arrives two tiles ahead by its
//              own bulk copy, the 16-byte tile descriptor two tiles ahead in registers: no fill starts with a chain
//              of dependent global loads.
//   consumers  LPR lanes per row.  A row of x is 8 chunks of 16 bytes; lane q of row r reads chunks
//              q + LPR*((r + j) % (8/LPR)), so the 8 lanes of every quarter-warp phase hit 8 different bank groups:
//              conflict-free 128-bit loads whatever the column positions are.  Warps never synchronise with each
//              other: a warp waits for `full`, does its rows, arrives on `empty`.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ int4 ld_meta(const int4* p) {
    // through inline asm so the compiler treats the descriptor as per-thread data: a plain uniform load is converted to
    // uniform registers right behind the load, which stalls every tile for a global-memory round trip
    int4 v;
    asm volatile("ld.global.nc.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

constexpr int kBlobHdr = 20;        // ints of a tile header (= kTileHdr of reorder.cu)
// TILE = rows per tile (fdb_system::tile_rows): 128 rows with one resident block per SM, or 64 rows with two - the two
// blocks of an SM drift apart, so one's fill overlaps the other's compute without a third stage buffer
constexpr int blob_producers(int tile) { return tile / 32; }   // producer warps
constexpr int blob_threads(int tile, int lpr) { return tile * lpr + blob_producers(tile) * 32; }

// CHEB (modal set-up, modal.cu): instead of y = G x the epilogue writes one step of the Chebyshev filter recurrence,
//     y[row] = c1 * (G x)[row] * cheb_d[row] + c2 * x[row] + c3 * cheb_xold[row]      (y may alias cheb_xold)
// with the row's own x taken from the staged tile: the filter step costs one vector read more than the SpMV instead of a
// second kernel with four vector passes.
// F32 (with CHEB): the vectors are single precision, 32 columns to the same 128-byte row - one pass filters twice the columns
// for the same bytes.  All columns share one shift (Scalars::w2[0]); the matrix entry is rounded to single once per entry.
template <bool CPX, bool DOT, bool OVERLAY, int LPR, int TILE, bool CHEB = false, bool F32 = false>
__global__ void __launch_bounds__(blob_threads(TILE, LPR), TILE == 64 ? 2 : 1)
k_spmv_blob(const uint16_t* __restrict__ sell_lidx, const double2* __restrict__ sell_hq,
            const int4* __restrict__ tile_meta, const int32_t* __restrict__ tile_cols, const int32_t* __restrict__ tile_hdr,
            const int32_t* __restrict__ ov_ptr, const int2* __restrict__ ov_cg, const double* __restrict__ ov_val,
            const double2* __restrict__ coef, const double* __restrict__ x, double* y, int64_t n_rows,
            int64_t n_tiles, int x_cap, int stage_entries, Scalars* __restrict__ sc, double* __restrict__ partial, int dbg, ChebEpilogue cheb) {
    static_assert(!(CHEB && (CPX || DOT || OVERLAY)), "the Chebyshev epilogue belongs to the real, dot-free SpMV of the modal set-up");
    static_assert(!F32 || CHEB, "single-precision vectors: the modal filter only");
    constexpr int F = CPX ? 8 : 16;
    constexpr int kBlobTile = TILE;
    constexpr int NCONS = kBlobTile * LPR;            // consumer threads
    constexpr int NPROD = TILE;                        // producer threads (TILE / 32 warps)
    constexpr int BLOCK = NCONS + NPROD;
    constexpr int NCH = 8 / LPR;      // 16-byte chunks of a row per lane
    constexpr int RPW = 32 / LPR;     // rows per warp
    static_assert(!(OVERLAY && !CPX), "a boundary term makes the system complex");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int cols_cap = (x_cap + 3) & ~3;
    const size_t x_bytes = (size_t)x_cap * 128, hq_bytes = (size_t)stage_entries * 16, lx_bytes = (size_t)stage_entries * 2;
    double* xs = reinterpret_cast<double*>(smem_raw);                                               // [2][x_cap][16]
    double2* s_hq = reinterpret_cast<double2*>(smem_raw + 2 * x_bytes);                             // [2][stage_entries]
    uint16_t* s_lx = reinterpret_cast<uint16_t*>(smem_raw + 2 * x_bytes + 2 * hq_bytes);            // [2][stage_entries]
    int32_t* s_hdr = reinterpret_cast<int32_t*>(smem_raw + 2 * x_bytes + 2 * hq_bytes + 2 * lx_bytes);   // [2][kBlobHdr]
    int32_t* s_cols = s_hdr + 2 * kBlobHdr;                                                          // [2][cols_cap]
    __shared__ uint64_t full[2], empty[2], cbar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full[i], NPROD + 1);       // every producer thread's cp.async completion + thread 0's expect_tx
            mbar_init(&empty[i], NCONS / 32);     // one arrival per consumer warp
            mbar_init(&cbar[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t G = gridDim.x;
    double dot[NCH][2];
#pragma unroll
    for (int j = 0; j < NCH; ++j) dot[j][0] = dot[j][1] = 0.0;

    if (tid >= NCONS) {
        // ===================== producers =====================
        const int ptid = tid - NCONS;
        auto list_bytes = [](const int4& m) { return (uint32_t)(((m.y & 0xffff) + 3) & ~3) * 4u; };
        auto copy_cols = [&](const int4& m, int slot) {      // producer thread 0
            mbar_expect_tx(&cbar[slot], list_bytes(m));
            bulk_g2s(s_cols + slot * cols_cap, tile_cols + m.x, list_bytes(m), &cbar[slot]);
        };
        const int4 zero4 = make_int4(0, 0, 0, 0);
        int64_t t = blockIdx.x;
        int4 m_cur = zero4, m_nxt = zero4, m_nx2 = zero4;
        if (t < n_tiles) m_cur = ld_meta(tile_meta + t);
        if (t + G < n_tiles) m_nxt = ld_meta(tile_meta + t + G);
        if (ptid == 0) {
            if (t < n_tiles) copy_cols(m_cur, 0);
            if (t + G < n_tiles) copy_cols(m_nxt, 1);
        }
        for (uint32_t it = 0; t < n_tiles; t += G, ++it) {
            const int buf = it & 1;
            if (t + 2 * G < n_tiles) m_nx2 = ld_meta(tile_meta + t + 2 * G);
            mbar_wait(&empty[buf], ((it >> 1) & 1u) ^ 1u);    // the consumers released this stage (passes at once for it < 2)
            if (ptid == 0) {
                const int64_t e0 = (int64_t)m_cur.z * 8;
                const uint32_t n = (dbg & 4) ? 0u : (uint32_t)m_cur.w;   // dbg: timing experiments only (FDB_BLOB_DBG)
                mbar_expect_tx(&full[buf], n * 18u + kBlobHdr * 4u);
                bulk_g2s(s_hdr + buf * kBlobHdr, tile_hdr + t * kBlobHdr, kBlobHdr * 4u, &full[buf]);
                if (n) {
                    bulk_g2s(s_hq + (size_t)buf * stage_entries, sell_hq + e0, n * 16u, &full[buf]);
                    bulk_g2s(s_lx + (size_t)buf * stage_entries, sell_lidx + e0, n * 2u, &full[buf]);
                }
            }
            mbar_wait(&cbar[buf], (it >> 1) & 1u);            // this tile's column list
            const int C = (dbg & 2) ? 0 : (m_cur.y & 0xffff);
            const int32_t* tc = s_cols + buf * cols_cap;
            double* dst = xs + (size_t)buf * x_cap * 16;
            for (int i = ptid; i < C * 8; i += NPROD) {
                const int c = i >> 3, ch = i & 7;
                cp_async16(dst + c * 16 + ch * 2, x + (int64_t)tc[c] * 16 + ch * 2);
            }
            cp_async_arrive_noinc(&full[buf]);
            // all producers are done reading the column list: its slot takes the list of the tile after next
            asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
            if (ptid == 0 && t + 2 * G < n_tiles) copy_cols(m_nx2, buf);
            m_cur = m_nxt;
            m_nxt = m_nx2;
        }
    } else {
        // ===================== consumers =====================
        const int q = lane % LPR, rl = lane / LPR;
        int choff[NCH];        // offset (in doubles) of the chunk this lane reads at step j
        double w2v[NCH][2];    // omega^2 of the frequencies in that chunk
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const int ch = q + LPR * ((rl + j) % NCH);
            choff[j] = 2 * ch;
            w2v[j][0] = sc->w2[CPX ? ch : 2 * ch];
            w2v[j][1] = sc->w2[CPX ? ch : 2 * ch + 1];
        }
        const int rin = warp * RPW + rl;
        uint32_t it = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += G, ++it) {
            const int buf = it & 1;
            const int64_t row = t * kBlobTile + rin;
            const bool valid = row < n_rows;
            int ob = 0, oe = 0;
            if (OVERLAY && valid) { ob = ov_ptr[row]; oe = ov_ptr[row + 1]; }
            mbar_wait(&full[buf], (it >> 1) & 1u);
            const int32_t* hd = s_hdr + buf * kBlobHdr;
            const int o = hd[rin >> 3];
            const int width = (valid && !(dbg & 1)) ? hd[(rin >> 3) + 1] - o : 0;
            const int self0 = hd[17];
            const double2* thq = s_hq + (size_t)buf * stage_entries + o * 8 + (rin & 7);
            const uint16_t* tlx = s_lx + (size_t)buf * stage_entries + o * 8 + (rin & 7);
            const double* xb = xs + (size_t)buf * x_cap * 16;
            double acc[NCH][2];
            float accf[NCH][4];
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                acc[j][0] = acc[j][1] = 0.0;
                accf[j][0] = accf[j][1] = accf[j][2] = accf[j][3] = 0.f;
            }
            // a row's ~15 entries: the deeper unrolling keeps more of the (value, position, x row) shared-memory loads in flight per warp
            // - the consumer warps' dependent chains, not bandwidth, were what the kernel waited on (127 -> 118 us at config 4)
            // (the filter variants, with their heavier epilogue, are faster at the old depth)
#pragma unroll (CHEB ? 4 : 32)
            for (int k = 0; k < width; ++k) {
                const double2 m = thq[k * 8];
                const int li = tlx[k * 8];
                const double* xr = xb + li * 16;
                if constexpr (F32) {
                    const float g = (float)(m.x - w2v[0][0] * m.y);
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        const float4 v = *reinterpret_cast<const float4*>(xr + choff[j]);
                        accf[j][0] = fmaf(g, v.x, accf[j][0]);
                        accf[j][1] = fmaf(g, v.y, accf[j][1]);
                        accf[j][2] = fmaf(g, v.z, accf[j][2]);
                        accf[j][3] = fmaf(g, v.w, accf[j][3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        const double2 v = *reinterpret_cast<const double2*>(xr + choff[j]);
                        if constexpr (CPX) {
                            const double g = m.x - w2v[j][0] * m.y;
                            acc[j][0] += g * v.x;
                            acc[j][1] += g * v.y;
                        } else {
                            acc[j][0] += (m.x - w2v[j][0] * m.y) * v.x;
                            acc[j][1] += (m.x - w2v[j][1] * m.y) * v.y;
                        }
                    }
                }
            }
            if constexpr (OVERLAY) {
                for (int oo = ob; oo < oe; ++oo) {
                    const int2 cg = ov_cg[oo];
                    const double a = ov_val[oo];
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        const double2 cf = coef[cg.y * F + (choff[j] >> 1)];
                        const double2 xv = *reinterpret_cast<const double2*>(x + (int64_t)cg.x * 16 + choff[j]);
                        const double2 tt = cmul(make_double2(cf.x * a, cf.y * a), xv);
                        acc[j][0] += tt.x;
                        acc[j][1] += tt.y;
                    }
                }
            }
            if (valid) {
                if constexpr (F32) {
                    // the same byte offsets as the double-precision rows: 128 bytes per row, 16 per chunk
                    const double* xo = xb + (self0 + rin) * 16;
                    const float sd = (float)(cheb.c1 * cheb.d[row]), c2 = (float)cheb.c2, c3 = (float)cheb.c3;
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        const float4 v = *reinterpret_cast<const float4*>(xo + choff[j]);
                        float4 o = make_float4(fmaf(sd, accf[j][0], c2 * v.x), fmaf(sd, accf[j][1], c2 * v.y), fmaf(sd, accf[j][2], c2 * v.z),
                                               fmaf(sd, accf[j][3], c2 * v.w));
                        if (cheb.c3 != 0.0) {
                            const float4 p = *reinterpret_cast<const float4*>(cheb.xold + row * 16 + choff[j]);
                            o.x = fmaf(c3, p.x, o.x); o.y = fmaf(c3, p.y, o.y); o.z = fmaf(c3, p.z, o.z); o.w = fmaf(c3, p.w, o.w);
                        }
                        *reinterpret_cast<float4*>(y + row * 16 + choff[j]) = o;
                    }
                } else if constexpr (CHEB) {
                    const double* xo = xb + (self0 + rin) * 16;
                    const double sd = cheb.c1 * cheb.d[row];
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        const double2 v = *reinterpret_cast<const double2*>(xo + choff[j]);
                        double2 o = make_double2(sd * acc[j][0] + cheb.c2 * v.x, sd * acc[j][1] + cheb.c2 * v.y);
                        if (cheb.c3 != 0.0) {
                            const double2 p = *reinterpret_cast<const double2*>(cheb.xold + row * 16 + choff[j]);
                            o.x += cheb.c3 * p.x;
                            o.y += cheb.c3 * p.y;
                        }
                        *reinterpret_cast<double2*>(y + row * 16 + choff[j]) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < NCH; ++j) *reinterpret_cast<double2*>(y + row * 16 + choff[j]) = make_double2(acc[j][0], acc[j][1]);
                }
                if (DOT) {
                    const double* xo = xb + (self0 + rin) * 16;
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        const double2 v = *reinterpret_cast<const double2*>(xo + choff[j]);
                        if constexpr (CPX) {
                            dot[j][0] += v.x * acc[j][0] - v.y * acc[j][1];
                            dot[j][1] += v.x * acc[j][1] + v.y * acc[j][0];
                        } else {
                            dot[j][0] += v.x * acc[j][0];
                            dot[j][1] += v.y * acc[j][1];
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[buf]);   // this warp is done with the stage
        }
    }
    if (DOT) {
        // per-thread sums -> per-frequency block sums in a fixed order (deterministic)
        __syncthreads();
        double* s_dot = xs;   // [NCONS][2 * NCH] = 16 KB <= one x buffer (x_cap >= 128 rows)
        if (tid < NCONS) {
#pragma unroll
            for (int j = 0; j < NCH; ++j) { s_dot[tid * 2 * NCH + 2 * j] = dot[j][0]; s_dot[tid * 2 * NCH + 2 * j + 1] = dot[j][1]; }
        }
        __syncthreads();
        if (tid < F) {
            const int ch = CPX ? tid : (tid >> 1);
            double s0 = 0.0, s1 = 0.0;
            for (int th = ch % LPR; th < NCONS; th += LPR) {                      // the lanes with q = ch % LPR hold this chunk ...
                const int jj = ((ch / LPR) - (th / LPR) % NCH + NCH) % NCH;       // ... at step jj
                if constexpr (CPX) {
                    s0 += s_dot[th * 2 * NCH + 2 * jj];
                    s1 += s_dot[th * 2 * NCH + 2 * jj + 1];
                } else {
                    s0 += s_dot[th * 2 * NCH + 2 * jj + (tid & 1)];
                }
            }
            partial[((int64_t)blockIdx.x * F + tid) * 2 + 0] = s0;
            partial[((int64_t)blockIdx.x * F + tid) * 2 + 1] = s1;
        }
        finalize_last_block<F, 2, BLOCK>(partial, &sc->ticket[0], [&](int ff, const double* tot) {
            sc->pq[ff] = make_double2(tot[0], tot[1]);
        });
    }
}

// ------------------------------------------------------------------------------------------------
// K6a: alpha = rho / pq ; r -= alpha q ; rho' = r^T (dinv r) ; rr = ||r||^2
// finalize: beta = rho'/rho, rho = rho', convergence flag, iteration count, alpha kept for K6b
// The x update of the step is DEFERRED to K6b, which streams p anyway: x += alpha p there costs a read and a write
// of x, here it would cost those plus a read of p (one vector pass less per iteration, same arithmetic).
// ------------------------------------------------------------------------------------------------
// Jacobi preconditioner z = r / diag(G_f).  ONFLY (no boundary term in the sweep): the diagonal is rebuilt from
// (H_rr, Q_rr) - 16 bytes per ROW instead of 16 bytes per (row, frequency); otherwise the precomputed 1/diag is read.
template <typename T, int F, bool ONFLY>
__device__ __forceinline__ double2 jacobi_load(const double2* __restrict__ dinv, const double2* __restrict__ diag_hq, int64_t i) {
    if constexpr (ONFLY) return __ldg(&diag_hq[i / F]);
    else return dinv[i];
}
template <typename T, bool ONFLY>
__device__ __forceinline__ T jacobi_mul(double2 d, double w2, T rv) {
    if constexpr (ONFLY) {
        const double g = d.x - w2 * d.y;
        const double inv = (g == 0.0) ? 1.0 : 1.0 / g;
        return r_mul(inv, rv);
    } else {
        return s_mul(d, rv);   // complex sweeps only
    }
}
template <typename T, int F, bool ONFLY>
__device__ __forceinline__ T jacobi_apply(const double2* __restrict__ dinv, const double2* __restrict__ diag_hq,
                                          int64_t i, double w2, T rv) {
    return jacobi_mul<T, ONFLY>(jacobi_load<T, F, ONFLY>(dinv, diag_hq, i), w2, rv);
}

// The streaming loops below handle kVecU elements per thread and round, ALL loads of a round issued before the first
// store: a thread of a real sweep moves 8 bytes per access, and with one element per round (the store to r ahead of
// the next round's loads, which the compiler must keep in order) a full SM has only ~32 KB in flight - 3 TB/s.
constexpr int kVecU = 4;

template <typename T, int F, bool ONFLY>
__global__ void __launch_bounds__(kBlock) k_update(T* __restrict__ r,
                                                    const T* __restrict__ q, const double2* __restrict__ dinv,
                                                    const double2* __restrict__ diag_hq, int64_t n_elems,
                                                    Scalars* __restrict__ sc, double* __restrict__ partial) {
    const int f = threadIdx.x % F;
    const bool act = sc->active[f] != 0;
    const double w2 = sc->w2[f];
    double2 alpha = make_double2(0.0, 0.0);
    if (act) {
        const double2 pq = sc->pq[f];
        if (pq.x != 0.0 || pq.y != 0.0) alpha = cdiv(sc->rho[f], pq);
    }
    double v[3] = {0.0, 0.0, 0.0};
    if (act) {
        const int64_t stride = (int64_t)gridDim.x * kBlock;   // a multiple of F: every element of a thread belongs to frequency f
        for (int64_t i0 = (int64_t)blockIdx.x * kBlock + threadIdx.x; i0 < n_elems; i0 += stride * kVecU) {
            T rv[kVecU], qv[kVecU];
            double2 d[kVecU];
#pragma unroll
            for (int u = 0; u < kVecU; ++u) {
                const int64_t i = i0 + u * stride;
                if (i < n_elems) {
                    rv[u] = r[i];
                    qv[u] = q[i];
                    d[u] = jacobi_load<T, F, ONFLY>(dinv, diag_hq, i);
                }
            }
#pragma unroll
            for (int u = 0; u < kVecU; ++u) {
                const int64_t i = i0 + u * stride;
                if (i < n_elems) {
                    const T rn = v_sub(rv[u], s_mul(alpha, qv[u]));
                    r[i] = rn;
                    const T z = jacobi_mul<T, ONFLY>(d[u], w2, rn);
                    const double2 rz = v_dot(rn, z);
                    v[0] += rz.x;
                    v[1] += rz.y;
                    v[2] += v_n2(rn);
                }
            }
        }
    }
    block_reduce_finalize<F, 3, kBlock>(v, partial, &sc->ticket[1], [&](int ff, const double* tot) { fin_update(sc, ff, tot); });
}

// K6b: x += alpha p (the step K6a took) ; p = z + beta p
template <typename T, int F, bool ONFLY>
__global__ void __launch_bounds__(kBlock) k_pupdate(T* __restrict__ x, T* __restrict__ p, const T* __restrict__ r,
                                                     const double2* __restrict__ dinv, const double2* __restrict__ diag_hq,
                                                     int64_t n_elems, const Scalars* __restrict__ sc) {
    const int f = threadIdx.x % F;
    if (!sc->xpend[f]) return;   // frozen frequency: nothing pending, p is not used while alpha = 0
    const bool act = sc->active[f] != 0;   // false: the slot converged in this iteration, only its last step is owed
    const double2 alpha = sc->alpha[f], beta = sc->beta[f];
    const double w2 = sc->w2[f];
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i0 = (int64_t)blockIdx.x * kBlock + threadIdx.x; i0 < n_elems; i0 += stride * kVecU) {
        T xv[kVecU], pv[kVecU], rv[kVecU];
        double2 d[kVecU];
#pragma unroll
        for (int u = 0; u < kVecU; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n_elems) {
                xv[u] = x[i];
                pv[u] = p[i];
                if (act) {
                    rv[u] = r[i];
                    d[u] = jacobi_load<T, F, ONFLY>(dinv, diag_hq, i);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kVecU; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n_elems) {
                x[i] = v_add(xv[u], s_mul(alpha, pv[u]));
                if (act) p[i] = v_add(jacobi_mul<T, ONFLY>(d[u], w2, rv[u]), s_mul(beta, pv[u]));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K6 with the BLOCK-JACOBI preconditioner (real sweeps, F <= 16): z = B_t r on every 128-row tile of the blob order, B_t the
// FP32 inverse of the tile's diagonal block of H (reorder.cu).  About half of point Jacobi's iterations for 0.5 GB more
// per iteration.  K6a becomes a tile kernel: r -= alpha q in FP64, the tile's new residual rounded to FP32 in shared
// memory, z = B r as a 128 x 128 x F FP32 product out of shared memory (B arrives by 16-byte cp.async while the
// residual is updated), z stored in FP32 (K6b reads it instead of r: the vector passes stay at 8), rho' = r^T z and
// ||r||^2 accumulated in FP64.  Fixed summation order everywhere: a sweep stays bit-reproducible.
// START: the same kernel loads a slot (r as it stands, no q): p = z, rho, ||r||^2 for the slots flagged in sc.reset.
// ------------------------------------------------------------------------------------------------
constexpr int kBjTile = 128;
constexpr int kBjPacked = kBjTile * (kBjTile + 1) / 2;   // floats of a packed lower triangle (reorder.cu)
// CPX: a complex batch of F slots is 2 F real columns per row (re, im interleaved): the product treats them as independent
// reals (B is real), only alpha and the unconjugated rho' = r^T z couple a slot's two columns (partner = lane ^ 1).
// MODAL: the modal coarse correction (modal.cu) completes z, so this kernel only leaves r and z_bj = B r (in z32, also when
// loading a slot) and reduces nothing: rho' = r^T z, ||r||^2 and the finalisation belong to k_modal_expand.
template <int F, bool START, bool CPX, bool MODAL = false>
__global__ void __launch_bounds__(256, 3)
k_update_bj(double* __restrict__ r, const double* __restrict__ q, const float* __restrict__ binv, float* __restrict__ z32,
            double* __restrict__ p_start, int64_t n_rows, int64_t n_tiles, Scalars* __restrict__ sc, double* __restrict__ partial) {
    constexpr int NC = CPX ? 2 * F : F;                 // real columns per row
    static_assert(NC >= 1 && NC <= 16, "block-Jacobi kernels cover 16 real columns per row");
    constexpr int EPT = (kBjTile * NC + 255) / 256;     // elements of a tile per thread
    constexpr int CPT = (NC >= 2) ? NC / 2 : 1;         // columns per thread in the product
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Bs = reinterpret_cast<float*>(smem_raw);                   // packed lower triangle: (i, k <= i) at i (i + 1) / 2 + k
    float* r32 = Bs + kBjPacked;                                      // [128][NC]
    float* zp = r32 + kBjTile * NC;                                   // [128][NC]
    const int tid = threadIdx.x;
    const int col = tid % NC;                     // 256 % NC == 0: every element tid + 256 u of a thread lies in this column
    const int f = CPX ? col / 2 : col;
    const bool im = CPX && (col & 1);
    const bool act = START ? (sc->reset[f] != 0) : (sc->active[f] != 0);
    double2 alpha = make_double2(0.0, 0.0);
    if (!START && act) {
        const double2 pq = sc->pq[f];
        if (pq.x != 0.0 || pq.y != 0.0) alpha = cdiv(sc->rho[f], pq);
    }
    const int rp = tid & 63, ch = (tid >> 6) & 1, kh = tid >> 7;
    const int c0 = ch * CPT;
    const bool prod = (NC >= 2) || ch == 0;
    double v[3] = {0.0, 0.0, 0.0};
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const float* bt = binv + t * (int64_t)kBjPacked;
        for (int i = tid; i < kBjPacked / 4; i += 256) cp_async16(Bs + i * 4, bt + i * 4);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const int64_t base = t * (int64_t)kBjTile * NC;
        const int n_el = (int)min((int64_t)kBjTile, n_rows - t * kBjTile) * NC;
        double rn[EPT];
        {
            double rv[EPT], qv[EPT];
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const int e = tid + 256 * u;
                rv[u] = 0.0; qv[u] = 0.0;
                if (e < n_el) {
                    rv[u] = r[base + e];
                    if (!START) qv[u] = q[base + e];
                }
            }
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const int e = tid + 256 * u;
                if constexpr (CPX) {
                    const double qo = __shfl_xor_sync(0xffffffffu, qv[u], 1);   // the slot's other component
                    rn[u] = rv[u] - (im ? alpha.x * qv[u] + alpha.y * qo : alpha.x * qv[u] - alpha.y * qo);
                } else {
                    rn[u] = rv[u] - alpha.x * qv[u];
                }
                if (e < kBjTile * NC) r32[e] = (act && e < n_el) ? (float)rn[u] : 0.0f;
                if (!START && act && e < n_el) r[base + e] = rn[u];
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        float a0[CPT], a1[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) a0[c] = a1[c] = 0.0f;
        if (prod) {
            // B(row, k) = B(k, row): whichever of the two lies in the stored triangle.  Lanes of a warp hold 32 consecutive
            // rows starting at a multiple of 32, so both the row walk (offsets T(row) + k) and the column walk (T(k) + row) hit
            // 32 different banks.
            const int k0 = kh * (kBjTile / 2);
            const int r0i = rp, r1i = rp + 64;
            const int t0 = r0i * (r0i + 1) / 2, t1 = r1i * (r1i + 1) / 2;
#pragma unroll 4
            for (int k = k0; k < k0 + kBjTile / 2; ++k) {
                const int tk = k * (k + 1) / 2;
                const float b0 = Bs[(k <= r0i) ? t0 + k : tk + r0i], b1 = Bs[(k <= r1i) ? t1 + k : tk + r1i];
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const float rr = r32[k * NC + c0 + c];
                    a0[c] += b0 * rr;
                    a1[c] += b1 * rr;
                }
            }
            if (kh == 1) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) { zp[rp * NC + c0 + c] = a0[c]; zp[(rp + 64) * NC + c0 + c] = a1[c]; }
            }
        }
        __syncthreads();
        if (prod && kh == 0) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                zp[rp * NC + c0 + c] = a0[c] + zp[rp * NC + c0 + c];
                zp[(rp + 64) * NC + c0 + c] = a1[c] + zp[(rp + 64) * NC + c0 + c];
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
            const int e = tid + 256 * u;
            const bool ok = act && e < n_el;
            const float z = (e < kBjTile * NC) ? zp[e] : 0.0f;
            if (ok) {
                if constexpr (START && !MODAL) p_start[base + e] = (double)z;   // first search direction of the slot
                else z32[base + e] = z;
            }
            if constexpr (MODAL) {
            } else if constexpr (CPX) {
                // (rr + j ri)(zr + j zi), accumulated by the lane that holds the real part
                const double ro = __shfl_xor_sync(0xffffffffu, rn[u], 1);
                const double zo = (double)__shfl_xor_sync(0xffffffffu, z, 1);
                if (ok) {
                    if (!im) {
                        v[0] += rn[u] * (double)z - ro * zo;
                        v[1] += rn[u] * zo + ro * (double)z;
                    }
                    v[2] += rn[u] * rn[u];
                }
            } else {
                if (ok) {
                    v[0] += rn[u] * (double)z;
                    v[2] += rn[u] * rn[u];
                }
            }
        }
        __syncthreads();   // the next tile overwrites Bs, r32, zp
    }
    auto fin = [&](int ff, const double* tot) {
        if constexpr (START) fin_start(sc, ff, tot); else fin_update(sc, ff, tot);
    };
    unsigned int* ticket = &sc->ticket[START ? 2 : 1];
    if constexpr (MODAL) {
    } else if constexpr (CPX) {
        // threads of slot f: columns 2f, 2f+1 of every group of NC threads; fixed order
        __shared__ double s_v[256][3];
#pragma unroll
        for (int i = 0; i < 3; ++i) s_v[tid][i] = v[i];
        __syncthreads();
        if (tid < F) {
            double sum[3] = {0.0, 0.0, 0.0};
            for (int th = 2 * tid; th < 256; th += NC) {
#pragma unroll
                for (int i = 0; i < 3; ++i) sum[i] += s_v[th][i] + s_v[th + 1][i];
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) partial[((int64_t)blockIdx.x * F + tid) * 3 + i] = sum[i];
        }
        finalize_last_block<F, 3, 256>(partial, ticket, fin);
    } else {
        block_reduce_finalize<F, 3, 256>(v, partial, ticket, fin);
    }
}

// K6b with z read from the FP32 array K6a left: x += alpha p ; p = z + beta p.  T = double (real sweep) or double2.
template <typename T, int F>
__global__ void __launch_bounds__(kBlock) k_pupdate_bj(T* __restrict__ x, T* __restrict__ p, const float* __restrict__ z32,
                                                        int64_t n_elems, const Scalars* __restrict__ sc) {
    constexpr bool CPX = std::is_same<T, double2>::value;
    const int f = threadIdx.x % F;
    if (!sc->xpend[f]) return;
    const bool act = sc->active[f] != 0;
    const double2 alpha = sc->alpha[f], beta = sc->beta[f];
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i0 = (int64_t)blockIdx.x * kBlock + threadIdx.x; i0 < n_elems; i0 += stride * kVecU) {
        T xv[kVecU], pv[kVecU], zv[kVecU];
#pragma unroll
        for (int u = 0; u < kVecU; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n_elems) {
                xv[u] = x[i];
                pv[u] = p[i];
                if (act) {
                    if constexpr (CPX) {
                        const float2 zf = reinterpret_cast<const float2*>(z32)[i];
                        zv[u] = make_double2((double)zf.x, (double)zf.y);
                    } else {
                        zv[u] = (double)z32[i];
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kVecU; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n_elems) {
                x[i] = v_add(xv[u], s_mul(alpha, pv[u]));
                if (act) p[i] = v_add(zv[u], s_mul(beta, pv[u]));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// slot (re)load kernels: a batch slot whose frequency converged is reloaded with the next frequency of the
// sweep while the other slots keep iterating ("continuous batching").  These touch only slots with reset[f] != 0.
// ------------------------------------------------------------------------------------------------
// dinv[row][f] = 1 / (H_rr - w2_f Q_rr + sum_ov,col==row coef[g][f] A)      (complex sweeps with a boundary term)
template <int F>
__global__ void __launch_bounds__(kBlock) k_jacobi(const double2* __restrict__ diag_hq, const int32_t* __restrict__ ov_ptr,
                                                    const int2* __restrict__ ov_cg, const double* __restrict__ ov_val,
                                                    const double2* __restrict__ coef, bool overlay, int64_t n_rows,
                                                    const Scalars* __restrict__ sc, double2* __restrict__ dinv) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t row = i / F;
    const int f = (int)(i % F);
    if (row >= n_rows || !sc->reset[f]) return;
    const double2 d = diag_hq[row];
    double2 g = make_double2(d.x - sc->w2[f] * d.y, 0.0);
    if (overlay) {
        for (int k = ov_ptr[row]; k < ov_ptr[row + 1]; ++k) {
            const int2 cg = ov_cg[k];
            if (cg.x == row) {
                const double2 cf = coef[cg.y * F + f];
                const double a = ov_val[k];
                g.x += cf.x * a;
                g.y += cf.y * a;
            }
        }
    }
    const bool zero = (g.x == 0.0 && g.y == 0.0);
    dinv[i] = zero ? make_double2(1.0, 0.0) : cdiv(make_double2(1.0, 0.0), g);
}

// reset slots: r = 0, and x = 0 unless the slot keeps its x as the starting guess (warm start, or a resumed slot: sc.keepx)
template <typename T, int F>
__global__ void __launch_bounds__(kBlock) k_slot_clear(T* __restrict__ x, T* __restrict__ r, bool keep_all, int64_t n_elems,
                                                        const Scalars* __restrict__ sc) {
    const int f = threadIdx.x % F;
    if (!sc->reset[f]) return;
    const bool keep_x = keep_all || sc->keepx[f] != 0;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n_elems; i += (int64_t)gridDim.x * kBlock) {
        r[i] = v_zero<T>();
        if (!keep_x) x[i] = v_zero<T>();
    }
}

// dst[node][f] = b = -1j*w_f*q_s (complex) or its imaginary part -w_f*q_s (real sweep); ALL: every slot
__device__ __forceinline__ void store_rhs(double2* dst, double2 scale, double2 q) { *dst = cmul(scale, q); }
__device__ __forceinline__ void store_rhs(double* dst, double2 scale, double2 q) { *dst = scale.y * q.x; }
template <typename T, int F>
__global__ void k_set_rhs(const int32_t* __restrict__ src_nodes, const double2* __restrict__ src_q, const int32_t* __restrict__ set_ptr,
                          int max_set, bool all, const Scalars* __restrict__ sc, T* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= max_set * F) return;
    const int j = i / F, f = i % F;
    if (!(all || sc->reset[f])) return;
    const int set = sc->src_set[f];
    const int e = set_ptr[set] + j;
    if (e < set_ptr[set + 1]) store_rhs(&dst[(int64_t)src_nodes[e] * F + f], sc->rhs_scale[f], src_q[e]);
}

// dst[node][f] += sum over the node's terms of rcoef[vec][f] * val   (one thread per (listed node, f): race-free, fixed order)
template <int F>
__global__ void k_add_rhs(const int32_t* __restrict__ rn_nodes, const int32_t* __restrict__ rn_ptr, const int32_t* __restrict__ rn_vec,
                          const double* __restrict__ rn_val, int n_nodes, const double2* __restrict__ rcoef, bool all,
                          const Scalars* __restrict__ sc, double2* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes * F) return;
    const int s = i / F, f = i % F;
    if (!(all || sc->reset[f])) return;
    double2 acc = dst[(int64_t)rn_nodes[s] * F + f];
    for (int k = rn_ptr[s]; k < rn_ptr[s + 1]; ++k) {
        const double2 c = rcoef[rn_vec[k] * F + f];
        const double v = rn_val[k];
        acc.x += c.x * v;
        acc.y += c.y * v;
    }
    dst[(int64_t)rn_nodes[s] * F + f] = acc;
}

// reset slots: r = r - q    (r holds b on entry; q = G x0)
template <typename T, int F>
__global__ void __launch_bounds__(kBlock) k_sub(T* __restrict__ r, const T* __restrict__ q, int64_t n,
                                                 const Scalars* __restrict__ sc) {
    const int f = threadIdx.x % F;
    if (!sc->reset[f]) return;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) r[i] = v_sub(r[i], q[i]);
}

// reset slots: p = z ; rho = r^T z ; rr = ||r||^2 ; activate the slot if its residual is above tolerance
template <typename T, int F, bool ONFLY>
__global__ void __launch_bounds__(kBlock) k_start(T* __restrict__ p, const T* __restrict__ r, const double2* __restrict__ dinv,
                                                   const double2* __restrict__ diag_hq, int64_t n_elems,
                                                   Scalars* __restrict__ sc, double* __restrict__ partial) {
    const int f = threadIdx.x % F;
    const bool rs = sc->reset[f] != 0;
    const double w2 = sc->w2[f];
    double v[3] = {0.0, 0.0, 0.0};
    if (rs) {
        for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n_elems; i += (int64_t)gridDim.x * kBlock) {
            const T rv = r[i];
            const T z = jacobi_apply<T, F, ONFLY>(dinv, diag_hq, i, w2, rv);
            p[i] = z;
            const double2 rz = v_dot(rv, z);
            v[0] += rz.x;
            v[1] += rz.y;
            v[2] += v_n2(rv);
        }
    }
    block_reduce_finalize<F, 3, kBlock>(v, partial, &sc->ticket[2], [&](int ff, const double* tot) { fin_start(sc, ff, tot); });
}

// true residual norm of what is returned: rt = ||b - q||^2 with b rebuilt, q = G x
template <typename T, int F>
__global__ void __launch_bounds__(kBlock) k_resnorm(const T* __restrict__ b, const T* __restrict__ q, int64_t n_elems,
                                                     Scalars* __restrict__ sc, double* __restrict__ partial) {
    double v[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n_elems; i += (int64_t)gridDim.x * kBlock)
        v[0] += v_n2(v_sub(b[i], q[i]));
    block_reduce_finalize<F, 1, kBlock>(v, partial, &sc->ticket[3], [&](int ff, const double* tot) { sc->rt[ff] = tot[0]; });
}

// tail of a sweep: move the still-running slots of a batch of width Fs into a narrower batch of width Fd
struct SlotMap { int src[kMaxF]; };
template <typename T>
__global__ void __launch_bounds__(kBlock) k_repack(const T* __restrict__ src, int Fs, T* __restrict__ dst, int Fd, SlotMap map,
                                                    int64_t n_rows) {
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n_rows * Fd; i += (int64_t)gridDim.x * kBlock) {
        const int64_t row = i / Fd;
        const int j = (int)(i % Fd);
        const int sj = map.src[j];
        dst[i] = (sj >= 0) ? src[row * Fs + sj] : v_zero<T>();
    }
}

// K7: pR[f][i] = sum_j w[i][j] x[node[i][j]][f]
template <typename T, int F>
__global__ void k_gather_rcv(const T* __restrict__ x, const int32_t* __restrict__ nodes, const double* __restrict__ w,
                             int n_rcv, double2* __restrict__ out /* [F][n_rcv] */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rcv * F) return;
    const int rc = i / F, f = i % F;
    T acc = v_zero<T>();
    bool first = true;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double wt = w[rc * 4 + j];
        if (wt == 0.0) continue;
        const T v = x[(int64_t)nodes[rc * 4 + j] * F + f];
        if (first && wt == 1.0) { acc = v; first = false; continue; }  // on-node receiver: exact copy
        acc = v_add(acc, r_mul(wt, v));
        first = false;
    }
    out[(int64_t)f * n_rcv + rc] = v_out(acc);
}

// pN export: stage[f][node] = x[perm[node]][f]  (back to the caller's node numbering)
template <typename T, int F>
__global__ void __launch_bounds__(kBlock) k_export(const T* __restrict__ x, const int32_t* __restrict__ perm, int64_t n_rows,
                                                    double2* __restrict__ stage) {
    __shared__ T tile[kBlock];
    constexpr int RPB = kBlock / F;
    const int64_t row0 = (int64_t)blockIdx.x * RPB;
    const int rr = threadIdx.x / F, ff = threadIdx.x % F;
    if (row0 + rr < n_rows) tile[threadIdx.x] = x[(int64_t)perm[row0 + rr] * F + ff];
    __syncthreads();
    const int f = threadIdx.x / RPB, rl = threadIdx.x % RPB;
    if (row0 + rl < n_rows) stage[(int64_t)f * n_rows + row0 + rl] = v_out(tile[rl * F + f]);
}

// stand-alone SpMV entry: caller's numbering <-> internal order, rows of F values
template <typename T>
__global__ void __launch_bounds__(kBlock) k_permute_rows(const T* __restrict__ src, T* __restrict__ dst, const int32_t* __restrict__ perm,
                                                          int F, int64_t n_rows, bool forward) {
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n_rows * F; i += (int64_t)gridDim.x * kBlock) {
        const int64_t row = i / F;
        const int f = (int)(i % F);
        const int64_t prow = perm[row];
        if (forward) dst[prow * F + f] = src[i]; else dst[i] = src[prow * F + f];
    }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
template <int F>
struct Launch {
    // ---- plain (LDG) SpMV: narrow batches (F = 1, 2 and the tail of a sweep) and the fallback when a mesh's rows are too
    //      long for the bulk-copy pipeline's shared-memory tiles
    template <typename T, int BLOCK, int MINB>
    static void spmv_cfg(fdb_system* sys, SweepWork* w, const T* x, T* y, bool dot, bool overlay) {
        if constexpr (F > 32) {
            throw Error("femder_b200: the plain SpMV kernel supports batch <= 32 (rows of this mesh are too long for the bulk-copy pipeline)");
        } else {
            const int64_t N = w->N;
            const int rpw = 32 / F;
            const int64_t groups = (N + rpw - 1) / rpw;
            const int64_t per_block = BLOCK / 32;
            int grid = (int)std::min<int64_t>((groups + per_block - 1) / per_block, (int64_t)sys->num_sms * MINB);
            auto args = [&](auto kern) {
                kern<<<grid, BLOCK, 0, sys->stream>>>(sys->sell_off.p, sys->sell_idx.p, sys->sell_hq.p, sys->ov_ptr.p, sys->ov_cg.p,
                                                     sys->ov_val.p, w->coef.p, x, y, N, w->sc.p, (double*)w->partial.p);
            };
            if constexpr (std::is_same<T, double2>::value) {
                if (dot) {
                    if (overlay) args(k_spmv<T, F, true, true, BLOCK, MINB>); else args(k_spmv<T, F, true, false, BLOCK, MINB>);
                } else {
                    if (overlay) args(k_spmv<T, F, false, true, BLOCK, MINB>); else args(k_spmv<T, F, false, false, BLOCK, MINB>);
                }
            } else {
                if (dot) args(k_spmv<T, F, true, false, BLOCK, MINB>); else args(k_spmv<T, F, false, false, BLOCK, MINB>);
            }
        }
    }
    template <typename T>
    static void spmv_ldg(fdb_system* sys, SweepWork* w, const T* x, T* y, bool dot, bool overlay) {
        spmv_cfg<T, kSpmvCfgs[0].block, kSpmvCfgs[0].per_sm>(sys, w, x, y, dot, overlay);
    }

    // ---- bulk-copy pipeline SpMV
    template <int TS>
    static int stage_entries(fdb_system* sys, SweepWork* w) {
        int& cached = w->tma_stage_entries[TS == 4 ? 2 : (TS == 8 ? 0 : 1)];
        if (cached == 0) {
            const int64_t S = sys->n_slices, n_tiles = (S + TS - 1) / TS;
            int64_t mx = 0;
            for (int64_t t = 0; t < n_tiles; ++t) {
                const int64_t a0 = sys->sell_off_host[t * TS], a1 = sys->sell_off_host[std::min<int64_t>((t + 1) * TS, S)];
                mx = std::max(mx, (a1 - a0) * 8);
            }
            cached = (int)std::max<int64_t>(mx, 8);
        }
        return cached;
    }
    template <bool CPX, int NWARP, int TS, int NST, int KB>
    static bool spmv_tma(fdb_system* sys, SweepWork* w, const void* x, void* y, bool dot, bool overlay) {
        static_assert(NWARP >= 1 && (NWARP + 1) * 32 <= 1024, "block too large");
        const int64_t N = w->N, S = sys->n_slices;
        const int64_t n_tiles = (S + TS - 1) / TS;
        const int se = stage_entries<TS>(sys, w);
        const size_t smem = (size_t)NST * se * 20;
        if (smem > 100 * 1024) return false;  // rows too long for this tile shape: use the plain kernel
        int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / ((NWARP + 1) * 32), (150 * 1024) / (smem + 8192)));
        if (w->spmv_bps > 0) per_sm = std::min(per_sm, w->spmv_bps);
        int grid = (int)std::min<int64_t>(n_tiles, (int64_t)sys->num_sms * per_sm);
        auto launch = [&](auto kern) {
            // largest dynamic shared memory already opted into, per (device, kernel): all instantiations share one
            // function-pointer TYPE, so the key must be the pointer VALUE
            static thread_local std::map<std::pair<int, const void*>, size_t> granted;
            size_t& have = granted[{sys->device, (const void*)kern}];
            if (smem > have) {
                FDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                have = smem;
            }
            kern<<<grid, (NWARP + 1) * 32, smem, sys->stream>>>(sys->sell_off.p, sys->sell_idx.p, sys->sell_hq.p, sys->ov_ptr.p,
                                                               sys->ov_cg.p, sys->ov_val.p, w->coef.p, (const double*)x, (double*)y, N, S,
                                                               se, w->sc.p, (double*)w->partial.p);
        };
        if constexpr (CPX) {
            if (dot) {
                if (overlay) launch(k_spmv_tma<F, true, true, true, NWARP, TS, NST, KB>);
                else launch(k_spmv_tma<F, true, true, false, NWARP, TS, NST, KB>);
            } else {
                if (overlay) launch(k_spmv_tma<F, true, false, true, NWARP, TS, NST, KB>);
                else launch(k_spmv_tma<F, true, false, false, NWARP, TS, NST, KB>);
            }
        } else {
            if (dot) launch(k_spmv_tma<F, false, true, false, NWARP, TS, NST, KB>);
            else launch(k_spmv_tma<F, false, false, false, NWARP, TS, NST, KB>);
        }
        return true;
    }
    template <bool CPX>
    static bool spmv_tma_cfg(fdb_system* sys, SweepWork* w, const void* x, void* y, bool dot, bool overlay) {
        constexpr int LPR = CPX ? F / 2 : F / 4;              // lanes per row
        if constexpr (LPR < 2) {
            return false;   // one lane per row scatters every 256-bit gather over 32 lines: the plain kernel is faster
        } else {
            constexpr int RPW = 32 / LPR;                        // rows per warp and pass
            constexpr int W8 = (64 / RPW < 16) ? (64 / RPW < 1 ? 1 : 64 / RPW) : 16;      // consumer warps for an 8-slice tile
            constexpr int W16 = (128 / RPW < 16) ? (128 / RPW < 1 ? 1 : 128 / RPW) : 16;  // ... for a 16-slice tile
            constexpr int W4 = (32 / RPW < 16) ? (32 / RPW < 1 ? 1 : 32 / RPW) : 16;      // ... for a 4-slice tile
            switch (w->spmv_cfg) {
                case 1: return spmv_tma<CPX, W8, 8, 2, 4>(sys, w, x, y, dot, overlay);
                case 7: return spmv_tma<CPX, W16, 16, 2, 8>(sys, w, x, y, dot, overlay);   // the L1-gather pipeline where k_spmv_blob would run
                case 2: return spmv_tma<CPX, W8, 8, 3, 8>(sys, w, x, y, dot, overlay);
                case 3: return spmv_tma<CPX, W8, 8, 2, 8>(sys, w, x, y, dot, overlay);
                default:
                    // 128-row tiles (measured best); smaller tiles when the rows are long (Tet10, unstructured meshes)
                    if (spmv_tma<CPX, W16, 16, 2, 8>(sys, w, x, y, dot, overlay)) return true;
                    if (spmv_tma<CPX, W8, 8, 2, 8>(sys, w, x, y, dot, overlay)) return true;
                    return spmv_tma<CPX, W4, 4, 2, 8>(sys, w, x, y, dot, overlay);
            }
        }
    }
    // ---- shared-memory x tiles (k_spmv_blob): 128-byte vector rows only
    template <bool CPX>
    static bool spmv_blob(fdb_system* sys, SweepWork* w, const void* x, void* y, bool dot, bool overlay) {
        if constexpr ((CPX && F != 8) || (!CPX && F != 16)) {
            return false;
        } else {
            if (sys->n_tiles == 0) return false;
            const int tile = sys->tile_rows;
            const int x_cap = std::max(sys->max_tile_cols, tile);
            const int se = sys->max_tile_entries;
            const size_t smem = (size_t)2 * x_cap * 128 + (size_t)2 * se * 18 + (size_t)2 * (((x_cap + 3) & ~3) + kBlobHdr) * 4;
            // long rows (Tet10) or a mesh without locality: the L1 kernels
            if (tile == 64 ? smem > 110 * 1024 : smem > 220 * 1024) return false;
            const int grid = (int)std::min<int64_t>(sys->n_tiles, (int64_t)sys->num_sms * (tile == 64 ? 2 : 1));
            auto launch = [&](auto kern, int threads) {
                static thread_local std::map<std::pair<int, const void*>, size_t> granted;
                size_t& have = granted[{sys->device, (const void*)kern}];
                if (smem > have) {
                    FDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    have = smem;
                }
                kern<<<grid, threads, smem, sys->stream>>>(sys->sell_lidx.p, sys->sell_hq.p, sys->tile_meta.p,
                                                             sys->tile_cols.p, sys->tile_hdr.p, sys->ov_ptr.p, sys->ov_cg.p, sys->ov_val.p,
                                                             w->coef.p, (const double*)x, (double*)y, w->N, sys->n_tiles,
                                                             x_cap, se, w->sc.p, (double*)w->partial.p, w->blob_dbg, w->cheb);
            };
            if constexpr (!CPX) {
                if (w->cheb.d != nullptr) {   // modal set-up: one filter step instead of y = G x (128-row tiles, 2 lanes per row)
                    if (tile != 128 || dot) return false;
                    if (w->cheb.f32) launch(k_spmv_blob<false, false, false, 2, 128, true, true>, blob_threads(128, 2));
                    else launch(k_spmv_blob<false, false, false, 2, 128, true>, blob_threads(128, 2));
                    return true;
                }
            }
            auto pick = [&](auto lpr, auto tl) {
                constexpr int L = decltype(lpr)::value, T = decltype(tl)::value;
                constexpr int TH = blob_threads(T, L);
                if constexpr (CPX) {
                    if (dot) {
                        if (overlay) launch(k_spmv_blob<true, true, true, L, T>, TH); else launch(k_spmv_blob<true, true, false, L, T>, TH);
                    } else {
                        if (overlay) launch(k_spmv_blob<true, false, true, L, T>, TH); else launch(k_spmv_blob<true, false, false, L, T>, TH);
                    }
                } else {
                    if (dot) launch(k_spmv_blob<false, true, false, L, T>, TH); else launch(k_spmv_blob<false, false, false, L, T>, TH);
                }
            };
            using I2 = std::integral_constant<int, 2>; using I4 = std::integral_constant<int, 4>;
            using T64 = std::integral_constant<int, 64>; using T128 = std::integral_constant<int, 128>;
            if (tile == 64) { if (w->blob_lpr == 4) pick(I4{}, T64{}); else pick(I2{}, T64{}); }
            else            { if (w->blob_lpr == 4) pick(I4{}, T128{}); else pick(I2{}, T128{}); }
            return true;
        }
    }
    // can a real-arithmetic sweep run at this batch width on this mesh?
    static bool real_supported(fdb_system* sys, SweepWork* w) {
        if (F <= 32) return true;                                          // the plain kernel covers it
        return (size_t)2 * stage_entries<4>(sys, w) * 20 <= 100 * 1024;   // wider: the smallest tile shape must fit
    }
    // ---- narrow real batches: one lane per row (k_spmv_rows)
    static bool spmv_rows(fdb_system* sys, SweepWork* w, const void* x, void* y, bool dot) {
        if constexpr (F == 2 || F == 4 || F == 8) {
            constexpr int BLOCK = (F == 2) ? 512 : 256;
            constexpr int MINB = (F == 2) ? 2 : (F == 4 ? 3 : 2);
            constexpr int UNR = (F == 8) ? 2 : 4;
            const int64_t groups = (w->N + 31) / 32;
            const int64_t per_block = BLOCK / 32;
            const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((groups + per_block - 1) / per_block, (int64_t)sys->num_sms * MINB));
            if (dot)
                k_spmv_rows<F, true, BLOCK, MINB, UNR><<<grid, BLOCK, 0, sys->stream>>>(sys->sell_off.p, sys->sell_idx.p, sys->sell_hq.p,
                                                                                      (const double*)x, (double*)y, w->N, w->sc.p, (double*)w->partial.p);
            else
                k_spmv_rows<F, false, BLOCK, MINB, UNR><<<grid, BLOCK, 0, sys->stream>>>(sys->sell_off.p, sys->sell_idx.p, sys->sell_hq.p,
                                                                                       (const double*)x, (double*)y, w->N, w->sc.p, (double*)w->partial.p);
            return true;
        } else {
            return false;
        }
    }
    static void spmv(fdb_system* sys, SweepWork* w, const void* x, void* y, bool dot, bool overlay) {
        if (w->real) {
            if (w->spmv_cfg == 0 && spmv_blob<false>(sys, w, x, y, dot, false)) return;
            if (w->spmv_cfg == 0 && w->rows_kernel && spmv_rows(sys, w, x, y, dot)) return;
            if (w->spmv_cfg < 8 && spmv_tma_cfg<false>(sys, w, x, y, dot, false)) return;
            spmv_ldg<double>(sys, w, (const double*)x, (double*)y, dot, false);
            return;
        }
        if (w->spmv_cfg == 0 && spmv_blob<true>(sys, w, x, y, dot, overlay)) return;
        if (w->spmv_cfg < 8 && spmv_tma_cfg<true>(sys, w, x, y, dot, overlay)) return;
        spmv_ldg<double2>(sys, w, (const double2*)x, (double2*)y, dot, overlay);
    }

    // block-Jacobi tile kernels: 48 KB of shared memory per block at F = 16 (packed FP32 block + two 128 x F tiles)
    static constexpr bool kBjReal = F <= 16, kBjCpx = F <= 8;   // 16 real columns per row at most
    static constexpr bool kModalReal = F == kModalCols, kModalCpx = 2 * F == kModalCols;   // the tensor-core kernels: exactly 16 real columns
    static void modal_supported(SweepWork* w, bool* ok) { *ok = w->real ? kModalReal : kModalCpx; }
    static size_t bj_smem(bool cpx) { return (size_t)kBjPacked * 4 + (size_t)2 * kBjTile * std::min(16, cpx ? 2 * F : F) * 4; }
    static int bj_grid(fdb_system* sys) { return (int)std::max<int64_t>(1, std::min<int64_t>(sys->n_tiles, (int64_t)sys->num_sms * 3)); }
    static void bj_supported(SweepWork* w, bool* ok) { *ok = w->real ? kBjReal : kBjCpx; }
    static void bj_prepare(fdb_system* sys, SweepWork* w) {
        build_block_prec(sys);
        w->z32.alloc((size_t)w->N * F * (w->real ? 1 : 2));
        static thread_local std::map<std::pair<int, bool>, bool> done;
        bool& d = done[{sys->device, w->real}];
        if (d) return;
        if (w->real) {
            if constexpr (kBjReal) {
                FDB_CUDA(cudaFuncSetAttribute(k_update_bj<F, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bj_smem(false)));
                FDB_CUDA(cudaFuncSetAttribute(k_update_bj<F, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bj_smem(false)));
                if constexpr (kModalReal) {
                    FDB_CUDA(cudaFuncSetAttribute(k_update_bj<F, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bj_smem(false)));
                    FDB_CUDA(cudaFuncSetAttribute(k_update_bj<F, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bj_smem(false)));
                }
            }
        } else {
            if constexpr (kBjCpx) {
                FDB_CUDA(cudaFuncSetAttribute(k_update_bj<F, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bj_smem(true)));
                FDB_CUDA(cudaFuncSetAttribute(k_update_bj<F, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bj_smem(true)));
                if constexpr (kModalCpx) {
                    FDB_CUDA(cudaFuncSetAttribute(k_update_bj<F, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bj_smem(true)));
                    FDB_CUDA(cudaFuncSetAttribute(k_update_bj<F, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bj_smem(true)));
                }
            }
        }
        d = true;
    }
    // K6a (start = false) or the slot load (start = true) of the block-Jacobi path; false if this width has no such kernel
    template <typename T>
    static bool bj_update(fdb_system* sys, SweepWork* w, bool start) {
        constexpr bool CPX = std::is_same<T, double2>::value;
        if constexpr (CPX ? kBjCpx : kBjReal) {
            double *r = (double*)w->r.p, *q = (double*)w->q.p, *p = (double*)w->p.p;
            if (w->precond == 2) {
                if constexpr (CPX ? kModalCpx : kModalReal) {
                    if (start)
                        k_update_bj<F, true, CPX, true><<<bj_grid(sys), 256, bj_smem(CPX), sys->stream>>>(r, nullptr, sys->blk_inv.p, w->z32.p, p, w->N,
                                                                                                         sys->n_tiles, w->sc.p, (double*)w->partial.p);
                    else
                        k_update_bj<F, false, CPX, true><<<bj_grid(sys), 256, bj_smem(CPX), sys->stream>>>(r, q, sys->blk_inv.p, w->z32.p, nullptr, w->N,
                                                                                                          sys->n_tiles, w->sc.p, (double*)w->partial.p);
                    return true;
                } else {
                    throw Error("femder_b200: the modal correction runs in batches of 16 real or 8 complex slots");
                }
            }
            if (start)
                k_update_bj<F, true, CPX><<<bj_grid(sys), 256, bj_smem(CPX), sys->stream>>>(r, nullptr, sys->blk_inv.p, w->z32.p, p, w->N, sys->n_tiles,
                                                                                           w->sc.p, (double*)w->partial.p);
            else
                k_update_bj<F, false, CPX><<<bj_grid(sys), 256, bj_smem(CPX), sys->stream>>>(r, q, sys->blk_inv.p, w->z32.p, nullptr, w->N, sys->n_tiles,
                                                                                            w->sc.p, (double*)w->partial.p);
            return true;
        } else {
            return false;
        }
    }
    // vector kernels: up to 8 blocks per SM, but at least 16 elements per thread - on a small mesh a thin grid leaves
    // the reducing kernels fewer block partials to sum in their last block
    static int egrid(SweepWork* w) {
        const int64_t want = (w->N * F + (int64_t)kBlock * 16 - 1) / ((int64_t)kBlock * 16);
        return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)w->grid));
    }
    // K6a / K6b of one iteration (which = 1 / 2), K5 is spmv() above
    template <typename T>
    static void vec_kernel_t(fdb_system* sys, SweepWork* w, bool overlay, int which) {
        const int64_t ne = w->N * F;
        T *x = (T*)w->x.p, *r = (T*)w->r.p, *p = (T*)w->p.p, *q = (T*)w->q.p;
        if (w->precond >= 1) {
            if (which == 1) { if (bj_update<T>(sys, w, false)) return; }
            else if constexpr (std::is_same<T, double2>::value ? kBjCpx : kBjReal) {
                static const bool old_k6b = tune_env("FDB_OLD_K6B") != nullptr;   // A/B: the grid-stride K6b
                if (!old_k6b && ((w->real && F == 16) || (!w->real && F == 8))) pupdate_tiles(sys, w);   // 128-byte rows
                else k_pupdate_bj<T, F><<<egrid(w), kBlock, 0, sys->stream>>>(x, p, w->z32.p, ne, w->sc.p);
                return;
            }
        }
        if (overlay) {
            if constexpr (std::is_same<T, double2>::value) {
                if (which == 1) k_update<T, F, false><<<egrid(w), kBlock, 0, sys->stream>>>(r, q, w->dinv.p, sys->diag_hq.p, ne, w->sc.p, (double*)w->partial.p);
                else k_pupdate<T, F, false><<<egrid(w), kBlock, 0, sys->stream>>>(x, p, r, w->dinv.p, sys->diag_hq.p, ne, w->sc.p);
            }
        } else {
            if (which == 1) k_update<T, F, true><<<egrid(w), kBlock, 0, sys->stream>>>(r, q, w->dinv.p, sys->diag_hq.p, ne, w->sc.p, (double*)w->partial.p);
            else k_pupdate<T, F, true><<<egrid(w), kBlock, 0, sys->stream>>>(x, p, r, w->dinv.p, sys->diag_hq.p, ne, w->sc.p);
        }
    }
    static void vec_kernel(fdb_system* sys, SweepWork* w, bool overlay, int which) {
        if (w->real) vec_kernel_t<double>(sys, w, false, which); else vec_kernel_t<double2>(sys, w, overlay, which);
    }
    template <typename T>
    static void iteration_t(fdb_system* sys, SweepWork* w, bool overlay) {
        spmv(sys, w, w->p.p, w->q.p, true, overlay);
        if (w->precond == 2 && w->fused_precond) {
            modal_precond_fused(sys, w, false);          // K6a + projection on the tensor cores
            // coefficients, expansion + scalars (tile-compressed modes: coarse projection, coefficients, coarse expansion, z += Phi u + scalars)
            modal_apply(sys, w, overlay, false, sys->modal_compress ? 15 : (2 | 4));
        } else {
            vec_kernel_t<T>(sys, w, overlay, 1);
            if (w->precond == 2) modal_apply(sys, w, overlay, false);
        }
        vec_kernel_t<T>(sys, w, overlay, 2);
    }
    static void iteration(fdb_system* sys, SweepWork* w, bool overlay) {
        if (w->real) iteration_t<double>(sys, w, false); else iteration_t<double2>(sys, w, overlay);
    }
    // (re)load the slots flagged in sc.reset
    template <typename T>
    // keep_all: every slot being loaded starts from the x it holds (warm start); any_keep: some slots are resumed (sc.keepx)
    static void reload_t(fdb_system* sys, SweepWork* w, bool overlay, bool keep_all, bool any_keep, const int32_t* nodes, const double2* q, int n_src,
                         int* n_launched) {
        const int64_t ne = w->N * F;
        T *x = (T*)w->x.p, *r = (T*)w->r.p, *p = (T*)w->p.p, *qq = (T*)w->q.p;
        int nk = 0;
        if (overlay) {
            k_jacobi<F><<<grid_for(ne, kBlock), kBlock, 0, sys->stream>>>(sys->diag_hq.p, sys->ov_ptr.p, sys->ov_cg.p, sys->ov_val.p,
                                                                       w->coef.p, overlay, w->N, w->sc.p, w->dinv.p);
            ++nk;
        }
        k_slot_clear<T, F><<<egrid(w), kBlock, 0, sys->stream>>>(x, r, keep_all, ne, w->sc.p);
        ++nk;
        if (n_src) {
            k_set_rhs<T, F><<<grid_for((int64_t)n_src * F, 128), 128, 0, sys->stream>>>(nodes, q, w->src_set_ptr.p, n_src, false, w->sc.p, r);
            ++nk;
        }
        if constexpr (std::is_same<T, double2>::value) {
            if (w->n_rn) {
                k_add_rhs<F><<<grid_for((int64_t)w->n_rn * F, 128), 128, 0, sys->stream>>>(w->rn_nodes.p, w->rn_ptr.p, w->rn_vec.p, w->rn_val.p,
                                                                                         w->n_rn, w->rcoef.p, false, w->sc.p, r);
                ++nk;
            }
        }
        if (keep_all || any_keep) {   // r = b - G x (x = 0 in the slots that keep nothing)
            spmv(sys, w, x, qq, false, overlay);
            k_sub<T, F><<<egrid(w), kBlock, 0, sys->stream>>>(r, qq, ne, w->sc.p);
            nk += 2;
        }
        bool started = false;
        if (w->precond == 2 && w->fused_precond) {
            modal_precond_fused(sys, w, true);
            modal_apply(sys, w, overlay, true, sys->modal_compress ? 15 : (2 | 4));
            nk += sys->modal_compress ? 4 : 2;
            started = true;
        } else {
            if (w->precond >= 1) started = bj_update<T>(sys, w, true);
            if (started && w->precond == 2) { modal_apply(sys, w, overlay, true); nk += 3; }
        }
        if (started) {
        } else if (overlay) {
            if constexpr (std::is_same<T, double2>::value)
                k_start<T, F, false><<<egrid(w), kBlock, 0, sys->stream>>>(p, r, w->dinv.p, sys->diag_hq.p, ne, w->sc.p, (double*)w->partial.p);
        } else {
            k_start<T, F, true><<<egrid(w), kBlock, 0, sys->stream>>>(p, r, w->dinv.p, sys->diag_hq.p, ne, w->sc.p, (double*)w->partial.p);
        }
        *n_launched = nk + 1;
    }
    static void reload(fdb_system* sys, SweepWork* w, bool overlay, bool keep_all, bool any_keep, const int32_t* nodes, const double2* q, int n_src,
                       int* n_launched) {
        if (w->real) reload_t<double>(sys, w, false, keep_all, any_keep, nodes, q, n_src, n_launched);
        else reload_t<double2>(sys, w, overlay, keep_all, any_keep, nodes, q, n_src, n_launched);
    }
    // rt[f] = ||b_f - G_f x_f||^2 for every slot (q and tmp are free between iterations)
    template <typename T>
    static void true_residual_t(fdb_system* sys, SweepWork* w, bool overlay, const int32_t* nodes, const double2* q, int n_src) {
        const int64_t ne = w->N * F;
        T *x = (T*)w->x.p, *qq = (T*)w->q.p, *tmp = (T*)w->tmp.p;
        spmv(sys, w, x, qq, false, overlay);
        FDB_CUDA(cudaMemsetAsync(tmp, 0, (size_t)ne * sizeof(T), sys->stream));
        if (n_src) k_set_rhs<T, F><<<grid_for((int64_t)n_src * F, 128), 128, 0, sys->stream>>>(nodes, q, w->src_set_ptr.p, n_src, true, w->sc.p, tmp);
        if constexpr (std::is_same<T, double2>::value) {
            if (w->n_rn)
                k_add_rhs<F><<<grid_for((int64_t)w->n_rn * F, 128), 128, 0, sys->stream>>>(w->rn_nodes.p, w->rn_ptr.p, w->rn_vec.p, w->rn_val.p,
                                                                                         w->n_rn, w->rcoef.p, true, w->sc.p, tmp);
        }
        k_resnorm<T, F><<<egrid(w), kBlock, 0, sys->stream>>>(tmp, qq, ne, w->sc.p, (double*)w->partial.p);
    }
    static void true_residual(fdb_system* sys, SweepWork* w, bool overlay, const int32_t* nodes, const double2* q, int n_src, int* n_launched) {
        if (w->real) true_residual_t<double>(sys, w, false, nodes, q, n_src); else true_residual_t<double2>(sys, w, overlay, nodes, q, n_src);
        *n_launched = 3;
    }
    static void gather(fdb_system* sys, SweepWork* w, const int32_t* nodes, const double* wts, int n_rcv) {
        const int g = grid_for((int64_t)n_rcv * F, 128);
        if (w->real) k_gather_rcv<double, F><<<g, 128, 0, sys->stream>>>((const double*)w->x.p, nodes, wts, n_rcv, w->rcv_out.p);
        else k_gather_rcv<double2, F><<<g, 128, 0, sys->stream>>>(w->x.p, nodes, wts, n_rcv, w->rcv_out.p);
    }
    static void export_pn(fdb_system* sys, SweepWork* w) {
        const int rpb = kBlock / F;
        if (w->real) k_export<double, F><<<grid_for(w->N, rpb), kBlock, 0, sys->stream>>>((const double*)w->x.p, sys->perm.p, w->N, w->stage.p);
        else k_export<double2, F><<<grid_for(w->N, rpb), kBlock, 0, sys->stream>>>(w->x.p, sys->perm.p, w->N, w->stage.p);
    }
    static void query_real(fdb_system* sys, SweepWork* w, bool* ok) { *ok = real_supported(sys, w); }
};

#define FDB_DISPATCH_F(F, CALL)                                   \
    switch (F) {                                                  \
        case 1: Launch<1>::CALL; break;                           \
        case 2: Launch<2>::CALL; break;                           \
        case 4: Launch<4>::CALL; break;                           \
        case 8: Launch<8>::CALL; break;                           \
        case 16: Launch<16>::CALL; break;                         \
        case 32: Launch<32>::CALL; break;                         \
        case 64: Launch<64>::CALL; break;                         \
        default: throw Error("femder_b200: batch must be 1, 2, 4, 8, 16, 32 or 64"); \
    }

static SweepWork* new_work(fdb_system* sys, int F);

static SweepWork* get_work(fdb_system* sys, int F) {
    FDB_REQUIRE(F == 1 || F == 2 || F == 4 || F == 8 || F == 16 || F == 32 || F == 64, "batch must be 1, 2, 4, 8, 16, 32 or 64");
    FDB_REQUIRE(sys->have_pattern && sys->have_hq, "pattern and H/Q values must be set before the sweep");
    SweepWork* w = sys->work;
    if (w && (w->F != F || w->N != sys->n_nodes || w->n_groups != sys->n_groups)) {
        free_sweep_work(w);
        w = sys->work = nullptr;
    }
    if (!w) w = sys->work = new_work(sys, F);
    return w;
}

static SweepWork* new_work(fdb_system* sys, int F) {
    SweepWork* w = nullptr;
    {
        w = new SweepWork();
        w->F = F;
        w->N = sys->n_nodes;
        w->n_groups = sys->n_groups;
        w->grid = sys->num_sms * 8;
#ifdef FDB_TUNING
        // Tuning knobs that change the code path (FDB_BLOB_DBG even makes results wrong on purpose) exist only in builds made
        // with -DFDB_TUNING (scripts/build_ab_lib.sh): a stray environment variable cannot change what the product computes.
        if (const char* e = getenv("FDB_SPMV_CFG")) w->spmv_cfg = std::max(0, std::min(12, atoi(e)));
        if (const char* e = getenv("FDB_SPMV_BPS")) w->spmv_bps = std::max(0, atoi(e));
        if (const char* e = getenv("FDB_BLOB_LPR")) w->blob_lpr = atoi(e) == 4 ? 4 : 2;
        if (const char* e = getenv("FDB_BLOB_DBG")) w->blob_dbg = atoi(e);
        if (getenv("FDB_NO_ROWS_KERNEL")) w->rows_kernel = false;
        if (getenv("FDB_NO_FUSED_PRECOND")) w->fused_precond = false;   // A/B: k_update_bj + k_modal_project instead of k_precond_fused
#endif
        const size_t n = (size_t)w->N * F;
        w->x.alloc(n); w->r.alloc(n); w->p.alloc(n); w->q.alloc(n); w->dinv.alloc(n); w->tmp.alloc(n);
        w->coef.alloc((size_t)std::max(1, sys->n_groups) * F);
        w->partial.alloc((size_t)w->grid * F * 2);  // 3 doubles per (block, f) fit in 2 double2
        w->sc.alloc(1);
        w->sc.zero(sys->stream);
    }
    // dummies so kernels always receive valid pointers
    if (!sys->ov_ptr.p) {
        sys->ov_ptr.alloc(sys->n_nodes + 1);
        sys->ov_ptr.zero(sys->stream);
        sys->ov_cg.alloc(1);
        sys->ov_val.alloc(1);
        sys->n_ov = 0;
    }
    return w;
}

// host mirror of one slot's frequency-dependent constants
static void fill_slot(Scalars& h, std::vector<double2>& coef, int F, int n_groups, int slot, double om, const double* mu /* [n_groups] complex */,
                      double src_norm2) {
    h.w2[slot] = om * om;
    h.rhs_scale[slot] = make_double2(0.0, -om);   // b = -1j*w*q
    h.bb[slot] = om * om * src_norm2;
    for (int g = 0; g < n_groups; ++g) {
        const double mr = mu ? mu[2 * g] : 0.0, mi = mu ? mu[2 * g + 1] : 0.0;
        coef[(size_t)g * F + slot] = make_double2(-om * mi, om * mr);  // 1j*om*(mr + 1j*mi)
    }
}

static void upload_scalars(fdb_system* sys, SweepWork* w, const Scalars& h, const std::vector<double2>& coef,
                           const std::vector<double2>* rcoef = nullptr) {
    FDB_CUDA(cudaMemcpyAsync(w->sc.p, &h, sizeof(h), cudaMemcpyHostToDevice, sys->stream));
    FDB_CUDA(cudaMemcpyAsync(w->coef.p, coef.data(), coef.size() * sizeof(double2), cudaMemcpyHostToDevice, sys->stream));
    if (rcoef && !rcoef->empty()) {
        w->rcoef.alloc(rcoef->size());
        FDB_CUDA(cudaMemcpyAsync(w->rcoef.p, rcoef->data(), rcoef->size() * sizeof(double2), cudaMemcpyHostToDevice, sys->stream));
    }
    FDB_CUDA(cudaStreamSynchronize(sys->stream));  // pageable host buffers: be done before they change
}

static void ensure_graph(fdb_system* sys, SweepWork* w, int F, int iters, bool overlay) {
    if (w->graph && w->graph_iters == iters && w->graph_F == F && w->graph_overlay == overlay && w->graph_real == w->real &&
        w->graph_precond == w->precond) return;
    if (w->graph) { cudaGraphExecDestroy(w->graph); w->graph = nullptr; }
    cudaGraph_t g = nullptr;
    FDB_CUDA(cudaStreamBeginCapture(sys->stream, cudaStreamCaptureModeThreadLocal));
    try {
        for (int i = 0; i < iters; ++i) { FDB_DISPATCH_F(F, iteration(sys, w, overlay)); }
    } catch (...) {
        cudaStreamEndCapture(sys->stream, &g);
        if (g) cudaGraphDestroy(g);
        throw;
    }
    FDB_CUDA(cudaStreamEndCapture(sys->stream, &g));
    cudaError_t e = cudaGraphInstantiate(&w->graph, g, 0);
    cudaGraphDestroy(g);
    FDB_CUDA(e);
    w->graph_iters = iters; w->graph_F = F; w->graph_overlay = overlay; w->graph_real = w->real; w->graph_precond = w->precond;
}

static void k_jacobi_launch(fdb_system* sys, SweepWork* w, int F) {
    const int64_t ne = w->N * F;
    auto go = [&](auto kern) {
        kern<<<grid_for(ne, kBlock), kBlock, 0, sys->stream>>>(sys->diag_hq.p, sys->ov_ptr.p, sys->ov_cg.p, sys->ov_val.p, w->coef.p, true,
                                                             w->N, w->sc.p, w->dinv.p);
    };
    switch (F) {
        case 1: go(k_jacobi<1>); break;
        case 2: go(k_jacobi<2>); break;
        case 4: go(k_jacobi<4>); break;
        case 8: go(k_jacobi<8>); break;
        case 16: go(k_jacobi<16>); break;
        case 32: go(k_jacobi<32>); break;
        case 64: go(k_jacobi<64>); break;
        default: throw Error("femder_b200: unsupported batch width");
    }
}

// The sweep.  F slots iterate together; a slot whose frequency has converged is reloaded with the next frequency
// of the list while the others keep going, so no bandwidth is spent on finished systems.
void run_sweep(fdb_system* sys, const fdb_sweep_params* prm, fdb_sweep_result* res) {
    FDB_REQUIRE(prm && res, "null argument");
    FDB_REQUIRE(prm->n_freq >= 0 && prm->n_src >= 0 && prm->n_rcv >= 0, "negative count");
    FDB_REQUIRE(prm->tol > 0 && prm->max_iter > 0, "tol and max_iter must be positive");
    int F = prm->batch;
    FDB_REQUIRE(F == 1 || F == 2 || F == 4 || F == 8 || F == 16 || F == 32 || F == 64, "batch must be 1, 2, 4, 8, 16, 32 or 64");
    const int64_t N = sys->n_nodes;
    const int ng = sys->n_groups;
    FDB_REQUIRE(ng == 0 || prm->mu != nullptr, "mu is required when the system has boundary groups");
    const int check_every = std::max(1, prm->check_every);
    const int64_t nf = prm->n_freq;
    cudaStream_t s = sys->stream;

    // inputs to the host (they may be device pointers)
    std::vector<double> omega(nf), mu((size_t)ng * nf * 2);
    if (nf) FDB_CUDA(cudaMemcpy(omega.data(), prm->omega, omega.size() * sizeof(double), cudaMemcpyDefault));
    if (!mu.empty()) FDB_CUDA(cudaMemcpy(mu.data(), prm->mu, mu.size() * sizeof(double), cudaMemcpyDefault));
    // a sweep whose admittances are all exactly zero (rigid walls) has no boundary term: the overlay is skipped
    bool any_mu = false;
    for (double v : mu) any_mu = any_mu || (v != 0.0);
    // never solve the rigid problem silently: non-zero admittances need the assembled surface matrices
    FDB_REQUIRE(!(ng > 0 && any_mu) || sys->have_overlay,
                "non-zero admittances but the surface matrices are not assembled (call fdb_assemble_surface / fdb_set_surface after the last mesh or pattern change)");
    const bool overlay = sys->have_overlay && sys->n_ov > 0 && ng > 0 && any_mu;
    std::vector<int32_t> src_nodes(prm->n_src);
    std::vector<double2> src_q(prm->n_src);
    if (prm->n_src) {
        FDB_CUDA(cudaMemcpy(src_nodes.data(), prm->src_nodes, prm->n_src * sizeof(int32_t), cudaMemcpyDefault));
        FDB_CUDA(cudaMemcpy(src_q.data(), prm->src_q, prm->n_src * sizeof(double2), cudaMemcpyDefault));
    }
    // Source sets = right-hand sides sharing G(w) (one set unless the caller batches source candidates).  Inside a
    // set q[node] = S.q is an assignment: a later source on the same node overwrites (FEM_3D.py:1298-1300).
    const int64_t n_sets = std::max<int64_t>(1, prm->n_src_sets);
    std::vector<int32_t> set_in(n_sets + 1, 0);
    if (prm->n_src_sets > 0) {
        FDB_REQUIRE(prm->src_set_ptr != nullptr, "src_set_ptr missing");
        FDB_CUDA(cudaMemcpy(set_in.data(), prm->src_set_ptr, (n_sets + 1) * sizeof(int32_t), cudaMemcpyDefault));
        FDB_REQUIRE(set_in[0] == 0 && set_in[n_sets] == prm->n_src, "src_set_ptr does not span the sources");
    } else {
        set_in[1] = (int32_t)prm->n_src;
    }
    std::vector<int32_t> u_nodes, set_ptr(n_sets + 1, 0);
    std::vector<double2> u_q;
    std::vector<double> set_norm2(n_sets, 0.0);
    int max_set = 0;
    for (int64_t st = 0; st < n_sets; ++st) {
        FDB_REQUIRE(set_in[st] <= set_in[st + 1], "src_set_ptr is not monotone");
        const size_t base = u_nodes.size();
        for (int32_t i = set_in[st]; i < set_in[st + 1]; ++i) {
            FDB_REQUIRE(src_nodes[i] >= 0 && src_nodes[i] < N, "source node out of range");
            bool found = false;
            for (size_t j = base; j < u_nodes.size(); ++j)
                if (u_nodes[j] == src_nodes[i]) { u_q[j] = src_q[i]; found = true; }
            if (!found) { u_nodes.push_back(src_nodes[i]); u_q.push_back(src_q[i]); }
        }
        set_ptr[st + 1] = (int32_t)u_nodes.size();
        max_set = std::max(max_set, (int)(u_nodes.size() - base));
        for (size_t j = base; j < u_nodes.size(); ++j) set_norm2[st] += u_q[j].x * u_q[j].x + u_q[j].y * u_q[j].y;
    }
    // the sweep's vectors live in the internal (blob) order: node lists are mapped once, here
    const std::vector<int32_t>& perm = sys->perm_host;
    FDB_REQUIRE(sys->have_order && (int64_t)perm.size() == N, "internal order missing");
    for (auto& v : u_nodes) v = perm[v];
    const int n_src = max_set;            // launch width of the right-hand-side kernels
    const int64_t n_sys = n_sets * nf;    // systems = (source set, frequency) pairs, index set*nf + frequency
    // extra right-hand-side vectors (velocity boundary conditions): regrouped per node so one thread owns a node
    const int64_t nrv = prm->n_rhs_vec;
    FDB_REQUIRE(nrv >= 0, "negative count");
    std::vector<int32_t> rn_nodes, rn_ptr, rn_vec;
    std::vector<double> rn_val, rhs_coef((size_t)nrv * nf * 2);
    if (nrv) {
        FDB_REQUIRE(prm->rhs_ptr && prm->rhs_coef, "right-hand-side vector arrays missing");
        std::vector<int32_t> rp(nrv + 1);
        FDB_CUDA(cudaMemcpy(rp.data(), prm->rhs_ptr, rp.size() * sizeof(int32_t), cudaMemcpyDefault));
        const int64_t ne = rp[nrv];
        FDB_REQUIRE(rp[0] == 0 && ne >= 0 && (ne == 0 || (prm->rhs_nodes && prm->rhs_vals)), "bad right-hand-side vector arrays");
        std::vector<int32_t> en(ne);
        std::vector<double> evl(ne);
        if (ne) {
            FDB_CUDA(cudaMemcpy(en.data(), prm->rhs_nodes, ne * sizeof(int32_t), cudaMemcpyDefault));
            FDB_CUDA(cudaMemcpy(evl.data(), prm->rhs_vals, ne * sizeof(double), cudaMemcpyDefault));
        }
        if (!rhs_coef.empty()) FDB_CUDA(cudaMemcpy(rhs_coef.data(), prm->rhs_coef, rhs_coef.size() * sizeof(double), cudaMemcpyDefault));
        std::vector<std::array<int64_t, 2>> ent;   // (node, entry index)
        for (int64_t k = 0; k < nrv; ++k) {
            FDB_REQUIRE(rp[k] <= rp[k + 1], "rhs_ptr is not monotone");
            for (int32_t e = rp[k]; e < rp[k + 1]; ++e) {
                FDB_REQUIRE(en[e] >= 0 && en[e] < N, "right-hand-side node out of range");
                ent.push_back({(int64_t)en[e], (int64_t)e});
            }
        }
        std::vector<int32_t> vec_of(ne);
        for (int64_t k = 0; k < nrv; ++k)
            for (int32_t e = rp[k]; e < rp[k + 1]; ++e) vec_of[e] = (int32_t)k;
        std::stable_sort(ent.begin(), ent.end(), [](const std::array<int64_t, 2>& a, const std::array<int64_t, 2>& b) { return a[0] < b[0]; });
        rn_ptr.push_back(0);
        for (size_t i = 0; i < ent.size(); ++i) {
            if (i == 0 || ent[i][0] != ent[i - 1][0]) {
                if (i) rn_ptr.push_back((int32_t)rn_vec.size());
                rn_nodes.push_back((int32_t)ent[i][0]);
            }
            rn_vec.push_back(vec_of[ent[i][1]]);
            rn_val.push_back(evl[ent[i][1]]);
        }
        if (!ent.empty()) rn_ptr.push_back((int32_t)rn_vec.size());
        for (auto& v : rn_nodes) v = perm[v];
    }
    const bool general_rhs = !rn_nodes.empty();

    // Real arithmetic when the systems are real (no boundary term) and the right-hand sides purely imaginary
    // (real q): p = 1j*u with G u = -w q.  Same COCG recurrences on half the bytes.
    bool real = !overlay && prm->arithmetic != 1 && !general_rhs;
    for (auto& v : u_q) real = real && (v.y == 0.0);
    // The modal coarse correction (precond 2) runs on 16 real columns per row: 16 real or 8 complex slots.
    const bool want_modal = prm->precond == 2;
    if (want_modal) {
        FDB_REQUIRE(sys->have_modes, "precond = 2 needs the room modes (fdb_modal_compute / fdb_modal_set)");
        FDB_REQUIRE(sys->n_tiles > 0 && sys->tile_rows == 128, "the modal correction needs the 128-row tiles of the internal order");
        F = real ? kModalCols : kModalCols / 2;
    } else {
        // no wider than the number of frequencies needs
        while (F / 2 >= 1 && F / 2 >= n_sys) F /= 2;
    }
    SweepWork* w = get_work(sys, F);
    w->real = false;
    if (real) {
        bool ok = false;
        FDB_DISPATCH_F(F, query_real(sys, w, &ok));
        real = ok;
    }
    w->real = real;
    res->real_arithmetic = real ? 1 : 0;
    // Preconditioner: block-Jacobi over the 128-row tiles of the internal order where it is implemented (real sweeps in
    // batches of up to 16, complex sweeps in batches of up to 8), point Jacobi otherwise.  prm->precond: 0 = Jacobi, 1 = block-Jacobi where possible.
#ifdef FDB_TUNING
    static const bool no_bj = getenv("FDB_NO_BLOCK_JACOBI") != nullptr;   // A/B aid (tuning builds only)
#else
    constexpr bool no_bj = false;
#endif
    bool use_bj = prm->precond >= 1 && !no_bj && sys->n_tiles > 0 && sys->tile_rows == 128;
    if (use_bj) { FDB_DISPATCH_F(F, bj_supported(w, &use_bj)); }   // 16 real columns per row at most: F <= 16 real, F <= 8 complex
    bool use_modal = want_modal && use_bj;
    if (want_modal) {
        FDB_REQUIRE(real == w->real, "the modal correction needs the real-arithmetic kernels at 16 slots");
        FDB_DISPATCH_F(F, modal_supported(w, &use_modal));
        FDB_REQUIRE(use_modal && use_bj, "the modal correction is not available for this batch shape");
    }
    w->precond = use_modal ? 2 : (use_bj ? 1 : 0);
    if (use_bj) { FDB_DISPATCH_F(F, bj_prepare(sys, w)); }
    if (use_modal) {
        modal_prepare(sys, w);
        if (overlay) modal_build_adiag(sys);
    }
    const double modal_delta = prm->modal_delta_hz > 0.0 ? prm->modal_delta_hz : 300.0;
    w->n_rn = (int)rn_nodes.size();
    if (general_rhs) {
        w->rn_nodes.alloc(rn_nodes.size()); w->rn_ptr.alloc(rn_ptr.size()); w->rn_vec.alloc(rn_vec.size()); w->rn_val.alloc(rn_val.size());
        FDB_CUDA(cudaMemcpy(w->rn_nodes.p, rn_nodes.data(), rn_nodes.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        FDB_CUDA(cudaMemcpy(w->rn_ptr.p, rn_ptr.data(), rn_ptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        FDB_CUDA(cudaMemcpy(w->rn_vec.p, rn_vec.data(), rn_vec.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        FDB_CUDA(cudaMemcpy(w->rn_val.p, rn_val.data(), rn_val.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    DevBuf<int32_t> d_src_nodes; DevBuf<double2> d_src_q;
    d_src_nodes.alloc(std::max<size_t>(1, u_nodes.size())); d_src_q.alloc(std::max<size_t>(1, u_nodes.size()));
    w->src_set_ptr.alloc(set_ptr.size());
    FDB_CUDA(cudaMemcpyAsync(w->src_set_ptr.p, set_ptr.data(), set_ptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    if (!u_nodes.empty()) {
        FDB_CUDA(cudaMemcpyAsync(d_src_nodes.p, u_nodes.data(), u_nodes.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        FDB_CUDA(cudaMemcpyAsync(d_src_q.p, u_q.data(), u_q.size() * sizeof(double2), cudaMemcpyHostToDevice, s));
    }
    const int n_rcv = (int)prm->n_rcv;
    DevBuf<int32_t> d_rcv_nodes; DevBuf<double> d_rcv_w;
    if (n_rcv) {
        FDB_REQUIRE(prm->rcv_nodes && prm->rcv_w, "receiver arrays missing");
        std::vector<int32_t> rn((size_t)n_rcv * 4);
        FDB_CUDA(cudaMemcpy(rn.data(), prm->rcv_nodes, rn.size() * sizeof(int32_t), cudaMemcpyDefault));
        for (auto& v : rn) {
            FDB_REQUIRE(v >= 0 && v < N, "receiver node out of range");
            v = perm[v];
        }
        d_rcv_nodes.alloc(rn.size()); d_rcv_w.alloc(rn.size());
        FDB_CUDA(cudaMemcpyAsync(d_rcv_nodes.p, rn.data(), rn.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        FDB_CUDA(cudaMemcpyAsync(d_rcv_w.p, prm->rcv_w, rn.size() * sizeof(double), cudaMemcpyDefault, s));
        w->rcv_out.alloc((size_t)F * n_rcv);
    }
    if (res->pN) w->stage.alloc((size_t)F * N);
    FDB_CUDA(cudaStreamSynchronize(s));

    struct Events {
        cudaEvent_t a = nullptr, b = nullptr;
        ~Events() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } evs;
    FDB_CUDA(cudaEventCreate(&evs.a));
    FDB_CUDA(cudaEventCreate(&evs.b));
    cudaEvent_t ev0 = evs.a, ev1 = evs.b;
    FDB_CUDA(cudaEventRecord(ev0, s));
    int64_t n_spmv = 0, n_kernels = 0;
    int n_unconv = 0, n_resumed = 0;

    Scalars h;
    memset(&h, 0, sizeof(h));
    h.tol2 = prm->tol * prm->tol;
    h.bb_from_r0 = general_rhs ? 1 : 0;
    std::vector<double2> rcoef((size_t)nrv * F, make_double2(0.0, 0.0));   // [n_rhs_vec][F]
    std::vector<double2> coef((size_t)std::max(1, ng) * F, make_double2(0.0, 0.0));
    std::vector<double> mu_f((size_t)std::max(1, ng) * 2);
    std::vector<int64_t> slot_freq(F, -1);
    std::vector<char> slot_used(F, 0);   // has the slot held a frequency before (its x is a usable starting guess)
    std::vector<int> slot_resumes(kMaxF, 0);   // times the slot's system was restarted from its true residual
    std::vector<double> slot_best(kMaxF, 1e300);   // modal path: best ||r||^2 / ||b||^2 the slot's recurrence has shown, and when
    std::vector<int> slot_best_it(kMaxF, 0);
    std::vector<char> slot_gave_up(kMaxF, 0);
    constexpr int kMaxResumes = 8, kStallWindow = 96;
    constexpr double kNearResonance = 1e-3;   // |lambda - w^2| / w^2 below which a system is dispatched ahead of the others
    std::vector<double2> rcv_host((size_t)F * std::max(1, n_rcv));
    int64_t next = 0;
    int n_busy = 0;

    // Frequencies are dispatched to the slots in descending order: the iteration count grows with frequency, and
    // starting the expensive systems first leaves the cheap ones for the tail where slots begin to idle.
    // With the modal correction the iteration count no longer depends on the frequency - except next to a resonance, where the
    // stored (approximate) mode cannot be deflated cleanly and a system may need several times the usual count: those go first,
    // so the long ones never start when the queue is about to run dry.
    std::vector<int64_t> order(n_sys);
    for (int64_t i = 0; i < n_sys; ++i) order[i] = i;
    std::vector<char> near_mode(nf, 0);
    if (use_modal) {
        const auto& lam = sys->modal_lam_host;
        for (int64_t i = 0; i < nf; ++i) {
            const double w2 = omega[i] * omega[i];
            auto it = std::lower_bound(lam.begin(), lam.end(), w2);
            double d = 1e300;
            if (it != lam.end()) d = std::min(d, std::fabs(*it - w2));
            if (it != lam.begin()) d = std::min(d, std::fabs(*(it - 1) - w2));
            near_mode[i] = d < kNearResonance * w2;
        }
    }
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        const int64_t fa = a % nf, fb = b % nf;
        if (near_mode[fa] != near_mode[fb]) return near_mode[fa] > near_mode[fb];
        return omega[fa] > omega[fb];
    });

    // the next system of the queue: this call's own, or the one the processes of a node share (fdb_sweep_params.shared_next)
    bool queue_dry = false;
    auto claim = [&]() -> int64_t {
        if (queue_dry) return -1;
        const int64_t k = prm->shared_next ? __atomic_fetch_add(prm->shared_next, (int64_t)1, __ATOMIC_RELAXED) : next++;
        if (k >= n_sys) { queue_dry = true; return -1; }
        return order[k];
    };
    if (res->solved) memset(res->solved, 0, (size_t)n_sys);
    auto assign = [&](int slot) {
        const int64_t si = claim();
        if (si < 0) {
            slot_freq[slot] = -1;
            h.active[slot] = 0;
            h.reset[slot] = 0;
            return false;
        }
        const int64_t fi = si % nf, st = si / nf;
        for (int g = 0; g < ng; ++g) {
            mu_f[2 * g] = mu[((size_t)g * nf + fi) * 2];
            mu_f[2 * g + 1] = mu[((size_t)g * nf + fi) * 2 + 1];
        }
        fill_slot(h, coef, F, ng, slot, omega[fi], ng ? mu_f.data() : nullptr, set_norm2[st]);
        h.src_set[slot] = (int)st;
        for (int64_t k = 0; k < nrv; ++k)
            rcoef[(size_t)k * F + slot] = make_double2(rhs_coef[((size_t)k * nf + fi) * 2], rhs_coef[((size_t)k * nf + fi) * 2 + 1]);
        slot_freq[slot] = si;
        h.reset[slot] = 1;
        h.keepx[slot] = 0;
        h.active[slot] = 0;
        h.iters[slot] = 0;
        slot_resumes[slot] = 0;
        slot_best[slot] = 1e300;
        slot_best_it[slot] = 0;
        slot_gave_up[slot] = 0;
        ++n_busy;
        return true;
    };
    // modes the coarse correction applies in the next launches: all below sqrt(f_max^2 + delta^2) of the busy slots
    auto set_modes_use = [&]() {
        if (!use_modal) { h.modes_use = 0; return; }
        double w2max = 0.0;
        for (int f = 0; f < F; ++f)
            if (slot_freq[f] >= 0) w2max = std::max(w2max, h.w2[f]);
        h.modes_use = modal_modes_for(sys, w2max, modal_delta);
    };

    // empty slots still need finite constants
    for (int f = 0; f < F; ++f) fill_slot(h, coef, F, ng, f, 1.0, nullptr, 0.0);
    bool any_reset = false;
    for (int f = 0; f < F; ++f) any_reset = assign(f) || any_reset;
    if (n_sys > 0) ensure_graph(sys, w, F, check_every, overlay);

    while (n_busy > 0) {
        if (any_reset) {
            // slots that held a solution before can start from it (warm start from the slot's previous frequency)
            bool keep_x = prm->warm_start != 0 && !general_rhs;   // ||b|| of a general right-hand side comes from r0 = b
            bool any_keep = false;
            for (int f = 0; f < F; ++f) {
                if (h.reset[f] && !slot_used[f]) keep_x = false;
                if (h.reset[f] && h.keepx[f]) any_keep = true;
            }
            set_modes_use();
            upload_scalars(sys, w, h, coef, &rcoef);
            int nk = 0;
            FDB_DISPATCH_F(F, reload(sys, w, overlay, keep_x, any_keep, d_src_nodes.p, d_src_q.p, n_src, &nk));
            n_kernels += nk;
            if (keep_x || any_keep) ++n_spmv;
            for (int f = 0; f < F; ++f)
                if (h.reset[f]) { slot_used[f] = 1; h.reset[f] = 0; h.keepx[f] = 0; }   // the device copy clears them as it loads the slots
            any_reset = false;
        } else {
            set_modes_use();
            upload_scalars(sys, w, h, coef, &rcoef);
        }
        FDB_CUDA(cudaGraphLaunch(w->graph, s));
        n_spmv += check_every;
        n_kernels += (use_modal ? (sys->modal_compress ? 7 : (w->fused_precond ? 5 : 6)) : 3) * (int64_t)check_every;
        FDB_CUDA(cudaMemcpyAsync(&h, w->sc.p, sizeof(Scalars), cudaMemcpyDeviceToHost, s));
        FDB_CUDA(cudaStreamSynchronize(s));
        FDB_CUDA(cudaGetLastError());

        // The modal correction makes the preconditioner indefinite (it inverts G on the modes below the drive frequency too): the
        // recurrence then converges like CG on a definite system, but an (unlucky) near-breakdown is possible.  A slot that has
        // not improved its residual for a long stretch, or has blown it up, is restarted from its true residual.
        if (use_modal) {
            for (int f = 0; f < F; ++f) {
                if (slot_freq[f] < 0 || !h.active[f]) continue;
                const double rrel = h.bb[f] > 0 ? h.rr[f] / h.bb[f] : 0.0;
                if (rrel < 0.25 * slot_best[f]) { slot_best[f] = rrel; slot_best_it[f] = h.iters[f]; continue; }
                const bool stalled = h.iters[f] - slot_best_it[f] >= kStallWindow || rrel > 1e8 * slot_best[f] || rrel != rrel;
                if (stalled && slot_resumes[f] >= kMaxResumes) {
                    // restarts exhausted: hand the slot back as it is (counted as unconverged below) instead of running to max_iter
                    h.active[f] = 0;
                    slot_gave_up[f] = 1;
                } else if (stalled && h.iters[f] < prm->max_iter) {
                    ++slot_resumes[f];
                    ++n_resumed;
                    h.reset[f] = 1;
                    h.keepx[f] = 1;
                    h.active[f] = 0;
                    slot_best[f] = 1e300;
                    slot_best_it[f] = h.iters[f];
                    any_reset = true;
                }
            }
        }
        bool finished_any = false;
        std::vector<char> done(kMaxF, 0);
        for (int f = 0; f < F; ++f) {
            if (slot_freq[f] < 0) continue;
            if (h.reset[f]) continue;   // being restarted (above)
            if (!h.active[f] || h.iters[f] >= prm->max_iter) { done[f] = 1; finished_any = true; }
        }
        if (!finished_any) continue;

        // true residual of what is returned, receivers, full pressure rows - only for the finished slots
        int nk = 0;
        FDB_DISPATCH_F(F, true_residual(sys, w, overlay, d_src_nodes.p, d_src_q.p, n_src, &nk));
        n_kernels += nk; n_spmv += 1;
        Scalars hf;
        FDB_CUDA(cudaMemcpyAsync(&hf, w->sc.p, sizeof(Scalars), cudaMemcpyDeviceToHost, s));
        if (n_rcv) {
            FDB_DISPATCH_F(F, gather(sys, w, d_rcv_nodes.p, d_rcv_w.p, n_rcv));
            n_kernels += 1;
            FDB_CUDA(cudaMemcpyAsync(rcv_host.data(), w->rcv_out.p, (size_t)F * n_rcv * sizeof(double2), cudaMemcpyDeviceToHost, s));
        }
        if (res->pN) {
            FDB_DISPATCH_F(F, export_pn(sys, w));
            n_kernels += 1;
            for (int f = 0; f < F; ++f)
                if (done[f])
                    FDB_CUDA(cudaMemcpyAsync(res->pN + (size_t)slot_freq[f] * N * 2, w->stage.p + (size_t)f * N, (size_t)N * sizeof(double2),
                                             cudaMemcpyDefault, s));
        }
        FDB_CUDA(cudaStreamSynchronize(s));
        for (int f = 0; f < F; ++f) {
            if (!done[f]) continue;
            const int64_t fi = slot_freq[f];
            const double rel = (hf.bb[f] > 0) ? sqrt(hf.rt[f] / hf.bb[f]) : 0.0;
            // Accepted on the TRUE residual only.  The recurrence stopped the slot but ||b - G x|| is still above tol: restart
            // it from x with r = b - G x (the iteration count keeps running) instead of returning it.
            if (!h.active[f] && !(rel <= prm->tol) && rel == rel && h.iters[f] < prm->max_iter && slot_resumes[f] < kMaxResumes && !slot_gave_up[f]) {
                ++slot_resumes[f];
                ++n_resumed;
                h.reset[f] = 1;
                h.keepx[f] = 1;
                any_reset = true;
                continue;
            }
            if (res->solved) res->solved[fi] = 1;
            if (res->iters) res->iters[fi] = h.iters[f];
            if (res->relres) res->relres[fi] = rel;
            if (!(rel <= prm->tol)) ++n_unconv;
            if (n_rcv && res->pR)
                FDB_CUDA(cudaMemcpy(res->pR + ((size_t)fi * n_rcv) * 2, rcv_host.data() + (size_t)f * n_rcv,
                                    (size_t)n_rcv * sizeof(double2), cudaMemcpyDefault));
            --n_busy;
            any_reset = assign(f) || any_reset;
        }

        // Tail of the sweep: the queue is empty and slots fall idle, but the fused SpMV costs the same whatever the
        // number of live slots.  Move the live slots into a narrower batch as soon as they fit one.
#ifdef FDB_TUNING
        static const bool no_narrow = getenv("FDB_NO_NARROW") != nullptr;   // debugging aid (tuning builds only)
#else
        constexpr bool no_narrow = false;
#endif
        // (the modal correction's tensor-core kernels are built for 16 real columns: its batches stay as they are)
        if (queue_dry && n_busy > 0 && !any_reset && !no_narrow && !use_modal) {
            int F2 = F;
            while (F2 / 2 >= 1 && F2 / 2 >= n_busy) F2 /= 2;
            if (F2 < F) {
                SweepWork* w2 = new_work(sys, F2);
                w2->real = w->real;
                w2->precond = w->precond;   // a narrower batch always has the kernels a wider one has
                if (w2->precond == 1) { FDB_DISPATCH_F(F2, bj_prepare(sys, w2)); }
                SlotMap map;
                Scalars h2;
                memset(&h2, 0, sizeof(h2));
                h2.tol2 = h.tol2;
                std::vector<double2> coef2((size_t)std::max(1, ng) * F2, make_double2(0.0, 0.0));
                std::vector<double2> rcoef2((size_t)nrv * F2, make_double2(0.0, 0.0));
                h2.bb_from_r0 = h.bb_from_r0;
                w2->n_rn = w->n_rn;
                std::swap(w2->src_set_ptr.p, w->src_set_ptr.p); std::swap(w2->src_set_ptr.n, w->src_set_ptr.n);
                std::swap(w2->rn_nodes.p, w->rn_nodes.p); std::swap(w2->rn_nodes.n, w->rn_nodes.n);
                std::swap(w2->rn_ptr.p, w->rn_ptr.p); std::swap(w2->rn_ptr.n, w->rn_ptr.n);
                std::swap(w2->rn_vec.p, w->rn_vec.p); std::swap(w2->rn_vec.n, w->rn_vec.n);
                std::swap(w2->rn_val.p, w->rn_val.p); std::swap(w2->rn_val.n, w->rn_val.n);
                std::vector<int64_t> freq2(F2, -1);
                std::vector<char> used2(F2, 0);
                int j = 0;
                for (int f = 0; f < kMaxF; ++f) map.src[f] = -1;
                for (int f = 0; f < F2; ++f) fill_slot(h2, coef2, F2, ng, f, 1.0, nullptr, 0.0);
                for (int f = 0; f < F; ++f) {
                    if (slot_freq[f] < 0) continue;
                    map.src[j] = f;
                    h2.rho[j] = h.rho[f]; h2.pq[j] = h.pq[f]; h2.beta[j] = h.beta[f]; h2.rr[j] = h.rr[f]; h2.bb[j] = h.bb[f];
                    h2.w2[j] = h.w2[f]; h2.rhs_scale[j] = h.rhs_scale[f]; h2.active[j] = h.active[f]; h2.iters[j] = h.iters[f];
                    h2.src_set[j] = h.src_set[f];
                    for (int g = 0; g < ng; ++g) coef2[(size_t)g * F2 + j] = coef[(size_t)g * F + f];
                    for (int64_t k = 0; k < nrv; ++k) rcoef2[(size_t)k * F2 + j] = rcoef[(size_t)k * F + f];
                    freq2[j] = slot_freq[f];
                    used2[j] = 1;
                    ++j;
                }
                const int rg = (int)std::min<int64_t>(grid_for(N * F2, kBlock), w->grid);
                if (w->real) {
                    k_repack<double><<<rg, kBlock, 0, s>>>((const double*)w->x.p, F, (double*)w2->x.p, F2, map, N);
                    k_repack<double><<<rg, kBlock, 0, s>>>((const double*)w->r.p, F, (double*)w2->r.p, F2, map, N);
                    k_repack<double><<<rg, kBlock, 0, s>>>((const double*)w->p.p, F, (double*)w2->p.p, F2, map, N);
                } else {
                    k_repack<double2><<<rg, kBlock, 0, s>>>(w->x.p, F, w2->x.p, F2, map, N);
                    k_repack<double2><<<rg, kBlock, 0, s>>>(w->r.p, F, w2->r.p, F2, map, N);
                    k_repack<double2><<<rg, kBlock, 0, s>>>(w->p.p, F, w2->p.p, F2, map, N);
                }
                n_kernels += 3;
                if (overlay) {
                    // the precomputed 1/diag depends on the slot's frequency: rebuild it for the moved slots
                    for (int f = 0; f < j; ++f) h2.reset[f] = 1;
                    upload_scalars(sys, w2, h2, coef2, &rcoef2);
                    k_jacobi_launch(sys, w2, F2);
                    n_kernels += 1;
                    for (int f = 0; f < j; ++f) h2.reset[f] = 0;
                }
                FDB_CUDA(cudaGetLastError());
                FDB_CUDA(cudaStreamSynchronize(s));
                free_sweep_work(w);
                sys->work = w = w2;
                F = F2;
                h = h2;
                coef.swap(coef2);
                rcoef.swap(rcoef2);
                slot_freq.swap(freq2);
                slot_used.swap(used2);
                if (n_rcv) { w->rcv_out.alloc((size_t)F * n_rcv); rcv_host.resize((size_t)F * n_rcv); }
                if (res->pN) w->stage.alloc((size_t)F * N);
                ensure_graph(sys, w, F, check_every, overlay);
            }
        }
    }
    FDB_CUDA(cudaEventRecord(ev1, s));
    FDB_CUDA(cudaEventSynchronize(ev1));
    float ms = 0;
    FDB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    res->seconds = ms * 1e-3;
    res->n_spmv = n_spmv;
    res->n_kernels = n_kernels;
    res->n_unconverged = n_unconv;
    res->n_resumed = n_resumed;
    res->precond_used = w->precond;
}

// one Chebyshev filter step fused into the SpMV (false: this mesh's tiles do not run k_spmv_blob - use the separate kernels)
// f32: x, xold and out hold 32 single-precision columns per row instead of 16 doubles (the same 128 bytes)
bool modal_spmv_cheb(fdb_system* sys, SweepWork* w, const double* x, const double* xold, double* out, const double* dinv, double c1, double c2,
                     double c3, bool f32) {
    w->cheb.d = dinv; w->cheb.xold = xold; w->cheb.c1 = c1; w->cheb.c2 = c2; w->cheb.c3 = c3; w->cheb.f32 = f32;
    const bool ok = Launch<kModalCols>::spmv_blob<false>(sys, w, x, out, false, false);
    w->cheb = ChebEpilogue();
    return ok;
}

// ---- the fused SpMV for the modal set-up (modal.cu): 16 real columns, one w2 for all of them ----
__global__ void k_set_w2(Scalars* sc, double w2) {
    if (threadIdx.x < kMaxF) sc->w2[threadIdx.x] = w2;
}
SweepWork* modal_spmv_begin(fdb_system* sys) {
    SweepWork* w = get_work(sys, kModalCols);
    bool ok = false;
    w->real = false;
    FDB_DISPATCH_F(kModalCols, query_real(sys, w, &ok));
    FDB_REQUIRE(ok, "the real-arithmetic SpMV is not available on this mesh");
    w->real = true;
    invalidate_graph(sys);
    return w;
}
void modal_spmv_w2(fdb_system* sys, SweepWork* w, double w2) { k_set_w2<<<1, kMaxF, 0, sys->stream>>>(w->sc.p, w2); }
void modal_spmv(fdb_system* sys, SweepWork* w, const double* x, double* y) { FDB_DISPATCH_F(kModalCols, spmv(sys, w, x, y, false, false)); }

// z = V diag(1 / (lam - w_f^2)) V^T r on caller data, through the solver's own kernels (parity tests of the tensor-core path)
__global__ void k_z32_to_f64(const float* __restrict__ z, int64_t n, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (double)z[i];
}
void run_modal_apply(fdb_system* sys, const double* omega, int n_use, const double* r, double* z) {
    FDB_REQUIRE(sys->have_modes, "no modes");
    SweepWork* w = modal_spmv_begin(sys);
    w->precond = 2;
    { bool okb = false; FDB_DISPATCH_F(kModalCols, bj_supported(w, &okb)); (void)okb; }
    if (sys->modal_compress) { FDB_DISPATCH_F(kModalCols, bj_prepare(sys, w)); }   // the fused kernel that forms t = Phi^T r also multiplies by B_t
    w->z32.alloc((size_t)w->N * kModalCols);
    modal_prepare(sys, w);
    std::vector<double> om(kModalCols);
    FDB_CUDA(cudaMemcpy(om.data(), omega, om.size() * sizeof(double), cudaMemcpyDefault));
    Scalars h;
    memset(&h, 0, sizeof(h));
    std::vector<double2> coef((size_t)std::max(1, w->n_groups) * kModalCols, make_double2(0.0, 0.0));
    for (int f = 0; f < kModalCols; ++f) {
        fill_slot(h, coef, kModalCols, w->n_groups, f, om[f], nullptr, 0.0);
        h.reset[f] = 1;   // tile-compressed modes: every column of r takes part in the slot-load flavour of the fused kernel
    }
    h.tol2 = 1.0;
    h.modes_use = std::min(((std::max(n_use, 1) + 15) / 16) * 16, sys->mode_chunks * 128);
    upload_scalars(sys, w, h, coef);
    const size_t n = (size_t)w->N * kModalCols;
    const int pg = (int)std::min<int64_t>(grid_for((int64_t)n, kBlock), w->grid);
    double* tmp = (double*)w->tmp.p;
    FDB_CUDA(cudaMemcpyAsync(tmp, r, n * sizeof(double), cudaMemcpyDefault, sys->stream));
    k_permute_rows<double><<<pg, kBlock, 0, sys->stream>>>(tmp, (double*)w->r.p, sys->perm.p, kModalCols, w->N, true);
    if (sys->modal_compress) modal_precond_fused(sys, w, true);   // t = Phi^T r (its block-Jacobi product is discarded)
    FDB_CUDA(cudaMemsetAsync(w->z32.p, 0, n * sizeof(float), sys->stream));
    modal_apply(sys, w, false, false);
    k_z32_to_f64<<<pg, kBlock, 0, sys->stream>>>(w->z32.p, (int64_t)n, (double*)w->q.p);
    k_permute_rows<double><<<pg, kBlock, 0, sys->stream>>>((const double*)w->q.p, tmp, sys->perm.p, kModalCols, w->N, false);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaMemcpyAsync(z, tmp, n * sizeof(double), cudaMemcpyDefault, sys->stream));
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
    w->precond = 0;
}

// one-batch constants for the stand-alone SpMV entry points
static void load_batch_scalars(fdb_system* sys, SweepWork* w, int F, const double* omega, const double* mu_batch /* [n_groups][F] complex */) {
    Scalars h;
    memset(&h, 0, sizeof(h));
    std::vector<double2> coef((size_t)std::max(1, w->n_groups) * F, make_double2(0.0, 0.0));
    std::vector<double> mu_f((size_t)std::max(1, w->n_groups) * 2);
    for (int f = 0; f < F; ++f) {
        for (int g = 0; g < w->n_groups; ++g) {
            mu_f[2 * g] = mu_batch[((size_t)g * F + f) * 2];
            mu_f[2 * g + 1] = mu_batch[((size_t)g * F + f) * 2 + 1];
        }
        fill_slot(h, coef, F, w->n_groups, f, omega[f], w->n_groups ? mu_f.data() : nullptr, 0.0);
    }
    h.tol2 = 1.0;
    upload_scalars(sys, w, h, coef);
}

static bool bench_real(fdb_system* sys, SweepWork* w, int F, int flags) {
    if (!(flags & 2)) return false;
    FDB_REQUIRE(!(flags & 1), "the real-arithmetic SpMV has no boundary term");
    bool ok = false;
    w->real = false;
    FDB_DISPATCH_F(F, query_real(sys, w, &ok));
    FDB_REQUIRE(ok, "real-arithmetic SpMV at this batch width needs tiles that fit shared memory");
    return true;
}

int64_t spmv_algorithmic_bytes(fdb_system* sys, int F, bool with_boundary, bool real) {
    // indptr + (indices + H + Q) once + overlay + x read and y write per frequency
    const int64_t N = sys->n_nodes, Z = sys->vol.nnz;
    const bool has_overlay = with_boundary && sys->have_overlay && sys->n_ov > 0 && sys->n_groups > 0;   // a rigid launch never reads the overlay
    // counted on the CSR sizes (the algorithm's bytes); the SELL-8 copy the kernel streams adds padding on top
    return 4 * (N + 1) + Z * 20 + (has_overlay ? 4 * (N + 1) + sys->n_ov * 16 : 0) + (real ? 16 : 32) * N * F;
}

void run_spmv(fdb_system* sys, int F, const double* omega, const double* mu, const double* x, double* y) {
    SweepWork* w = get_work(sys, F);
    const int ng = sys->n_groups;
    const bool has_overlay = sys->have_overlay && sys->n_ov > 0 && ng > 0;
    std::vector<double> om(F), m((size_t)std::max(1, ng) * F * 2, 0.0);
    FDB_CUDA(cudaMemcpy(om.data(), omega, F * sizeof(double), cudaMemcpyDefault));
    if (ng) {
        FDB_REQUIRE(mu != nullptr, "mu is required when the system has boundary groups");
        FDB_CUDA(cudaMemcpy(m.data(), mu, (size_t)ng * F * 2 * sizeof(double), cudaMemcpyDefault));
    }
    w->real = false;
    load_batch_scalars(sys, w, F, om.data(), m.data());
    const size_t n = (size_t)w->N * F;
    const int pg = (int)std::min<int64_t>(grid_for((int64_t)n, kBlock), w->grid);
    FDB_CUDA(cudaMemcpyAsync(w->tmp.p, x, n * sizeof(double2), cudaMemcpyDefault, sys->stream));
    k_permute_rows<double2><<<pg, kBlock, 0, sys->stream>>>(w->tmp.p, w->p.p, sys->perm.p, F, w->N, true);
    FDB_DISPATCH_F(F, spmv(sys, w, w->p.p, w->q.p, false, has_overlay));
    k_permute_rows<double2><<<pg, kBlock, 0, sys->stream>>>(w->q.p, w->tmp.p, sys->perm.p, F, w->N, false);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaMemcpyAsync(y, w->tmp.p, n * sizeof(double2), cudaMemcpyDefault, sys->stream));
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
}

__global__ void k_fill_pattern(double2* __restrict__ v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double t = (double)(i % 1024) * (1.0 / 1024.0);
        v[i] = make_double2(0.5 + t, 0.25 - t);
    }
}

void run_spmv_enqueue(fdb_system* sys, int F, int reps, int flags, int64_t* bytes) {
    SweepWork* w = get_work(sys, F);
    const bool with_boundary = (flags & 1) != 0;
    w->real = bench_real(sys, w, F, flags);
    const bool has_overlay = with_boundary && sys->have_overlay && sys->n_ov > 0 && sys->n_groups > 0;
    for (int i = 0; i < reps; ++i) { FDB_DISPATCH_F(F, spmv(sys, w, w->p.p, w->q.p, true, has_overlay)); }
    FDB_CUDA(cudaGetLastError());
    if (bytes) *bytes = spmv_algorithmic_bytes(sys, F, with_boundary, w->real);
}

// One kernel of a solver iteration, `reps` times on resident synthetic vectors, enqueued on the system's stream without
// synchronising so the caller brackets it with its own CUDA events.  which: -1 = prepare (fill the vectors, every slot
// active), 0 = K5 (SpMV with its dot-product epilogue), 1 = K6a (k_update / k_update_bj), 2 = K6b (k_pupdate / k_pupdate_bj).
// flags: FDB_SPMV_REAL, FDB_ITER_BLOCK_JACOBI.  bytes = the kernel's algorithmic bytes per launch.
void run_iteration_enqueue(fdb_system* sys, int F, int reps, int flags, int which, int64_t* bytes) {
    FDB_REQUIRE(!(flags & 1), "the iteration bench covers rigid-wall sweeps");
    SweepWork* w = get_work(sys, F);
    w->real = bench_real(sys, w, F, flags);
    const bool modal = (flags & 8) != 0;
    if (modal) FDB_REQUIRE(sys->have_modes && w->real && F == kModalCols, "the modal iteration bench needs the modes and 16 real slots");
    const bool bj = ((flags & 4) || modal) && w->real && F <= 16 && sys->n_tiles > 0 && sys->tile_rows == 128;   // the bench covers real sweeps
    w->precond = modal ? 2 : (bj ? 1 : 0);
    if (bj) { FDB_DISPATCH_F(F, bj_prepare(sys, w)); }
    if (modal) modal_prepare(sys, w);
    const int modes_all = modal ? std::min(((sys->n_modes + 15) / 16) * 16, sys->mode_chunks * 128) : 0;
    const int64_t N = w->N, ne = N * F;
    const int64_t vb = (w->real ? 8 : 16) * ne;   // bytes of one vector pass
    if (which < 0) {
        Scalars h;
        memset(&h, 0, sizeof(h));
        std::vector<double2> coef((size_t)std::max(1, w->n_groups) * F, make_double2(0.0, 0.0));
        for (int f = 0; f < F; ++f) {
            fill_slot(h, coef, F, w->n_groups, f, 2.0 * 3.141592653589793 * (100.0 + f), nullptr, 0.0);
            h.active[f] = 1; h.xpend[f] = 1;
            h.rho[f] = make_double2(1.0, 0.0); h.pq[f] = make_double2(1e30, 0.0);   // alpha ~ 0: repeated launches stay finite
            h.alpha[f] = make_double2(1e-6, 0.0); h.beta[f] = make_double2(0.5, 0.0);
            h.bb[f] = 0.0;   // never "converged": rr > tol2 * 0
        }
        h.tol2 = 1.0;
        h.modes_use = modes_all;
        upload_scalars(sys, w, h, coef);
        const int64_t n2 = w->real ? (ne + 1) / 2 : ne;   // buffers are sized in double2
        k_fill_pattern<<<kNumSMs * 4, 256, 0, sys->stream>>>(w->p.p, n2);
        k_fill_pattern<<<kNumSMs * 4, 256, 0, sys->stream>>>(w->r.p, n2);
        k_fill_pattern<<<kNumSMs * 4, 256, 0, sys->stream>>>(w->x.p, n2);
        k_fill_pattern<<<kNumSMs * 4, 256, 0, sys->stream>>>(w->q.p, n2);
        if (bj) FDB_CUDA(cudaMemsetAsync(w->z32.p, 0, (size_t)ne * sizeof(float), sys->stream));
        FDB_CUDA(cudaGetLastError());
        return;
    }
    for (int i = 0; i < reps; ++i) {
        if (which == 0) { FDB_DISPATCH_F(F, spmv(sys, w, w->p.p, w->q.p, true, false)); }
        else if (which == 1 && modal && w->fused_precond) modal_precond_fused(sys, w, false);
        else if (which == 3) modal_apply(sys, w, false, false, 1);
        else if (which == 4) modal_apply(sys, w, false, false, 4);
        else if (which == 5) modal_apply(sys, w, false, false, 2);
        else if (which == 6) modal_apply(sys, w, false, false, 8);
        else { FDB_DISPATCH_F(F, vec_kernel(sys, w, false, which)); }
    }
    FDB_CUDA(cudaGetLastError());
    if (bytes) {
        if (which == 0) *bytes = spmv_algorithmic_bytes(sys, F, false, w->real);
        else if (which == 1 && modal && w->fused_precond) *bytes = modal_fused_bytes(sys, w, modes_all);
        else if (which == 3) *bytes = modal_bytes(sys, w, modes_all, 1);
        else if (which == 4) *bytes = modal_bytes(sys, w, modes_all, 4);
        else if (which == 5) *bytes = modal_bytes(sys, w, modes_all, 2);
        else if (which == 6) *bytes = sys->modal_compress ? modal_bytes(sys, w, modes_all, 8) : 0;
        else if (which == 1) *bytes = bj ? 3 * vb + 4 * ne + sys->n_tiles * (int64_t)kBjPacked * 4 : 3 * vb + 16 * N;
        else *bytes = bj ? 4 * vb + 4 * ne : 5 * vb + 16 * N;
    }
}

void run_spmv_bench(fdb_system* sys, int F, int reps, int flags, double* ms_per_launch, int64_t* bytes) {
    SweepWork* w = get_work(sys, F);
    const bool with_boundary = (flags & 1) != 0;
    w->real = bench_real(sys, w, F, flags);
    const bool has_overlay = with_boundary && sys->have_overlay && sys->n_ov > 0 && sys->n_groups > 0;
    std::vector<double> om(F), m((size_t)std::max(1, sys->n_groups) * F * 2, 0.0);
    for (int f = 0; f < F; ++f) om[f] = 2.0 * 3.141592653589793 * (100.0 + f);
    for (size_t i = 0; i < m.size(); i += 2) m[i] = 1e-4;
    load_batch_scalars(sys, w, F, om.data(), m.data());
    const int64_t n = w->N * F;
    k_fill_pattern<<<kNumSMs * 4, 256, 0, sys->stream>>>(w->p.p, n);
    for (int i = 0; i < 3; ++i) { FDB_DISPATCH_F(F, spmv(sys, w, w->p.p, w->q.p, true, has_overlay)); }
    cudaEvent_t e0, e1;
    FDB_CUDA(cudaEventCreate(&e0));
    FDB_CUDA(cudaEventCreate(&e1));
    FDB_CUDA(cudaEventRecord(e0, sys->stream));
    for (int i = 0; i < reps; ++i) { FDB_DISPATCH_F(F, spmv(sys, w, w->p.p, w->q.p, true, has_overlay)); }
    FDB_CUDA(cudaEventRecord(e1, sys->stream));
    FDB_CUDA(cudaEventSynchronize(e1));
    FDB_CUDA(cudaGetLastError());
    float ms = 0;
    FDB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_per_launch) *ms_per_launch = ms / reps;
    if (bytes) *bytes = spmv_algorithmic_bytes(sys, F, with_boundary, w->real);
}

}  // namespace fdb

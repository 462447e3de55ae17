// Internal node order of the sweep ("blob order") and the tile tables of the shared-memory SpMV (k_spmv_blob).
//
// The reference numbers nodes as gmsh emits them and hands G to a direct solver (femder/FEM_3D.py:1307-1319); nothing
// downstream depends on the order the Krylov vectors are stored in.  The sweep therefore works in its own order:
//   * nodes are grouped into compact blobs of 128 rows grown breadth-first over the matrix graph, so the columns a
//     tile of 128 rows touches are its own rows plus a thin halo (about 2.5 x 128 rows on a tetrahedral mesh);
//   * inside a tile the rows are sorted by descending length, so the 8-row SELL slices are homogeneous (no padding)
//   * per tile the sorted list of distinct columns is stored once; the matrix stream carries 16-bit positions in it.
// K5 then stages the tile's x rows in shared memory once and serves the ~15 gathers per row from there instead of
// from L1.  Everything the C ABI exchanges stays in the caller's numbering: run_sweep maps
// node lists through perm, the pN export gathers through it.
#include "common.cuh"
#include <algorithm>
#include <thread>

namespace fdb {

constexpr int kMaxTileRows = 128;   // rows per tile: 128, or 64 (two resident blocks per SM); fdb_system::tile_rows
constexpr int kTileHdr = 20;    // ints of a tile's header: 17 relative slice offsets, position of the first own row, column count

// Breadth-first blobs: a blob starts at the oldest still-free node on the boundary of what is already ordered and
// grows until the current tile is full; what its queue still holds becomes boundary for later blobs.
static void blob_order(int64_t n, const int32_t* indptr, const int32_t* indices, int kTileRows, std::vector<int32_t>& order /* new -> old */) {
    order.resize(n);
    std::vector<int32_t> mark(n, -1);
    std::vector<uint8_t> done(n, 0);
    std::vector<int32_t> frontier;
    frontier.reserve(n);
    size_t fpos = 0;
    int64_t next_seed = 0, cnt = 0;
    int32_t blob = 0;
    std::vector<int32_t> q;
    q.reserve(8 * kMaxTileRows);
    while (cnt < n) {
        int32_t seed = -1;
        while (fpos < frontier.size()) {
            const int32_t c = frontier[fpos++];
            if (!done[c]) { seed = c; break; }
        }
        if (seed < 0) {
            while (done[next_seed]) ++next_seed;
            seed = (int32_t)next_seed;
        }
        q.clear();
        q.push_back(seed);
        mark[seed] = blob;
        size_t qh = 0;
        int got = 0;
        const int want = kTileRows - (int)(cnt % kTileRows);
        while (got < want && qh < q.size()) {
            const int32_t u = q[qh++];
            order[cnt++] = u;
            done[u] = 1;
            ++got;
            for (int32_t k = indptr[u]; k < indptr[u + 1]; ++k) {
                const int32_t v = indices[k];
                if (!done[v] && mark[v] != blob) { mark[v] = blob; q.push_back(v); }
            }
        }
        for (; qh < q.size(); ++qh) frontier.push_back(q[qh]);
        ++blob;
    }
}

template <typename Fn>
static void parallel_for(int64_t n, Fn fn) {
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::thread::hardware_concurrency(), (int64_t)16, n / 64 + 1}));
    if (nt == 1) { fn(0, n, 0); return; }
    std::vector<std::thread> th;
    const int64_t per = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        const int64_t a = t * per, b = std::min(n, a + per);
        if (a < b) th.emplace_back([=, &fn] { fn(a, b, t); });
    }
    for (auto& x : th) x.join();
}

void build_order(fdb_system* sys) {
    const int64_t N = sys->n_nodes, Z = sys->vol.nnz;
    cudaStream_t s = sys->stream;
    std::vector<int32_t> indptr(N + 1), indices(std::max<int64_t>(Z, 1));
    FDB_CUDA(cudaMemcpyAsync(indptr.data(), sys->vol.indptr.p, (N + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (Z) FDB_CUDA(cudaMemcpyAsync(indices.data(), sys->vol.indices.p, Z * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    FDB_CUDA(cudaStreamSynchronize(s));

    int kTileRows = kMaxTileRows;
#ifdef FDB_TUNING
    if (const char* e = getenv("FDB_TILE_ROWS")) kTileRows = (atoi(e) == 64) ? 64 : 128;   // tuning aid (tuning builds only)
#endif
    sys->tile_rows = kTileRows;
    std::vector<int32_t> order;
    blob_order(N, indptr.data(), indices.data(), kTileRows, order);
    const int64_t n_tiles = (N + kTileRows - 1) / kTileRows;
    // rows of a tile by descending length (stable: ties keep the breadth-first order)
    parallel_for(n_tiles, [&](int64_t a, int64_t b, int) {
        for (int64_t t = a; t < b; ++t) {
            int32_t* o0 = order.data() + t * kTileRows;
            int32_t* o1 = order.data() + std::min<int64_t>(N, (t + 1) * kTileRows);
            std::stable_sort(o0, o1, [&](int32_t x, int32_t y) { return indptr[x + 1] - indptr[x] > indptr[y + 1] - indptr[y]; });
        }
    });
    std::vector<int32_t>& perm = sys->perm_host;
    perm.resize(N);
    for (int64_t i = 0; i < N; ++i) perm[order[i]] = (int32_t)i;

    // SELL-8 structure in the new order
    const int64_t S = (N + 7) / 8;
    std::vector<int32_t>& off = sys->sell_off_host;
    off.assign(S + 1, 0);
    for (int64_t sl = 0; sl < S; ++sl) {
        int w = 0;
        for (int64_t r = sl * 8; r < std::min<int64_t>(N, sl * 8 + 8); ++r) w = std::max(w, indptr[order[r] + 1] - indptr[order[r]]);
        off[sl + 1] = off[sl] + w;
        FDB_REQUIRE((int64_t)off[sl + 1] * 8 < 2147483647LL - 64, "SELL layout exceeds int32 indexing");
    }
    const int64_t n_ent = (int64_t)off[S] * 8;
    std::vector<int32_t> sidx(std::max<int64_t>(n_ent, 1)), ssrc(std::max<int64_t>(n_ent, 1));
    std::vector<uint16_t> lidx(std::max<int64_t>(n_ent, 1));
    std::vector<int32_t> cptr(n_tiles + 1, 0);
    std::vector<std::vector<int32_t>> tcols(n_tiles);
    std::vector<int> t_entries(n_tiles, 0);
    std::vector<int32_t> tself(n_tiles, 0);
    parallel_for(n_tiles, [&](int64_t a, int64_t b, int) {
        std::vector<int32_t> cols;
        for (int64_t t = a; t < b; ++t) {
            const int64_t r0 = t * kTileRows, r1 = std::min<int64_t>(N, r0 + kTileRows);
            cols.clear();
            for (int64_t r = r0; r < r1; ++r) {
                cols.push_back((int32_t)r);   // own rows are always staged (dot-product epilogue, padding entries)
                const int32_t o = order[r];
                for (int32_t k = indptr[o]; k < indptr[o + 1]; ++k) cols.push_back(perm[indices[k]]);
            }
            std::sort(cols.begin(), cols.end());
            cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
            tself[t] = (int32_t)(std::lower_bound(cols.begin(), cols.end(), (int32_t)r0) - cols.begin());
            for (int64_t r = r0; r < r1; ++r) {
                const int64_t sl = r >> 3;
                const int32_t o = order[r];
                const int len = indptr[o + 1] - indptr[o], width = off[sl + 1] - off[sl];
                const uint16_t self = (uint16_t)(std::lower_bound(cols.begin(), cols.end(), (int32_t)r) - cols.begin());
                for (int k = 0; k < width; ++k) {
                    const int64_t e = ((int64_t)off[sl] + k) * 8 + (r & 7);
                    if (k < len) {
                        const int32_t c = perm[indices[indptr[o] + k]];
                        sidx[e] = c;
                        ssrc[e] = indptr[o] + k;
                        lidx[e] = (uint16_t)(std::lower_bound(cols.begin(), cols.end(), c) - cols.begin());
                    } else {
                        sidx[e] = (int32_t)r;
                        ssrc[e] = -1;
                        lidx[e] = self;
                    }
                }
            }
            // rows past N in the last slice: padding that points at the last row
            if (r1 == N && (N & 7)) {
                const int64_t sl = (N - 1) >> 3;
                const int width = off[sl + 1] - off[sl];
                const uint16_t self = (uint16_t)(std::lower_bound(cols.begin(), cols.end(), (int32_t)(N - 1)) - cols.begin());
                for (int64_t r = N; r < sl * 8 + 8; ++r)
                    for (int k = 0; k < width; ++k) {
                        const int64_t e = ((int64_t)off[sl] + k) * 8 + (r & 7);
                        sidx[e] = (int32_t)(N - 1);
                        ssrc[e] = -1;
                        lidx[e] = self;
                    }
            }
            const int64_t s0 = t * (kTileRows / 8), s1 = std::min<int64_t>(S, s0 + kTileRows / 8);
            t_entries[t] = (off[s1] - off[s0]) * 8;
            tcols[t] = cols;
        }
    });
    // column lists back to back, each padded to a multiple of 4 entries (16 bytes: the kernel fetches a list with one
    // bulk copy); tile_meta[t] = (start of the list, count | position of the first own row << 16, first SELL slice
    // offset, SELL entries of the tile)
    int max_cols = 0, max_ent = 8;
    std::vector<int4> meta(n_tiles);
    for (int64_t t = 0; t < n_tiles; ++t) {
        const int c = (int)tcols[t].size();
        FDB_REQUIRE(c < 65536 && t_entries[t] < (1 << 30), "tile too wide for the 16-bit column positions");
        cptr[t + 1] = cptr[t] + ((c + 3) & ~3);
        max_cols = std::max<int>(max_cols, c);
        max_ent = std::max(max_ent, t_entries[t]);
        meta[t] = make_int4(cptr[t], c | (tself[t] << 16), off[t * (kTileRows / 8)], t_entries[t]);
    }
    std::vector<int32_t> allcols(std::max<int64_t>(cptr[n_tiles], 4));
    std::vector<int32_t> hdr((size_t)n_tiles * kTileHdr);
    parallel_for(n_tiles, [&](int64_t a, int64_t b, int) {
        for (int64_t t = a; t < b; ++t) {
            std::copy(tcols[t].begin(), tcols[t].end(), allcols.begin() + cptr[t]);
            for (int32_t i = cptr[t] + (int32_t)tcols[t].size(); i < cptr[t + 1]; ++i) allcols[i] = tcols[t].back();
            // header the consumer warps read: the tile's 17 SELL slice offsets relative to its first slice, ...
            const int64_t s0 = t * (kTileRows / 8);
            int32_t* hd = hdr.data() + t * kTileHdr;
            for (int i = 0; i <= kTileRows / 8; ++i) hd[i] = off[std::min<int64_t>(S, s0 + i)] - off[s0];
            hd[17] = tself[t];
            hd[18] = (int32_t)tcols[t].size();
            hd[19] = 0;
        }
    });

    sys->n_slices = S;
    sys->sell_entries = n_ent;
    sys->n_tiles = n_tiles;
    sys->max_tile_cols = max_cols;
    sys->max_tile_entries = max_ent;
    sys->perm.alloc(N);
    sys->iperm.alloc(N);
    sys->sell_off.alloc(S + 1);
    sys->sell_idx.alloc(sidx.size());
    sys->sell_src.alloc(ssrc.size());
    sys->sell_lidx.alloc(lidx.size());
    sys->tile_meta.alloc(n_tiles);
    sys->tile_hdr.alloc(hdr.size());
    FDB_CUDA(cudaMemcpyAsync(sys->tile_hdr.p, hdr.data(), hdr.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    sys->tile_cols.alloc(allcols.size());
    FDB_CUDA(cudaMemcpyAsync(sys->perm.p, perm.data(), N * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    FDB_CUDA(cudaMemcpyAsync(sys->iperm.p, order.data(), N * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    FDB_CUDA(cudaMemcpyAsync(sys->sell_off.p, off.data(), (S + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    FDB_CUDA(cudaMemcpyAsync(sys->sell_idx.p, sidx.data(), sidx.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    FDB_CUDA(cudaMemcpyAsync(sys->sell_src.p, ssrc.data(), ssrc.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    FDB_CUDA(cudaMemcpyAsync(sys->sell_lidx.p, lidx.data(), lidx.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
    FDB_CUDA(cudaMemcpyAsync(sys->tile_meta.p, meta.data(), n_tiles * sizeof(int4), cudaMemcpyHostToDevice, s));
    FDB_CUDA(cudaMemcpyAsync(sys->tile_cols.p, allcols.data(), allcols.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    FDB_CUDA(cudaStreamSynchronize(s));   // the host vectors die here
    sys->have_order = true;
}

// values of the SELL stream from the CSR values (every re-assembly; the structure above is per pattern)
__global__ void k_sell_values(const int32_t* __restrict__ src, const double2* __restrict__ hq, int64_t n, double2* __restrict__ sell_hq) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int32_t z = src[e];
    sell_hq[e] = (z >= 0) ? hq[z] : make_double2(0.0, 0.0);
}

void build_sell(fdb_system* sys) {
    FDB_REQUIRE(sys->have_order, "internal order missing");
    const int64_t n = sys->sell_entries;
    sys->have_blk = false;   // the block-Jacobi inverses belong to the previous values
    modal_reset(sys);        // ... and so do the modes
    sys->sell_hq.alloc(n ? n : 1);
    if (n) k_sell_values<<<grid_for(n, 256), 256, 0, sys->stream>>>(sys->sell_src.p, sys->hq.p, n, sys->sell_hq.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
}

// ------------------------------------------------------------------------------------------------
// Block-Jacobi preconditioner of the sweep: B_t = ((H + shift Q) restricted to the 128 rows of tile t)^-1, one dense symmetric block
// per tile of the blob order.  The diagonal blocks of H are principal submatrices of a positive semi-definite matrix
// whose only null vector is the constant, hence positive definite: Gauss-Jordan without pivoting, FP64 in shared
// memory, result symmetrised and stored in FP32 as a packed lower triangle (a preconditioner needs no more: a quarter
// of the bytes of a full FP64 block).
// Frequency independent: at the mesh sizes a sweep resolves (k h <= 0.3) the -w^2 Q part of a 128-node block is a small
// perturbation of its stiffness (tests/research/block_jacobi_prototype.py: same iteration counts as the exact
// (H - w^2 Q) block, about half of point Jacobi's).
// ------------------------------------------------------------------------------------------------
constexpr int kBlk = 128;
constexpr int kBlkPacked = kBlk * (kBlk + 1) / 2;   // 8256 floats = 33 024 bytes per tile
__global__ void __launch_bounds__(256) k_block_inverse(const int32_t* __restrict__ sell_off, const int32_t* __restrict__ sell_idx,
                                                        const double2* __restrict__ sell_hq, int64_t n_rows, double shift, float* __restrict__ out,
                                                        uint16_t* __restrict__ out16) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* A = reinterpret_cast<double*>(smem_raw);          // [128][128]
    double* colk = A + kBlk * kBlk;                            // [128]
    double* rowk = colk + kBlk;                                // [128]
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * kBlk;
    const int rows = (int)min((int64_t)kBlk, n_rows - row0);
    for (int i = tid; i < kBlk * kBlk; i += 256) A[i] = 0.0;
    __syncthreads();
    if (tid < kBlk) {
        if (tid < rows) {
            const int64_t r = row0 + tid;
            const int64_t slice = r >> 3;
            const int off = sell_off[slice];
            const int width = sell_off[slice + 1] - off;
            for (int k = 0; k < width; ++k) {
                const int64_t e = ((int64_t)off + k) * 8 + (r & 7);
                const int64_t c = sell_idx[e];
                if (c >= row0 && c < row0 + kBlk) {
                    const double2 m = sell_hq[e];   // padding entries add 0
                    A[tid * kBlk + (int)(c - row0)] += m.x + shift * m.y;
                }
            }
        } else {
            A[tid * kBlk + tid] = 1.0;   // rows past the end of the last tile
        }
    }
    __syncthreads();
    for (int k = 0; k < kBlk; ++k) {
        if (tid < kBlk) { colk[tid] = A[tid * kBlk + k]; rowk[tid] = A[k * kBlk + tid]; }
        __syncthreads();
        const double pv = colk[k];
        const double pinv = (pv != 0.0) ? 1.0 / pv : 1.0;
        for (int idx = tid; idx < kBlk * kBlk; idx += 256) {
            const int i = idx >> 7, j = idx & (kBlk - 1);
            double v;
            if (i == k) v = (j == k) ? pinv : rowk[j] * pinv;
            else if (j == k) v = -colk[i] * pinv;
            else v = A[idx] - colk[i] * (rowk[j] * pinv);
            A[idx] = v;
        }
        __syncthreads();
    }
    // lower triangle, row-major: entry (i, k <= i) at i (i + 1) / 2 + k - the product kernel reads it conflict-free both
    // along rows and along columns (triangular numbers mod 32 are a permutation of 0..31)
    float* o = out + (int64_t)blockIdx.x * kBlkPacked;
    for (int idx = tid; idx < kBlk * kBlk; idx += 256) {
        const int i = idx >> 7, j = idx & (kBlk - 1);
        if (j <= i) o[i * (i + 1) / 2 + j] = (float)(0.5 * (A[i * kBlk + j] + A[j * kBlk + i]));
    }
    // the full (symmetric) block in bf16 for the tensor-core product of the modal path: 8 x 8 core matrices, entry (i, k) at byte
    // (k / 8) * 2048 + (i / 8) * 128 + (i % 8) * 16 + (k % 8) * 2 - the layout of a tile of modes, rows = nodes, K = columns
    uint16_t* o16 = out16 + (int64_t)blockIdx.x * (kBlk * kBlk);
    for (int idx = tid; idx < kBlk * kBlk; idx += 256) {
        const int i = idx >> 7, k = idx & (kBlk - 1);
        const float v = (float)(0.5 * (A[i * kBlk + k] + A[k * kBlk + i]));
        uint32_t u = __float_as_uint(v);
        u += 0x7fffu + ((u >> 16) & 1u);
        o16[((k >> 3) * 2048 + (i >> 3) * 128 + (i & 7) * 16 + (k & 7) * 2) >> 1] = (uint16_t)(u >> 16);
    }
}

void build_block_prec(fdb_system* sys) {
    if (sys->have_blk) return;
    FDB_REQUIRE(sys->have_order && sys->tile_rows == kBlk && sys->n_tiles > 0, "block-Jacobi needs the 128-row tiles of the internal order");
    sys->blk_inv.alloc((size_t)sys->n_tiles * kBlkPacked);
    sys->blk_inv16.alloc((size_t)sys->n_tiles * kBlk * kBlk);
    const size_t smem = (size_t)(kBlk * kBlk + 2 * kBlk) * sizeof(double);
    FDB_CUDA(cudaFuncSetAttribute(k_block_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // H + shift Q with a small positive shift (w = 2 pi 50 Hz, far below a block's own eigenvalues): a tile that holds a whole
    // connected component (meshes of up to 128 nodes) has a singular H block, and H + shift Q is positive definite always
    const double shift = (2.0 * 3.14159265358979323846 * 50.0) * (2.0 * 3.14159265358979323846 * 50.0);
    k_block_inverse<<<(unsigned)sys->n_tiles, 256, smem, sys->stream>>>(sys->sell_off.p, sys->sell_idx.p, sys->sell_hq.p, sys->n_nodes,
                                                                         shift, sys->blk_inv.p, sys->blk_inv16.p);
    FDB_CUDA(cudaGetLastError());
    sys->have_blk = true;
}

}  // namespace fdb

// K4 - CSR pattern, element->nnz map and nnz->contribution map, built on the GPU once per connectivity.
//
// Replaces the COO->CSC conversion scipy performs inside csc_matrix((data,(i,j))) at
// femder/fem_numerical.py:337-343 and coo_matrix(...).tocsc() at femder/FEM_3D.py:413-416, 519-522.
// The result must be bit-identical to scipy's canonical pattern (sorted column indices, no duplicates,
// int32).  The pattern of an FEM operator is symmetric, so this CSR is also the reference's CSC.
//
// Everything is deterministic: atomics are only used for counting and for a fill whose result is sorted
// afterwards, so the arrays do not depend on thread scheduling.
#include "common.cuh"

namespace fdb {

// ------------------------------------------------------------------------------------------------
// exclusive scan (int32 in, int32 out[n+1], int64 total): block sums -> single-block scan -> apply
// ------------------------------------------------------------------------------------------------
constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long* total, long long* smem /*[33]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : 0;
        long long wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        smem[lane] = wi - w;  // exclusive warp offsets
        if (lane == 31) smem[32] = wi;
    }
    __syncthreads();
    long long excl = smem[warp] + incl - v;
    *total = smem[32];
    __syncthreads();
    return excl;
}

__global__ void __launch_bounds__(kScanBlock) k_scan_block_sums(const int32_t* __restrict__ in, int64_t n,
                                                                  long long* __restrict__ block_sums) {
    __shared__ long long smem[33];
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    long long s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) s += in[base + i];
    long long total;
    block_exclusive_scan(s, &total, smem);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_block_offsets(long long* __restrict__ block_sums, int nb,
                                                               long long* __restrict__ total_out) {
    __shared__ long long smem[33];
    long long carry = 0;
    for (int base = 0; base < nb; base += blockDim.x) {
        int i = base + threadIdx.x;
        long long v = (i < nb) ? block_sums[i] : 0;
        long long total;
        long long excl = block_exclusive_scan(v, &total, smem);
        if (i < nb) block_sums[i] = carry + excl;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(kScanBlock) k_scan_apply(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                             int64_t n, const long long* __restrict__ block_off,
                                                             const long long* __restrict__ total) {
    __shared__ long long smem[33];
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int32_t v[kScanItems];
    long long s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    long long tot;
    long long excl = block_exclusive_scan(s, &tot, smem) + block_off[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = (int32_t)excl;
        excl += v[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (int32_t)(*total);
}

void exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int64_t* total_host, cudaStream_t s) {
    int nb = (int)((n + kScanTile - 1) / kScanTile);
    if (nb == 0) nb = 1;
    DevBuf<long long> sums;
    sums.alloc((size_t)nb + 1);
    k_scan_block_sums<<<nb, kScanBlock, 0, s>>>(in, n, sums.p);
    k_scan_block_offsets<<<1, 1024, 0, s>>>(sums.p, nb, sums.p + nb);
    k_scan_apply<<<nb, kScanBlock, 0, s>>>(in, out, n, sums.p, sums.p + nb);
    FDB_CUDA(cudaGetLastError());
    long long tot = 0;
    FDB_CUDA(cudaMemcpyAsync(&tot, sums.p + nb, sizeof(long long), cudaMemcpyDeviceToHost, s));
    FDB_CUDA(cudaStreamSynchronize(s));
    FDB_REQUIRE(tot >= 0 && tot < 2147483647LL, "size exceeds the int32 index range scipy uses for this pattern");
    if (total_host) *total_host = tot;
}

// ------------------------------------------------------------------------------------------------
// node -> incident elements (sorted ascending per node)
// ------------------------------------------------------------------------------------------------
__global__ void k_check_conn(const int32_t* __restrict__ conn, int64_t n, int64_t n_rows, int* __restrict__ bad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int32_t c = conn[i];
        if (c < 0 || c >= n_rows) atomicExch(bad, 1);
    }
}

__global__ void k_count_deg(const int32_t* __restrict__ conn, int64_t n, int32_t* __restrict__ deg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&deg[conn[i]], 1);
}

__global__ void k_fill_n2e(const int32_t* __restrict__ conn, int64_t n, int nper, const int32_t* __restrict__ n2e_ptr,
                           int32_t* __restrict__ cursor, int32_t* __restrict__ n2e) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int32_t node = conn[i];
        int32_t pos = atomicAdd(&cursor[node], 1);
        n2e[n2e_ptr[node] + pos] = (int32_t)(i / nper);
    }
}

__global__ void k_sort_segments(const int32_t* __restrict__ ptr, int32_t* __restrict__ data, int64_t n_rows) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int32_t b = ptr[r], e = ptr[r + 1];
    for (int32_t i = b + 1; i < e; ++i) {  // insertion sort: segments are short (elements around one node)
        int32_t v = data[i];
        int32_t j = i - 1;
        while (j >= b && data[j] > v) {
            data[j + 1] = data[j];
            --j;
        }
        data[j + 1] = v;
    }
}

// one thread per row: sorted, duplicate-free column list in the row's scratch segment
__global__ void k_row_unique(const int32_t* __restrict__ conn, int nper, const int32_t* __restrict__ n2e_ptr,
                             const int32_t* __restrict__ n2e, int64_t n_rows, int32_t* __restrict__ scratch,
                             int32_t* __restrict__ rowlen) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int32_t b = n2e_ptr[r], e = n2e_ptr[r + 1];
    int32_t* list = scratch + (int64_t)b * nper;
    int len = 0;
    for (int32_t k = b; k < e; ++k) {
        const int32_t* c = conn + (int64_t)n2e[k] * nper;
        for (int a = 0; a < nper; ++a) {
            int32_t col = c[a];
            int lo = 0, hi = len;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (list[mid] < col) lo = mid + 1; else hi = mid;
            }
            if (lo < len && list[lo] == col) continue;
            for (int j = len; j > lo; --j) list[j] = list[j - 1];
            list[lo] = col;
            ++len;
        }
    }
    rowlen[r] = len;
}

__global__ void k_copy_cols(const int32_t* __restrict__ n2e_ptr, int nper, const int32_t* __restrict__ scratch,
                            const int32_t* __restrict__ indptr, int32_t* __restrict__ indices, int64_t n_rows) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int32_t* list = scratch + (int64_t)n2e_ptr[r] * nper;
    int32_t b = indptr[r], e = indptr[r + 1];
    for (int32_t k = b; k < e; ++k) indices[k] = list[k - b];
}

// one thread per nnz (row r, col c): walk the elements around r in ascending order and record every
// (e, a, b) with conn[e][a] == r, conn[e][b] == c.  PASS 0 counts, PASS 1 fills c_src and elem2nnz.
template <int PASS>
__global__ void k_contrib(const int32_t* __restrict__ conn, int nper, const int32_t* __restrict__ n2e_ptr,
                          const int32_t* __restrict__ n2e, const int32_t* __restrict__ indptr,
                          const int32_t* __restrict__ indices, int64_t n_rows, int32_t* __restrict__ cnt,
                          const int32_t* __restrict__ c_ptr, int32_t* __restrict__ c_src,
                          int32_t* __restrict__ elem2nnz) {
    // a warp walks rows; lanes take the row's nnz
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= n_rows) return;
    int32_t r = (int32_t)warp;
    int32_t zb = indptr[r], ze = indptr[r + 1];
    int32_t eb = n2e_ptr[r], ee = n2e_ptr[r + 1];
    const int n2 = nper * nper;
    for (int32_t z = zb + lane; z < ze; z += 32) {
        int32_t c = indices[z];
        int32_t n = 0;
        int32_t out = (PASS == 1) ? c_ptr[z] : 0;
        for (int32_t k = eb; k < ee; ++k) {
            int32_t e = n2e[k];
            const int32_t* cn = conn + (int64_t)e * nper;
            int a = -1, b = -1;
            for (int i = 0; i < nper; ++i) {
                int32_t v = cn[i];
                if (v == r && a < 0) a = i;
                if (v == c && b < 0) b = i;
            }
            if (b >= 0) {
                if (PASS == 1) {
                    c_src[out + n] = e * n2 + a * nper + b;
                    elem2nnz[(int64_t)e * n2 + a * nper + b] = z;
                }
                ++n;
            }
        }
        if (PASS == 0) cnt[z] = n;
    }
}

void build_pattern(Pattern& pat, int64_t n_rows, int64_t n_elem, int nper, const int32_t* conn_dev, cudaStream_t s) {
    FDB_REQUIRE(n_rows > 0 && n_rows < 2147483647LL, "n_nodes out of range");
    FDB_REQUIRE(n_elem >= 0 && n_elem * nper * nper < 2147483647LL, "n_elem * nper^2 exceeds int32");
    pat.n_rows = n_rows;
    pat.n_elem = n_elem;
    pat.nper = nper;
    const int64_t nc = n_elem * nper;
    const int B = 256;

    DevBuf<int> bad;
    bad.alloc(1);
    bad.zero(s);
    if (nc) k_check_conn<<<grid_for(nc, B), B, 0, s>>>(conn_dev, nc, n_rows, bad.p);
    int bad_h = 0;
    FDB_CUDA(cudaMemcpyAsync(&bad_h, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    FDB_CUDA(cudaStreamSynchronize(s));
    FDB_REQUIRE(!bad_h, "connectivity holds a node index outside [0, n_nodes)");

    DevBuf<int32_t> deg, cursor, scratch, rowlen;
    DevBuf<int32_t>& n2e_ptr = pat.n2e_ptr;   // kept: receiver location walks the elements around a node
    DevBuf<int32_t>& n2e = pat.n2e;
    deg.alloc(n_rows);
    deg.zero(s);
    if (nc) k_count_deg<<<grid_for(nc, B), B, 0, s>>>(conn_dev, nc, deg.p);
    n2e_ptr.alloc(n_rows + 1);
    int64_t tot = 0;
    exclusive_scan_i32(deg.p, n2e_ptr.p, n_rows, &tot, s);
    n2e.alloc(nc ? nc : 1);
    cursor.alloc(n_rows);
    cursor.zero(s);
    if (nc) k_fill_n2e<<<grid_for(nc, B), B, 0, s>>>(conn_dev, nc, nper, n2e_ptr.p, cursor.p, n2e.p);
    k_sort_segments<<<grid_for(n_rows, 128), 128, 0, s>>>(n2e_ptr.p, n2e.p, n_rows);

    scratch.alloc(nc ? nc * nper : 1);
    rowlen.alloc(n_rows);
    k_row_unique<<<grid_for(n_rows, 128), 128, 0, s>>>(conn_dev, nper, n2e_ptr.p, n2e.p, n_rows, scratch.p, rowlen.p);
    pat.indptr.alloc(n_rows + 1);
    int64_t nnz = 0;
    exclusive_scan_i32(rowlen.p, pat.indptr.p, n_rows, &nnz, s);
    pat.nnz = nnz;
    pat.indices.alloc(nnz ? nnz : 1);
    k_copy_cols<<<grid_for(n_rows, 128), 128, 0, s>>>(n2e_ptr.p, nper, scratch.p, pat.indptr.p, pat.indices.p, n_rows);
    scratch.release();

    // nnz -> contributions, element -> nnz
    DevBuf<int32_t> cnt;
    cnt.alloc(nnz ? nnz : 1);
    pat.c_ptr.alloc(nnz + 1);
    pat.c_src.alloc(nc ? nc * nper : 1);
    pat.elem2nnz.alloc(nc ? nc * nper : 1);
    const int64_t threads = n_rows * 32;
    k_contrib<0><<<grid_for(threads, B), B, 0, s>>>(conn_dev, nper, n2e_ptr.p, n2e.p, pat.indptr.p, pat.indices.p,
                                                    n_rows, cnt.p, nullptr, nullptr, nullptr);
    int64_t ncontrib = 0;
    exclusive_scan_i32(cnt.p, pat.c_ptr.p, nnz, &ncontrib, s);
    FDB_REQUIRE(ncontrib == nc * nper, "element connectivity repeats a node inside one element");
    k_contrib<1><<<grid_for(threads, B), B, 0, s>>>(conn_dev, nper, n2e_ptr.p, n2e.p, pat.indptr.p, pat.indices.p,
                                                    n_rows, nullptr, pat.c_ptr.p, pat.c_src.p, pat.elem2nnz.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaStreamSynchronize(s));
}

void check_external_pattern(const int32_t* indptr, const int32_t* indices, int64_t n_rows, int64_t nnz, cudaStream_t) {
    FDB_REQUIRE(indptr[0] == 0 && indptr[n_rows] == nnz, "indptr does not span [0, nnz]");
    for (int64_t r = 0; r < n_rows; ++r) {
        FDB_REQUIRE(indptr[r] <= indptr[r + 1], "indptr is not monotone");
        for (int32_t k = indptr[r]; k < indptr[r + 1]; ++k) {
            FDB_REQUIRE(indices[k] >= 0 && indices[k] < n_rows, "column index out of range");
            FDB_REQUIRE(k == indptr[r] || indices[k - 1] < indices[k], "column indices must be sorted and unique");
        }
    }
}

}  // namespace fdb

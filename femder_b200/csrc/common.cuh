// femder_b200 internal declarations shared by pattern.cu / assemble.cu / sweep.cu / capi.cu.
// sm_100a only.  Nothing here is part of the C ABI (include/femder_b200.h is).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace fdb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

void set_last_error(const std::string& msg);

#define FDB_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            char _buf[512];                                                                              \
            snprintf(_buf, sizeof(_buf), "%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,              \
                     cudaGetErrorString(_e));                                                            \
            throw fdb::Error(_buf);                                                                      \
        }                                                                                                \
    } while (0)

#define FDB_REQUIRE(cond, msg)                                                                           \
    do {                                                                                                 \
        if (!(cond)) throw fdb::Error(std::string("femder_b200: ") + (msg));                             \
    } while (0)

// Owning device buffer (cudaMalloc / cudaFree on the system's device).
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    // Stream-ordered allocation out of the device's default memory pool (its release threshold is raised in fdb_system_create, so
    // freed blocks stay with the process): a set-up that allocates and frees tens of gigabytes of temporaries per call - the modal
    // set-up does - otherwise spends a noticeable part of its time mapping and unmapping device memory inside cudaMalloc / cudaFree.
    // Semantics are those of cudaMalloc / cudaFree: the free waits for the device, the memory is usable on every stream at once.
    void release() {
        if (p) {
            cudaDeviceSynchronize();
            cudaFreeAsync(p, (cudaStream_t)0);
        }
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count) {
        if (count == n && p) return;
        release();
        if (count) {
            FDB_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), (cudaStream_t)0));
            FDB_CUDA(cudaStreamSynchronize((cudaStream_t)0));
        }
        n = count;
    }
    void zero(cudaStream_t s) {
        if (n) FDB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
    size_t bytes() const { return n * sizeof(T); }
};

// CSR pattern + the maps that make assembly a deterministic segmented sum.
struct Pattern {
    int64_t n_rows = 0;
    int64_t nnz = 0;
    int64_t n_elem = 0;
    int nper = 0;
    DevBuf<int32_t> conn;      // [n_elem][nper]
    DevBuf<int32_t> indptr;    // [n_rows+1]
    DevBuf<int32_t> indices;   // [nnz]
    DevBuf<int32_t> elem2nnz;  // [n_elem][nper][nper]
    DevBuf<int32_t> c_ptr;     // [nnz+1]  contributions of nnz z are c_src[c_ptr[z] .. c_ptr[z+1])
    DevBuf<int32_t> c_src;     // [n_elem*nper*nper]  key e*nper^2 + a*nper + b, ascending within each nnz
    DevBuf<int32_t> n2e_ptr;   // [n_rows+1]  elements around node n are n2e[n2e_ptr[n] .. n2e_ptr[n+1]), ascending
    DevBuf<int32_t> n2e;       // [n_elem*nper]
    void reset() {
        n_rows = nnz = n_elem = 0;
        nper = 0;
        conn.release(); indptr.release(); indices.release(); elem2nnz.release(); c_ptr.release(); c_src.release();
        n2e_ptr.release(); n2e.release();
    }
};

struct SweepWork;  // sweep.cu

}  // namespace fdb

struct fdb_system {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = fdb::kNumSMs;

    // volume
    int64_t n_nodes = 0;
    fdb::DevBuf<double> nos;  // [n_nodes][3]
    fdb::Pattern vol;
    bool have_pattern = false;
    fdb::DevBuf<double> He, Qe;      // [nper*nper][n_elem]: formed on request (fdb_get_element_matrices) and for the Tet10 assembly
    bool have_elem_matrices = false;
    fdb::DevBuf<double2> hq;         // [nnz] (H, Q) interleaved: the layout the SpMV streams
    bool have_hq = false;
    fdb::DevBuf<double2> diag_hq;    // [n_nodes] diagonal of H and Q
    // SELL-8 copy of (indices, hq) the SpMV streams, in the internal order: slice s = rows 8s..8s+7, entry k of row r
    // at (sell_off[s] + k)*8 + r%8; rows shorter than the slice width are padded with (col = row, 0, 0)
    int64_t n_slices = 0;
    int64_t sell_entries = 0;
    fdb::DevBuf<int32_t> sell_off;   // [n_slices+1] in units of 8 entries
    fdb::DevBuf<int32_t> sell_idx;   // [sell_entries]
    fdb::DevBuf<double2> sell_hq;    // [sell_entries]
    std::vector<int32_t> sell_off_host;
    // Internal node order of the sweep (reorder.cu): new = perm[old], old = iperm[new].  The SELL stream, diag_hq and
    // the overlay are stored in the NEW order; the C ABI speaks the caller's numbering.
    bool have_order = false;
    fdb::DevBuf<int32_t> perm, iperm;
    std::vector<int32_t> perm_host;
    fdb::DevBuf<int32_t> sell_src;     // [sell_entries] CSR position behind each SELL entry, -1 = padding
    // 128-row tiles of k_spmv_blob: sorted distinct columns per tile, 16-bit position of every entry's column in it
    int64_t n_tiles = 0;
    int tile_rows = 128;               // rows per tile (128, or 64 with two resident blocks per SM)
    int max_tile_cols = 0, max_tile_entries = 0;
    fdb::DevBuf<int4> tile_meta;       // [n_tiles] (list start, count | first own row's position << 16, SELL offset, SELL entries)
    fdb::DevBuf<int32_t> tile_cols;    // column lists (new-order ids, ascending), each padded to a multiple of 4
    fdb::DevBuf<int32_t> tile_hdr;     // [n_tiles][20] relative SELL slice offsets (17), first own row's position, column count
    fdb::DevBuf<uint16_t> sell_lidx;   // [sell_entries]
    // block-Jacobi preconditioner: FP32 inverse of the tile's diagonal block of H (128 x 128, symmetric), per re-assembly
    fdb::DevBuf<float> blk_inv;        // [n_tiles][8256] packed lower triangles
    fdb::DevBuf<uint16_t> blk_inv16;   // [n_tiles][128 x 128] the same blocks in bf16, as 8 x 8 core matrices (tensor-core operand, modal.cu)
    bool have_blk = false;

    // surface
    int64_t n_tri = 0;
    int nper_s = 0;
    int n_groups = 0;
    fdb::Pattern surf;
    bool have_surf_pattern = false;
    fdb::DevBuf<int32_t> tri_group;  // [n_tri]
    fdb::DevBuf<double> areas;       // [n_tri]
    fdb::DevBuf<double> A_vals;      // [n_groups][nnz_b]
    fdb::DevBuf<uint8_t> A_touched;  // [n_groups][nnz_b]
    bool have_A = false;
    // overlay the SpMV reads: per row, entries (col, group, value) of all groups
    int64_t n_ov = 0;
    fdb::DevBuf<int32_t> ov_ptr;     // [n_nodes+1]
    fdb::DevBuf<int2> ov_cg;         // [n_ov] (col, group)
    fdb::DevBuf<double> ov_val;      // [n_ov]
    bool have_overlay = false;

    // Low room modes (modal.cu): H v = lambda Q v, V^T Q V = I.  Coarse correction of the sweep's preconditioner
    // (the modal path, femder/FEM_3D.py:1699-1768, as a deflation basis).
    int n_modes = 0;
    int mode_chunks = 0;               // 128-mode chunks stored per tile
    std::vector<double> modal_lam_host;
    fdb::DevBuf<double> modal_lam;     // [n_modes] ascending
    fdb::DevBuf<double> modal_res;     // [n_modes] ||H v - lambda Q v|| / ||Q v||: how far each stored vector is from an eigenvector
    std::vector<double> modal_res_host;
    fdb::DevBuf<uint16_t> modal_v;     // bf16 [n_tiles][mode_chunks][16 mode blocks][16 node blocks][8 nodes][8 modes]
    // tile-compressed modes (modal.cu): V_t ~ Phi_t C_t per 128-node tile
    bool modal_compress = false;       // the sweep streams Phi and C instead of V
    double modal_fit_worst = 0.0;      // worst relative fit error of a mode in the compressed form
    bool have_coarse_v = false;        // modal_phi / modal_vc are built
    bool have_fine_v = false;          // modal_v is packed (always without compression; on demand with it)
    int64_t nc_rows = 0, n_ctiles = 0; // coarse unknowns (16 per tile) and their 128-row tiles
    fdb::DevBuf<uint16_t> modal_phi;   // bf16 [n_tiles][2 function blocks][16 node blocks][8 nodes][8 functions]
    fdb::DevBuf<uint16_t> modal_vc;    // bf16 coefficient tiles C in modal_v's layout over the coarse unknowns
    fdb::DevBuf<double> modal_x;       // FP64 modes [ceil(n_modes / 16)][n_nodes (internal order)][16], kept on request
    fdb::DevBuf<double> modal_adiag;   // [n_groups][n_modes] v_i^T A_g v_i (damping of mode i by group g), built on demand
    bool have_modes = false, have_adiag = false;
    double rho0 = 0.0, c0 = 0.0;       // of the last fdb_assemble_volume (0: matrices given from outside)
    void* cublas = nullptr;            // cublasHandle_t / cusolverDnHandle_t of the modal set-up (dense Gram products, Rayleigh-Ritz)
    void* cusolver = nullptr;
    void* comm = nullptr;              // RankComm* (comm.cu): the ranks sharing the modal set-up; nullptr = this process alone
    int modal_outer = 0, modal_block = 0, modal_order = 0;   // diagnostics of the last fdb_modal_compute: outer iterations, block width, mass order
    double modal_resid = 0.0;               // ... worst relative residual of the kept pairs

    fdb::SweepWork* work = nullptr;
};

namespace fdb {

// ---- pattern.cu ----
void exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int64_t* total_host, cudaStream_t s);
void build_pattern(Pattern& pat, int64_t n_rows, int64_t n_elem, int nper, const int32_t* conn_dev, cudaStream_t s);
void check_external_pattern(const int32_t* indptr, const int32_t* indices, int64_t n_rows, int64_t nnz, cudaStream_t s);

// ---- assemble.cu ----
void assemble_volume(fdb_system* sys, double rho0, double c0);
void element_matrices(fdb_system* sys);
void heron_areas(fdb_system* sys);
void assemble_surface(fdb_system* sys);
void build_overlay(fdb_system* sys);
void extract_diagonal(fdb_system* sys);
void build_sell(fdb_system* sys);    // reorder.cu: SELL values from the CSR values
void build_block_prec(fdb_system* sys);   // reorder.cu: inverses of the tiles' diagonal blocks of H (block-Jacobi), on demand
void build_order(fdb_system* sys);   // reorder.cu: internal order + SELL structure + tile tables, once per pattern
void locate_points(fdb_system* sys, int64_t n_pts, const double* pts_dev, double tol, int32_t* nodes_dev, double* weights_dev);
void interleave_hq(fdb_system* sys, const double* H, const double* Q);
void deinterleave_hq(fdb_system* sys, double* H, double* Q);

// ---- modal.cu ----
void modal_reset(fdb_system* sys);   // the modes belong to the previous H, Q
void modal_free_handles(fdb_system* sys);

// ---- sweep.cu ----
void free_sweep_work(SweepWork* w);
void invalidate_graph(fdb_system* sys);  // call whenever a device pointer the captured graph holds may change

inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

}  // namespace fdb

// extern "C" entry points declared in include/femder_b200.h.  Exceptions never cross the boundary:
// every call returns a status and records a message for fdb_last_error().
#include "solver.cuh"

namespace fdb {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

void run_sweep(fdb_system* sys, const fdb_sweep_params* prm, fdb_sweep_result* res);
void run_spmv(fdb_system* sys, int F, const double* omega, const double* mu, const double* x, double* y);
void run_spmv_bench(fdb_system* sys, int F, int reps, int flags, double* ms_per_launch, int64_t* bytes);
void run_spmv_enqueue(fdb_system* sys, int F, int reps, int flags, int64_t* bytes);
void run_iteration_enqueue(fdb_system* sys, int F, int reps, int flags, int which, int64_t* bytes);

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) {
        FDB_CUDA(cudaGetDevice(&prev));
        if (prev != dev) FDB_CUDA(cudaSetDevice(dev));
        cur = dev;
    }
    ~DeviceGuard() {
        if (prev != cur) cudaSetDevice(prev);
    }
    int cur;
};
}  // namespace fdb

using namespace fdb;

#define FDB_API_BEGIN try {
#define FDB_API_END                                    \
    return 0;                                          \
    }                                                  \
    catch (const std::exception& e) {                  \
        fdb::set_last_error(e.what());                 \
        return 1;                                      \
    }                                                  \
    catch (...) {                                      \
        fdb::set_last_error("unknown C++ exception");  \
        return 2;                                      \
    }

#define FDB_SYS(sys)                                   \
    FDB_REQUIRE((sys) != nullptr, "null system handle"); \
    fdb::DeviceGuard _guard((sys)->device)
// for calls that (re)allocate buffers a captured sweep graph may point to
#define FDB_SYS_MUT(sys) \
    FDB_SYS(sys);        \
    fdb::invalidate_graph(sys)

extern "C" {

const char* fdb_last_error(void) { return fdb::g_last_error.c_str(); }
int fdb_abi_version(void) { return 2; }

int fdb_device_count(int* count) {
    FDB_API_BEGIN
    FDB_REQUIRE(count != nullptr, "null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    FDB_API_END
}

int fdb_system_create(int device, void* stream, fdb_system** out) {
    FDB_API_BEGIN
    FDB_REQUIRE(out != nullptr, "null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        throw Error("femder_b200: no CUDA device available (this library has no CPU fallback)");
    }
    FDB_REQUIRE(device >= 0 && device < n, "device index out of range");
    fdb::DeviceGuard guard(device);
    cudaDeviceProp prop;
    FDB_CUDA(cudaGetDeviceProperties(&prop, device));
    FDB_REQUIRE(prop.major >= 10, "femder_b200 is built for sm_100a (Blackwell B200) only");
    {   // keep freed device memory in the process (see DevBuf)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    fdb_system* s = new fdb_system();
    s->device = device;
    s->num_sms = prop.multiProcessorCount;
    if (stream) {
        s->stream = (cudaStream_t)stream;
        s->own_stream = false;
    } else {
        FDB_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        s->own_stream = true;
    }
    *out = s;
    FDB_API_END
}

int fdb_system_destroy(fdb_system* sys) {
    FDB_API_BEGIN
    if (!sys) return 0;
    {
        fdb::DeviceGuard guard(sys->device);
        cudaStreamSynchronize(sys->stream);
        fdb::free_sweep_work(sys->work);
        sys->work = nullptr;
        fdb::modal_free_handles(sys);
        fdb::comm_destroy((fdb::RankComm*)sys->comm);
        sys->comm = nullptr;
        if (sys->own_stream) cudaStreamDestroy(sys->stream);
        delete sys;
    }
    FDB_API_END
}

int fdb_synchronize(fdb_system* sys) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
    FDB_API_END
}

int fdb_set_volume_mesh(fdb_system* sys, int64_t n_nodes, const double* nos, int64_t n_elem, int nper, const int32_t* conn) {
    FDB_API_BEGIN
    FDB_SYS_MUT(sys);
    FDB_REQUIRE(nper == 4 || nper == 10, "nper must be 4 (Tet4) or 10 (Tet10)");
    FDB_REQUIRE(n_nodes > 0 && n_elem > 0 && nos && conn, "empty volume mesh");
    sys->n_nodes = n_nodes;
    sys->nos.alloc((size_t)n_nodes * 3);
    FDB_CUDA(cudaMemcpyAsync(sys->nos.p, nos, (size_t)n_nodes * 3 * sizeof(double), cudaMemcpyDefault, sys->stream));
    sys->vol.conn.alloc((size_t)n_elem * nper);
    FDB_CUDA(cudaMemcpyAsync(sys->vol.conn.p, conn, (size_t)n_elem * nper * sizeof(int32_t), cudaMemcpyDefault, sys->stream));
    sys->have_pattern = false;
    sys->have_hq = false;
    sys->have_elem_matrices = false;
    sys->have_order = false;
    sys->have_overlay = false;   // stored in the internal order of the previous pattern
    fdb::build_pattern(sys->vol, n_nodes, n_elem, nper, sys->vol.conn.p, sys->stream);
    fdb::build_order(sys);
    sys->have_pattern = true;
    FDB_API_END
}

int fdb_set_nodes(fdb_system* sys, int64_t n_nodes, const double* nos) {
    FDB_API_BEGIN
    FDB_SYS_MUT(sys);
    FDB_REQUIRE(n_nodes > 0 && nos, "no nodes");
    sys->n_nodes = n_nodes;
    sys->nos.alloc((size_t)n_nodes * 3);
    FDB_CUDA(cudaMemcpyAsync(sys->nos.p, nos, (size_t)n_nodes * 3 * sizeof(double), cudaMemcpyDefault, sys->stream));
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
    // whatever volume pattern, matrices, order and modes the system held belonged to another mesh
    sys->have_pattern = sys->have_hq = sys->have_order = sys->have_overlay = false;
    sys->have_A = sys->have_surf_pattern = false;
    sys->have_modes = sys->have_adiag = false;
    FDB_API_END
}

int fdb_pattern_nnz(fdb_system* sys, int64_t* nnz) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(sys->have_pattern && nnz, "no pattern");
    *nnz = sys->vol.nnz;
    FDB_API_END
}

int fdb_pattern_get(fdb_system* sys, int32_t* indptr, int32_t* indices) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(sys->have_pattern && indptr && indices, "no pattern");
    FDB_CUDA(cudaMemcpyAsync(indptr, sys->vol.indptr.p, (size_t)(sys->n_nodes + 1) * sizeof(int32_t), cudaMemcpyDefault, sys->stream));
    FDB_CUDA(cudaMemcpyAsync(indices, sys->vol.indices.p, (size_t)sys->vol.nnz * sizeof(int32_t), cudaMemcpyDefault, sys->stream));
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
    FDB_API_END
}

int fdb_elem2nnz_get(fdb_system* sys, int32_t* elem2nnz) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(sys->have_pattern && sys->vol.elem2nnz.p && elem2nnz, "no element->nnz map (external pattern?)");
    const size_t n = (size_t)sys->vol.n_elem * sys->vol.nper * sys->vol.nper;
    FDB_CUDA(cudaMemcpyAsync(elem2nnz, sys->vol.elem2nnz.p, n * sizeof(int32_t), cudaMemcpyDefault, sys->stream));
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
    FDB_API_END
}

int fdb_set_pattern(fdb_system* sys, int64_t n_nodes, int64_t nnz, const int32_t* indptr, const int32_t* indices) {
    FDB_API_BEGIN
    FDB_SYS_MUT(sys);
    FDB_REQUIRE(n_nodes > 0 && nnz >= 0 && nnz < 2147483647LL && indptr && indices, "bad pattern arguments");
    std::vector<int32_t> hp(n_nodes + 1), hi(nnz);
    FDB_CUDA(cudaMemcpy(hp.data(), indptr, hp.size() * sizeof(int32_t), cudaMemcpyDefault));
    if (nnz) FDB_CUDA(cudaMemcpy(hi.data(), indices, hi.size() * sizeof(int32_t), cudaMemcpyDefault));
    fdb::check_external_pattern(hp.data(), hi.data(), n_nodes, nnz, sys->stream);
    sys->n_nodes = n_nodes;
    sys->vol.reset();
    sys->vol.n_rows = n_nodes;
    sys->vol.nnz = nnz;
    sys->vol.indptr.alloc(n_nodes + 1);
    sys->vol.indices.alloc(nnz ? nnz : 1);
    FDB_CUDA(cudaMemcpy(sys->vol.indptr.p, hp.data(), hp.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (nnz) FDB_CUDA(cudaMemcpy(sys->vol.indices.p, hi.data(), hi.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    sys->have_order = false;
    sys->have_overlay = false;   // stored in the internal order of the previous pattern
    fdb::build_order(sys);
    sys->have_pattern = true;
    sys->have_hq = false;
    FDB_API_END
}

int fdb_assemble_volume(fdb_system* sys, double rho0, double c0) {
    FDB_API_BEGIN
    FDB_SYS_MUT(sys);
    fdb::assemble_volume(sys, rho0, c0);
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
    FDB_API_END
}

int fdb_get_HQ(fdb_system* sys, double* H_vals, double* Q_vals) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(sys->have_hq && H_vals && Q_vals, "H/Q values not available");
    fdb::deinterleave_hq(sys, H_vals, Q_vals);
    FDB_API_END
}

int fdb_set_HQ(fdb_system* sys, const double* H_vals, const double* Q_vals) {
    FDB_API_BEGIN
    FDB_SYS_MUT(sys);
    FDB_REQUIRE(sys->have_pattern && H_vals && Q_vals, "set the pattern before the values");
    fdb::interleave_hq(sys, H_vals, Q_vals);
    FDB_API_END
}

int fdb_get_element_matrices(fdb_system* sys, double* He, double* Qe) {
    FDB_API_BEGIN
    FDB_SYS_MUT(sys);
    FDB_REQUIRE(sys->have_hq && sys->vol.n_elem > 0 && He && Qe, "no element matrices (call fdb_assemble_volume)");
    fdb::element_matrices(sys);
    FDB_CUDA(cudaMemcpyAsync(He, sys->He.p, sys->He.bytes(), cudaMemcpyDefault, sys->stream));
    FDB_CUDA(cudaMemcpyAsync(Qe, sys->Qe.p, sys->Qe.bytes(), cudaMemcpyDefault, sys->stream));
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
    FDB_API_END
}

int fdb_set_surface_mesh(fdb_system* sys, int64_t n_tri, int nper, const int32_t* conn, const int32_t* group_of_tri, int n_groups) {
    FDB_API_BEGIN
    FDB_SYS_MUT(sys);
    FDB_REQUIRE(nper == 3 || nper == 6, "nper must be 3 (Tri3) or 6 (Tri6)");
    FDB_REQUIRE(sys->n_nodes > 0 && sys->nos.p, "fdb_set_volume_mesh must be called first");
    FDB_REQUIRE(n_tri >= 0 && n_groups >= 0 && (n_tri == 0 || (conn && group_of_tri)), "bad surface arguments");
    sys->n_tri = n_tri;
    sys->nper_s = nper;
    sys->n_groups = n_groups;
    sys->have_A = sys->have_overlay = sys->have_surf_pattern = false;
    sys->surf.conn.alloc(n_tri ? (size_t)n_tri * nper : 1);
    sys->tri_group.alloc(n_tri ? n_tri : 1);
    if (n_tri) {
        std::vector<int32_t> grp(n_tri);
        FDB_CUDA(cudaMemcpy(grp.data(), group_of_tri, n_tri * sizeof(int32_t), cudaMemcpyDefault));
        for (auto g : grp) FDB_REQUIRE(g >= -1 && g < n_groups, "group_of_tri entry out of range");
        FDB_CUDA(cudaMemcpyAsync(sys->surf.conn.p, conn, (size_t)n_tri * nper * sizeof(int32_t), cudaMemcpyDefault, sys->stream));
        FDB_CUDA(cudaMemcpyAsync(sys->tri_group.p, grp.data(), n_tri * sizeof(int32_t), cudaMemcpyHostToDevice, sys->stream));
        FDB_CUDA(cudaStreamSynchronize(sys->stream));
    }
    fdb::build_pattern(sys->surf, sys->n_nodes, n_tri, nper, sys->surf.conn.p, sys->stream);
    sys->have_surf_pattern = true;
    FDB_API_END
}

int fdb_heron_areas(fdb_system* sys, double* areas) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(areas != nullptr || sys->n_tri == 0, "null argument");
    fdb::heron_areas(sys);
    if (sys->n_tri) FDB_CUDA(cudaMemcpyAsync(areas, sys->areas.p, (size_t)sys->n_tri * sizeof(double), cudaMemcpyDefault, sys->stream));
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
    FDB_API_END
}

int fdb_assemble_surface(fdb_system* sys) {
    FDB_API_BEGIN
    FDB_SYS_MUT(sys);
    fdb::assemble_surface(sys);
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
    FDB_API_END
}

int fdb_surface_nnz(fdb_system* sys, int64_t* nnz_b) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(sys->have_surf_pattern && nnz_b, "no surface pattern");
    *nnz_b = sys->surf.nnz;
    FDB_API_END
}

int fdb_surface_pattern_get(fdb_system* sys, int32_t* indptr, int32_t* indices) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(sys->have_surf_pattern && indptr && indices, "no surface pattern");
    FDB_CUDA(cudaMemcpyAsync(indptr, sys->surf.indptr.p, (size_t)(sys->n_nodes + 1) * sizeof(int32_t), cudaMemcpyDefault, sys->stream));
    if (sys->surf.nnz)
        FDB_CUDA(cudaMemcpyAsync(indices, sys->surf.indices.p, (size_t)sys->surf.nnz * sizeof(int32_t), cudaMemcpyDefault, sys->stream));
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
    FDB_API_END
}

int fdb_surface_get(fdb_system* sys, double* A_vals, uint8_t* touched) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(sys->have_A, "fdb_assemble_surface must be called first");
    const size_t n = (size_t)sys->n_groups * sys->surf.nnz;
    if (n) {
        if (A_vals) FDB_CUDA(cudaMemcpyAsync(A_vals, sys->A_vals.p, n * sizeof(double), cudaMemcpyDefault, sys->stream));
        if (touched) FDB_CUDA(cudaMemcpyAsync(touched, sys->A_touched.p, n, cudaMemcpyDefault, sys->stream));
    }
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
    FDB_API_END
}

int fdb_set_surface(fdb_system* sys, int n_groups, int64_t nnz_b, const int32_t* indptr, const int32_t* indices, const double* A_vals) {
    FDB_API_BEGIN
    FDB_SYS_MUT(sys);
    FDB_REQUIRE(sys->n_nodes > 0, "set the volume pattern first");
    FDB_REQUIRE(n_groups >= 0 && nnz_b >= 0 && indptr && (nnz_b == 0 || (indices && A_vals)), "bad surface arguments");
    const int64_t N = sys->n_nodes;
    std::vector<int32_t> hp(N + 1), hi(nnz_b);
    FDB_CUDA(cudaMemcpy(hp.data(), indptr, hp.size() * sizeof(int32_t), cudaMemcpyDefault));
    if (nnz_b) FDB_CUDA(cudaMemcpy(hi.data(), indices, hi.size() * sizeof(int32_t), cudaMemcpyDefault));
    fdb::check_external_pattern(hp.data(), hi.data(), N, nnz_b, sys->stream);
    sys->surf.reset();
    sys->surf.n_rows = N;
    sys->surf.nnz = nnz_b;
    sys->surf.indptr.alloc(N + 1);
    sys->surf.indices.alloc(nnz_b ? nnz_b : 1);
    FDB_CUDA(cudaMemcpy(sys->surf.indptr.p, hp.data(), hp.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (nnz_b) FDB_CUDA(cudaMemcpy(sys->surf.indices.p, hi.data(), hi.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    sys->n_groups = n_groups;
    sys->n_tri = 0;
    const size_t tot = (size_t)n_groups * nnz_b;
    sys->A_vals.alloc(tot ? tot : 1);
    sys->A_touched.alloc(tot ? tot : 1);
    if (tot) {
        FDB_CUDA(cudaMemcpy(sys->A_vals.p, A_vals, tot * sizeof(double), cudaMemcpyDefault));
        // every non-zero value is an entry of its group's matrix
        std::vector<double> hv(tot);
        FDB_CUDA(cudaMemcpy(hv.data(), sys->A_vals.p, tot * sizeof(double), cudaMemcpyDeviceToHost));
        std::vector<uint8_t> ht(tot);
        for (size_t i = 0; i < tot; ++i) ht[i] = hv[i] != 0.0;
        FDB_CUDA(cudaMemcpy(sys->A_touched.p, ht.data(), tot, cudaMemcpyHostToDevice));
    }
    sys->have_surf_pattern = false;  // no triangle data behind this pattern
    sys->have_A = true;
    fdb::build_overlay(sys);
    FDB_API_END
}

int fdb_locate_points(fdb_system* sys, int64_t n_pts, const double* pts, double tol, int32_t* nodes, double* weights) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(n_pts >= 0 && (n_pts == 0 || (pts && nodes && weights)), "bad arguments");
    if (n_pts) {
        fdb::DevBuf<double> d_pts, d_w;
        fdb::DevBuf<int32_t> d_nodes;
        d_pts.alloc((size_t)n_pts * 3);
        d_w.alloc((size_t)n_pts * 4);
        d_nodes.alloc((size_t)n_pts * 4);
        FDB_CUDA(cudaMemcpyAsync(d_pts.p, pts, (size_t)n_pts * 3 * sizeof(double), cudaMemcpyDefault, sys->stream));
        fdb::locate_points(sys, n_pts, d_pts.p, tol, d_nodes.p, d_w.p);
        FDB_CUDA(cudaMemcpy(nodes, d_nodes.p, (size_t)n_pts * 4 * sizeof(int32_t), cudaMemcpyDefault));
        FDB_CUDA(cudaMemcpy(weights, d_w.p, (size_t)n_pts * 4 * sizeof(double), cudaMemcpyDefault));
    }
    FDB_API_END
}

int fdb_sweep(fdb_system* sys, const fdb_sweep_params* prm, fdb_sweep_result* res) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    fdb::run_sweep(sys, prm, res);
    FDB_API_END
}

int fdb_sizeof_sweep_params(void) { return (int)sizeof(fdb_sweep_params); }
int fdb_sizeof_sweep_result(void) { return (int)sizeof(fdb_sweep_result); }

int fdb_modal_compute(fdb_system* sys, double f_cut_hz, double f_acc_hz, int n_modes_hint, double tol, int max_outer, int* n_modes_found) {
    FDB_API_BEGIN
    FDB_SYS_MUT(sys);
    fdb::modal_compute(sys, f_cut_hz, f_acc_hz, n_modes_hint, tol, max_outer, n_modes_found);
    FDB_API_END
}

int fdb_modal_compression(fdb_system* sys, int on) {
    FDB_API_BEGIN
    FDB_SYS_MUT(sys);
    FDB_REQUIRE(sys->have_modes, "no modes");
    fdb::modal_set_compression(sys, on != 0, on == 2);
    FDB_API_END
}

int fdb_comm_unique_id(char* out128) {
    FDB_API_BEGIN
    FDB_REQUIRE(out128 != nullptr, "null argument");
    fdb::comm_unique_id(out128);
    FDB_API_END
}

int fdb_comm_init(fdb_system* sys, const char* id128, int rank, int world) {
    FDB_API_BEGIN
    FDB_SYS_MUT(sys);
    FDB_REQUIRE(id128 != nullptr, "null argument");
    fdb::comm_destroy((fdb::RankComm*)sys->comm);
    sys->comm = nullptr;
    sys->comm = fdb::comm_create(id128, rank, world);
    FDB_API_END
}

int fdb_comm_free(fdb_system* sys) {
    FDB_API_BEGIN
    FDB_SYS_MUT(sys);
    fdb::comm_destroy((fdb::RankComm*)sys->comm);
    sys->comm = nullptr;
    FDB_API_END
}

int fdb_modal_estimate(fdb_system* sys, double f_hz, int* n_modes) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(n_modes != nullptr, "null argument");
    *n_modes = fdb::modal_estimate(sys, f_hz);
    FDB_API_END
}

int fdb_modal_info(fdb_system* sys, int* outer_iterations, int* block_columns, int* mass_order, double* worst_residual, double* residuals) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(sys->have_modes, "no modes");
    if (outer_iterations) *outer_iterations = sys->modal_outer;
    if (block_columns) *block_columns = sys->modal_block;
    if (mass_order) *mass_order = sys->modal_order;
    if (worst_residual) *worst_residual = sys->modal_resid;
    if (residuals) FDB_CUDA(cudaMemcpy(residuals, sys->modal_res_host.data(), sys->n_modes * sizeof(double), cudaMemcpyDefault));
    FDB_API_END
}

int fdb_modal_set(fdb_system* sys, int n_modes, const double* lam, const double* V) {
    FDB_API_BEGIN
    FDB_SYS_MUT(sys);
    fdb::modal_set(sys, n_modes, lam, V);
    FDB_API_END
}

int fdb_modal_count(fdb_system* sys, int* n_modes) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(n_modes != nullptr, "null argument");
    *n_modes = sys->have_modes ? sys->n_modes : 0;
    FDB_API_END
}

int fdb_modal_get(fdb_system* sys, double* lam, double* V) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    fdb::modal_get(sys, lam, V);
    FDB_API_END
}

int fdb_modal_apply(fdb_system* sys, const double* omega, int n_modes_use, const double* r, double* z) {
    FDB_API_BEGIN
    FDB_SYS_MUT(sys);
    FDB_REQUIRE(omega && r && z, "null argument");
    fdb::run_modal_apply(sys, omega, n_modes_use, r, z);
    FDB_API_END
}

int fdb_spmv(fdb_system* sys, int batch, const double* omega, const double* mu, const double* x, double* y) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(omega && x && y, "null argument");
    fdb::run_spmv(sys, batch, omega, mu, x, y);
    FDB_API_END
}

int fdb_spmv_bench(fdb_system* sys, int batch, int reps, int flags, double* ms_per_launch, int64_t* algorithmic_bytes) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(reps > 0, "reps must be positive");
    fdb::run_spmv_bench(sys, batch, reps, flags, ms_per_launch, algorithmic_bytes);
    FDB_API_END
}

int fdb_spmv_enqueue(fdb_system* sys, int batch, int reps, int flags, int64_t* algorithmic_bytes) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(reps > 0, "reps must be positive");
    fdb::run_spmv_enqueue(sys, batch, reps, flags, algorithmic_bytes);
    FDB_API_END
}

int fdb_iteration_enqueue(fdb_system* sys, int batch, int reps, int flags, int which, int64_t* algorithmic_bytes) {
    FDB_API_BEGIN
    FDB_SYS(sys);
    FDB_REQUIRE(which < 0 || reps > 0, "reps must be positive");
    FDB_REQUIRE(which >= -1 && which <= 6, "which must be -1 (prepare), 0 (K5), 1 (K6a), 2 (K6b), 3 / 4 / 5 / 6 (modal projection / expansion / coefficients / z += Phi u)");
    fdb::run_iteration_enqueue(sys, batch, reps, flags, which, algorithmic_bytes);
    FDB_API_END
}

}  // extern "C"

// K1 (Tet4), K2 (Tet10), K3 (Tri3 / Tri6 + Heron areas): FP64 element matrices and their deterministic
// segmented sum into the CSR value arrays.
//
// Reference routines restated here (paths relative to the reference root):
//   Tet4   femder/fem_numerical.py:129-150 (ja_pre, arg_stiff_pre, det) and 267-345 (H_e, Q_e)
//   Tet10  femder/FEM_3D.py:244-254, 281-291 (shape functions, gradients), 888-919 (4-point rule)
//   Heron  femder/FEM_3D.py:745-771 (area_normal_ph)
//   Tri3   femder/FEM_3D.py:440-459, 475-523     Tri6  femder/FEM_3D.py:256-280, 541-554, 1053-1092
// The reference stores Tet4 H/Q in complex64 (fem_numerical.py:307-312); this path keeps FP64.
#include "common.cuh"

namespace fdb {

// 4-point Gauss rule, fem_numerical.py:153-171 / FEM_3D.py:897-901
constexpr double kGa = 0.5854101966249685;
constexpr double kGb = 0.1381966011250105;

struct Tet4Const {
    double S[16];  // sum over Gauss points of N_a N_b
};
struct Tet10Const {
    double N[4][10];      // shape functions at the Gauss points
    double G[4][3][10];   // gradients at the Gauss points
};
struct TriConst {
    double D3[9];         // Tri3 unit matrix (area multiplies it)
    double NN6[3][36];    // Tri6: N_a N_b at each Gauss point
    double w6;            // Tri6 weight 1/6*2
};

__constant__ Tet4Const c_tet4;
__constant__ Tet10Const c_tet10;
__constant__ TriConst c_tri;

static void tet10_shape_host(const double q[3], double N[10], double G[3][10]) {
    double t1 = q[0], t2 = q[1], t3 = q[2], t4 = 1 - q[0] - q[1] - q[2];
    double n[10] = {t4 * (2 * t4 - 1), t1 * (2 * t1 - 1), t2 * (2 * t2 - 1), t3 * (2 * t3 - 1), 4 * t1 * t4,
                    4 * t1 * t2,       4 * t2 * t4,       4 * t3 * t4,       4 * t2 * t3,       4 * t3 * t1};
    for (int i = 0; i < 10; ++i) N[i] = n[i];
    double u1 = 4 * q[0], u2 = 4 * q[1], u3 = 4 * q[2];
    double d[10][3] = {{u1 + u2 + u3 - 3, u1 + u2 + u3 - 3, u1 + u2 + u3 - 3},
                       {u1 - 1, 0, 0},
                       {0, u2 - 1, 0},
                       {0, 0, u3 - 1},
                       {4 - u2 - u3 - 2 * u1, -u1, -u1},
                       {u2, u1, 0},
                       {-u2, 4 - 2 * u2 - u3 - u1, -u2},
                       {-u3, -u3, 4 - u2 - 2 * u3 - u1},
                       {0, u3, u2},
                       {u3, 0, u1}};
    for (int i = 0; i < 10; ++i)
        for (int k = 0; k < 3; ++k) G[k][i] = d[i][k];
}

static void upload_constants() {
    static bool done[64] = {};
    int dev = 0;
    FDB_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && done[dev]) return;
    const double px[4] = {kGa, kGb, kGb, kGb}, py[4] = {kGb, kGa, kGb, kGb}, pz[4] = {kGb, kGb, kGa, kGb};
    Tet4Const t4{};
    for (int g = 0; g < 4; ++g) {
        double N[4] = {1 - px[g] - py[g] - pz[g], px[g], py[g], pz[g]};
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b) t4.S[a * 4 + b] += N[a] * N[b];
    }
    FDB_CUDA(cudaMemcpyToSymbol(c_tet4, &t4, sizeof(t4)));
    Tet10Const t10{};
    for (int g = 0; g < 4; ++g) {
        double q[3] = {px[g], py[g], pz[g]};
        tet10_shape_host(q, t10.N[g], t10.G[g]);
    }
    FDB_CUDA(cudaMemcpyToSymbol(c_tet10, &t10, sizeof(t10)));
    TriConst tc{};
    {
        // FEM_3D.py:494-503 + 440-459: shape_function[ix][r][c], damp += sf[ix] @ sf[ix].T / 9
        const double tx[3] = {1.0 / 6, 1.0 / 6, 2.0 / 3}, ty[3] = {1.0 / 6, 2.0 / 3, 1.0 / 6};
        double sf[3][3][3];
        // the three stacked (3,3) arrays are M0[j][k] = ptx[k], M1[j][k] = pty[j], M2[j][k] = 1 - ptx[k] - pty[j];
        // stacked S[i][j][k], transpose(2,0,1) -> sf[k][i][j] = S[i][j][k].
        for (int k = 0; k < 3; ++k)
            for (int j = 0; j < 3; ++j) {
                sf[k][0][j] = tx[k];
                sf[k][1][j] = ty[j];
                sf[k][2][j] = 1 - tx[k] - ty[j];
            }
        for (int i = 0; i < 9; ++i) tc.D3[i] = 0;
        for (int ix = 0; ix < 3; ++ix)
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) {
                    double dot = 0;
                    for (int j = 0; j < 3; ++j) dot += sf[ix][a][j] * sf[ix][b][j];
                    tc.D3[a * 3 + b] += dot * (1.0 / 9);
                }
        // FEM_3D.py:1069-1090 with Triangle10N (256-280)
        for (int g = 0; g < 3; ++g) {
            double q0 = tx[g], q1 = ty[g];
            double L0 = -q0 - q1 + 1;
            double N[6] = {L0 * (2 * L0 - 1), q0 * (2 * q0 - 1), q1 * (2 * q1 - 1), 4 * q0 * q1, 4 * q1 * L0, 4 * q0 * L0};
            for (int a = 0; a < 6; ++a)
                for (int b = 0; b < 6; ++b) tc.NN6[g][a * 6 + b] = N[a] * N[b];
        }
        tc.w6 = 1.0 / 6 * 2;
    }
    FDB_CUDA(cudaMemcpyToSymbol(c_tri, &tc, sizeof(tc)));
    if (dev < 64) done[dev] = true;
}

// ------------------------------------------------------------------------------------------------
// K1: Tet4, one thread per element.  conn is read as one 128-bit load per thread (coalesced).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void inv3(const double J[3][3], double det, double inv[3][3]) {
    const double id = 1.0 / det;
    inv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * id;
    inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    inv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * id;
    inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    inv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * id;
    inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
}

__device__ __forceinline__ double det3(const double J[3][3]) {
    return J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
           J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
}

__global__ void __launch_bounds__(128) k_tet4_elem(const int4* __restrict__ conn, const double* __restrict__ nos,
                                                    int64_t n_elem, double inv_rho, double inv_rhoc2,
                                                    double* __restrict__ He, double* __restrict__ Qe) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elem) return;
    const int4 c = conn[e];
    const int id[4] = {c.x, c.y, c.z, c.w};
    double X[4][3];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) X[a][k] = nos[(int64_t)id[a] * 3 + k];
    double J[3][3];  // row i = X[i+1] - X[0]   (ja_pre)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) J[i][k] = X[i + 1][k] - X[0][k];
    const double det = det3(J);  // signed: negatively oriented tets subtract
    double Ji[3][3];
    inv3(J, det, Ji);
    double B[3][4];  // B = J^-1 GNi, GNi = [[-1,1,0,0],[-1,0,1,0],[-1,0,0,1]]
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        B[i][0] = -(Ji[i][0] + Ji[i][1] + Ji[i][2]);
        B[i][1] = Ji[i][0];
        B[i][2] = Ji[i][1];
        B[i][3] = Ji[i][2];
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double btb = B[0][a] * B[0][b] + B[1][a] * B[1][b] + B[2][a] * B[2][b];
            const double arg = btb * det;                       // arg_stiff
            He[(int64_t)(a * 4 + b) * n_elem + e] = inv_rho * arg / 6;
            Qe[(int64_t)(a * 4 + b) * n_elem + e] = c_tet4.S[a * 4 + b] * det * inv_rhoc2 / 24;
        }
}

// ------------------------------------------------------------------------------------------------
// K2: Tet10.  blockIdx.y = row a of the element matrix; a thread produces H_e[a][0..9], Q_e[a][0..9] of one
// element.  The tile's connectivity is staged in shared memory with coalesced loads.
// ------------------------------------------------------------------------------------------------
constexpr int kT10Tile = 128;

__global__ void __launch_bounds__(kT10Tile) k_tet10_elem(const int32_t* __restrict__ conn,
                                                          const double* __restrict__ nos, int64_t n_elem,
                                                          double inv_rho, double inv_rhoc2, double* __restrict__ He,
                                                          double* __restrict__ Qe) {
    __shared__ int32_t s_conn[kT10Tile * 10];
    const int64_t e0 = (int64_t)blockIdx.x * kT10Tile;
    const int64_t n_here = min((int64_t)kT10Tile, n_elem - e0);
    for (int i = threadIdx.x; i < n_here * 10; i += kT10Tile) s_conn[i] = conn[e0 * 10 + i];
    __syncthreads();
    if (threadIdx.x >= n_here) return;
    const int64_t e = e0 + threadIdx.x;
    const int a = blockIdx.y;
    double X[10][3];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int64_t n = s_conn[threadIdx.x * 10 + i];
#pragma unroll
        for (int k = 0; k < 3; ++k) X[i][k] = nos[n * 3 + k];
    }
    double h[10], q[10];
#pragma unroll
    for (int b = 0; b < 10; ++b) h[b] = q[b] = 0.0;
    const double w = 1.0 / 24;
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
        double J[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double s = 0;
#pragma unroll
                for (int n = 0; n < 10; ++n) s += c_tet10.G[g][i][n] * X[n][k];
                J[i][k] = s;
            }
        const double det = det3(J);
        double Ji[3][3];
        inv3(J, det, Ji);
        double Ba[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
            Ba[i] = Ji[i][0] * c_tet10.G[g][0][a] + Ji[i][1] * c_tet10.G[g][1][a] + Ji[i][2] * c_tet10.G[g][2][a];
        const double Na = c_tet10.N[g][a];
#pragma unroll
        for (int b = 0; b < 10; ++b) {
            double btb = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double Bb = Ji[i][0] * c_tet10.G[g][0][b] + Ji[i][1] * c_tet10.G[g][1][b] + Ji[i][2] * c_tet10.G[g][2][b];
                btb += Ba[i] * Bb;
            }
            h[b] += w * (inv_rho * btb * det);                 // He + weigths*argHe1
            q[b] += w * (inv_rhoc2 * (Na * c_tet10.N[g][b]) * det);
        }
    }
#pragma unroll
    for (int b = 0; b < 10; ++b) {
        He[(int64_t)(a * 10 + b) * n_elem + e] = h[b];
        Qe[(int64_t)(a * 10 + b) * n_elem + e] = q[b];
    }
}

// ------------------------------------------------------------------------------------------------
// deterministic segmented sum: one thread per nnz adds its contributions in ascending (element, ab) order
// ------------------------------------------------------------------------------------------------
template <int N2>
__global__ void __launch_bounds__(256) k_reduce_hq(const int32_t* __restrict__ c_ptr, const int32_t* __restrict__ c_src,
                                                    const double* __restrict__ He, const double* __restrict__ Qe,
                                                    int64_t n_elem, int64_t nnz, double2* __restrict__ hq) {
    int64_t z = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= nnz) return;
    double h = 0, q = 0;
    const int32_t b = c_ptr[z], e = c_ptr[z + 1];
    for (int32_t k = b; k < e; ++k) {
        const int32_t key = c_src[k];
        const int64_t src = (int64_t)(key % N2) * n_elem + key / N2;
        h += He[src];
        q += Qe[src];
    }
    hq[z] = make_double2(h, q);
}

// Tet4 without the element-matrix round trip: one thread per nnz recomputes each of its contributions (element e, entry (a, b))
// from the element's four node coordinates - the same expressions as k_tet4_elem, added in the same ascending (element, ab)
// order, so the values are those of the two-kernel path - instead of reading them back from 2 x 16 E doubles that another
// kernel wrote (1.5 GB out and in again at config 4).  Connectivity and coordinates of neighbouring nnz are the same few
// elements: they come out of L1 / L2.
__global__ void __launch_bounds__(256) k_assemble_tet4_direct(const int32_t* __restrict__ c_ptr, const int32_t* __restrict__ c_src,
                                                               const int4* __restrict__ conn, const double* __restrict__ nos, int64_t nnz,
                                                               double inv_rho, double inv_rhoc2, double2* __restrict__ hq) {
    int64_t z = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= nnz) return;
    double h = 0, q = 0;
    const int32_t kb = c_ptr[z], ke = c_ptr[z + 1];
    for (int32_t k = kb; k < ke; ++k) {
        const int32_t key = c_src[k];
        const int e = key >> 4, a = (key >> 2) & 3, b = key & 3;
        const int4 c = __ldg(&conn[e]);
        const int id[4] = {c.x, c.y, c.z, c.w};
        double X[4][3];
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int d = 0; d < 3; ++d) X[n][d] = __ldg(&nos[(int64_t)id[n] * 3 + d]);
        double J[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int d = 0; d < 3; ++d) J[i][d] = X[i + 1][d] - X[0][d];
        const double det = det3(J);
        double Ji[3][3];
        inv3(J, det, Ji);
        double Ba[3], Bb[3];   // columns a and b of B = J^-1 GNi
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double b0 = -(Ji[i][0] + Ji[i][1] + Ji[i][2]);
            Ba[i] = a == 0 ? b0 : (a == 1 ? Ji[i][0] : (a == 2 ? Ji[i][1] : Ji[i][2]));
            Bb[i] = b == 0 ? b0 : (b == 1 ? Ji[i][0] : (b == 2 ? Ji[i][1] : Ji[i][2]));
        }
        const double btb = Ba[0] * Bb[0] + Ba[1] * Bb[1] + Ba[2] * Bb[2];
        const double arg = btb * det;
        h += inv_rho * arg / 6;
        q += c_tet4.S[a * 4 + b] * det * inv_rhoc2 / 24;
    }
    hq[z] = make_double2(h, q);
}

// the element matrices themselves (fdb_get_element_matrices; the Tet10 assembly): K1 / K2
void element_matrices(fdb_system* sys) {
    FDB_REQUIRE(sys->have_pattern && sys->vol.n_elem > 0 && sys->rho0 != 0.0 && sys->c0 != 0.0, "fdb_assemble_volume must be called first");
    if (sys->have_elem_matrices) return;
    upload_constants();
    Pattern& P = sys->vol;
    const int nper = P.nper, n2 = nper * nper;
    const int64_t E = P.n_elem;
    sys->He.alloc((size_t)n2 * E);
    sys->Qe.alloc((size_t)n2 * E);
    const double inv_rho = 1.0 / sys->rho0;
    const double inv_rhoc2 = 1.0 / (sys->rho0 * (sys->c0 * sys->c0));
    if (nper == 4) {
        k_tet4_elem<<<grid_for(E, 128), 128, 0, sys->stream>>>((const int4*)P.conn.p, sys->nos.p, E, inv_rho, inv_rhoc2,
                                                               sys->He.p, sys->Qe.p);
    } else {
        dim3 grid(grid_for(E, kT10Tile), 10);
        k_tet10_elem<<<grid, kT10Tile, 0, sys->stream>>>(P.conn.p, sys->nos.p, E, inv_rho, inv_rhoc2, sys->He.p, sys->Qe.p);
    }
    FDB_CUDA(cudaGetLastError());
    sys->have_elem_matrices = true;
}

void assemble_volume(fdb_system* sys, double rho0, double c0) {
    FDB_REQUIRE(sys->have_pattern && sys->vol.n_elem > 0 && sys->vol.conn.p, "fdb_set_volume_mesh must be called first");
    FDB_REQUIRE(rho0 != 0.0 && c0 != 0.0, "rho0 and c0 must be non-zero");
    sys->rho0 = rho0;   // kept for Weyl's mode-count estimate (modal.cu)
    sys->c0 = c0;
    upload_constants();
    Pattern& P = sys->vol;
    const int nper = P.nper;
    const int64_t E = P.n_elem;
    sys->have_elem_matrices = false;   // of another fluid or geometry
    sys->hq.alloc(P.nnz);
    if (nper == 4) {
        // Tet4: straight from the coordinates (the element matrices are formed only if somebody asks for them)
        sys->He.release();
        sys->Qe.release();
        k_assemble_tet4_direct<<<grid_for(P.nnz, 256), 256, 0, sys->stream>>>(P.c_ptr.p, P.c_src.p, (const int4*)P.conn.p, sys->nos.p, P.nnz,
                                                                             1.0 / rho0, 1.0 / (rho0 * (c0 * c0)), sys->hq.p);
    } else {
        // Tet10: 4 Gauss points x 10 nodes per entry are too much to recompute per contribution
        element_matrices(sys);
        k_reduce_hq<100><<<grid_for(P.nnz, 256), 256, 0, sys->stream>>>(P.c_ptr.p, P.c_src.p, sys->He.p, sys->Qe.p, E, P.nnz, sys->hq.p);
    }
    FDB_CUDA(cudaGetLastError());
    sys->have_hq = true;
    extract_diagonal(sys);
}

// ------------------------------------------------------------------------------------------------
// (H, Q) interleave for the SpMV stream, diagonal for the Jacobi preconditioner
// ------------------------------------------------------------------------------------------------
__global__ void k_interleave(const double* __restrict__ H, const double* __restrict__ Q, int64_t n, double2* __restrict__ hq) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hq[i] = make_double2(H[i], Q[i]);
}
__global__ void k_deinterleave(const double2* __restrict__ hq, int64_t n, double* __restrict__ H, double* __restrict__ Q) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        double2 v = hq[i];
        H[i] = v.x;
        Q[i] = v.y;
    }
}
__global__ void k_diag(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                       const double2* __restrict__ hq, int64_t n_rows, const int32_t* __restrict__ perm, double2* __restrict__ diag) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int lo = indptr[r], hi = indptr[r + 1];
    double2 d = make_double2(0.0, 0.0);
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        int c = indices[mid];
        if (c == r) { d = hq[mid]; break; }
        if (c < r) lo = mid + 1; else hi = mid;
    }
    diag[perm[r]] = d;   // stored in the sweep's internal order
}

void interleave_hq(fdb_system* sys, const double* H, const double* Q) {
    const int64_t nnz = sys->vol.nnz;
    DevBuf<double> dH, dQ;
    dH.alloc(nnz);
    dQ.alloc(nnz);
    FDB_CUDA(cudaMemcpyAsync(dH.p, H, nnz * sizeof(double), cudaMemcpyDefault, sys->stream));
    FDB_CUDA(cudaMemcpyAsync(dQ.p, Q, nnz * sizeof(double), cudaMemcpyDefault, sys->stream));
    sys->hq.alloc(nnz);
    k_interleave<<<grid_for(nnz, 256), 256, 0, sys->stream>>>(dH.p, dQ.p, nnz, sys->hq.p);
    FDB_CUDA(cudaGetLastError());
    sys->have_hq = true;
    extract_diagonal(sys);
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
}

void deinterleave_hq(fdb_system* sys, double* H, double* Q) {
    const int64_t nnz = sys->vol.nnz;
    DevBuf<double> dH, dQ;
    dH.alloc(nnz);
    dQ.alloc(nnz);
    k_deinterleave<<<grid_for(nnz, 256), 256, 0, sys->stream>>>(sys->hq.p, nnz, dH.p, dQ.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaMemcpyAsync(H, dH.p, nnz * sizeof(double), cudaMemcpyDefault, sys->stream));
    FDB_CUDA(cudaMemcpyAsync(Q, dQ.p, nnz * sizeof(double), cudaMemcpyDefault, sys->stream));
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
}

void extract_diagonal(fdb_system* sys) {
    FDB_REQUIRE(sys->have_order, "internal order missing");
    sys->diag_hq.alloc(sys->n_nodes);
    k_diag<<<grid_for(sys->n_nodes, 256), 256, 0, sys->stream>>>(sys->vol.indptr.p, sys->vol.indices.p, sys->hq.p,
                                                                  sys->n_nodes, sys->perm.p, sys->diag_hq.p);
    FDB_CUDA(cudaGetLastError());
    build_sell(sys);
}

// ------------------------------------------------------------------------------------------------
// K3: Heron areas (IEEE operations in the reference's order, no FMA contraction) and A_g values
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double edge_len(const double* __restrict__ nos, int64_t i, int64_t j) {
    const double dx = __dsub_rn(nos[i * 3 + 0], nos[j * 3 + 0]);
    const double dy = __dsub_rn(nos[i * 3 + 1], nos[j * 3 + 1]);
    const double dz = __dsub_rn(nos[i * 3 + 2], nos[j * 3 + 2]);
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

__global__ void k_heron(const int32_t* __restrict__ conn, int nper, const double* __restrict__ nos, int64_t n_tri,
                        double* __restrict__ areas) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tri) return;
    const int64_t i0 = conn[t * nper + 0], i1 = conn[t * nper + 1], i2 = conn[t * nper + 2];
    const double a = edge_len(nos, i0, i1);
    const double b = edge_len(nos, i1, i2);
    const double c = edge_len(nos, i2, i0);
    const double p = __ddiv_rn(__dadd_rn(__dadd_rn(a, b), c), 2.0);
    const double prod = __dmul_rn(__dmul_rn(__dmul_rn(p, __dsub_rn(p, a)), __dsub_rn(p, b)), __dsub_rn(p, c));
    areas[t] = fabs(__dsqrt_rn(prod));  // NaN for slivers with a negative product, as np.sqrt gives
}

template <int NPER>
__global__ void __launch_bounds__(256) k_reduce_surface(const int32_t* __restrict__ c_ptr, const int32_t* __restrict__ c_src,
                                                         const int32_t* __restrict__ tri_group,
                                                         const double* __restrict__ areas, int64_t nnz_b,
                                                         double* __restrict__ A_vals, uint8_t* __restrict__ touched) {
    constexpr int N2 = NPER * NPER;
    int64_t z = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= nnz_b) return;
    const int32_t b = c_ptr[z], e = c_ptr[z + 1];
    for (int32_t k = b; k < e; ++k) {
        const int32_t key = c_src[k];
        const int32_t t = key / N2, ab = key % N2;
        const int g = tri_group[t];
        if (g < 0) continue;
        const double area = areas[t];
        double ae;
        if (NPER == 3) {
            ae = __dmul_rn(c_tri.D3[ab], area);                // _damp_elem * area_elem
        } else {
            ae = 0.0;
#pragma unroll
            for (int gp = 0; gp < 3; ++gp)                     // Ae + weight*(Ni Ni^T * detJa)
                ae = __dadd_rn(ae, __dmul_rn(c_tri.w6, __dmul_rn(c_tri.NN6[gp][ab], area)));
        }
        const int64_t o = (int64_t)g * nnz_b + z;
        A_vals[o] = __dadd_rn(A_vals[o], ae);
        touched[o] = 1;
    }
}

void heron_areas(fdb_system* sys) {
    FDB_REQUIRE(sys->have_surf_pattern, "fdb_set_surface_mesh must be called first");
    sys->areas.alloc(sys->n_tri ? sys->n_tri : 1);
    if (sys->n_tri)
        k_heron<<<grid_for(sys->n_tri, 256), 256, 0, sys->stream>>>(sys->surf.conn.p, sys->nper_s, sys->nos.p, sys->n_tri,
                                                                     sys->areas.p);
    FDB_CUDA(cudaGetLastError());
}

void assemble_surface(fdb_system* sys) {
    FDB_REQUIRE(sys->have_surf_pattern, "fdb_set_surface_mesh must be called first");
    upload_constants();
    heron_areas(sys);
    const int64_t Zb = sys->surf.nnz;
    const size_t tot = (size_t)sys->n_groups * Zb;
    sys->A_vals.alloc(tot ? tot : 1);
    sys->A_touched.alloc(tot ? tot : 1);
    sys->A_vals.zero(sys->stream);
    sys->A_touched.zero(sys->stream);
    if (Zb && sys->n_groups) {
        if (sys->nper_s == 3)
            k_reduce_surface<3><<<grid_for(Zb, 256), 256, 0, sys->stream>>>(sys->surf.c_ptr.p, sys->surf.c_src.p, sys->tri_group.p,
                                                                             sys->areas.p, Zb, sys->A_vals.p, sys->A_touched.p);
        else
            k_reduce_surface<6><<<grid_for(Zb, 256), 256, 0, sys->stream>>>(sys->surf.c_ptr.p, sys->surf.c_src.p, sys->tri_group.p,
                                                                             sys->areas.p, Zb, sys->A_vals.p, sys->A_touched.p);
    }
    FDB_CUDA(cudaGetLastError());
    sys->have_A = true;
    // the overlay is the sweep's copy of the boundary matrices in the volume pattern's internal order: nodes-only systems
    // (fdb_set_nodes: areas and boundary matrices on their own) have neither
    if (sys->have_order) build_overlay(sys);
}

// ------------------------------------------------------------------------------------------------
// K7 (set-up half): batched receiver location.  Replaces closest_node + prob_elem + which_tetra + coord_interpolation
// (femder/FEM_3D.py:173-224) and the nearest-node rule of evaluate() (2077-2086) for many points at once.
// ------------------------------------------------------------------------------------------------
// one block per query point: argmin over the nodes of |x_n - p|^2, ties to the lowest index (np.argmin)
__global__ void __launch_bounds__(256) k_nearest_node(const double* __restrict__ nos, int64_t n_nodes, const double* __restrict__ pts,
                                                       int32_t* __restrict__ nearest, double* __restrict__ dist2) {
    const int64_t q = blockIdx.x;
    const double px = pts[q * 3 + 0], py = pts[q * 3 + 1], pz = pts[q * 3 + 2];
    double best = 1.7976931348623157e308;
    int64_t bi = 0x7fffffff;
    for (int64_t n = threadIdx.x; n < n_nodes; n += blockDim.x) {
        const double dx = nos[n * 3 + 0] - px, dy = nos[n * 3 + 1] - py, dz = nos[n * 3 + 2] - pz;
        const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (d < best) { best = d; bi = n; }   // ascending n per thread: the first minimum is kept
    }
    __shared__ double s_d[256];
    __shared__ long long s_i[256];
    s_d[threadIdx.x] = best;
    s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const double d2 = s_d[threadIdx.x + o];
            const long long i2 = s_i[threadIdx.x + o];
            if (d2 < s_d[threadIdx.x] || (d2 == s_d[threadIdx.x] && i2 < s_i[threadIdx.x])) { s_d[threadIdx.x] = d2; s_i[threadIdx.x] = i2; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { nearest[q] = (int32_t)s_i[0]; dist2[q] = s_d[0]; }
}

// one thread per query point: node stencil within `tol`, else the LAST element around the nearest node whose
// barycentric coordinates pass 0 <= l <= 1, sum <= 1 (which_tetra: later hits overwrite), else the last candidate
// (the reference's -1 sentinel); linear shape functions of the element's 4 corner nodes
__global__ void k_locate(const double* __restrict__ nos, const int32_t* __restrict__ conn, int nper, const int32_t* __restrict__ n2e_ptr,
                         const int32_t* __restrict__ n2e, const double* __restrict__ pts, const int32_t* __restrict__ nearest,
                         const double* __restrict__ dist2, double tol, int64_t n_pts, int32_t* __restrict__ nodes,
                         double* __restrict__ weights) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_pts) return;
    const int32_t n = nearest[q];
    if (sqrt(dist2[q]) < tol || n2e_ptr[n] == n2e_ptr[n + 1]) {
        for (int j = 0; j < 4; ++j) { nodes[q * 4 + j] = n; weights[q * 4 + j] = (j == 0) ? 1.0 : 0.0; }
        return;
    }
    const double p[3] = {pts[q * 3 + 0], pts[q * 3 + 1], pts[q * 3 + 2]};
    int32_t hit = -1;
    double lh[3] = {0.0, 0.0, 0.0};
    const int32_t kb = n2e_ptr[n], ke = n2e_ptr[n + 1];
    for (int32_t k = kb; k < ke; ++k) {
        const int32_t e = n2e[k];
        const int32_t* c = conn + (int64_t)e * nper;
        double o[3], M[3][3];   // columns of M = edge vectors from corner 0
        for (int d = 0; d < 3; ++d) o[d] = nos[(int64_t)c[0] * 3 + d];
        for (int j = 0; j < 3; ++j)
            for (int d = 0; d < 3; ++d) M[d][j] = nos[(int64_t)c[j + 1] * 3 + d] - o[d];
        const double det = det3(M);
        double Mi[3][3];
        inv3(M, det, Mi);
        double l[3];
        for (int i = 0; i < 3; ++i) l[i] = Mi[i][0] * (p[0] - o[0]) + Mi[i][1] * (p[1] - o[1]) + Mi[i][2] * (p[2] - o[2]);
        const bool inside = l[0] >= 0 && l[1] >= 0 && l[2] >= 0 && l[0] <= 1 && l[1] <= 1 && l[2] <= 1 && (l[0] + l[1] + l[2]) <= 1;
        if (inside || (hit < 0 && k == ke - 1)) {
            if (inside || hit < 0) { hit = e; lh[0] = l[0]; lh[1] = l[1]; lh[2] = l[2]; }
        }
    }
    const int32_t* c = conn + (int64_t)hit * nper;
    for (int j = 0; j < 4; ++j) nodes[q * 4 + j] = c[j];
    weights[q * 4 + 0] = 1.0 - lh[0] - lh[1] - lh[2];
    weights[q * 4 + 1] = lh[0];
    weights[q * 4 + 2] = lh[1];
    weights[q * 4 + 3] = lh[2];
}

void locate_points(fdb_system* sys, int64_t n_pts, const double* pts_dev, double tol, int32_t* nodes_dev, double* weights_dev) {
    FDB_REQUIRE(sys->have_pattern && sys->vol.conn.p && sys->vol.n2e_ptr.p && sys->nos.p, "fdb_set_volume_mesh must be called first");
    if (n_pts == 0) return;
    DevBuf<int32_t> nearest;
    DevBuf<double> dist2;
    nearest.alloc(n_pts);
    dist2.alloc(n_pts);
    k_nearest_node<<<(unsigned)n_pts, 256, 0, sys->stream>>>(sys->nos.p, sys->n_nodes, pts_dev, nearest.p, dist2.p);
    k_locate<<<grid_for(n_pts, 128), 128, 0, sys->stream>>>(sys->nos.p, sys->vol.conn.p, sys->vol.nper, sys->vol.n2e_ptr.p, sys->vol.n2e.p,
                                                           pts_dev, nearest.p, dist2.p, tol, n_pts, nodes_dev, weights_dev);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
}

// ------------------------------------------------------------------------------------------------
// overlay for the SpMV: per row, (col, group, value) of every touched (group, entry), ordered by entry then group
// ------------------------------------------------------------------------------------------------
template <int PASS>
__global__ void k_overlay(const int32_t* __restrict__ b_indptr, const int32_t* __restrict__ b_indices,
                          const double* __restrict__ A_vals, const uint8_t* __restrict__ touched, int n_groups,
                          int64_t nnz_b, int64_t n_rows, const int32_t* __restrict__ perm, const int32_t* __restrict__ iperm,
                          int32_t* __restrict__ cnt, const int32_t* __restrict__ ov_ptr,
                          int2* __restrict__ ov_cg, double* __restrict__ ov_val) {
    // one thread per row of the sweep's internal order; r = that row in the caller's numbering
    const int64_t rn = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (rn >= n_rows) return;
    const int64_t r = iperm[rn];
    int n = 0;
    int32_t out = (PASS == 1) ? ov_ptr[rn] : 0;
    for (int32_t z = b_indptr[r]; z < b_indptr[r + 1]; ++z)
        for (int g = 0; g < n_groups; ++g)
            if (touched[(int64_t)g * nnz_b + z]) {
                if (PASS == 1) {
                    ov_cg[out + n] = make_int2(perm[b_indices[z]], g);
                    ov_val[out + n] = A_vals[(int64_t)g * nnz_b + z];
                }
                ++n;
            }
    if (PASS == 0) cnt[rn] = n;
}

void build_overlay(fdb_system* sys) {
    FDB_REQUIRE(sys->have_order, "set the volume pattern before the surface matrices");
    const int64_t N = sys->n_nodes, Zb = sys->surf.nnz;
    DevBuf<int32_t> cnt;
    cnt.alloc(N);
    sys->ov_ptr.alloc(N + 1);
    k_overlay<0><<<grid_for(N, 256), 256, 0, sys->stream>>>(sys->surf.indptr.p, sys->surf.indices.p, sys->A_vals.p,
                                                           sys->A_touched.p, sys->n_groups, Zb, N, sys->perm.p, sys->iperm.p, cnt.p, nullptr, nullptr, nullptr);
    int64_t n_ov = 0;
    exclusive_scan_i32(cnt.p, sys->ov_ptr.p, N, &n_ov, sys->stream);
    sys->n_ov = n_ov;
    sys->ov_cg.alloc(n_ov ? n_ov : 1);
    sys->ov_val.alloc(n_ov ? n_ov : 1);
    k_overlay<1><<<grid_for(N, 256), 256, 0, sys->stream>>>(sys->surf.indptr.p, sys->surf.indices.p, sys->A_vals.p,
                                                           sys->A_touched.p, sys->n_groups, Zb, N, sys->perm.p, sys->iperm.p, nullptr, sys->ov_ptr.p,
                                                           sys->ov_cg.p, sys->ov_val.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaStreamSynchronize(sys->stream));
    sys->have_overlay = true;
}

}  // namespace fdb

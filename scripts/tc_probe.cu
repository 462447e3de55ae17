// Bring-up probe for the tcgen05 building blocks of femder_b200/csrc/modal.cu (not part of the product).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/tc_probe scripts/tc_probe.cu && scripts/tc_probe
//
// One CTA computes D[128 x 32] = A[128 x K] B[K x 32] (bf16 in, fp32 accumulate in TMEM) from hand-laid shared-memory
// operand images without swizzle, for the two operand forms the modal kernels use:
//   case 0  A K-major  (rows = nodes, K = modes):   the expansion  z += V e
//   case 1  A MN-major (rows = modes, K = nodes):   the projection c = V^T r   - the SAME bytes of V
// B is MN-major in both.  Each case is run with (LBO, SBO) as derived in modal.cu and with the two swapped, so one run on
// the GPU tells which reading of the shared-memory descriptor the hardware implements.  Waits are bounded: a wrong
// descriptor must not hang the GPU.
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    // tcgen05 shared-memory matrix descriptor, no swizzle: start address, leading / stride byte offsets in 16-byte units,
    // version 1 (bits 46-47), base offset 0, layout type 0 (bits 61-63)
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__global__ void __launch_bounds__(128) probe(const uint4* a_img, int a_bytes, const uint4* b_img, int b_bytes, float* d_out, uint32_t idesc,
                                             uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, uint32_t a_kstep, uint32_t b_kstep,
                                             int ksteps, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sa = smem;
    uint8_t* sb = smem + 65536;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < a_bytes / 16; i += 128) reinterpret_cast<uint4*>(sa)[i] = a_img[i];
    for (int i = tid; i < b_bytes / 16; i += 128) reinterpret_cast<uint4*>(sb)[i] = b_img[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base;
    if (warp == 0 && lane == 0) {
        for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t da = make_desc(smem_u32(sa) + ks * a_kstep, a_lbo, a_sbo);
            const uint64_t db = make_desc(smem_u32(sb) + ks * b_kstep, b_lbo, b_sbo);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
                "}\n" ::"r"(tbase), "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // bounded wait on the commit
    bool ok = false;
    for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(&bar)), "r"(0u)
            : "memory");
        ok = done != 0;
    }
    ok = __all_sync(0xffffffffu, ok);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
        uint32_t v[32];
        const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
            "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) d_out[(warp * 32 + lane) * 32 + j] = __uint_as_float(v[j]);
    } else if (tid == 0) {
        *status = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(32));
}

static uint16_t bf16_of(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return (uint16_t)(u >> 16);   // the test values are small integers: exact
}

int main() {
    const int M = 128, N = 32, K = 32;
    std::vector<float> A((size_t)M * K), B((size_t)K * N), D((size_t)M * N, 0.f);
    uint32_t s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((int)((s >> 24) % 9) - 4); };
    for (auto& v : A) v = rnd();
    for (auto& v : B) v = rnd();
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float acc = 0;
            for (int k = 0; k < K; ++k) acc += A[m * K + k] * B[k * N + n];
            D[m * N + n] = acc;
        }
    // B image, MN-major: core (kb, cb) at cb * S_cb + kb * 128, inside (k % 8) * 16 + (n % 8) * 2
    const uint32_t S_cb = (K / 8) * 128;
    std::vector<uint16_t> bimg((size_t)K * N);
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) bimg[((n / 8) * S_cb + (k / 8) * 128 + (k % 8) * 16 + (n % 8) * 2) / 2] = bf16_of(B[k * N + n]);
    uint4 *d_a, *d_b;
    float* d_d;
    int* d_status;
    CK(cudaMalloc(&d_a, 65536));
    CK(cudaMalloc(&d_b, 8192));
    CK(cudaMalloc(&d_d, M * N * 4));
    CK(cudaMalloc(&d_status, 4));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 8192));
    int all_ok = 1;
    for (int cs = 0; cs < 2; ++cs) {
        // A image.  Physical form of a tile of V: core (nb, mb) = 8 nodes x 8 modes at nb * S_n + mb * S_m, inside node * 16 + mode * 2
        std::vector<uint16_t> aimg((size_t)M * K, 0);
        uint32_t S_n, S_m, lbo, sbo, kstep;
        if (cs == 0) {   // rows (M) = nodes, K = modes: K-major
            S_m = 128; S_n = (K / 8) * 128;
            for (int m = 0; m < M; ++m)
                for (int k = 0; k < K; ++k) aimg[((m / 8) * S_n + (k / 8) * S_m + (m % 8) * 16 + (k % 8) * 2) / 2] = bf16_of(A[m * K + k]);
            sbo = S_n; lbo = S_m; kstep = 2 * S_m;
        } else {         // rows (M) = modes, K = nodes: MN-major
            S_m = 128; S_n = (M / 8) * 128;
            for (int m = 0; m < M; ++m)
                for (int k = 0; k < K; ++k) aimg[((k / 8) * S_n + (m / 8) * S_m + (k % 8) * 16 + (m % 8) * 2) / 2] = bf16_of(A[m * K + k]);
            sbo = S_m; lbo = S_n; kstep = 2 * S_n;
        }
        CK(cudaMemcpy(d_a, aimg.data(), aimg.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_b, bimg.data(), bimg.size() * 2, cudaMemcpyHostToDevice));
        // instruction descriptor: c = F32 (1 << 4), a = b = BF16 (1 << 7, 1 << 10), a_major (bit 15), b_major (bit 16), N >> 3 at 17, M >> 4 at 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((cs == 1 ? 1u : 0u) << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
                               ((uint32_t)(M >> 4) << 24);
        for (int variant = 0; variant < 4; ++variant) {
            const bool swap_a = variant & 1, swap_b = variant & 2;
            const uint32_t al = swap_a ? sbo : lbo, as = swap_a ? lbo : sbo;
            const uint32_t bl = swap_b ? S_cb : 128u, bs = swap_b ? 128u : S_cb;
            CK(cudaMemset(d_d, 0xff, M * N * 4));
            CK(cudaMemset(d_status, 0, 4));
            probe<<<1, 128, 65536 + 8192>>>(d_a, (int)(aimg.size() * 2), d_b, (int)(bimg.size() * 2), d_d, idesc, al, as, bl, bs, kstep, 256u, K / 16, d_status);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("case %d variant %d: CUDA error %s\n", cs, variant, cudaGetErrorString(e)); return 3; }
            std::vector<float> out((size_t)M * N);
            int st = 0;
            CK(cudaMemcpy(out.data(), d_d, out.size() * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (size_t i = 0; i < out.size(); ++i) bad += (out[i] != D[i]);
            printf("case %d (%s) variant %d (A %s, B %s): status %d, mismatches %d / %d   D[0][0..3] = %g %g %g %g (want %g %g %g %g)  D[9][17] = %g (want %g)\n",
                   cs, cs == 0 ? "A K-major" : "A MN-major", variant, swap_a ? "swapped" : "derived", swap_b ? "swapped" : "derived", st, bad, M * N,
                   out[0], out[1], out[2], out[3], D[0], D[1], D[2], D[3], out[9 * N + 17], D[9 * N + 17]);
            if (variant == 0 && (bad || st)) all_ok = 0;
        }
    }
    printf(all_ok ? "PROBE OK: derived descriptors are right\n" : "PROBE: derived descriptors are NOT right - see the variants above\n");
    return 0;
}

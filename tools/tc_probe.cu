// Bring-up probe for hand-written tcgen05 (sm_100a): D[M x N] (fp32, TMEM) = A[M x K] * B[N x K]^T
// with fp16 operands written to shared memory by ordinary threads in the canonical
// no-swizzle ("interleave") core-matrix layouts, for K-major and MN-major operands.
// Tries the descriptor conventions that cannot be verified without hardware and prints
// which ones reproduce a host reference.   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Variant {
    int a_mn_major;    // 0: A stored K-major, 1: MN-major
    int b_mn_major;
    int swap_lbo_sbo;  // 0: LBO = K-direction core stride, SBO = MN-direction; 1: swapped
    int version_bit;   // descriptor bits 46-48 = 0b001 ?
    int M, N, K;
};

// element offset (in halves) of (mn, k) inside a tile whose core matrices are ordered
// [mn/8][k/8] (K-major storage) or [k/8][mn/8] (MN-major storage); returns also strides
__host__ __device__ inline int off_kmajor(int mn, int k, int K) {        // core = 8 rows(mn) x 8 k, row 16 B
    return ((mn >> 3) * (K >> 3) + (k >> 3)) * 64 + (mn & 7) * 8 + (k & 7);
}
__host__ __device__ inline int off_mnmajor(int mn, int k, int MN) {      // core = 8 k-rows x 8 mn, row 16 B
    return ((k >> 3) * (MN >> 3) + (mn >> 3)) * 64 + (k & 7) * 8 + (mn & 7);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, int version_bit) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    if (version_bit) d |= (uint64_t)1 << 46;
    return d;   // swizzle mode 0 (bits 61-63), base offset 0
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D, Variant v) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __half* sA = reinterpret_cast<__half*>(smem);
    __half* sB = sA + v.M * v.K;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int M = v.M, N = v.N, K = v.K;

    for (int i = tid; i < M * K; i += 128) {
        const int m = i / K, k = i % K;
        sA[v.a_mn_major ? off_mnmajor(m, k, M) : off_kmajor(m, k, K)] = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        sB[v.b_mn_major ? off_mnmajor(n, k, N) : off_kmajor(n, k, K)] = B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(s32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // generic-proxy smem writes -> visible to the async proxy (tensor core)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;

    if (tid == 0) {
        // instruction descriptor: c_format F32 (1) @4, a/b format F16 (0), a_major @15, b_major @16,
        // n_dim = N>>3 @17, m_dim = M>>4 @24
        uint32_t idesc = (1u << 4) | ((uint32_t)v.a_mn_major << 15) | ((uint32_t)v.b_mn_major << 16) |
                         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        // strides between core matrices (bytes)
        const uint32_t a_kdir = v.a_mn_major ? (uint32_t)(M >> 3) * 128 : 128;
        const uint32_t a_mdir = v.a_mn_major ? 128 : (uint32_t)(K >> 3) * 128;
        const uint32_t b_kdir = v.b_mn_major ? (uint32_t)(N >> 3) * 128 : 128;
        const uint32_t b_ndir = v.b_mn_major ? 128 : (uint32_t)(K >> 3) * 128;
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t a_addr = s32(sA) + ks * 2 * a_kdir;     // 16 k = 2 core matrices along K
            const uint32_t b_addr = s32(sB) + ks * 2 * b_kdir;
            const uint64_t da = v.swap_lbo_sbo ? make_desc(a_addr, a_mdir, a_kdir, v.version_bit)
                                               : make_desc(a_addr, a_kdir, a_mdir, v.version_bit);
            const uint64_t db = v.swap_lbo_sbo ? make_desc(b_addr, b_ndir, b_kdir, v.version_bit)
                                               : make_desc(b_addr, b_kdir, b_ndir, v.version_bit);
            const uint32_t acc = ks > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    }
    // wait for the MMAs
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE_%=;\n\tbra W_%=;\n\tDONE_%=:\n\t}"
        ::"r"(s32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w reads TMEM lanes [32w, 32w+32), 16 columns at a time
    const int row = tid;     // M = 128 rows
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < M)
            for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

int run(Variant v, const char* name) {
    const int M = v.M, N = v.N, K = v.K;
    std::vector<__half> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K), ref(M * N), out(M * N);
    srand(1234);
    for (int i = 0; i < M * K; ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half(x); fA[i] = __half2float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half(x); fB[i] = __half2float(hB[i]); }
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k];
            ref[m * N + n] = (float)s;
        }
    __half *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, M * K * 2)); CK(cudaMalloc(&dB, N * K * 2)); CK(cudaMalloc(&dD, M * N * 4));
    CK(cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, M * N * 4));
    const size_t smem = (size_t)(M + N) * K * 2 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, v);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s CUDA ERROR %s\n", name, cudaGetErrorString(e)); return 2; }
    CK(cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs((double)out[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
    printf("%-28s a_mn=%d b_mn=%d swap=%d ver=%d M=%d N=%d K=%d : max err %.4g (max ref %.3g) %s\n", name, v.a_mn_major,
           v.b_mn_major, v.swap_lbo_sbo, v.version_bit, M, N, K, maxerr, maxref, maxerr < 1e-2 * maxref ? "MATCH" : "mismatch");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return maxerr < 1e-2 * maxref ? 0 : 1;
}

int main(int argc, char** argv) {
    int which = argc > 1 ? atoi(argv[1]) : -1;
    int id = 0;
    for (int ver = 1; ver >= 0; --ver)
        for (int swap = 0; swap < 2; ++swap)
            for (int amn = 0; amn < 2; ++amn)
                for (int bmn = 0; bmn < 2; ++bmn) {
                    struct { int M, N, K; } shapes[2] = {{128, 96, 208}, {128, 208, 96}};
                    for (int s = 0; s < 2; ++s, ++id) {
                        if (which >= 0 && which != id) continue;
                        Variant v{amn, bmn, swap, ver, shapes[s].M, shapes[s].N, shapes[s].K};
                        if ((amn && v.M % 8) || (bmn && v.N % 8) || v.K % 16) continue;
                        char name[64];
                        snprintf(name, sizeof name, "variant %d", id);
                        int rc = run(v, name);
                        if (rc == 2) { printf("context poisoned, stop (rerun with a single id)\n"); return 0; }
                    }
                }
    return 0;
}

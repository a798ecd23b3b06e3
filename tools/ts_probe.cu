// Bring-up probe + micro-benchmark for tcgen05.mma with the A operand in TENSOR MEMORY (".ts" form):
//   D[128 x N] (fp32, TMEM) = A[128 x K] (fp16, TMEM, written by tcgen05.st, lane = row) * B[N x K]^T (fp16, smem image)
// Part 1 checks the A layout assumption (column c of lane m holds elements K = 2c (low half) and 2c + 1 (high half))
// against a host reference; part 2 measures cycles per MMA (N = 96, K = 16) next to the smem-descriptor form.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o tools/ts_probe tools/ts_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../vln_ver_b200/csrc/tcgen05.cuh"
void ver_set_error(const char*, ...) {}
std::atomic<int64_t> g_ver_launches{0};

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(db),
        "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int M = 128, N = 96, K = 208;
constexpr int A_COL = 256;     // TMEM column of the A operand (accumulator at column 0)

// swap_halves = 0: low half = even k
__global__ void __launch_bounds__(128, 1) probe(const __half* __restrict__ A, const __half* __restrict__ B,
                                                float* __restrict__ D, int swap_halves) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __half* sB = reinterpret_cast<__half*>(smem);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        sB[img_off(n, k, K / 8)] = B[i];
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    // my row of A -> TMEM lane tid, 8 columns (16 halves) per store
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16) + A_COL;
    for (int c = 0; c < K / 16; ++c) {
        uint32_t r[8];
        for (int j = 0; j < 8; ++j) {
            const uint16_t lo = __half_as_ushort(A[tid * K + c * 16 + 2 * j]);
            const uint16_t hi = __half_as_ushort(A[tid * K + c * 16 + 2 * j + 1]);
            r[j] = swap_halves ? ((uint32_t)lo << 16 | hi) : ((uint32_t)hi << 16 | lo);
        }
        tmem_st8(lane_base + c * 8, r);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        constexpr uint32_t idesc = umma_idesc(128, N, 0, 0);
        for (int ks = 0; ks < K / 16; ++ks)
            umma_f16_ts(tmem, tmem + A_COL + ks * 8, umma_desc(smem_u32(sB) + ks * 256, 128, (K / 8) * 128), idesc, ks > 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 16; ++j) D[tid * N + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// cycles per MMA: mode 0 = A from smem (descriptor), 1 = A from TMEM; nmma MMAs per commit
__global__ void __launch_bounds__(160, 1) bench(int mode, int iters, int nmma, unsigned long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (65536 + N * 512) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(&s_tmem, 512);
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (warp < 4) {            // zero the A columns
        uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 13; ++c) tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + A_COL + c * 8, z);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 128) {
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 65536);
        constexpr uint32_t idesc = umma_idesc(128, N, 0, 0);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int ks = 0; ks < nmma; ++ks) {
                if (mode == 0)
                    umma_f16(tmem, umma_desc(a_addr + ks * 256, 128, 26 * 128), umma_desc(b_addr + ks * 256, 128, 26 * 128),
                             idesc, ks > 0);
                else
                    umma_f16_ts(tmem, tmem + A_COL + ks * 8, umma_desc(b_addr + ks * 256, 128, 26 * 128), idesc, ks > 0);
            }
            umma_commit(&bar);
            mbar_wait(&bar, it & 1);
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 512);
}

int main() {
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
    __half *dA, *dB;
    float* dD;
    cudaMalloc(&dA, M * K * 2);
    cudaMalloc(&dB, N * K * 2);
    cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
    const int smem = N * K * 2 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int swap = 0; swap < 2; ++swap) {
        cudaMemset(dD, 0xff, M * N * 4);
        probe<<<1, 128, smem>>>(dA, dB, dD, swap);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("ts probe swap=%d: CUDA ERROR %s\n", swap, cudaGetErrorString(e)); return 0; }
        cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs((double)out[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
        printf("ts probe swap_halves=%d: max err %.4g (max ref %.3g) %s\n", swap, maxerr, maxref,
               maxerr < 1e-2 * maxref ? "MATCH" : "mismatch");
    }
    unsigned long long* dout;
    cudaMalloc(&dout, 16);
    const int bsmem = 65536 + N * 512;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, bsmem);
    for (int mode = 0; mode < 2; ++mode)
        for (int nmma : {13, 4, 1}) {
            cudaMemset(dout, 0, 16);
            const int iters = 200;
            bench<<<148, 160, bsmem>>>(mode, iters, nmma, dout);
            cudaError_t e = cudaDeviceSynchronize();
            unsigned long long h = 0;
            cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
            printf("N=96 K=16 A from %s, %2d MMAs per commit: %7.1f cycles/MMA (%s)\n", mode ? "TMEM" : "smem", nmma,
                   (double)h / (iters * nmma), cudaGetErrorString(e));
        }
    return 0;
}

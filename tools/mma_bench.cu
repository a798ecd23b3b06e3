// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, fp16) for the operand layouts / shapes the samplers use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o tools/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cstdlib>
#include "../vln_ver_b200/csrc/tcgen05.cuh"
void ver_set_error(const char*, ...) {}
std::atomic<int64_t> g_ver_launches{0};

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                      // LBO (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;            // SBO = 8 rows x 128 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
    return d;
}

// mode 0: no swizzle, 1: swizzle 128B; hammer: 8 extra warps doing LDS/STS on a disjoint smem region
template <int N>
__global__ void __launch_bounds__(288, 1) bench(int mode, int hammer, int iters, int nmma, unsigned long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    __shared__ int s_stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (65536 + N * 512) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        s_stop = 0;
    }
    if (warp == 8) tmem_alloc(&s_tmem, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 65536);
    if (warp == 8) {
        if ((tid & 31) == 0) {
            const uint32_t idesc = umma_idesc(128, N, 0, 0);
            const int G = 26;
            proxy_fence();
            long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                for (int ks = 0; ks < nmma; ++ks) {
                    if (mode == 0)
                        umma_f16(tmem, umma_desc(a_addr + ks * 256, 128, G * 128), umma_desc(b_addr + ks * 256, 128, G * 128),
                                 idesc, ks > 0);
                    else
                        umma_f16(tmem, desc_sw128(a_addr + (ks >> 2) * 16384 + (ks & 3) * 32),
                                 desc_sw128(b_addr + (ks >> 2) * N * 128 + (ks & 3) * 32), idesc, ks > 0);
                }
                umma_commit(&bar);
                mbar_wait(&bar, it & 1);
            }
            long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
            *(volatile int*)&s_stop = 1;
        }
    } else if (hammer) {
        // LDS.U16 + STS.U16 read-modify-write stream on a region behind the operands
        const uint32_t base = smem_u32(smem + 65536 + N * 512) + ((tid >> 3) & 15) * 3328 + (tid & 7) * 16;
        uint32_t x = tid * 2654435761u;
        unsigned long long n = 0;
        while (!*(volatile int*)&s_stop) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                x = x * 1664525u + 1013904223u;
                const uint32_t k = (x >> 8) % 208;
                const uint32_t a = base + (k >> 3) * 128 + (k & 7) * 2;
                uint16_t v;
                asm volatile("ld.shared.b16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
                v += 1;
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(v) : "memory");
            }
            n += 8;
        }
        if (blockIdx.x == 0 && tid == 0) out[1] = n;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 256);
}

template <int N>
void run(int mode, int hammer, int nmma) {
    unsigned long long* out;
    cudaMalloc(&out, 16);
    cudaMemset(out, 0, 16);
    const int smem = 65536 + N * 512 + 16 * 3328;
    cudaFuncSetAttribute(bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 200;
    bench<N><<<148, 288, smem>>>(mode, hammer, iters, nmma, out);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    unsigned long long h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d %s hammer=%d nmma/batch=%2d: %7.1f cycles/MMA (%s)  rmw/thread=%llu\n", N, mode ? "sw128 " : "noswz", hammer,
           nmma, (double)h[0] / (iters * nmma), cudaGetErrorString(e), h[1]);
    cudaFree(out);
}

int main() {
    for (int hammer = 0; hammer < 2; ++hammer) {
        run<96>(0, hammer, 13);
        run<96>(1, hammer, 13);
        run<192>(0, hammer, 13);
        run<192>(1, hammer, 13);
        run<96>(0, hammer, 4);
        run<96>(0, hammer, 1);
    }
    return 0;
}

// Shared device/host helpers for libver_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/ver_b200.h"

// ---------------------------------------------------------------- host: errors
void ver_set_error(const char* fmt, ...);

#define VER_CHECK_ARG(cond, ...)                         \
    do {                                                 \
        if (!(cond)) {                                   \
            ver_set_error(__VA_ARGS__);                  \
            return VER_ERR_INVALID_ARG;                  \
        }                                                \
    } while (0)

#define VER_CHECK_CUDA(expr)                                                      \
    do {                                                                          \
        cudaError_t e__ = (expr);                                                 \
        if (e__ != cudaSuccess) {                                                 \
            ver_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                          __FILE__, __LINE__);                                    \
            return VER_ERR_CUDA;                                                  \
        }                                                                         \
    } while (0)

#define VER_CHECK_LAUNCH() VER_CHECK_CUDA(cudaGetLastError())

extern std::atomic<int64_t> g_ver_launches;
// tensor-core (tcgen05) sampler, sca_tc.cu
int ver_tc_supported(int Ncam, int S, int Dh, int NP);
int ver_sca_forward_tc(const void* vimg, const float* logits, int ld, const float* rpc, const uint32_t* vis_bits,
                       void* slots, int B, int Ncam, int Z, int H, int W, int Sh, int Sw, int NH, int Dh, int NP,
                       cudaStream_t st);
int ver_sca_backward_tc(const void* vimg, const float* logits, int ld, const float* rpc, const uint32_t* vis_bits,
                        const int32_t* counts, const int32_t* index, const void* gslots, float* gvalue,
                        float* glogits, int B, int Ncam, int Nq, int Sh, int Sw, int NH, int Dh, int NP,
                        cudaStream_t st);
// second-generation tensor-core backward, sca_bwd_tc2.cu
int ver_bwd_tc2_supported(int Ncam, int Sh, int Sw, int Dh, int NP, int ld);
int ver_sca_backward_tc2(const void* vimg, const float* logits, int ld, const float* rpc, const uint32_t* vis_bits,
                         const int32_t* counts, const int32_t* index, const void* gslots, float* gvalue,
                         float* glogits, int B, int Ncam, int Nq, int Sh, int Sw, int NH, int Dh, int NP,
                         cudaStream_t st);
// TMEM-operand sorted-row forward, sca_tc4.cu
int ver_tc4_supported(int Ncam, int S, int Dh, int NP);
int ver_sca_forward_tc4(const void* vimg, const float* logits, int ld, const float* rpc, const int32_t* order,
                        const uint32_t* smask, const uint32_t* tile_union, void* slots, int B, int Ncam, int Nq,
                        int Sh, int Sw, int NH, int Dh, int NP, cudaStream_t st);
// two-threads-per-row builder in front of the TMEM A operand, 16-cell image rows, sca_tc7.cu
int ver_tc7_supported(int Ncam, int Sh, int Sw, int Dh, int NP);
int ver_sca_forward_tc7(const void* vimg16, const float* logits, int ld_logits, const float* rpc, const int32_t* order,
                        const uint32_t* smask, const uint32_t* tile_union, void* slots, int B, int Ncam, int Nq, int Sh,
                        int Sw, int NH, int Dh, int NP, cudaStream_t st);
// three-operand software-pipelined sorted-row forward, sca_tc5.cu
int ver_tc5_supported(int Ncam, int S, int Dh, int NP);
int ver_sca_forward_tc5(const void* vimg, const float* logits, int ld, const float* rpc, const int32_t* order,
                        const uint32_t* smask, const uint32_t* tile_union, void* slots, int B, int Ncam, int Nq,
                        int Sh, int Sw, int NH, int Dh, int NP, cudaStream_t st);
int ver_device_sm_count();
int ver_device_max_smem_optin();

// ---------------------------------------------------------------- device helpers
#define VER_FULL_MASK 0xffffffffu

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// mbarrier + bulk async copy (TMA engine, 1-D form: SASS UBLKCP)
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-B aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// fp32 accumulate of an fp16 x fp16 product (exact product), SASS FHFMA (sm_100a)
__device__ __forceinline__ float fhfma(uint16_t a, uint16_t b, float c) {
    float r;
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r) : "h"(a), "h"(b), "f"(c));
    return r;
}

template <typename T>
struct ElemTraits;
template <>
struct ElemTraits<float> {
    static constexpr int kDtype = VER_F32;
};
template <>
struct ElemTraits<__half> {
    static constexpr int kDtype = VER_F16;
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
__device__ __forceinline__ void from_f32(float& d, float v) { d = v; }
__device__ __forceinline__ void from_f32(__half& d, float v) { d = __float2half_rn(v); }

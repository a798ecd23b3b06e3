// Counter-based dropout masks shared by the row kernels (fused_norm.cu) and the GEMM epilogue (gemm_tc.cu): the
// mask of element e of a tensor depends only on (seed, e), so forward and backward -- and a fused and an unfused
// implementation of the same step -- regenerate identical masks without storing them.
#pragma once
#include "common.cuh"

namespace {

// ---------------------------------------------------------------- Philox4x32-7
// (7 rounds: the smallest round count of Philox4x32 that passes BigCrush (Salmon et al., SC'11); 10 is the library
// default margin.  The masks only have to be unbiased and reproducible between forward and backward.)
__device__ __forceinline__ uint4 philox4x32(uint64_t ctr, uint64_t seed) {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// keep-mask for 8 consecutive elements starting at flat element index e0 (e0 % 8 == 0):
// one Philox call yields 128 random bits = 8 x 16-bit uniforms
__device__ __forceinline__ uint32_t keep8(uint64_t e0, uint64_t seed, uint32_t thr16) {
    const uint4 r = philox4x32(e0 >> 3, seed);
    uint32_t m = 0;
    m |= ((r.x & 0xffff) >= thr16) << 0;
    m |= ((r.x >> 16) >= thr16) << 1;
    m |= ((r.y & 0xffff) >= thr16) << 2;
    m |= ((r.y >> 16) >= thr16) << 3;
    m |= ((r.z & 0xffff) >= thr16) << 4;
    m |= ((r.z >> 16) >= thr16) << 5;
    m |= ((r.w & 0xffff) >= thr16) << 6;
    m |= ((r.w >> 16) >= thr16) << 7;
    return m;
}

}  // namespace

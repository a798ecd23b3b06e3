// Dense projections of the path as hand-written tcgen05 GEMMs with fused epilogues:
//
//     out[M, N] = epilogue( A[M, K] (fp16, row-major)  x  W[N, K]^T (fp16, row-major = nn.Linear.weight)  + bias[N] )
//
// i.e. nn.Linear as the reference uses it for value_proj (M/spatial_cross_attention.py:336), sampling_offsets (+)
// attention_weights (:340-343), output_proj (:174) and the FFN (M/custom_base_transformer_layer.py:157-158,
// vocc.py:134-135).  Epilogues: bias -> fp16 | bias -> fp32 (the offset / weight logits stay fp32) |
// bias + ReLU + dropout -> fp16 (mmcv FFN's Linear -> ReLU -> Dropout: the separate ver_relu_dropout_fwd pass over the
// (rows, 1536) tensor disappears; the mask is the same counter-based Philox mask, philox.cuh).
//
// Structure (one persistent CTA per SM, warp specialised):
//   warp 0      TMA producer: 128 x 64 tile of A and BN x 64 tile of W per k-block (cp.async.bulk.tensor.2d, 128-byte
//               swizzle, mbarrier complete_tx), STAGES-deep ring (full / empty barriers)
//   warp 1      MMA issuer (one lane): 4 x tcgen05.mma (M = 128, N = BN, K = 16) per k-block from the swizzled tiles
//               (K-major descriptors, SBO = 1024 B, +32 B per K step inside the swizzle atom), fp32 accumulator in
//               TMEM; tcgen05.commit frees the smem stage / publishes the accumulator
//   warps 2-9   epilogue (two warps per TMEM lane quarter, each half of the columns): tcgen05.ld (lane = row, 32
//               columns at a time) -> bias / activation / dropout -> a padded staging tile in shared memory -> whole
//               rows as full 128-byte lines (the first version stored 16 bytes per lane at a 3 KB lane stride: 5.3 us
//               of the 8.7 us per tile, r02f); the two accumulator stages (2 x BN TMEM columns) let the epilogue of
//               tile i run under the main loop of tile i + 1
// Tiles are walked m-block major / n-block minor, so the CTAs resident at any time share a few A row blocks (L2).
// Roofline: tensor pipe.  M = 128, N = 256, K = 16 is 128 cycles on the tensor pipe (B300_MICROARCH tcgen05 floor);
// a k-block moves 48 KB of operands through shared memory in those 512 cycles (96 B / clk of the 128 B / clk port).
#include <cuda.h>

#include "philox.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int kGM = 128, kGK = 64, kGStages = 3;
constexpr int kGEpiWarps = 8;                    // two per TMEM lane quarter: each takes half of the tile's columns
constexpr int kGThreads = 64 + 32 * kGEpiWarps;

enum GemmEpilogue { kEpiBiasF16 = 0, kEpiBiasF32 = 1, kEpiBiasReluDropoutF16 = 2, kEpiMaskColsumF16 = 3 };

template <int BN, int ESIZE>
struct GemmSmem {
    static constexpr int a_bytes = kGM * kGK * 2, b_bytes = BN * kGK * 2, stage_bytes = a_bytes + b_bytes;
    // output staging tile [128 rows][BN] of ESIZE-byte elements, rows padded by 16 bytes: a thread writes its row in
    // 16-byte pieces (pitch = 4 words mod 32 banks: the 8 lanes of a quarter-warp phase cover all 32 banks), then whole
    // rows leave as full 128-byte lines
    static constexpr int c_pitch = BN * ESIZE + 16, c_bytes = kGM * c_pitch;
    static constexpr int off_c = kGStages * stage_bytes;
    static constexpr int off_bias = off_c + c_bytes;                         // BN floats
    static constexpr int total = off_bias + BN * 4 + 1024;                  // + slack for the 1024-byte alignment
};

// K-major operand tile with 128-byte swizzle: rows of 64 fp16 (128 B), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t gemm_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                     // leading byte offset (unused for swizzled K-major): 1
    d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                     // sm_100 descriptor version
    d |= (uint64_t)2 << 61;                     // layout type: SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// bounded wait: a protocol error traps (the launch fails) instead of hanging the GPU
__device__ __forceinline__ void gemm_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t tries = 0;; ++tries) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(10000u)
            : "memory");
        if (ok) return;
        if (tries > 400000u) __trap();
    }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const float* __restrict__ bias, void* __restrict__ out, int M, int N, int K, int ldo, float p_drop,
               uint64_t seed, const unsigned long long* __restrict__ seed_epoch) {
    using S = GemmSmem<BN, (EPI == kEpiBiasF32 ? 4 : 2)>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[kGStages], bar_empty[kGStages], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < kGStages; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_tfull[i], 1);
            mbar_init(&bar_tempty[i], kGEpiWarps);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int n_blocks = N / BN, m_blocks = (M + kGM - 1) / kGM, tiles = n_blocks * m_blocks, kblocks = K / kGK;

    if (warp == 0) {
        // ================================================================ TMA producer
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int mb = t / n_blocks, nb = t % n_blocks;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t st = it % kGStages, use = it / kGStages;
                    if (use) gemm_wait(&bar_empty[st], (use - 1) & 1);
                    unsigned char* sa = smem + st * S::stage_bytes;
                    mbar_expect_tx(&bar_full[st], S::stage_bytes);
                    tma_load_2d(sa, &map_a, kb * kGK, mb * kGM, &bar_full[st]);
                    tma_load_2d(sa + S::a_bytes, &map_b, kb * kGK, nb * BN, &bar_full[st]);
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(kGM, BN, 0, 0);
            uint32_t it = 0, tile_i = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++tile_i) {
                const uint32_t as = tile_i & 1, ause = tile_i >> 1;
                if (ause) gemm_wait(&bar_tempty[as], (ause - 1) & 1);
                tc_fence_after();
                const uint32_t d_addr = tmem + as * BN;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t st = it % kGStages, use = it / kGStages;
                    gemm_wait(&bar_full[st], use & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + st * S::stage_bytes);
                    const uint64_t da = gemm_desc_sw128(sa), db = gemm_desc_sw128(sa + S::a_bytes);
#pragma unroll
                    for (int k = 0; k < kGK / 16; ++k)      // +32 bytes per K step inside the 128-byte swizzle atom
                        umma_f16(d_addr, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(&bar_empty[st]);
                }
                umma_commit(&bar_tfull[as]);
            }
        }
    } else {
        // ================================================================ epilogue
        constexpr int ESIZE = (EPI == kEpiBiasF32) ? 4 : 2;
        const int q = warp & 3, row_in_tile = q * 32 + lane, ew = warp - 2;
        constexpr int kColsPerWarp = BN / (kGEpiWarps / 4);
        const int col_lo = (ew >> 2) * kColsPerWarp;                   // my share of the tile's columns
        const uint32_t thr16 = (uint32_t)(p_drop * 65536.f);
        const float scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
        if (EPI == kEpiBiasReluDropoutF16 && seed_epoch) seed += *seed_epoch;
        float* s_bias = reinterpret_cast<float*>(smem + S::off_bias);
        unsigned char* s_c = smem + S::off_c;
        uint32_t tile_i = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++tile_i) {
            const int mb = t / n_blocks, nb = t % n_blocks;
            const uint32_t as = tile_i & 1, ause = tile_i >> 1;
            // bias of this n-block -> shared (the staging tile and s_bias are free: the previous tile ended on a barrier)
            for (int i = tid - 64; i < BN; i += 32 * kGEpiWarps) s_bias[i] = bias ? __ldg(bias + nb * BN + i) : 0.f;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kGEpiWarps) : "memory");
            gemm_wait(&bar_tfull[as], ause & 1);
            tc_fence_after();
            const int row = mb * kGM + row_in_tile;
            const uint32_t taddr = tmem + as * BN + ((uint32_t)(q * 32) << 16);
            // ---- phase 1: accumulator -> registers -> epilogue math -> my row of the staging tile
#pragma unroll 1
            for (int c0 = col_lo; c0 < col_lo + kColsPerWarp; c0 += 32) {
                float v[32];
                tmem_ld32(taddr + c0, v);
                if (c0 + 32 >= col_lo + kColsPerWarp) {       // my last read of this accumulator stage: hand it back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_tempty[as]);
                }
                const int col = nb * BN + c0;
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += s_bias[c0 + i];
                unsigned char* dst = s_c + row_in_tile * S::c_pitch + c0 * ESIZE;
                if (EPI == kEpiBiasF32) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                } else {
                    if (EPI == kEpiBiasReluDropoutF16) {
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            const uint32_t m = p_drop > 0.f ? keep8((uint64_t)row * N + col + g8 * 8, seed, thr16) : 0xffu;
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float x = v[g8 * 8 + e];
                                v[g8 * 8 + e] = (((m >> e) & 1u) && x > 0.f) ? x * scale : 0.f;
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 u;
                        __half2 h;
                        h = __floats2half2_rn(v[8 * i], v[8 * i + 1]);
                        u.x = *reinterpret_cast<const uint32_t*>(&h);
                        h = __floats2half2_rn(v[8 * i + 2], v[8 * i + 3]);
                        u.y = *reinterpret_cast<const uint32_t*>(&h);
                        h = __floats2half2_rn(v[8 * i + 4], v[8 * i + 5]);
                        u.z = *reinterpret_cast<const uint32_t*>(&h);
                        h = __floats2half2_rn(v[8 * i + 6], v[8 * i + 7]);
                        u.w = *reinterpret_cast<const uint32_t*>(&h);
                        reinterpret_cast<uint4*>(dst)[i] = u;
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kGEpiWarps) : "memory");
            // ---- phase 2: whole rows out, 16 bytes per lane: full 128-byte lines
            constexpr int kRowBytes = BN * ESIZE;
            unsigned char* out_b = reinterpret_cast<unsigned char*>(out);
            for (int r = ew; r < kGM; r += kGEpiWarps) {
                const int grow = mb * kGM + r;
                if (grow >= M) break;
                unsigned char* gdst = out_b + ((size_t)grow * ldo + (size_t)nb * BN) * ESIZE;
                const unsigned char* src = s_c + r * S::c_pitch;
#pragma unroll
                for (int o = 0; o < kRowBytes; o += 512)
                    if (o + lane * 16 < kRowBytes)
                        *reinterpret_cast<uint4*>(gdst + o + lane * 16) = *reinterpret_cast<const uint4*>(src + o + lane * 16);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kGEpiWarps) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// ================================================================================================================
// Two-CTA form (tcgen05 cta_group::2): a cluster of two CTAs owns a 256 x 256 output tile.  CTA r holds rows
// [128 r, 128 r + 128) of the A tile and rows [128 r, 128 r + 128) of the W tile (half of the N = 256 columns), the
// leader CTA issues ONE tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16) for the pair, and each CTA's tensor
// memory receives its 128 accumulator rows.  Per SM and k-block that is 32 KB of operands written and read instead
// of 48 KB: the single-CTA kernel above is shared-memory-bandwidth bound at ~70 % of the tensor peak
// (profiles/r02f_gemm_check.txt), this form is what lifts that bound.
//   * both CTAs' TMA loads complete on the LEADER's full barrier (shared::cluster address with the CTA-rank bit
//     cleared), which expects the bytes of both;
//   * tcgen05.commit ... multicast::cluster arrives on the empty / accumulator-full barriers of BOTH CTAs;
//   * the epilogue warps of both CTAs arrive on the LEADER's accumulator-empty barrier (remote mbarrier.arrive).
constexpr int kG2Stages = 4, kG2N = 256;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;        // clears the CTA-rank bit of a shared::cluster address (rank 0 = leader)

struct Gemm2Smem {
    static constexpr int a_bytes = kGM * kGK * 2, b_bytes = (kG2N / 2) * kGK * 2, stage_bytes = a_bytes + b_bytes;
    // output staging: two buffers of one column half (128 rows x 128 columns fp16 = two 128-byte-swizzled TMA boxes)
    static constexpr int half_bytes = 2 * kGM * 128, c_bytes = 2 * half_bytes;
    static constexpr int off_c = kG2Stages * stage_bytes, off_bias = off_c + c_bytes;
    static constexpr int off_cs = off_bias + kG2N * 4;                       // [2][4 row groups][128] fp32 column sums
    static constexpr int total = off_cs + 2 * 4 * 128 * 4 + 1024;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db),
        "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                     "r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGThreads, 1)
gemm2_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_h,
                const float* __restrict__ bias, int M, int N, int K, int ldo, float p_drop,
                uint64_t seed, const unsigned long long* __restrict__ seed_epoch, const __half* __restrict__ aux,
                float* __restrict__ colsum_part) {
    static_assert(EPI != kEpiBiasF32, "the two-CTA kernel writes fp16");
    using S = Gemm2Smem;
    constexpr int BN = kG2N;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[kG2Stages], bar_empty[kG2Stages], bar_tfull[2], bar_tempty[2], bar_h[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    if (tid == 0) {
        for (int i = 0; i < kG2Stages; ++i) {
            mbar_init(&bar_full[i], 1);                     // leader's: one arrive.expect_tx (+ the bytes of both CTAs)
            mbar_init(&bar_empty[i], 1);                    // multicast commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_tfull[i], 1);                    // multicast commit
            mbar_init(&bar_tempty[i], 2 * kGEpiWarps);      // leader's: the epilogue warps of both CTAs
            mbar_init(&bar_h[i], 1);                        // mask operand of a half landed in staging buffer i
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                     // the peer's barriers are initialised too
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int n_blocks = N / BN, m_blocks = (M + 2 * kGM - 1) / (2 * kGM), tiles = n_blocks * m_blocks, kblocks = K / kGK;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0) {
        // ================================================================ TMA producer (both CTAs)
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = cluster_id; t < tiles; t += n_clusters) {
                const int mb = t / n_blocks, nb = t % n_blocks;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t st = it % kG2Stages, use = it / kG2Stages;
                    if (use) gemm_wait(&bar_empty[st], (use - 1) & 1);
                    unsigned char* sa = smem + st * S::stage_bytes;
                    if (rank == 0) mbar_expect_tx(&bar_full[st], 2 * S::stage_bytes);
                    tma_load_2d_2sm(sa, &map_a, kb * kGK, mb * 2 * kGM + (int)rank * kGM, &bar_full[st]);
                    tma_load_2d_2sm(sa + S::a_bytes, &map_b, kb * kGK, nb * BN + (int)rank * (BN / 2), &bar_full[st]);
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer (leader CTA, one lane)
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc(2 * kGM, BN, 0, 0);
            uint32_t it = 0, tile_i = 0;
            for (int t = cluster_id; t < tiles; t += n_clusters, ++tile_i) {
                const uint32_t as = tile_i & 1, ause = tile_i >> 1;
                if (ause) gemm_wait(&bar_tempty[as], (ause - 1) & 1);
                tc_fence_after();
                const uint32_t d_addr = tmem + as * BN;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t st = it % kG2Stages, use = it / kG2Stages;
                    gemm_wait(&bar_full[st], use & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + st * S::stage_bytes);
                    const uint64_t da = gemm_desc_sw128(sa), db = gemm_desc_sw128(sa + S::a_bytes);
#pragma unroll
                    for (int k = 0; k < kGK / 16; ++k)
                        umma_f16_2sm(d_addr, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit_2sm(&bar_empty[st]);
                }
                umma_commit_2sm(&bar_tfull[as]);
            }
        }
    } else {
        // ================================================================ epilogue (both CTAs, own 128 rows)
        // The tile leaves in two column halves of 128: a half is staged in shared memory as two TMA boxes
        // (128 rows x 64 columns, 128-byte swizzle: a thread writing its row's 16-byte pieces is conflict free) and
        // stored by ONE asynchronous TMA store per box; two staging buffers, so the store of a half runs under the
        // epilogue math of the next one (the first version copied rows out with the epilogue warps themselves and was
        // epilogue bound: profiles/r02h_gemm_check.txt).  Warp (quarter q, box x): rows 32 q .., box x of the half.
        const int q = warp & 3, row_in_tile = q * 32 + lane, ew = warp - 2, box = ew >> 2;
        const uint32_t thr16 = (uint32_t)(p_drop * 65536.f);
        const float scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
        if (EPI == kEpiBiasReluDropoutF16 && seed_epoch) seed += *seed_epoch;
        float* s_bias = reinterpret_cast<float*>(smem + S::off_bias);
        float* s_cs = reinterpret_cast<float*>(smem + S::off_cs);              // [2 (half parity)][4 row groups][128]
        const bool issuer = tid == 64;
        // (mask epilogue) the saved activation's half tile is TMA-loaded into the staging buffer the half will be
        // written to, one half ahead: coalesced, asynchronous, and read back at the same swizzled positions
        auto load_h = [&](int tt, int half, uint32_t buf) {
            const int mb2 = tt / n_blocks, nb2 = tt % n_blocks;
            unsigned char* dstb = smem + S::off_c + buf * S::half_bytes;
            mbar_expect_tx(&bar_h[buf], S::half_bytes);
#pragma unroll
            for (int x = 0; x < 2; ++x)
                tma_load_2d(dstb + x * (kGM * 128), &map_h, nb2 * BN + half * 128 + x * 64, mb2 * 2 * kGM + (int)rank * kGM,
                            &bar_h[buf]);
        };
        if (EPI == kEpiMaskColsumF16 && issuer && cluster_id < tiles) load_h(cluster_id, 0, 0);
        uint32_t tile_i = 0, half_i = 0;
        for (int t = cluster_id; t < tiles; t += n_clusters, ++tile_i) {
            const int mb = t / n_blocks, nb = t % n_blocks;
            const uint32_t as = tile_i & 1, ause = tile_i >> 1;
            const int row0 = mb * 2 * kGM + (int)rank * kGM, row = row0 + row_in_tile;
            const uint32_t taddr = tmem + as * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int half = 0; half < 2; ++half, ++half_i) {
                unsigned char* s_h = smem + S::off_c + (half_i & 1) * S::half_bytes;          // staging buffer of this half
                // buffer (half_i & 1) was last stored two halves ago: its TMA store must have finished READING it
                // (mask epilogue: already waited for when the mask operand was loaded into it)
                if (EPI != kEpiMaskColsumF16 && issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                if (half == 0)
                    for (int i = tid - 64; i < BN; i += 32 * kGEpiWarps) s_bias[i] = bias ? __ldg(bias + nb * BN + i) : 0.f;
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kGEpiWarps) : "memory");
                if (half == 0) {
                    gemm_wait(&bar_tfull[as], ause & 1);
                    tc_fence_after();
                }
                unsigned char* s_box = s_h + box * (kGM * 128) + row_in_tile * 128;
#pragma unroll 1
                for (int cc = 0; cc < 64; cc += 32) {
                    const int c0 = half * 128 + box * 64 + cc;                 // column inside the tile
                    float v[32];
                    tmem_ld32(taddr + c0, v);
                    if (half == 1 && cc == 32) {                               // my last read of this accumulator stage
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader(&bar_tempty[as]);
                    }
                    const int col = nb * BN + c0;
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += s_bias[c0 + i];
                    if (EPI == kEpiBiasReluDropoutF16) {
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            const uint32_t m = p_drop > 0.f ? keep8((uint64_t)row * N + col + g8 * 8, seed, thr16) : 0xffu;
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float x = v[g8 * 8 + e];
                                v[g8 * 8 + e] = (((m >> e) & 1u) && x > 0.f) ? x * scale : 0.f;
                            }
                        }
                    }
                    if (EPI == kEpiMaskColsumF16) {
                        // ReLU / dropout backward: the saved activation is > 0 exactly where the unit was kept and
                        // positive; its tile sits in this staging buffer (rows past M: zero filled by TMA)
                        if (cc == 0) gemm_wait(&bar_h[half_i & 1], (half_i >> 1) & 1);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int chunk = (cc >> 3) + i;
                            const uint4 hv = *reinterpret_cast<const uint4*>(s_box + ((chunk ^ (row_in_tile & 7)) << 4));
                            const __half2* hh = reinterpret_cast<const __half2*>(&hv);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __half22float2(hh[e]);
                                v[8 * i + 2 * e] = f.x > 0.f ? v[8 * i + 2 * e] * scale : 0.f;
                                v[8 * i + 2 * e + 1] = f.y > 0.f ? v[8 * i + 2 * e + 1] * scale : 0.f;
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 u;
                        __half2 h;
                        h = __floats2half2_rn(v[8 * i], v[8 * i + 1]);
                        u.x = *reinterpret_cast<const uint32_t*>(&h);
                        h = __floats2half2_rn(v[8 * i + 2], v[8 * i + 3]);
                        u.y = *reinterpret_cast<const uint32_t*>(&h);
                        h = __floats2half2_rn(v[8 * i + 4], v[8 * i + 5]);
                        u.z = *reinterpret_cast<const uint32_t*>(&h);
                        h = __floats2half2_rn(v[8 * i + 6], v[8 * i + 7]);
                        u.w = *reinterpret_cast<const uint32_t*>(&h);
                        const int chunk = (cc >> 3) + i;                        // 16-byte piece 0..7 of my 128-byte row
                        *reinterpret_cast<uint4*>(s_box + ((chunk ^ (row_in_tile & 7)) << 4)) = u;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // staged bytes -> visible to the TMA engine
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kGEpiWarps) : "memory");
                if (issuer) {
                    const int gc = nb * BN + half * 128;
#pragma unroll
                    for (int x = 0; x < 2; ++x)
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_c),
                                     "r"(smem_u32(s_h + x * (kGM * 128))), "r"(gc + x * 64), "r"(row0)
                                     : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    if (EPI == kEpiMaskColsumF16) {
                        // next half's mask operand -> the other buffer, once the store that last used it was read out
                        const int nt = half == 0 ? t : t + n_clusters, nh = half ^ 1;
                        if (nt < tiles) {
                            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            load_h(nt, nh, (half_i + 1) & 1);
                        }
                    }
                }
                if (EPI == kEpiMaskColsumF16) {
                    // column sums of the staged half (rows past M hold zeros): warp = (box ew & 1, row group ew >> 1);
                    // a lane keeps one 16-byte piece (8 columns) and walks 4 rows per step -- conflict free
                    const int cbox = ew & 1, rg = ew >> 1, piece = lane & 7;
                    float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int r = rg * 32 + k * 4 + (lane >> 3);
                        const uint4 u = *reinterpret_cast<const uint4*>(s_h + cbox * (kGM * 128) + r * 128 + ((piece ^ (r & 7)) << 4));
                        const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = __half22float2(hh[e]);
                            cs[2 * e] += f.x;
                            cs[2 * e + 1] += f.y;
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        cs[e] += __shfl_xor_sync(VER_FULL_MASK, cs[e], 8);
                        cs[e] += __shfl_xor_sync(VER_FULL_MASK, cs[e], 16);
                    }
                    float* dst = s_cs + ((half_i & 1) * 4 + rg) * 128 + cbox * 64 + piece * 8;
                    if (lane < 8) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) dst[e] = cs[e];
                    }
                    asm volatile("bar.sync 2, %0;" ::"n"(32 * kGEpiWarps) : "memory");
                    if (tid - 64 < 128) {
                        const int c = tid - 64;
                        const float* src = s_cs + (half_i & 1) * 4 * 128 + c;
                        colsum_part[(size_t)(mb * 2 + (int)rank) * N + nb * BN + half * 128 + c] =
                            src[0] + src[128] + src[256] + src[384];
                    }
                }
            }
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // every store landed before the CTA leaves
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                     // nobody leaves while the peer may still signal its barriers
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---------------------------------------------------------------- host: tensor maps through the driver entry point
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// rows x K fp16 row-major (leading dimension ld elements) -> boxes of box_rows x 64, 128-byte swizzle
int make_map(CUtensorMap* map, const void* ptr, int rows, int K, int ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        ver_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return VER_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kGK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ver_set_error("cuTensorMapEncodeTiled failed (%d) for a %d x %d matrix, ld %d", (int)r, rows, K, ld);
        return VER_ERR_CUDA;
    }
    return VER_OK;
}

template <int BN, int EPI>
int launch_gemm(const void* a, int lda, const void* w, int ldw, const float* bias, void* out, int ldo, int M, int N,
                int K, float p_drop, uint64_t seed, const uint64_t* seed_epoch, cudaStream_t st) {
    CUtensorMap ma, mb;
    int rc = make_map(&ma, a, M, K, lda, kGM);
    if (rc) return rc;
    rc = make_map(&mb, w, N, K, ldw, BN);
    if (rc) return rc;
    auto kern = gemm_tn_kernel<BN, EPI>;
    using S = GemmSmem<BN, (EPI == kEpiBiasF32 ? 4 : 2)>;
    VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total));
    const int tiles = (N / BN) * ((M + kGM - 1) / kGM);
    const int sms = ver_device_sm_count();
    kern<<<tiles < sms ? tiles : sms, kGThreads, S::total, st>>>(ma, mb, bias, out, M, N, K, ldo, p_drop, seed,
                                                                          (const unsigned long long*)seed_epoch);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

template <int EPI>
int launch_gemm2(const void* a, int lda, const void* w, int ldw, const float* bias, void* out, int ldo, int M, int N,
                 int K, float p_drop, uint64_t seed, const uint64_t* seed_epoch, cudaStream_t st,
                 const void* aux = nullptr, float* colsum_part = nullptr) {
    CUtensorMap ma, mb;
    int rc = make_map(&ma, a, M, K, lda, kGM);
    if (rc) return rc;
    rc = make_map(&mb, w, N, K, ldw, kG2N / 2);
    if (rc) return rc;
    CUtensorMap mc;                                     // output [M, N] fp16: 128 x 64 boxes, same swizzle as the staging
    rc = make_map(&mc, out, M, N, ldo, kGM);
    if (rc) return rc;
    CUtensorMap mh = mc;                                // mask operand (same shape as the output) of the backward epilogue
    if (aux) {
        rc = make_map(&mh, aux, M, N, ldo, kGM);
        if (rc) return rc;
    }
    auto kern = gemm2_tn_kernel<EPI>;
    VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Smem::total));
    const int tiles = (N / kG2N) * ((M + 2 * kGM - 1) / (2 * kGM));
    int ctas = ver_device_sm_count() & ~1;
    if (2 * tiles < ctas) ctas = 2 * tiles;
    kern<<<ctas, kGThreads, Gemm2Smem::total, st>>>(ma, mb, mc, mh, bias, M, N, K, ldo, p_drop, seed,
                                                   (const unsigned long long*)seed_epoch, (const __half*)aux, colsum_part);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

// 0 = choose (two-CTA kernel where it applies), 1 = single-CTA kernel, 2 = two-CTA kernel (tests / tools A-B)
int g_gemm_variant = 0;

}  // namespace

extern "C" int ver_debug_gemm_variant(int v) {
    g_gemm_variant = v;
    return VER_OK;
}

/* da = (dy @ w^T) * [h > 0] / (1 - p_drop) and the column sums of da, see include/ver_b200.h */
extern "C" int ver_linear_relu_dropout_bwd_f16(const void* dy, int lddy, const void* w, int ldw, const void* h, void* da,
                                               int ld, float* colsum_part, int M, int N, int K, float p_drop,
                                               ver_stream_t stream) {
    VER_CHECK_ARG(dy && w && h && da && colsum_part, "null pointer");
    VER_CHECK_ARG(M > 0 && K > 0 && K % kGK == 0 && N > 0 && N % kG2N == 0, "needs K %% 64 == 0 and N %% 256 == 0 (got N=%d K=%d)", N, K);
    VER_CHECK_ARG(lddy >= K && ldw >= K && ld >= N && lddy % 8 == 0 && ldw % 8 == 0 && ld % 8 == 0, "leading dimensions");
    VER_CHECK_ARG(((uintptr_t)dy & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)h & 15) == 0 && ((uintptr_t)da & 15) == 0,
                  "operands must be 16-byte aligned");
    VER_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "bad dropout probability");
    return launch_gemm2<kEpiMaskColsumF16>(dy, lddy, w, ldw, nullptr, da, ld, M, N, K, p_drop, 0, nullptr,
                                           (cudaStream_t)stream, h, colsum_part);
}

extern "C" int ver_linear_bwd_colsum_rows(int M) { return 2 * ((M + 2 * kGM - 1) / (2 * kGM)); }

extern "C" int ver_linear_supported(int M, int N, int K) {
    return M > 0 && K > 0 && K % kGK == 0 && N > 0 && (N % 256 == 0 || N % 192 == 0 || N % 128 == 0);
}

/* out = epilogue(a @ w^T + bias), see include/ver_b200.h */
extern "C" int ver_linear_f16(int epilogue, const void* a, int lda, const void* w, int ldw, const float* bias, void* out,
                              int ldo, int M, int N, int K, float p_drop, uint64_t seed, const uint64_t* seed_epoch,
                              ver_stream_t stream) {
    VER_CHECK_ARG(a && w && out, "null pointer");
    VER_CHECK_ARG(epilogue >= 0 && epilogue <= 2, "epilogue must be 0 (bias -> fp16), 1 (bias -> fp32) or 2 (bias + ReLU + dropout -> fp16)");
    VER_CHECK_ARG(ver_linear_supported(M, N, K), "unsupported shape M=%d N=%d K=%d (K %% 64 == 0, N %% 128 == 0 or N %% 192 == 0)", M, N, K);
    VER_CHECK_ARG(lda >= K && ldw >= K && ldo >= N && lda % 8 == 0 && ldw % 8 == 0, "leading dimensions");
    VER_CHECK_ARG(((uintptr_t)a & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
                      ldo % (epilogue == 1 ? 4 : 8) == 0,
                  "operands must be 16-byte aligned");
    VER_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "bad dropout probability");
    cudaStream_t st = (cudaStream_t)stream;
#define GEMM_BN(BN)                                                                                                       \
    (epilogue == 0   ? launch_gemm<BN, kEpiBiasF16>(a, lda, w, ldw, bias, out, ldo, M, N, K, 0.f, 0, nullptr, st)         \
     : epilogue == 1 ? launch_gemm<BN, kEpiBiasF32>(a, lda, w, ldw, bias, out, ldo, M, N, K, 0.f, 0, nullptr, st)         \
                     : launch_gemm<BN, kEpiBiasReluDropoutF16>(a, lda, w, ldw, bias, out, ldo, M, N, K, p_drop, seed,     \
                                                               seed_epoch, st))
    if (N % 256 == 0 && epilogue != 1 && g_gemm_variant != 1) {       // two CTAs per 256 x 256 tile
        if (epilogue == 0)
            return launch_gemm2<kEpiBiasF16>(a, lda, w, ldw, bias, out, ldo, M, N, K, 0.f, 0, nullptr, st);
        return launch_gemm2<kEpiBiasReluDropoutF16>(a, lda, w, ldw, bias, out, ldo, M, N, K, p_drop, seed, seed_epoch, st);
    }
    // (the fp32 staging tile of a 256-column block does not fit next to the operand ring)
    if (N % 256 == 0 && epilogue != 1) return GEMM_BN(256);
    if (N % 192 == 0) return GEMM_BN(192);
    return GEMM_BN(128);
#undef GEMM_BN
}

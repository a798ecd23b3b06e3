// Fused SCA sampler forward, fourth generation: the interpolation matrix A goes to the tensor cores through
// TENSOR MEMORY (tcgen05.mma with the A operand in TMEM), rows are built in thread-private linear scratch rows.
//
//     slots[b, n, h, :] = 1/max(count,1) * sum_{cam sees n, ascending} A_cam[n, :] V_{b,cam,h}[:, :]
//
// Replaces SpatialCrossAttention.forward's rebatch / sampling / scatter-mean
// (M/spatial_cross_attention.py:138-173, MSDeformableAttention3D :340-374).  Same tiling as sca_tc3.cu
// (visibility-sorted 128-row tiles of one panorama, one thread per row, persistent CTAs over
// (panorama, 256-row chunk, head) items); what changed, and why (profiles/r01e: the 8 worker warps were the
// critical path -- 1370 issued instructions per (row-warp, camera), 15 % of their time in the epilogue, 14 %
// waiting for the MMAs that read their shared-memory A image to retire):
//   * a row of A is built in a LINEAR scratch row (element = pixel index): the four corners of a tap are at
//     +0, +2, +2 Sw, +2 Sw + 2 bytes of one base address, out-of-map corners are folded into the weights
//     (clamped base cell, zero weight), so there is no per-corner address arithmetic and no trash cell;
//   * the finished row is copied scratch -> registers -> TMEM (tcgen05.st, lane = row) and the scratch chunk is
//     zeroed by the same loop: the scratch never waits for the tensor cores, only the TMEM copy does;
//   * the epilogue (TMEM -> slots) runs on four extra warps, off the builders' critical path;
//   * the next item's logits, reference points and ids are prefetched with cp.async into a per-thread slot;
//   * bilinear weights are the branch-free tent max(0, 1 - |coordinate - cell|) on a clamped 2x2 cell block.
// Measured (B200, 8 x 18 views, 16x40x40, profiles/r01j, r01k): 417-421 us per launch against 433-437 us for
// sca_fwd_tc3_kernel.  tcgen05.st / tcgen05.ld queue behind the MMAs already issued by the CTA, so a group's copy
// waits for the other group's batch (copy phase 2750 cycles per camera); walking the two groups in lock step
// removes that wait but puts the epilogue on the critical path (465 us) -- kept as independent groups.
//
// Roles: warps 0-3 = group 0 (rows 0..127 of the chunk), warps 4-7 = group 1, warps 8-11 = epilogue (TMEM lane
// quarter = warp % 4, both groups), warps 12 / 13 = control (one lane each): MMA issue of group 0 / 1; warp 12 also issues
// the TMA of the value images.
// TMEM columns: [0, 2 DH) the two accumulators, [2 DH, 2 DH + SP) the two A operands (SP / 2 columns each).
// Hand-offs (mbarriers):
//     bar_built[g]    group g wrote A_g for its next camera into TMEM            (4 arrivals, one per warp)
//     bar_mma[g]      tcgen05.commit: the MMAs reading A_g retired -> A_g may be overwritten
//     bar_full[g]     tcgen05.commit after the item's last camera: accumulator g is complete
//     bar_free[g]     the epilogue warps drained accumulator g                  (4 arrivals)
//     bar_v[buf] / bar_vfree[buf]   value image landed (transaction bytes) / all MMAs reading it retired
#include <type_traits>

#include "sampler.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int kF4Workers = 256;
constexpr int kF4Threads = kF4Workers + 128 + 64;        // 8 worker warps, 4 epilogue warps, 2 control warps
constexpr int kF4Rows = 128;                  // rows per group = UMMA M
constexpr int kF4ChunkRows = 2 * kF4Rows;
// per-worker landing slot of the prefetches (cp.async): 24 fp32 logits of (row, head) | n, camera mask, tile
// union of the item after | reference points of the first 4 cameras that see the row
constexpr int kF4SlotLogits = 0, kF4SlotIds = 96, kF4SlotRefs = 112, kF4SlotBytes = 144;
constexpr int kF4SlotRefCams = 4;

struct F4Smem {
    int v_bytes, warp_scratch, off_v[2], off_scratch, off_slots, total;
    __host__ __device__ F4Smem(int Dh, int SP) {
        v_bytes = Dh * SP * 2;
        // scratch rows of one warp, lane-interleaved: 32-bit word w (cells 2w, 2w + 1) of lane l at (w * 32 + l) * 4
        // -> lane l only ever touches bank l: every scratch access of a warp is conflict free, whatever the taps
        warp_scratch = (SP / 2) * 32 * 4;
        off_v[0] = 0;
        off_v[1] = v_bytes;
        off_scratch = 2 * v_bytes;
        off_slots = off_scratch + (kF4Workers / 32) * warp_scratch;
        total = off_slots + kF4Workers * kF4SlotBytes;
    }
};

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(db),
        "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16b(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ uint16_t lds16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.b16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts16(uint32_t a, uint16_t v) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(v) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

struct F4Item {
    int b, chunk, h;
};
__device__ __forceinline__ F4Item f4_item(int item, int NH, int chunks_per_b) {
    F4Item it;
    it.h = item % NH;
    const int r = item / NH;
    it.chunk = r % chunks_per_b;
    it.b = r / chunks_per_b;
    return it;
}

// phase timers (debug; enabled through ver_debug_tc4_timing, read by tools/tc_timing.py, never by the product)
__device__ unsigned long long g_f4_timing[32];
__device__ int g_f4_timing_on = 0;
struct F4Timer {
    bool on;
    long long t;
    __device__ __forceinline__ F4Timer(bool active) : on(active && g_f4_timing_on), t(0) {
        if (on) t = clock64();
    }
    __device__ __forceinline__ void lap(int slot) {
        if (on) {
            const long long n = clock64();
            atomicAdd(&g_f4_timing[slot], (unsigned long long)(n - t));
            t = n;
        }
    }
};

template <int DH, int NP>
__global__ void __launch_bounds__(kF4Threads, 1)
sca_fwd_tc4_kernel(const __half* __restrict__ vimg, const float* __restrict__ logits, int ld,
                   const float* __restrict__ rpc, const int32_t* __restrict__ order,
                   const uint32_t* __restrict__ smask, const uint32_t* __restrict__ tile_union,
                   __half* __restrict__ slots, int B, int Ncam, int Nq, int Sh, int Sw, int SP, int NH,
                   int chunks_per_b, int n_items) {
    const int G = SP >> 3;                       // 8-pixel groups per row of the V image
    extern __shared__ __align__(128) unsigned char smem[];
    const F4Smem L(DH, SP);
    __shared__ __align__(8) uint64_t bar_built[2], bar_mma[2], bar_full[2], bar_free[2], bar_v[2], bar_vfree[2];
    __shared__ uint32_t s_tmem;
    __shared__ volatile uint32_t s_kmask[2][2][4];     // [group][batch parity][warp]: K chunks (16 pixels) that hold taps

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_built[i], 4);
            mbar_init(&bar_mma[i], 1);
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_free[i], 4);
            mbar_init(&bar_v[i], 1);
            mbar_init(&bar_vfree[i], 2);
        }
        mbar_fence_init();
    }
    if (warp == 12) tmem_alloc(&s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int tiles_per_b = (Nq + kF4Rows - 1) / kF4Rows;      // order.cu's tile_union row length
    const size_t v_elems = (size_t)DH * SP;                    // halves per (view, head) image
    const int nchunks = SP >> 4;
    const uint32_t tm_a0 = tmem + 2 * DH;                      // A operand of group g at + g * (SP / 2) columns

    if (warp >= 12) {
        // ================================================================ control: one issuing thread per row group
        // (warp 12: group 0 and the TMA of the value images, warp 13: group 1).  Both walk the same sequence of
        // (item, camera) steps.  A single thread issuing both groups' batches one after the other delayed each group's
        // MMAs by the other's issue time (round 2, profiles/r03c: the split took the sibling kernel sca_fwd_tc7_kernel
        // from 409 to 389 us).
        // (the group index is a compile-time constant inside: barrier and mask addresses are static, which is also what
        // compute-sanitizer's barrier tracking wants)
        auto control = [&](auto cg_const) {
            constexpr int cg = decltype(cg_const)::value;
            constexpr uint32_t idesc = umma_idesc(128, DH, 0, 0);
            int nx_item = (int)blockIdx.x - (int)gridDim.x;
            uint32_t nx_rest = 0, nx_ug = 0;
            int nx_b = 0, nx_h = 0, nx_cam = 0;
            auto advance = [&]() -> bool {
                while (true) {
                    if (nx_rest) {
                        nx_cam = __ffs(nx_rest) - 1;
                        nx_rest &= nx_rest - 1;
                        return true;
                    }
                    nx_item += gridDim.x;
                    if (nx_item >= n_items) return false;
                    const F4Item it = f4_item(nx_item, NH, chunks_per_b);
                    nx_b = it.b;
                    nx_h = it.h;
                    const uint32_t* tu = tile_union + (size_t)it.b * tiles_per_b + 2 * it.chunk;
                    const uint32_t u0 = tu[0], u1 = (2 * it.chunk + 1 < tiles_per_b) ? tu[1] : 0u;
                    nx_ug = cg ? u1 : u0;
                    nx_rest = u0 | u1;
                }
            };
            auto load_v = [&](int buf) {
                mbar_expect_tx(&bar_v[buf], L.v_bytes);
                bulk_g2s(smem + (buf ? L.off_v[1] : L.off_v[0]), vimg + ((size_t)(nx_b * Ncam + nx_cam) * NH + nx_h) * v_elems,
                         L.v_bytes, &bar_v[buf]);
            };
            F4Timer tc(cg == 0);
            bool has_next = advance();
            if (cg == 0 && has_next) load_v(0);
            uint32_t kk = 0, itg = 0, acc_items = 0;
            const uint64_t dv_0 = umma_desc(smem_u32(smem + L.off_v[0]), 128, G * 128);
            const uint64_t dv_1 = umma_desc(smem_u32(smem + L.off_v[1]), 128, G * 128);
            const uint32_t chunk_mask = (1u << nchunks) - 1u;
            const uint32_t d_addr = tmem + cg * DH, a_addr = tm_a0 + cg * (SP >> 1);
            while (has_next) {
                const int cam = nx_cam;
                const uint32_t ug = nx_ug;
                has_next = advance();                  // nx_* now describe step kk + 1
                const int buf = kk & 1;
                // EVERY step, work or not: a parity wait is only valid for a waiter that observes every phase of the
                // barrier -- a control thread whose group idles through some steps would otherwise run ahead of the
                // value-image loads, pass a later wait early and desynchronise bar_vfree
                mbar_wait_park(&bar_v[buf], (kk >> 1) & 1);
                tc.lap(10);                            // control: wait for the value image
                if ((ug >> cam) & 1u) {
                    const bool first_cam = !(ug & ((1u << cam) - 1u));         // lowest camera of this tile overwrites
                    const bool last_cam = !(ug >> (cam + 1));
                    mbar_wait_park(&bar_built[cg], itg & 1);
                    tc.lap(9);                         // control: wait for a built A
                    if (first_cam && acc_items) mbar_wait_park(&bar_free[cg], (acc_items - 1) & 1);
                    tc_fence_after();
                    tc.lap(13);                        // control: wait for a drained accumulator
                    const int par = itg & 1;
                    uint32_t km = (s_kmask[cg][par][0] | s_kmask[cg][par][1] | s_kmask[cg][par][2] | s_kmask[cg][par][3]) & chunk_mask;
                    if (first_cam && !km) km = 1u;     // (an all-zero chunk zeroes the accumulator)
                    const uint64_t dvb = buf ? dv_1 : dv_0;
                    // straight-line issue (a find-first-set loop that rebuilds the descriptor costs ~150 cycles per MMA on
                    // the uniform datapath; an indexed descriptor array lives in local memory: one LDL per MMA), the
                    // accumulate flag a compile-time constant except on an item's first camera
                    if (first_cam) {
                        const uint32_t first_bit = km & (0u - km);
#pragma unroll
                        for (int ks = 0; ks < 16; ++ks)
                            if ((km >> ks) & 1u)
                                umma_f16_ts(d_addr, a_addr + ks * 8, dvb + (uint64_t)(ks * 16), idesc, (first_bit >> ks) & 1u ? 0u : 1u);
                    } else {
#pragma unroll
                        for (int ks = 0; ks < 16; ++ks)
                            if ((km >> ks) & 1u) umma_f16_ts(d_addr, a_addr + ks * 8, dvb + (uint64_t)(ks * 16), idesc, 1u);
                    }
                    umma_commit(&bar_mma[cg]);
                    if (last_cam) {
                        umma_commit(&bar_full[cg]);
                        ++acc_items;
                    }
                    tc.lap(11);                        // control: MMA issue
                    ++itg;
                }
                // two arrivals per step, one from each control thread: behind my MMAs if I issued any, else a plain arrive
                if ((ug >> cam) & 1u) umma_commit(&bar_vfree[buf]);
                else mbar_arrive(&bar_vfree[buf]);
                if (cg == 0 && has_next) {
                    // value image of step kk + 1 -> the other buffer, once step kk - 1 stopped reading it
                    if (kk >= 1) mbar_wait_park(&bar_vfree[(kk + 1) & 1], ((kk - 1) >> 1) & 1);
                    load_v((kk + 1) & 1);
                    tc.lap(12);                        // control: wait for a free value buffer + TMA issue
                }
                ++kk;
            }
            // drain: the last commits must have arrived before the CTA tears TMEM / smem down
            if (kk >= 1) mbar_wait_park(&bar_vfree[(kk - 1) & 1], ((kk - 1) >> 1) & 1);
        };
        if (lane == 0) {
            if (warp == 12) control(std::integral_constant<int, 0>{});
            else control(std::integral_constant<int, 1>{});
        }
    } else if (warp >= 8) {         // (warps 8-11; the control warps were taken above)
        // ================================================================ epilogue: TMEM -> slots
        const int q = warp & 3, r = q * 32 + lane;              // TMEM lane quarter / row inside a group
        uint32_t full_seen[2] = {0, 0};
        int n_nx[2] = {-1, -1};
        uint32_t m_nx[2] = {0, 0}, u_nx[2] = {0, 0};
        auto load_ids = [&](int item) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                n_nx[g] = -1;
                m_nx[g] = u_nx[g] = 0;
                if (item >= n_items) continue;
                const F4Item it = f4_item(item, NH, chunks_per_b);
                const int tile = 2 * it.chunk + g, i = tile * kF4Rows + r;
                if (tile < tiles_per_b) u_nx[g] = __ldg(tile_union + (size_t)it.b * tiles_per_b + tile);
                if (i < Nq) {
                    n_nx[g] = __ldg(order + (size_t)it.b * Nq + i);
                    m_nx[g] = __ldg(smask + (size_t)it.b * Nq + i);
                }
            }
        };
        F4Timer te(tid == 256);
        load_ids(blockIdx.x);
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const F4Item it = f4_item(item, NH, chunks_per_b);
            const int n[2] = {n_nx[0], n_nx[1]};
            const uint32_t m[2] = {m_nx[0], m_nx[1]}, u[2] = {u_nx[0], u_nx[1]};
            load_ids(item + gridDim.x);                 // in flight during this item's epilogue
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const float inv_cnt = 1.f / (float)max(__popc(m[g]), 1);
                __half* dst = (n[g] >= 0) ? slots + (((size_t)it.b * Nq + n[g]) * NH + it.h) * DH : nullptr;
                if (u[g]) {                             // warp-uniform (tile property)
                    mbar_wait_park(&bar_full[g], full_seen[g] & 1);
                    ++full_seen[g];
                    tc_fence_after();
                    te.lap(16);                         // epilogue: wait for a complete accumulator
                    // 32 columns at a time: load, wait, scale, store (keeps the warp at 32 live accumulator registers)
#pragma unroll
                    for (int c0 = 0; c0 < DH; c0 += 32) {
                        float vv[32];
                        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + g * DH + c0, vv);
                        if (c0 + 32 >= DH) {                       // last read of accumulator g: hand it back
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&bar_free[g]);
                        }
                        if (dst) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) vv[i] *= inv_cnt;
                            store_channels16<32>(dst + c0, vv);
                        }
                    }
                    te.lap(17);                         // epilogue: TMEM -> registers -> slots
                } else if (dst) {                       // no camera sees this tile: zeros
#pragma unroll
                    for (int i = 0; i < DH / 8; ++i) reinterpret_cast<uint4*>(dst)[i] = make_uint4(0, 0, 0, 0);
                }
            }
        }
    } else {
        // ================================================================ workers: one thread = one row
        const int g = warp >> 2, r = (warp & 3) * 32 + lane;
        // my scratch row: cell k at mybase + cell_off(k)
        const uint32_t* my_words = reinterpret_cast<const uint32_t*>(smem + L.off_scratch + (size_t)warp * L.warp_scratch) + lane;
        const uint32_t mybase = smem_u32(smem + L.off_scratch) + (uint32_t)warp * L.warp_scratch + (uint32_t)lane * 4u;
        const uint32_t slot = smem_u32(smem + L.off_slots) + (uint32_t)tid * kF4SlotBytes;
        const uint32_t tm_row = tm_a0 + g * (SP >> 1) + ((uint32_t)((warp & 3) * 32) << 16);    // my lane, A_g
        const float fSw = (float)Sw, fSh = (float)Sh;
        const float pix_bias = 8388608.f - (float)(Sw + 1);
        const uint32_t row_half = (uint32_t)(Sw >> 1) << 7, sw_odd = (uint32_t)Sw & 1u;
        const float2* rp2 = reinterpret_cast<const float2*>(rpc);
        uint32_t it = 0, seen = 0;                // MMA batches handed over / observed retired (this group)
        uint32_t dirty = 0;                       // chunks of my warp's TMEM lanes that may be non-zero
        bool tapped = false;                      // my scratch row holds the taps recorded in ua / ub
        uint32_t ua[8], ub[8];                    // addresses of the upper-left / lower-left cell of each point's tap

        // scratch row and TMEM operand start out all zero
        for (int w = 0; w < SP / 2; ++w) asm volatile("st.shared.b32 [%0], %1;" ::"r"(mybase + w * 128), "r"(0) : "memory");
        {
            const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int c = 0; c < nchunks; ++c) tmem_st8(tm_row + c * 8, z);
            tmem_st_wait();
        }

        // ---- prefetch pipeline: while item j is processed, its successor's logits / reference points and the
        // ids (voxel, camera mask, tile union) of the item after that are in flight into my slot
        auto issue_ids = [&](int item) {           // -> slot ids; zeros when the item / tile / row does not exist
            if (item >= n_items) return;
            const F4Item q = f4_item(item, NH, chunks_per_b);
            const int tile = 2 * q.chunk + g, i = tile * kF4Rows + r;
            if (tile < tiles_per_b) cp_async4(slot + kF4SlotIds + 8, tile_union + (size_t)q.b * tiles_per_b + tile);
            if (i < Nq) {
                cp_async4(slot + kF4SlotIds, order + (size_t)q.b * Nq + i);
                cp_async4(slot + kF4SlotIds + 4, smask + (size_t)q.b * Nq + i);
            }
        };
        auto ids_exist = [&](int item, bool& has_tile, bool& has_row) {
            has_tile = has_row = false;
            if (item >= n_items) return;
            const F4Item q = f4_item(item, NH, chunks_per_b);
            const int tile = 2 * q.chunk + g;
            has_tile = tile < tiles_per_b;
            has_row = tile * kF4Rows + r < Nq;
        };
        auto issue_row = [&](int item, int n, uint32_t m) {      // logits + first reference points of `item`
            if (item >= n_items || n < 0) return;
            const F4Item q = f4_item(item, NH, chunks_per_b);
            const float* row = logits + ((size_t)q.b * Nq + n) * ld;
            const float* po = row + q.h * NP * 2;
            const float* pl = row + NH * NP * 2 + q.h * NP;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i * 2 < NP) cp_async16b(slot + kF4SlotLogits + i * 16, po + i * 4);
            cp_async16b(slot + kF4SlotLogits + 64, pl);
            if (NP > 4) cp_async16b(slot + kF4SlotLogits + 80, pl + 4);
            uint32_t rest = m;
#pragma unroll
            for (int k = 0; k < kF4SlotRefCams; ++k) {
                if (!rest) break;
                const int c = __ffs(rest) - 1;
                rest &= rest - 1;
                cp_async8(slot + kF4SlotRefs + k * 8, rp2 + ((size_t)c * B + q.b) * Nq + n);
            }
        };

        F4Timer tw(tid == 0);
        int item = blockIdx.x;
        int n_nx = -1;
        uint32_t m_nx = 0, u_nx = 0;
        {   // ids of the first item: plain loads; then prime the pipeline
            bool ht, hr;
            ids_exist(item, ht, hr);
            const F4Item q = f4_item(item < n_items ? item : 0, NH, chunks_per_b);
            const int tile = 2 * q.chunk + g, i = tile * kF4Rows + r;
            if (ht) u_nx = __ldg(tile_union + (size_t)q.b * tiles_per_b + tile);
            if (hr) {
                n_nx = __ldg(order + (size_t)q.b * Nq + i);
                m_nx = __ldg(smask + (size_t)q.b * Nq + i);
            }
            issue_ids(item + gridDim.x);
            issue_row(item, n_nx, m_nx);
        }
        tw.lap(0);                                      // setup
        for (; item < n_items; item += gridDim.x) {
            const F4Item q = f4_item(item, NH, chunks_per_b);
            const int n = n_nx;
            const uint32_t m = m_nx, u = u_nx;
            // ---- my slot now holds this item's logits / reference points and the next item's ids
            cp_async_commit_wait_all();
            float ox[8], oy[8], aw[8];
            float2 refs[kF4SlotRefCams];
            {
                float4 raw[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) raw[i] = lds_f4(slot + kF4SlotLogits + i * 16);
#pragma unroll
                for (int k = 0; k < kF4SlotRefCams; ++k) refs[k] = lds_f2(slot + kF4SlotRefs + k * 8);
                bool ht, hr;
                ids_exist(item + gridDim.x, ht, hr);
                n_nx = hr ? (int)lds_u32(slot + kF4SlotIds) : -1;
                m_nx = hr ? lds_u32(slot + kF4SlotIds + 4) : 0u;
                u_nx = ht ? lds_u32(slot + kF4SlotIds + 8) : 0u;
                // prefetch: ids of the item after next, logits / reference points of the next item
                issue_ids(item + 2 * gridDim.x);
                issue_row(item + gridDim.x, n_nx, m_nx);
                float mx = -INFINITY;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const float4 o4 = raw[p >> 1];
                    ox[p] = (p & 1) ? o4.z : o4.x;
                    oy[p] = (p & 1) ? o4.w : o4.y;
                    const float4 l4 = raw[4 + (p >> 2)];
                    const float lg = (p & 3) == 0 ? l4.x : (p & 3) == 1 ? l4.y : (p & 3) == 2 ? l4.z : l4.w;
                    aw[p] = (p < NP) ? lg : -INFINITY;
                    mx = fmaxf(mx, aw[p]);
                }
                float s = 0.f;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    aw[p] = (p < NP) ? __expf(aw[p] - mx) : 0.f;
                    s += aw[p];
                }
                const float inv = 1.f / s;
#pragma unroll
                for (int p = 0; p < 8; ++p) aw[p] *= inv;
            }
            tw.lap(1);                                  // item top: slot -> registers, prefetch issue, softmax

            // ---- cameras of my tile, ascending (= the reference's accumulation order, :166-168)
            int kvis = 0;                               // visible cameras of my row walked so far
            for (uint32_t rest = u; rest; rest &= rest - 1) {
                const int cam = __ffs(rest) - 1;
                const bool vis = (m >> cam) & 1u;
                uint32_t kmask = 0;
                if (vis) {
                    float2 ref;
                    if (kvis < kF4SlotRefCams) {
                        ref = refs[0];
#pragma unroll
                        for (int k = 1; k < kF4SlotRefCams; ++k)
                            if (kvis == k) ref = refs[k];
                    } else {
                        ref = __ldg(rp2 + ((size_t)cam * B + q.b) * Nq + n);
                    }
                    ++kvis;
                    // ---- the 8 points: bilinear cell + weights, read-modify-write of the four cells.  floor()
                    // through the 2^23 trick (add with round-down): no conversion-pipe instructions.  The base
                    // cell is clamped into the map; a corner that is outside gets weight 0 on a real cell.
                    const float rx1 = fmaf(ref.x, fSw, 0.5f), ry1 = fmaf(ref.y, fSh, 0.5f);      // pixel coordinate + 1
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        // t = pixel coordinate + 1; c = clamp(floor(t), 1, S - 1) = left / upper cell + 1 of a 2x2 block
                        // that lies inside the map; a cell's weight is the bilinear tent max(0, 1 - |coordinate - cell|),
                        // which is the corner weight for in-map corners and 0 for everything else -- branch free
                        const float tx = rx1 + ox[p], ty = ry1 + oy[p];
                        const float flx = __fadd_rd(tx, 8388608.f) - 8388608.f, fly = __fadd_rd(ty, 8388608.f) - 8388608.f;
                        const float cx = fminf(fmaxf(flx, 1.f), fSw - 1.f), cy = fminf(fmaxf(fly, 1.f), fSh - 1.f);
                        const float dx = tx - cx, dy = ty - cy;
                        const float a = aw[p];
                        const float wxa = fmaxf(1.f - fabsf(dx), 0.f), wxb = fmaxf(1.f - fabsf(dx - 1.f), 0.f);
                        const float wya = a * fmaxf(1.f - fabsf(dy), 0.f), wyb = a * fmaxf(1.f - fabsf(dy - 1.f), 0.f);
                        const __half2 wa = __floats2half2_rn(wya * wxa, wya * wxb);
                        const __half2 wb = __floats2half2_rn(wyb * wxa, wyb * wxb);
                        // pix = (cy - 1) * Sw + (cx - 1), exact small integer in fp32 -> int through the 2^23 trick
                        const int pix = __float_as_int(fmaf(cy, fSw, cx) + pix_bias) - 0x4B000000;
                        const uint32_t c0 = (uint32_t)pix >> 4, c1 = (uint32_t)(pix + Sw + 1) >> 4;
                        kmask |= (2u << c1) - (1u << c0);
                        // right neighbour of cell k: same word (+2) if k is even, next word (+126) if odd; the cell
                        // below is Sw cells on: Sw / 2 words, plus one more cell if Sw is odd
                        const uint32_t odd = (uint32_t)pix & 1u;
                        const uint32_t step0 = odd ? 126u : 2u;
                        const uint32_t a0 = mybase + (((uint32_t)pix >> 1) << 7) + (odd << 1), a0r = a0 + step0;
                        const uint32_t a1 = a0 + row_half + sw_odd * step0, a1r = a1 + ((odd ^ sw_odd) ? 126u : 2u);
                        ua[p] = a0;
                        ub[p] = a1;
                        const uint16_t h0 = lds16(a0), h1 = lds16(a0r), h2 = lds16(a1), h3 = lds16(a1r);
                        sts16(a0, __half_as_ushort(__hadd(__ushort_as_half(h0), __low2half(wa))));
                        sts16(a0r, __half_as_ushort(__hadd(__ushort_as_half(h1), __high2half(wa))));
                        sts16(a1, __half_as_ushort(__hadd(__ushort_as_half(h2), __low2half(wb))));
                        sts16(a1r, __half_as_ushort(__hadd(__ushort_as_half(h3), __high2half(wb))));
                    }
                    tapped = true;
                }
                tw.lap(4);                              // taps: arithmetic + read-modify-writes
                kmask = __reduce_or_sync(VER_FULL_MASK, kmask);
                const uint32_t copy = kmask | dirty;    // chunks of my warp's lanes that change in TMEM
                dirty = kmask;
                if (seen < it) {                        // MMAs of my previous batch retired -> A_g is mine again
                    mbar_wait_park(&bar_mma[g], seen & 1);
                    ++seen;
                    tc_fence_after();
                }
                tw.lap(2);                              // wait: my previous MMA batch retired
                // scratch -> registers -> TMEM (plain loads: the compiler may run the next chunk's ahead of the store)
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    if (c < nchunks && ((copy >> c) & 1u)) {
                        uint32_t rr[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) rr[j] = my_words[(c * 8 + j) * 32];
                        tmem_st8(tm_row + c * 8, rr);
                    }
                }
                asm volatile("" ::: "memory");            // the loads above stay above the un-tap stores
                if (tapped) {                           // un-tap: my scratch row is all zero again
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        sts16(ua[p], 0);
                        sts16(ua[p] + ((ua[p] & 2u) ? 126u : 2u), 0);
                        sts16(ub[p], 0);
                        sts16(ub[p] + ((ub[p] & 2u) ? 126u : 2u), 0);
                    }
                    tapped = false;
                }
                tmem_st_wait();
                tw.lap(5);                              // copy scratch -> TMEM, un-tap
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    s_kmask[g][it & 1][warp & 3] = kmask;
                    mbar_arrive(&bar_built[g]);
                }
                ++it;
                tw.lap(6);                              // fences + arrive
            }
            tw.lap(7);                                  // item end
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tmem, 512);
}

template <int DH, int NP>
int launch_fwd_tc4(const __half* vimg, const float* logits, int ld, const float* rpc, const int32_t* order,
                   const uint32_t* smask, const uint32_t* tile_union, __half* slots, int B, int Ncam, int Nq,
                   int Sh, int Sw, int SP, int NH, cudaStream_t st) {
    const F4Smem L(DH, SP);
    auto kern = sca_fwd_tc4_kernel<DH, NP>;
    VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    const int chunks_per_b = (Nq + kF4ChunkRows - 1) / kF4ChunkRows;
    const int n_items = B * NH * chunks_per_b;
    const int sms = ver_device_sm_count();
    const int grid = n_items < sms ? n_items : sms;
    kern<<<grid, kF4Threads, L.total, st>>>(vimg, logits, ld, rpc, order, smask, tile_union, slots, B, Ncam, Nq, Sh,
                                            Sw, SP, NH, chunks_per_b, n_items);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

}  // namespace

extern "C" int ver_debug_tc4_timing(int enable, unsigned long long* host_out32) {
    if (host_out32) VER_CHECK_CUDA(cudaMemcpyFromSymbol(host_out32, g_f4_timing, sizeof(unsigned long long) * 32));
    unsigned long long zero[32] = {0};
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f4_timing, zero, sizeof(zero)));
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f4_timing_on, &enable, sizeof(int)));
    return VER_OK;
}

// shapes the TMEM-operand kernel covers: the rest of ver_tc3_supported's shapes take sca_fwd_tc3_kernel
int ver_tc4_supported(int Ncam, int S, int Dh, int NP) {
    if (!(Ncam <= 32 && (NP == 4 || NP == 8) && S <= 256 && (Dh == 32 || Dh == 64 || Dh == 96))) return 0;
    const int SP = (S + 15) / 16 * 16;
    if (2 * Dh + SP > 512) return 0;                   // TMEM columns: two accumulators + two A operands
    return F4Smem(Dh, SP).total + 1024 <= ver_device_max_smem_optin();
}

int ver_sca_forward_tc4(const void* vimg, const float* logits, int ld, const float* rpc, const int32_t* order,
                        const uint32_t* smask, const uint32_t* tile_union, void* slots, int B, int Ncam, int Nq,
                        int Sh, int Sw, int NH, int Dh, int NP, cudaStream_t st) {
    const int SP = (Sh * Sw + 15) / 16 * 16;
#define FWD4(D)                                                                                                  \
    (NP == 8 ? launch_fwd_tc4<D, 8>((const __half*)vimg, logits, ld, rpc, order, smask, tile_union, (__half*)slots, B, \
                                    Ncam, Nq, Sh, Sw, SP, NH, st)                                                 \
             : launch_fwd_tc4<D, 4>((const __half*)vimg, logits, ld, rpc, order, smask, tile_union, (__half*)slots, B, \
                                    Ncam, Nq, Sh, Sw, SP, NH, st))
    switch (Dh) {
        case 32: return FWD4(32);
        case 64: return FWD4(64);
        default: return FWD4(96);
    }
#undef FWD4
}

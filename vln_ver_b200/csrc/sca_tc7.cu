// Fused SCA sampler forward, seventh generation: the two-threads-per-row builder of sca_tc6.cu (image-row parity split,
// 16-cell image rows, aligned 32-bit read-modify-writes) in front of the TENSOR-MEMORY A operand of sca_tc4.cu.
//
//     slots[b, n, h, :] = 1/max(count,1) * sum_{cam sees n, ascending} A_cam[n, :] V_{b,cam,h}[:, :]
//
// Replaces SpatialCrossAttention.forward's rebatch / sampling / scatter-mean
// (M/spatial_cross_attention.py:138-173, MSDeformableAttention3D :340-374).  Generation 6 showed that an A tile the
// tensor core can read from shared memory costs 4-way bank conflicts on every read-modify-write and ~200-cycle MMAs;
// this one keeps the rows in a LANE-INTERLEAVED scratch (word w of lane l at (w * 32 + l) * 4: lane l only ever touches
// bank l, conflict free whatever the taps) shared by the two threads of a row, which own disjoint words, and copies
// scratch -> registers -> TMEM (tcgen05.st, lane = row) like generation 4 -- but each thread copies only the image rows
// of its parity (K chunk = image row = 8 TMEM columns), so the copy is split over 16 warps as well.
//
// Roles: warps 0-7 = builders of even image rows (warp w: rows 32 w .. 32 w + 31 of the chunk), warps 8-15 = builders of
// odd image rows (warp 8 + w: the same rows), warps 16-19 = epilogue (TMEM lane quarter = warp % 4), warps 20 / 21 = control
// (one lane each): MMA issue of row group 0 / 1; warp 20 also issues the TMA of the value images.  Row group g = rows 128 g .. 128 g + 127 = UMMA M.
// TMEM columns: [0, 2 DH) the two accumulators, [2 DH, 2 DH + 16 Sh) the two A operands (8 Sh columns each).
// Hand-offs (mbarriers):
//     bar_built[g]    group g wrote A_g for its next camera into TMEM            (8 arrivals, one per warp)
//     bar_mma[g]      tcgen05.commit: the MMAs reading A_g retired -> A_g may be overwritten
//     bar_full[g]     tcgen05.commit after the item's last camera: accumulator g is complete
//     bar_free[g]     the epilogue warps drained accumulator g                  (4 arrivals)
//     bar_v[buf] / bar_vfree[buf]   value image landed (transaction bytes) / all MMAs reading it retired (2 commits)
//   named barrier 1 + w (64 threads): the two builder warps of rows 32 w .. share one prefetch slot (logits of the item)
#include <type_traits>

#include "sampler.cuh"
#include "tap16.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int kF7Builders = 512;
constexpr int kF7Threads = kF7Builders + 128 + 64;       // 16 builder warps, 4 epilogue warps, 2 control warps
constexpr int kF7Rows = 128;                  // rows per group = UMMA M
constexpr int kF7ChunkRows = 2 * kF7Rows;
constexpr int kF7SlotUnits = 6;               // 16-byte units per row: 4 x offsets (8 points x 2), 2 x attention logits
constexpr int kF7SlotWarpBytes = kF7SlotUnits * 512;
constexpr float kF7Magic = 8388608.f;         // 2^23: floor() and float -> int through round-down adds

struct F7Smem {
    int v_bytes, words, warp_scratch, off_v[2], off_scratch, off_slots, total;
    __host__ __device__ F7Smem(int Dh, int Sh) {
        v_bytes = Dh * Sh * 16 * 2;           // [Dh / 8][2 Sh][8][8] halves
        words = Sh * 8 + 2;                   // 32-bit words of a scratch row: 8 per image row + one sink word per parity thread
        warp_scratch = words * 128;           // 32 rows, lane interleaved
        off_v[0] = 0;
        off_v[1] = v_bytes;
        off_scratch = 2 * v_bytes;
        off_slots = off_scratch + 8 * warp_scratch;
        total = off_slots + 8 * kF7SlotWarpBytes;
    }
};

__device__ __forceinline__ void umma_f16_ts7(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(db),
        "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_st8_7(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
}
__device__ __forceinline__ void tmem_st_wait7() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 bits_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }

struct F7Item {
    int b, chunk, h;
};
__device__ __forceinline__ F7Item f7_item(int item, int NH, int chunks_per_b) {
    F7Item it;
    it.h = item % NH;
    const int r = item / NH;
    it.chunk = r % chunks_per_b;
    it.b = r / chunks_per_b;
    return it;
}

// Bounded mbarrier wait (as sca_tc5.cu's): after kF7WaitLimit failed try_waits the waiter records what it was waiting
// for, raises g_f7_abort and returns; every other wait of the grid returns at its next wake-up, the kernel runs off its
// (garbage) end and traps there: the launch FAILS instead of hanging.  ver_debug_tc7(…) reads the record.
__device__ unsigned int g_f7_abort = 0;
__device__ unsigned int g_f7_diag[8];
__device__ int g_f7_flags = 0;                         // bit 0: phase timers, bit 1: do not trap on a failed wait
constexpr unsigned int kF7WaitLimit = 200000;
__device__ __noinline__ void f7_wait_failed(uint32_t code, uint32_t a, uint32_t b) {
    if (atomicExch(&g_f7_abort, 1u) == 0u) {
        g_f7_diag[0] = code;
        g_f7_diag[1] = blockIdx.x;
        g_f7_diag[2] = threadIdx.x;
        g_f7_diag[3] = a;
        g_f7_diag[4] = b;
        __threadfence();
    }
}
__device__ __forceinline__ void f7_wait(uint64_t* bar, uint32_t parity, uint32_t code, uint32_t a, uint32_t b) {
    const uint32_t addr = smem_u32(bar);
    for (unsigned int tries = 0;; ++tries) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(20000u)
            : "memory");
        if (ok) return;
        __nanosleep(tries < 4 ? 40u : 160u);
        if (tries >= 64 && *(volatile unsigned int*)&g_f7_abort) return;
        if (tries >= kF7WaitLimit) {
            f7_wait_failed(code, a, b);
            return;
        }
    }
}

// phase timers (debug; enabled through ver_debug_tc7, read by tools/tc_timing.py, never by the product)
__device__ unsigned long long g_f7_timing[32];
struct F7Timer {
    bool on;
    long long t;
    __device__ __forceinline__ F7Timer(bool active) : on(active && (g_f7_flags & 1)), t(0) {
        if (on) t = clock64();
    }
    __device__ __forceinline__ void lap(int slot) {
        if (on) {
            const long long n = clock64();
            atomicAdd(&g_f7_timing[slot], (unsigned long long)(n - t));
            t = n;
        }
    }
};

template <int DH, int NP>
__global__ void __launch_bounds__(kF7Threads, 1)
sca_fwd_tc7_kernel(const __half* __restrict__ vimg, const float* __restrict__ logits, int ld,
                   const float* __restrict__ rpc, const int32_t* __restrict__ order,
                   const uint32_t* __restrict__ smask, const uint32_t* __restrict__ tile_union,
                   __half* __restrict__ slots, int B, int Ncam, int Nq, int Sh, int Sw, int NH,
                   int chunks_per_b, int n_items) {
    extern __shared__ __align__(128) unsigned char smem[];
    const F7Smem L(DH, Sh);
    __shared__ __align__(8) uint64_t bar_built[2], bar_mma[2], bar_full[2], bar_free[2], bar_v[2], bar_vfree[2];
    __shared__ uint32_t s_tmem;
    __shared__ volatile uint32_t s_kmask[2][2][8];     // [group][batch parity][warp of the group]: image rows that hold taps

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_built[i], 8);
            mbar_init(&bar_mma[i], 1);
            mbar_init(&bar_v[i], 1);
            mbar_init(&bar_vfree[i], 2);
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_free[i], 4);
        }
        mbar_fence_init();
    }
    if (warp == 20) tmem_alloc(&s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int tiles_per_b = (Nq + kF7Rows - 1) / kF7Rows;      // order.cu's tile_union row length
    const size_t v_elems = (size_t)DH * Sh * 16;               // halves per (view, head) image
    const int G = 2 * Sh;                                      // 8-cell K groups of an image
    const uint32_t tm_a0 = tmem + 2 * DH;                      // A operand of group g at + g * 8 Sh columns

    if (warp >= 20) {
        // ================================================================ control: one issuing thread per row group
        // (warp 20: group 0 and the TMA of the value images, warp 21: group 1).  Both walk the same sequence of
        // (item, camera) steps; a single thread issuing both groups' batches one after the other delayed each group's
        // MMAs by the other's issue time.
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
                    const F7Item it = f7_item(nx_item, NH, chunks_per_b);
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
            F7Timer tc(cg == 0);
            bool has_next = advance();
            if (cg == 0 && has_next) load_v(0);
            uint32_t kk = 0, itg = 0, acc_items = 0;
            // V image descriptors of image row 0 (K stride 128 B, N stride G * 128 B); image row y is + y * 16 in the
            // address field; the A operand of group g sits in TMEM at tm_a0 + g * 8 Sh, image row y at + 8 y columns
            const uint64_t dv_0 = umma_desc(smem_u32(smem + L.off_v[0]), 128, G * 128);
            const uint64_t dv_1 = umma_desc(smem_u32(smem + L.off_v[1]), 128, G * 128);
            const uint32_t row_mask = (1u << Sh) - 1u;
            const uint32_t d_addr = tmem + cg * DH, a_addr = tm_a0 + cg * (Sh * 8);
            while (has_next) {
                const int cam = nx_cam;
                const uint32_t ug = nx_ug;
                has_next = advance();                  // nx_* now describe step kk + 1
                const int buf = kk & 1;
                // EVERY step, work or not: a parity wait is only valid for a waiter that observes every phase of the
                // barrier -- a control thread whose group idles through some steps would otherwise run ahead of the
                // value-image loads, pass a later wait early and desynchronise bar_vfree
                f7_wait(&bar_v[buf], (kk >> 1) & 1, 2, kk, 0);
                tc.lap(10);                            // control: wait for the value image
                if ((ug >> cam) & 1u) {
                    const bool first_cam = !(ug & ((1u << cam) - 1u));         // lowest camera of this tile overwrites
                    const bool last_cam = !(ug >> (cam + 1));
                    f7_wait(&bar_built[cg], itg & 1, 1, kk, cg);
                    tc.lap(9);                         // control: wait for a built A
                    if (first_cam && acc_items) f7_wait(&bar_free[cg], (acc_items - 1) & 1, 3, kk, cg);
                    tc_fence_after();
                    tc.lap(13);                        // control: wait for a drained accumulator
                    const int par = itg & 1;
                    uint32_t km = 0;
#pragma unroll
                    for (int w = 0; w < 8; ++w) km |= s_kmask[cg][par][w];
                    km &= row_mask;
                    if (first_cam && !km) km = 1u;     // (an all-zero image row zeroes the accumulator)
                    const uint64_t dvb = buf ? dv_1 : dv_0;        // (no indexed arrays here: they would live in local memory)
                    // the issue loop is this thread's critical path and shares its scheduler with five other warps:
                    // straight-line code, the accumulate flag a compile-time constant except on an item's first camera
                    if (first_cam) {
                        const uint32_t first_bit = km & (0u - km);
#pragma unroll
                        for (int y = 0; y < 14; ++y)
                            if ((km >> y) & 1u)
                                umma_f16_ts7(d_addr, a_addr + y * 8, dvb + (uint64_t)(y * 16), idesc, (first_bit >> y) & 1u ? 0u : 1u);
                    } else {
#pragma unroll
                        for (int y = 0; y < 14; ++y)
                            if ((km >> y) & 1u) umma_f16_ts7(d_addr, a_addr + y * 8, dvb + (uint64_t)(y * 16), idesc, 1u);
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
                // (a tcgen05.commit with nothing in flight before it was observed never to arrive)
                if ((ug >> cam) & 1u) umma_commit(&bar_vfree[buf]);
                else mbar_arrive(&bar_vfree[buf]);
                if (cg == 0 && has_next) {
                    // value image of step kk + 1 -> the other buffer, once step kk - 1 stopped reading it
                    if (kk >= 1) f7_wait(&bar_vfree[(kk + 1) & 1], ((kk - 1) >> 1) & 1, 4, kk, 0);
                    load_v((kk + 1) & 1);
                    tc.lap(12);                        // control: wait for a free value buffer + TMA issue
                }
                ++kk;
            }
            // drain: the last commits must have arrived before the CTA tears TMEM / smem down
            if (kk >= 1) f7_wait(&bar_vfree[(kk - 1) & 1], ((kk - 1) >> 1) & 1, 5, kk, 0);
        };
        if (lane == 0) {
            if (warp == 20) control(std::integral_constant<int, 0>{});
            else control(std::integral_constant<int, 1>{});
        }
    } else if (warp >= 16) {        // (warps 16-19; the control warps were taken above)
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
                const F7Item it = f7_item(item, NH, chunks_per_b);
                const int tile = 2 * it.chunk + g, i = tile * kF7Rows + r;
                if (tile < tiles_per_b) u_nx[g] = __ldg(tile_union + (size_t)it.b * tiles_per_b + tile);
                if (i < Nq) {
                    n_nx[g] = __ldg(order + (size_t)it.b * Nq + i);
                    m_nx[g] = __ldg(smask + (size_t)it.b * Nq + i);
                }
            }
        };
        F7Timer te(tid == kF7Builders);
        load_ids(blockIdx.x);
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const F7Item it = f7_item(item, NH, chunks_per_b);
            const int n[2] = {n_nx[0], n_nx[1]};
            const uint32_t m[2] = {m_nx[0], m_nx[1]}, u[2] = {u_nx[0], u_nx[1]};
            load_ids(item + gridDim.x);                 // in flight during this item's epilogue
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const float inv_cnt = 1.f / (float)max(__popc(m[g]), 1);
                __half* dst = (n[g] >= 0) ? slots + (((size_t)it.b * Nq + n[g]) * NH + it.h) * DH : nullptr;
                if (u[g]) {                             // warp-uniform (tile property)
                    f7_wait(&bar_full[g], full_seen[g] & 1, 6, (uint32_t)item, g);
                    ++full_seen[g];
                    tc_fence_after();
                    te.lap(16);                         // epilogue: wait for a complete accumulator
#pragma unroll
                    for (int c0 = 0; c0 < DH; c0 += 32) {
                        float vv[32];
                        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + g * DH + c0, vv);
                        if (c0 + 32 >= DH) {                       // last read of this accumulator: hand it back
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
        // ================================================================ builders: two threads per row
        const int pi = warp >> 3, rw = warp & 7, g = rw >> 2, rr = (rw & 3) * 32 + lane;
        // my scratch row (shared with the thread of the other parity): word w at mybase + w * 128; image row y = words
        // 8 y .. 8 y + 7, then one sink word per parity (second pair of a point clamped to X = Sw + 1: always zero)
        const uint32_t* my_words = reinterpret_cast<const uint32_t*>(smem + L.off_scratch + (size_t)rw * L.warp_scratch) + lane;
        const uint32_t mybase = smem_u32(smem + L.off_scratch) + (uint32_t)rw * L.warp_scratch + (uint32_t)lane * 4u;
        const uint32_t dummy = mybase + (uint32_t)(Sh * 8 + pi) * 128u;
        const uint32_t slot = smem_u32(smem + L.off_slots) + (uint32_t)rw * kF7SlotWarpBytes + (uint32_t)lane * 16u;
        const uint32_t tm_row = tm_a0 + g * (Sh * 8) + ((uint32_t)((rw & 3) * 32) << 16);       // my TMEM lane, A_g
        const float fSw = (float)Sw, fSh = (float)Sh, fpi = (float)pi;
        const float xmax = (float)(Sw + 1);
        const float jtop = kF7Magic + (float)((Sh - pi + 1) / 2 - 1);       // last image row of my parity, magic domain
        // word index (2 j + pi) * 8 + e / 2 with j and e / 2 taken as 0x4B000000 + integer (magic floats)
        const uint32_t base_adj = mybase + (uint32_t)pi * 1024u - ((0x4B000000u * 17u) << 7);
        const float2* rp2 = reinterpret_cast<const float2*>(rpc);
        uint32_t it = 0, seen = 0;                // MMA batches handed over / observed retired (this group)
        uint32_t dirty = 0;                       // image rows (bit 2 j) of my parity that may be non-zero in TMEM
        bool tapped = false;                      // my words hold the taps recorded in ua / ub
        uint32_t ua[8];                           // address of the first pair each point updated (+ flag, see below)

        // my words of the scratch row and my image rows of the TMEM operand start out all zero
        for (int w = pi; w < L.words; w += 2) sts32(mybase + (uint32_t)w * 128u, 0u);
        {
            const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int y = pi; y < Sh; y += 2) tmem_st8_7(tm_row + y * 8, z);
            tmem_st_wait7();
        }
        named_bar_sync(1 + rw, 64);               // (the other parity zeroed the other half of the words)

        auto load_ids = [&](int item, int& n, uint32_t& m, uint32_t& u) {
            n = -1;
            m = u = 0;
            if (item >= n_items) return;
            const F7Item q = f7_item(item, NH, chunks_per_b);
            const int tile = 2 * q.chunk + g, i = tile * kF7Rows + rr;
            if (tile < tiles_per_b) u = __ldg(tile_union + (size_t)q.b * tiles_per_b + tile);
            if (i < Nq) {
                n = __ldg(order + (size_t)q.b * Nq + i);
                m = __ldg(smask + (size_t)q.b * Nq + i);
            }
        };
        auto issue_row = [&](int item, int n) {                 // logits of (row, head) of `item` -> the pair's slot
            if (item >= n_items || n < 0) return;
            const F7Item q = f7_item(item, NH, chunks_per_b);
            const float* row = logits + ((size_t)q.b * Nq + n) * ld;
            const float* po = row + q.h * NP * 2;
            const float* pl = row + NH * NP * 2 + q.h * NP;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i * 2 < NP) cp_async16(slot + i * 512, po + i * 4);
            cp_async16(slot + 4 * 512, pl);
            if (NP > 4) cp_async16(slot + 5 * 512, pl + 4);
        };

        F7Timer tw(tid == 0);
        int item = blockIdx.x;
        int n_nx, n_n2;
        uint32_t m_nx, u_nx, m_n2, u_n2;
        load_ids(item, n_nx, m_nx, u_nx);
        load_ids(item + gridDim.x, n_n2, m_n2, u_n2);
        if (pi == 0) issue_row(item, n_nx);
        float2 ref_nx = make_float2(0.f, 0.f);
        int ref_item = -1;                               // item whose first camera's reference point ref_nx holds
        tw.lap(0);                                      // setup
        for (; item < n_items; item += gridDim.x) {
            const F7Item q = f7_item(item, NH, chunks_per_b);
            const int n = n_nx;
            const uint32_t m = m_nx, u = u_nx;
            n_nx = n_n2;
            m_nx = m_n2;
            u_nx = u_n2;
            load_ids(item + 2 * gridDim.x, n_n2, m_n2, u_n2);       // consumed two items from now
            // ---- the pair's slot holds this item's logits
            if (pi == 0) cp_async_wait_all();
            named_bar_sync(1 + rw, 64);
            float ox[8], oy[8], aw[8];
            {
                float4 raw[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) raw[i] = lds128(slot + i * 512);
                named_bar_sync(1 + rw, 64);             // both threads hold the logits: the slot may be refilled
                if (pi == 0) issue_row(item + gridDim.x, n_nx);
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
            if (u && ref_item != item) {                // not prefetched (first item, or the item before had no camera)
                const int cam = __ffs(u) - 1;
                if ((m >> cam) & 1u) ref_nx = __ldg(rp2 + ((size_t)cam * B + q.b) * Nq + n);
            }
            tw.lap(1);                                  // item top: slot -> registers, prefetch issue, softmax

            // ---- cameras of my tile, ascending (= the reference's accumulation order, :166-168)
            for (uint32_t rest = u; rest;) {
                const int cam = __ffs(rest) - 1;
                rest &= rest - 1;
                const bool vis = (m >> cam) & 1u;
                const float2 ref = ref_nx;
                // reference point of my next batch: next camera of this tile, else the first camera of the next item
                if (rest) {
                    const int cam2 = __ffs(rest) - 1;
                    if ((m >> cam2) & 1u) ref_nx = __ldg(rp2 + ((size_t)cam2 * B + q.b) * Nq + n);
                } else if (u_nx) {
                    const int cam2 = __ffs(u_nx) - 1;
                    if ((m_nx >> cam2) & 1u) {
                        const F7Item qn = f7_item(item + gridDim.x, NH, chunks_per_b);
                        ref_nx = __ldg(rp2 + ((size_t)cam2 * B + qn.b) * Nq + n_nx);
                    }
                    ref_item = item + gridDim.x;
                }
                // ---- the 8 points: my image row of each, three cell weights, two 32-bit read-modify-writes
                uint32_t kmask = 0;
                if (vis) {
                    const float rx1 = fmaf(ref.x, fSw, 0.5f);                     // X = pixel x + 1
                    const float ry1 = fmaf(ref.y, fSh, 0.5f) - fpi;               // Y - pi, Y = pixel y + 1
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        // aligned pair base + three x weights, my image row + its weight: tap16.cuh (checked on the CPU)
                        const Tap16 tp = tap16(rx1 + ox[p], ry1 + oy[p], aw[p], xmax, jtop);
                        const float w0 = tp.w0, w1 = tp.w1, w2 = tp.w2, wy = tp.wy, hh = tp.hh, jm = tp.jm;
                        const __half2 h01 = __floats2half2_rn(wy * w0, wy * w1), h2 = __floats2half2_rn(wy * w2, 0.f);
                        const uint32_t hb = __float_as_uint(hh), jb = __float_as_uint(jm);
                        const uint32_t ad = base_adj + ((jb * 16u + hb) << 7);
                        const uint32_t ad2 = (hb == 0x4B000007u) ? dummy : ad + 128u;
                        kmask |= 1u << ((jb * 2u) & 15u);
                        ua[p] = ad | (hb == 0x4B000007u ? 1u : 0u);          // bit 0: the second pair went to the sink word
                        const uint32_t v0 = lds32(ad), v1 = lds32(ad2);
                        sts32(ad, h2_bits(__hadd2(bits_h2(v0), h01)));
                        sts32(ad2, h2_bits(__hadd2(bits_h2(v1), h2)));
                    }
                    tapped = true;
                }
                tw.lap(4);                              // taps: arithmetic + read-modify-writes
                kmask = __reduce_or_sync(VER_FULL_MASK, kmask);          // bit 2 j: image row 2 j + pi of my warp's rows
                const uint32_t copy = kmask | dirty;    // image rows of my parity that change in TMEM
                dirty = kmask;
                // ---- the MMAs of my group's previous batch retired -> A_g in TMEM is ours again
                if (seen < it) {
                    f7_wait(&bar_mma[g], seen & 1, 7, it, (uint32_t)warp);
                    ++seen;
                    tc_fence_after();
                }
                tw.lap(2);                              // wait: my previous MMA batch retired
                // scratch -> registers -> TMEM, image row by image row (8 words = 8 columns)
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const int y = 2 * j + pi;
                    if (y < Sh && ((copy >> (2 * j)) & 1u)) {
                        uint32_t rr8[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) rr8[c] = my_words[(y * 8 + c) * 32];
                        tmem_st8_7(tm_row + y * 8, rr8);
                    }
                }
                asm volatile("" ::: "memory");            // the loads above stay above the un-tap stores
                if (tapped) {                           // un-tap: my words are all zero again
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        const uint32_t a0 = ua[p] & ~3u;
                        sts32(a0, 0u);
                        if (!(ua[p] & 1u)) sts32(a0 + 128u, 0u);           // (the sink word is never read)
                    }
                    tapped = false;
                }
                tmem_st_wait7();
                tw.lap(3);                              // copy scratch -> TMEM, un-tap
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    s_kmask[g][it & 1][(rw & 3) + 4 * pi] = kmask << pi;      // bit y: image row y holds taps
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
    if (warp == 20) tmem_dealloc(tmem, 512);
    if (*(volatile unsigned int*)&g_f7_abort && !(g_f7_flags & 2)) __trap();      // fail loudly (see f7_wait)
}

template <int DH, int NP>
int launch_fwd_tc7(const __half* vimg, const float* logits, int ld, const float* rpc, const int32_t* order,
                   const uint32_t* smask, const uint32_t* tile_union, __half* slots, int B, int Ncam, int Nq,
                   int Sh, int Sw, int NH, cudaStream_t st) {
    const F7Smem L(DH, Sh);
    auto kern = sca_fwd_tc7_kernel<DH, NP>;
    VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    const int chunks_per_b = (Nq + kF7ChunkRows - 1) / kF7ChunkRows;
    const int n_items = B * NH * chunks_per_b;
    const int sms = ver_device_sm_count();
    const int grid = n_items < sms ? n_items : sms;
    kern<<<grid, kF7Threads, L.total, st>>>(vimg, logits, ld, rpc, order, smask, tile_union, slots, B, Ncam, Nq, Sh,
                                            Sw, NH, chunks_per_b, n_items);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

}  // namespace

// debug hook of tools/ and tests/: flags bit 0 = phase timers on, bit 1 = report a failed wait instead of trapping;
// returns the phase timers and the record of the first failed wait (diag[0] != 0), and clears both
extern "C" int ver_debug_tc7(int flags, unsigned long long* timing32, unsigned int* diag8) {
    unsigned int abort_flag = 0;
    VER_CHECK_CUDA(cudaMemcpyFromSymbol(&abort_flag, g_f7_abort, sizeof(abort_flag)));
    if (timing32) VER_CHECK_CUDA(cudaMemcpyFromSymbol(timing32, g_f7_timing, sizeof(unsigned long long) * 32));
    if (diag8) {
        VER_CHECK_CUDA(cudaMemcpyFromSymbol(diag8, g_f7_diag, sizeof(unsigned int) * 8));
        if (!abort_flag) diag8[0] = 0;
    }
    unsigned long long zero[32] = {0};
    unsigned int zero8[8] = {0};
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f7_timing, zero, sizeof(zero)));
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f7_diag, zero8, sizeof(zero8)));
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f7_abort, zero8, sizeof(unsigned int)));
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f7_flags, &flags, sizeof(int)));
    return abort_flag ? 1 : 0;
}

// shapes the kernel covers: image rows of at most 14 pixels (16 cells with the zero padding), at most 14 image rows
int ver_tc7_supported(int Ncam, int Sh, int Sw, int Dh, int NP) {
    if (!(Ncam <= 32 && (NP == 4 || NP == 8) && Sh >= 2 && Sh <= 14 && Sw >= 2 && Sw <= 14 &&
          (Dh == 32 || Dh == 64 || Dh == 96)))
        return 0;
    if (2 * Dh + 16 * Sh > 512) return 0;                  // TMEM columns: two accumulators + two A operands
    return F7Smem(Dh, Sh).total + 1024 <= ver_device_max_smem_optin();
}

int ver_sca_forward_tc7(const void* vimg16, const float* logits, int ld_logits, const float* rpc, const int32_t* order,
                        const uint32_t* smask, const uint32_t* tile_union, void* slots, int B, int Ncam, int Nq, int Sh,
                        int Sw, int NH, int Dh, int NP, cudaStream_t st) {
#define FWD7(D)                                                                                                      \
    (NP == 8 ? launch_fwd_tc7<D, 8>((const __half*)vimg16, logits, ld_logits, rpc, order, smask, tile_union,         \
                                    (__half*)slots, B, Ncam, Nq, Sh, Sw, NH, st)                                     \
             : launch_fwd_tc7<D, 4>((const __half*)vimg16, logits, ld_logits, rpc, order, smask, tile_union,         \
                                    (__half*)slots, B, Ncam, Nq, Sh, Sw, NH, st))
    switch (Dh) {
        case 32: return FWD7(32);
        case 64: return FWD7(64);
        default: return FWD7(96);
    }
#undef FWD7
}

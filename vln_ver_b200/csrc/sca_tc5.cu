// Fused SCA sampler forward, fifth generation: sca_tc4.cu's formulation (interpolation matrix A through TENSOR
// MEMORY into tcgen05.mma, visibility-sorted 128-row tiles, one thread per row, persistent CTAs over
// (panorama, 256-row chunk, head) items) with the builder / tensor-core hand-off taken off the critical path.
//
//     slots[b, n, h, :] = 1/max(count,1) * sum_{cam sees n, ascending} A_cam[n, :] V_{b,cam,h}[:, :]
//
// Replaces SpatialCrossAttention.forward's rebatch / sampling / scatter-mean
// (M/spatial_cross_attention.py:138-173, MSDeformableAttention3D :340-374).
//
// What profiles/r01x_tc_phase_timers.txt showed for sca_fwd_tc4_kernel (742 k cycles per CTA): a row group
// spent 197 k cycles waiting for the MMAs that read its single A operand to retire and 248 k in the
// scratch -> TMEM copy, most of it in tcgen05.wait::st behind the other group's MMA batch -- the builders were
// idle more than half of the time while the tensor pipe was 13 % busy.  Changes:
//   * THREE A operands in TMEM (columns [2 DH, 2 DH + 3 SP/2)), handed out round-robin over ONE global batch
//     sequence J = 0, 1, 2, ... per CTA (item by item, cameras of u0 | u1 ascending, group 0 before group 1): batch J
//     uses operand J % 3.  Every role derives J from the two tile unions of the item, nobody communicates it.
//     A group therefore never waits for its own previous MMA batch, only for batch J - 3.  mbarrier waits are
//     parity waits, so a waiter must never fall two phases behind a barrier: ALL eight builder warps walk the
//     global sequence -- at a batch of the other group a warp only observes "batch J - 3 retired" and arrives on
//     bar_built (8 arrivals per batch) -- which makes every warp see every phase of every barrier, in order,
//     and keeps bar_mma[o] from completing a phase before everybody saw the previous one;
//   * a builder copies batch n (scratch -> registers -> tcgen05.st), un-taps its scratch row, publishes, and builds
//     the taps of batch n + 1 while the MMAs of batch n run; the other group's positions in front of batch n + 1
//     are passed early when that does not block (test_wait);
//   * per-point instruction count: tent weights as one FADD.SAT each (1 - |d| saturated), the K-chunk mask from
//     the min / max tap cell of the row instead of per point, even map widths take constant row strides.
//
//   * the MMA issue loop was the serial bottleneck of tc4 (profiles/r02a: 221 cycles per tcgen05.mma on the control
//     thread against 13 % tensor-pipe activity -- a find-first-set loop with the shared-memory descriptor rebuilt on
//     the uniform datapath for every instruction): each group now has its own issuing thread, the K chunks are a
//     counted loop over [first, last] chunk with the descriptor advanced by a constant, and the value-image TMA
//     moved to a thread of its own, so a free value buffer never waits for an issuing thread.
//
//   * the warp scheduler prefers the highest warp id among the ready warps and the waiting roles poll their barriers:
//     the builders are now the HIGHEST warps of the CTA, and a failed wait backs off with nanosleep instead of
//     re-polling (profiles/r02a: 45 % of the executed instructions were the epilogue warps' polling loop).
//
// Roles: warps 0-3 = epilogue (TMEM lane quarter = warp % 4, both groups), warps 4 / 5 = MMA issue for group 0 / 1
// (one lane each), warp 6 = TMA of the value images (one lane), warp 7 idle, warps 8-11 = builders of group 0
// (rows 0..127 of the chunk), warps 12-15 = builders of group 1.
// TMEM columns: [0, 2 DH) the two accumulators, then three A operands of SP / 2 columns.
// Hand-offs (mbarriers):
//     bar_built[o]    batch J (J % 3 == o): the 4 warps of the owning group wrote operand o, the 4 warps of the
//                     other group passed position J                             (8 arrivals)
//     bar_mma[o]      tcgen05.commit: the MMAs reading operand o retired -> batch J + 3 may overwrite it
//     bar_full[g]     tcgen05.commit after the item's last camera: accumulator g is complete
//     bar_free[g]     the epilogue warps drained accumulator g                  (4 arrivals)
//     bar_v[buf] / bar_vfree[buf]   value image landed (transaction bytes) / all MMAs reading it retired
//     bar_done        every MMA of the CTA retired
#include "sampler.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int kF5Workers = 256;
constexpr int kF5Threads = 512;
constexpr int kF5FirstWorker = kF5Threads - kF5Workers;      // builder threads are the LAST 256 threads
constexpr int kF5Rows = 128;                  // rows per group = UMMA M
constexpr int kF5ChunkRows = 2 * kF5Rows;
constexpr int kF5Ops = 3;                     // A operands in TMEM
// per-warp landing zone of the prefetches (cp.async), LANE-INTERLEAVED per field so that a warp's copy of one field
// is one contiguous run (profiles/r02a: with per-thread 144-byte slots every 16-byte cp.async of a warp cost 32
// shared-memory wavefronts, 39 % of all wavefronts of the kernel): six 16-byte fields = 24 fp32 logits of (row,
// head) | four 4-byte fields = n, camera mask, union of my tile, union of the other group's tile (all of the item
// after next) | four 8-byte fields = reference points of the first 4 cameras that see the row
constexpr int kF5SlotLogits = 0, kF5SlotIds = 6 * 512, kF5SlotRefs = kF5SlotIds + 4 * 128;
constexpr int kF5SlotRefCams = 4;
constexpr int kF5SlotWarpBytes = kF5SlotRefs + kF5SlotRefCams * 256;       // 4608 = 144 bytes per thread

struct F5Smem {
    int v_bytes, warp_scratch, off_v[2], off_scratch, off_slots, total;
    __host__ __device__ F5Smem(int Dh, int SP) {
        v_bytes = Dh * SP * 2;
        // scratch rows of one warp, lane-interleaved: 32-bit word w (cells 2w, 2w + 1) of lane l at (w * 32 + l) * 4
        // -> lane l only ever touches bank l: every scratch access of a warp is conflict free, whatever the taps
        warp_scratch = (SP / 2) * 32 * 4;
        off_v[0] = 0;
        off_v[1] = v_bytes;
        off_scratch = 2 * v_bytes;
        off_slots = off_scratch + (kF5Workers / 32) * warp_scratch;
        total = off_slots + (kF5Workers / 32) * kF5SlotWarpBytes;
    }
};

__device__ __forceinline__ void f5_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(db),
        "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void f5_tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
}
__device__ __forceinline__ void f5_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void f5_cp4(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void f5_cp8(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void f5_cp16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void f5_cp_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ uint16_t f5_lds16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.b16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void f5_sts16(uint32_t a, uint16_t v) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(v) : "memory");
}
__device__ __forceinline__ float4 f5_lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float2 f5_lds_f2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t f5_lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void f5_named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct F5Item {
    int b, chunk, h;
};
__device__ __forceinline__ F5Item f5_item(int item, int NH, int chunks_per_b) {
    F5Item it;
    it.h = item % NH;
    const int r = item / NH;
    it.chunk = r % chunks_per_b;
    it.b = r / chunks_per_b;
    return it;
}
// position of batch (camera `cam`, group g) inside its item's batch sequence: cameras of u0 | u1 ascending,
// group 0 before group 1
__device__ __forceinline__ uint32_t f5_batch_index(uint32_t u0, uint32_t u1, int cam, int g) {
    const uint32_t below = (1u << cam) - 1u;
    return __popc(u0 & below) + __popc(u1 & below) + (g ? ((u0 >> cam) & 1u) : 0u);
}

// Bounded mbarrier wait.  A protocol error must not hang the GPU: after kF5WaitLimit failed try_waits (each
// parks the thread for up to 20 us, then sleeps 40-160 ns) the waiter records what it was waiting for in g_f5_diag, raises g_f5_abort
// and returns; every other wait of the grid then returns at its next wake-up, the kernel runs off its (garbage)
// end and traps there, so the launch FAILS instead of hanging.  ver_debug_tc5_diag() reads the record.
__device__ unsigned int g_f5_abort = 0;
__device__ unsigned int g_f5_diag[8];
constexpr unsigned int kF5WaitLimit = 200000;       // x (<= 20 us park + 160 ns sleep): seconds
__device__ __noinline__ void f5_wait_failed(uint32_t code, uint32_t a, uint32_t b) {
    if (atomicExch(&g_f5_abort, 1u) == 0u) {
        g_f5_diag[0] = code;
        g_f5_diag[1] = blockIdx.x;
        g_f5_diag[2] = threadIdx.x;
        g_f5_diag[3] = a;
        g_f5_diag[4] = b;
        __threadfence();
    }
}
__device__ __forceinline__ void f5_wait(uint64_t* bar, uint32_t parity, uint32_t code, uint32_t a, uint32_t b) {
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
        // back off instead of re-polling: the scheduler favours the highest ready warp, a polling waiter
        // takes issue slots from the warps it waits for
        __nanosleep(tries < 4 ? 40u : 160u);
        if (tries >= 64 && *(volatile unsigned int*)&g_f5_abort) return;
        if (tries >= kF5WaitLimit) {
            f5_wait_failed(code, a, b);
            return;
        }
    }
}

// phase timers (debug; enabled through ver_debug_tc5_timing, read by tools/tc_timing.py, never by the product)
__device__ unsigned long long g_f5_timing[32];
__device__ int g_f5_timing_on = 0;
struct F5Timer {
    bool on;
    long long t;
    __device__ __forceinline__ F5Timer(bool active) : on(active && (g_f5_timing_on & 1)), t(0) {
        if (on) t = clock64();
    }
    __device__ __forceinline__ void lap(int slot) {
        if (on) {
            const long long n = clock64();
            atomicAdd(&g_f5_timing[slot], (unsigned long long)(n - t));
            t = n;
        }
    }
};

template <int DH, int NP, bool SW_EVEN>
__global__ void __launch_bounds__(kF5Threads, 1)
sca_fwd_tc5_kernel(const __half* __restrict__ vimg, const float* __restrict__ logits, int ld,
                   const float* __restrict__ rpc, const int32_t* __restrict__ order,
                   const uint32_t* __restrict__ smask, const uint32_t* __restrict__ tile_union,
                   __half* __restrict__ slots, int B, int Ncam, int Nq, int Sh, int Sw, int SP, int NH,
                   int chunks_per_b, int n_items) {
    const int G = SP >> 3;                       // 8-pixel groups per row of the V image
    extern __shared__ __align__(128) unsigned char smem[];
    const F5Smem L(DH, SP);
    __shared__ __align__(8) uint64_t bar_built[kF5Ops], bar_mma[kF5Ops], bar_full[2], bar_free[2], bar_v[2], bar_vfree[2], bar_done;
    __shared__ uint32_t s_tmem;
    // [operand][lane quarter]: K chunks (16 pixels) of the operand's lanes that hold taps of its latest batch;
    // the control thread ORs the four quarters for the MMA loop, the next writer of the operand reads its
    // quarter's entry as the set of chunks it has to overwrite with zeros
    __shared__ volatile uint32_t s_kmask[kF5Ops][4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < kF5Ops; ++i) {
            mbar_init(&bar_built[i], 8);
            mbar_init(&bar_mma[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_free[i], 4);
            mbar_init(&bar_v[i], 1);
            mbar_init(&bar_vfree[i], 2);
        }
        mbar_init(&bar_done, 2);
        mbar_fence_init();
    }
    if (tid < kF5Ops * 4) s_kmask[tid >> 2][tid & 3] = 0;
    if (warp == 4) tmem_alloc(&s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int tiles_per_b = (Nq + kF5Rows - 1) / kF5Rows;      // order.cu's tile_union row length
    const size_t v_elems = (size_t)DH * SP;                    // halves per (view, head) image
    const int nchunks = SP >> 4;
    const uint32_t op_cols = (uint32_t)SP >> 1;                // TMEM columns of one A operand
    const uint32_t tm_a0 = tmem + 2 * DH;                      // A operand o at + o * op_cols

    if (warp >= 4 && warp < 8) {
        // ================================================================ control: MMA issue (warp 12), value TMA (warp 13)
        if (lane == 0 && warp < 7) {
            // all three threads walk the same steps: (item, camera) for the cameras of u0 | u1, ascending
            int nx_item = (int)blockIdx.x - (int)gridDim.x;
            uint32_t nx_rest = 0, nx_u0 = 0, nx_u1 = 0;
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
                    const F5Item it = f5_item(nx_item, NH, chunks_per_b);
                    nx_b = it.b;
                    nx_h = it.h;
                    const uint32_t* tu = tile_union + (size_t)it.b * tiles_per_b + 2 * it.chunk;
                    nx_u0 = tu[0];
                    nx_u1 = (2 * it.chunk + 1 < tiles_per_b) ? tu[1] : 0u;
                    nx_rest = nx_u0 | nx_u1;
                }
            };
            if (warp == 6) {
                // ---- value images: step kk -> buffer kk & 1, as soon as step kk - 2 stopped reading it
                for (uint32_t kk = 0; advance(); ++kk) {
                    const int buf = kk & 1;
                    if (kk >= 2) f5_wait(&bar_vfree[buf], ((kk - 2) >> 1) & 1, 4, kk, 0);
                    mbar_expect_tx(&bar_v[buf], L.v_bytes);
                    bulk_g2s(smem + L.off_v[buf], vimg + ((size_t)(nx_b * Ncam + nx_cam) * NH + nx_h) * v_elems,
                             L.v_bytes, &bar_v[buf]);
                }
            } else {
                // ---- MMA issue for group cg.  Both issuing threads walk ALL batches in global order and observe
                // every phase of bar_v / bar_built (parity waits must not skip phases); each issues only the
                // MMAs of its own group, so the MMAs on one accumulator stay in one thread's program order.
                const int cg = warp - 4;
                constexpr uint32_t idesc = umma_idesc(128, DH, 0, 0);
                const uint64_t vdesc[2] = {umma_desc(smem_u32(smem + L.off_v[0]), 128, G * 128),
                                           umma_desc(smem_u32(smem + L.off_v[1]), 128, G * 128)};
                const uint32_t d_addr = tmem + cg * DH;
                F5Timer tc(cg == 0);
                uint32_t kk = 0, J = 0, acc_items = 0;
                while (advance()) {
                    const int cam = nx_cam;
                    const uint32_t u[2] = {nx_u0, nx_u1};
                    const int buf = kk & 1;
                    f5_wait(&bar_v[buf], (kk >> 1) & 1, 2, J, kk);
                    tc.lap(10);                            // control: wait for the value image
#pragma unroll
                    for (int gg = 0; gg < 2; ++gg) {
                        if (!((u[gg] >> cam) & 1u)) continue;
                        const uint32_t op = J % kF5Ops, use = J / kF5Ops;
                        f5_wait(&bar_built[op], use & 1, 1, J, cg);
                        tc.lap(9);                         // control: wait for a built A
                        if (gg == cg) {
                            const bool first_cam = !(u[gg] & ((1u << cam) - 1u));  // lowest camera of this tile overwrites
                            const bool last_cam = !(u[gg] >> (cam + 1));
                            if (first_cam && acc_items) f5_wait(&bar_free[cg], (acc_items - 1) & 1, 3, J, cg);
                            tc_fence_after();
                            tc.lap(13);                    // control: wait for a drained accumulator
                            uint32_t km = s_kmask[op][0] | s_kmask[op][1] | s_kmask[op][2] | s_kmask[op][3];
                            uint32_t acc = first_cam ? 0u : 1u;
                            if (!acc && !km) km = 1u;      // (an all-zero chunk zeroes the accumulator)
                            if (km) {
                                // K chunks [ks, ke): 8 operand columns and 256 bytes of the value image (16 descriptor
                                // units) apiece.  Chunks inside the range that hold no tap are zero in TMEM.
                                int ks = __ffs(km) - 1;
                                const int ke = 32 - __clz(km);
                                uint32_t a_addr = tm_a0 + op * op_cols + ks * 8;
                                uint64_t db = vdesc[buf] + (uint64_t)(ks * 16);
                                if (g_f5_timing_on & 4) ks = ke;      // (debug: bottleneck experiments, results invalid)
                                for (; ks < ke; ++ks, a_addr += 8, db += 16) {
                                    f5_umma_ts(d_addr, a_addr, db, idesc, acc);
                                    acc = 1u;
                                }
                            }
                            umma_commit(&bar_mma[op]);
                            if (last_cam) {
                                umma_commit(&bar_full[cg]);
                                ++acc_items;
                            }
                            tc.lap(11);                    // control: MMA issue
                        }
                        ++J;
                    }
                    umma_commit(&bar_vfree[buf]);          // (2 arrivals: one per issuing thread)
                    ++kk;
                }
                // drain: every MMA retired before the CTA tears TMEM / smem down
                umma_commit(&bar_done);
                f5_wait(&bar_done, 0, 5, J, kk);
            }
        }
    } else if (warp < 4) {
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
                const F5Item it = f5_item(item, NH, chunks_per_b);
                const int tile = 2 * it.chunk + g, i = tile * kF5Rows + r;
                if (tile < tiles_per_b) u_nx[g] = __ldg(tile_union + (size_t)it.b * tiles_per_b + tile);
                if (i < Nq) {
                    n_nx[g] = __ldg(order + (size_t)it.b * Nq + i);
                    m_nx[g] = __ldg(smask + (size_t)it.b * Nq + i);
                }
            }
        };
        F5Timer te(tid == 0);
        load_ids(blockIdx.x);
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const F5Item it = f5_item(item, NH, chunks_per_b);
            const int n[2] = {n_nx[0], n_nx[1]};
            const uint32_t m[2] = {m_nx[0], m_nx[1]}, u[2] = {u_nx[0], u_nx[1]};
            load_ids(item + gridDim.x);                 // in flight during this item's epilogue
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const float inv_cnt = 1.f / (float)max(__popc(m[g]), 1);
                __half* dst = (n[g] >= 0) ? slots + (((size_t)it.b * Nq + n[g]) * NH + it.h) * DH : nullptr;
                if (u[g]) {                             // warp-uniform (tile property)
                    f5_wait(&bar_full[g], full_seen[g] & 1, 6, (uint32_t)item, g);
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
        const int ww = warp - kF5FirstWorker / 32;               // builder warp 0..7
        const int g = ww >> 2, q = ww & 3, r = q * 32 + lane;
        // my scratch row: 32-bit word w at my_words[w * 32]
        const uint32_t* my_words = reinterpret_cast<const uint32_t*>(smem + L.off_scratch + (size_t)ww * L.warp_scratch) + lane;
        const uint32_t mybase = smem_u32(smem + L.off_scratch) + (uint32_t)ww * L.warp_scratch + (uint32_t)lane * 4u;
        // my landing zone: field f of k bytes at slot + f * 32 * k + lane * k (the lane term is folded into the bases)
        const uint32_t slot = smem_u32(smem + L.off_slots) + (uint32_t)ww * kF5SlotWarpBytes;
        const uint32_t slot_lg = slot + kF5SlotLogits + lane * 16, slot_id = slot + kF5SlotIds + lane * 4,
                       slot_rf = slot + kF5SlotRefs + lane * 8;
        const uint32_t tm_lane = tm_a0 + ((uint32_t)(q * 32) << 16);          // my lane quarter, operand 0
        const float fSw = (float)Sw, fSh = (float)Sh;
        const float pix_bias = 8388608.f - (float)(Sw + 1);
        const uint32_t row_half = (uint32_t)(Sw >> 1) << 7, sw_odd = (uint32_t)Sw & 1u;
        const float2* rp2 = reinterpret_cast<const float2*>(rpc);

        // scratch row and (group 0 only: both groups own the same TMEM lanes) the three operands start out zero
        for (int w = 0; w < SP / 2; ++w) asm volatile("st.shared.b32 [%0], %1;" ::"r"(mybase + w * 128), "r"(0) : "memory");
        if (g == 0) {
            const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int o = 0; o < kF5Ops; ++o)
                for (int c = 0; c < nchunks; ++c) f5_tmem_st8(tm_lane + o * op_cols + c * 8, z);
            f5_tmem_st_wait();
        }
        tc_fence_before();
        f5_named_barrier(1, kF5Workers);
        tc_fence_after();

        // ---- prefetch pipeline: while item j is processed, its successor's logits / reference points and the
        // ids (voxel, camera mask, the two tile unions) of the item after that are in flight into my slot
        // (b, chunk, h) of item, item + gridDim.x, item + 2 gridDim.x, advanced without divisions
        const F5Item stride = f5_item((int)gridDim.x, NH, chunks_per_b);
        auto next_pos = [&](F5Item t) {
            t.h += stride.h;
            t.chunk += stride.chunk;
            t.b += stride.b;
            if (t.h >= NH) {
                t.h -= NH;
                ++t.chunk;
            }
            if (t.chunk >= chunks_per_b) {
                t.chunk -= chunks_per_b;
                ++t.b;
            }
            return t;
        };
        F5Item pos0 = f5_item((int)blockIdx.x, NH, chunks_per_b), pos1 = next_pos(pos0), pos2 = next_pos(pos1);
        auto issue_ids = [&](int item, const F5Item& t) {           // -> slot ids; stale when the item / tile / row does not exist
            if (item >= n_items) return;
            const int tile = 2 * t.chunk + g, i = tile * kF5Rows + r;
            const uint32_t* tu = tile_union + (size_t)t.b * tiles_per_b;
            if (tile < tiles_per_b) f5_cp4(slot_id + 2 * 128, tu + tile);
            if ((tile ^ 1) < tiles_per_b) f5_cp4(slot_id + 3 * 128, tu + (tile ^ 1));
            if (i < Nq) {
                f5_cp4(slot_id, order + (size_t)t.b * Nq + i);
                f5_cp4(slot_id + 128, smask + (size_t)t.b * Nq + i);
            }
        };
        auto issue_row = [&](int item, const F5Item& t, int n, uint32_t m) {      // logits + first reference points of `item`
            if (item >= n_items || n < 0) return;
            const float* row = logits + ((size_t)t.b * Nq + n) * ld;
            const float* po = row + t.h * NP * 2;
            const float* pl = row + NH * NP * 2 + t.h * NP;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i * 2 < NP) f5_cp16(slot_lg + i * 512, po + i * 4);
            f5_cp16(slot_lg + 4 * 512, pl);
            if (NP > 4) f5_cp16(slot_lg + 5 * 512, pl + 4);
            uint32_t rest = m;
#pragma unroll
            for (int k = 0; k < kF5SlotRefCams; ++k) {
                if (!rest) break;
                const int c = __ffs(rest) - 1;
                rest &= rest - 1;
                f5_cp8(slot_rf + k * 256, rp2 + ((size_t)c * B + t.b) * Nq + n);
            }
        };
        // ids of item `item` as (n, m, u_mine, u_other) from global memory (first item only)
        auto ids_direct = [&](int item, const F5Item& t, int& n, uint32_t& m, uint32_t& um, uint32_t& uo) {
            n = -1;
            m = um = uo = 0;
            if (item >= n_items) return;
            const int tile = 2 * t.chunk + g, i = tile * kF5Rows + r;
            const uint32_t* tu = tile_union + (size_t)t.b * tiles_per_b;
            if (tile < tiles_per_b) um = __ldg(tu + tile);
            if ((tile ^ 1) < tiles_per_b) uo = __ldg(tu + (tile ^ 1));
            if (i < Nq) {
                n = __ldg(order + (size_t)t.b * Nq + i);
                m = __ldg(smask + (size_t)t.b * Nq + i);
            }
        };

        F5Timer tw(tid == kF5FirstWorker);
        // ---- state of the item whose cameras are being walked
        int item = blockIdx.x, it_b = 0, n = -1, kvis = 0;
        uint32_t m = 0, u_mine = 0, u_oth = 0, rest = 0, Jbase = 0;
        float ox[8], oy[8], aw[8];
        float2 refs[kF5SlotRefCams];
        int n_nx;                                   // ids of item + gridDim.x (landed in the slot / loaded directly)
        uint32_t m_nx, um_nx, uo_nx;
        ids_direct(item, pos0, n_nx, m_nx, um_nx, uo_nx);
        issue_ids(item + gridDim.x, pos1);
        issue_row(item, pos0, n_nx, m_nx);
        bool item_loaded = false;                   // state above describes `item`
        // ---- the pending batch: its taps are in my scratch row, not yet in TMEM
        bool pending = false;
        uint32_t p_kmask = 0, p_J = 0;
        bool tapped = false;
        uint32_t ua[8], ub[8];                      // cell addresses (upper-left, lower-left) of the pending taps
        uint32_t walked = 0;                        // global positions [0, walked) passed by this warp
        // a position is passed once "batch j - 3 retired" has been observed; at a batch of the other group the
        // warp also arrives on bar_built (the control thread waits for all eight builder warps)
        auto observe = [&](uint32_t j) {
            if (j >= kF5Ops) f5_wait(&bar_mma[j % kF5Ops], ((j / kF5Ops) - 1) & 1, 7, j, (uint32_t)warp);
        };
        auto try_observe = [&](uint32_t j) -> bool {          // warp-uniform, never blocks
            if (j < kF5Ops) return true;
            uint32_t ok;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(smem_u32(&bar_mma[j % kF5Ops])), "r"(((j / kF5Ops) - 1) & 1)
                : "memory");
            return __all_sync(VER_FULL_MASK, ok != 0);
        };
        auto walk_foreign = [&](uint32_t upto) {
            for (uint32_t j = walked; j < upto; ++j) {
                observe(j);
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_built[j % kF5Ops]);
            }
            walked = upto;
        };
        tw.lap(0);                                  // setup

        while (true) {
            // ---- (1) pending batch: scratch -> registers -> TMEM (tcgen05.st issued, not waited for), un-tap
            uint32_t p_op = 0;
            if (pending) {
                p_op = p_J % kF5Ops;
                walk_foreign(p_J);                  // the other group's batches since my previous one
                observe(p_J);                       // the MMAs of batch J - 3 retired -> the operand is mine
                walked = p_J + 1;
                tc_fence_after();
                tw.lap(2);                          // wait: operand free
                const uint32_t copy = (g_f5_timing_on & 8) ? 0u : (p_kmask | s_kmask[p_op][q]);   // (& 8: debug)
                const uint32_t tm_row = tm_lane + p_op * op_cols;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    if (c < nchunks && ((copy >> c) & 1u)) {
                        uint32_t rr[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) rr[j] = my_words[(c * 8 + j) * 32];
                        f5_tmem_st8(tm_row + c * 8, rr);
                    }
                }
                asm volatile("" ::: "memory");      // the loads above stay above the un-tap stores
                if (tapped) {                       // un-tap: my scratch row is all zero again
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        const uint32_t step = (ua[p] & 2u) ? 126u : 2u;
                        f5_sts16(ua[p], 0);
                        f5_sts16(ua[p] + step, 0);
                        if (SW_EVEN) {
                            f5_sts16(ua[p] + row_half, 0);
                            f5_sts16(ua[p] + row_half + step, 0);
                        } else {
                            f5_sts16(ub[p], 0);
                            f5_sts16(ub[p] + ((ub[p] & 2u) ? 126u : 2u), 0);
                        }
                    }
                    tapped = false;
                }
                tw.lap(5);                          // copy issue + un-tap
                // publish: the tcgen05.st drained while the row was un-tapped
                f5_tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    s_kmask[p_op][q] = p_kmask;
                    mbar_arrive(&bar_built[p_op]);
                }
                tw.lap(6);                          // st drain + fences + arrive
            }
            // ---- (2) next batch of my group: next camera of the item, or the first camera of the next item
            bool have_next = false;
            int cam = 0;
            while (true) {
                if (rest) {
                    cam = __ffs(rest) - 1;
                    rest &= rest - 1;
                    have_next = true;
                    break;
                }
                if (item_loaded) {                  // leave the item
                    Jbase += __popc(u_mine) + __popc(u_oth);
                    item += gridDim.x;
                    pos0 = pos1;
                    pos1 = pos2;
                    pos2 = next_pos(pos2);
                    item_loaded = false;
                }
                if (item >= n_items) break;
                // ---- item top: my slot holds this item's logits / reference points and the next item's ids
                it_b = pos0.b;
                n = n_nx;
                m = m_nx;
                u_mine = um_nx;
                u_oth = uo_nx;
                f5_cp_wait_all();
                {
                    float4 raw[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) raw[i] = f5_lds_f4(slot_lg + i * 512);
#pragma unroll
                    for (int k = 0; k < kF5SlotRefCams; ++k) refs[k] = f5_lds_f2(slot_rf + k * 256);
                    // ids of the next item (stale slot words are masked by existence)
                    const int nitem = item + gridDim.x;
                    n_nx = -1;
                    m_nx = um_nx = uo_nx = 0;
                    if (nitem < n_items) {
                        const int tile = 2 * pos1.chunk + g;
                        if (tile < tiles_per_b) um_nx = f5_lds_u32(slot_id + 2 * 128);
                        if ((tile ^ 1) < tiles_per_b) uo_nx = f5_lds_u32(slot_id + 3 * 128);
                        if (tile * kF5Rows + r < Nq) {
                            n_nx = (int)f5_lds_u32(slot_id);
                            m_nx = f5_lds_u32(slot_id + 128);
                        }
                    }
                    // prefetch: ids of the item after next, logits / reference points of the next item
                    issue_ids(item + 2 * gridDim.x, pos2);
                    issue_row(nitem, pos1, n_nx, m_nx);
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
                rest = u_mine;
                kvis = 0;
                item_loaded = true;
                tw.lap(1);                          // item top: slot -> registers, prefetch issue, softmax
            }
            uint32_t kmask = 0, J = 0;
            if (have_next) {
                J = Jbase + (g ? f5_batch_index(u_oth, u_mine, cam, 1) : f5_batch_index(u_mine, u_oth, cam, 0));
                // positions of the other group in front of my next batch: pass them now if that does not block
                // (the issuing thread waits for all eight builder warps at every batch)
                while (walked < J && try_observe(walked)) {
                    if (lane == 0) mbar_arrive(&bar_built[walked % kF5Ops]);
                    ++walked;
                }
                // ---- taps of (my row, cam) into the scratch row.  floor() through the 2^23 trick (add with
                // round-down): no conversion-pipe instructions.  The base cell is clamped into the map; a corner
                // that is outside gets weight 0 on a real cell (tent weights, branch free).
                if ((m >> cam) & 1u) {
                    float2 ref;
                    if (kvis < kF5SlotRefCams) {
                        ref = refs[0];
#pragma unroll
                        for (int k = 1; k < kF5SlotRefCams; ++k)
                            if (kvis == k) ref = refs[k];
                    } else {
                        ref = __ldg(rp2 + ((size_t)cam * B + it_b) * Nq + n);
                    }
                    ++kvis;
                    const float rx1 = fmaf(ref.x, fSw, 0.5f), ry1 = fmaf(ref.y, fSh, 0.5f);      // pixel coordinate + 1
                    int pmin = 0x7fffffff, pmax = 0;
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        // t = pixel coordinate + 1; c = clamp(floor(t), 1, S - 1) = left / upper cell + 1 of a 2x2
                        // block inside the map; cell weight = tent max(0, 1 - |coordinate - cell|)
                        const float tx = rx1 + ox[p], ty = ry1 + oy[p];
                        const float flx = __fadd_rd(tx, 8388608.f) - 8388608.f, fly = __fadd_rd(ty, 8388608.f) - 8388608.f;
                        const float cx = fminf(fmaxf(flx, 1.f), fSw - 1.f), cy = fminf(fmaxf(fly, 1.f), fSh - 1.f);
                        const float dx = tx - cx, dy = ty - cy;
                        const float a = aw[p];
                        const float wxa = __saturatef(1.f - fabsf(dx)), wxb = __saturatef(1.f - fabsf(dx - 1.f));
                        const float wya = a * __saturatef(1.f - fabsf(dy)), wyb = a * __saturatef(1.f - fabsf(dy - 1.f));
                        const __half2 wa = __floats2half2_rn(wya * wxa, wya * wxb);
                        const __half2 wb = __floats2half2_rn(wyb * wxa, wyb * wxb);
                        // pix = (cy - 1) * Sw + (cx - 1), exact small integer in fp32 -> int through the 2^23 trick
                        const int pix = __float_as_int(fmaf(cy, fSw, cx) + pix_bias) - 0x4B000000;
                        pmin = min(pmin, pix);
                        pmax = max(pmax, pix);
                        // right neighbour of cell k: same word (+2) if k is even, next word (+126) if odd; the cell
                        // below is Sw cells on: Sw / 2 words, plus one more cell if Sw is odd
                        const uint32_t odd = (uint32_t)pix & 1u;
                        const uint32_t step0 = odd ? 126u : 2u;
                        const uint32_t a0 = mybase + (((uint32_t)pix >> 1) << 7) + (odd << 1), a0r = a0 + step0;
                        uint32_t a1, a1r;
                        if (SW_EVEN) {
                            a1 = a0 + row_half;
                            a1r = a1 + step0;
                        } else {
                            a1 = a0 + row_half + sw_odd * step0;
                            a1r = a1 + ((odd ^ sw_odd) ? 126u : 2u);
                        }
                        ua[p] = a0;
                        ub[p] = a1;
                        const uint16_t h0 = f5_lds16(a0), h1 = f5_lds16(a0r), h2 = f5_lds16(a1), h3 = f5_lds16(a1r);
                        f5_sts16(a0, __half_as_ushort(__hadd(__ushort_as_half(h0), __low2half(wa))));
                        f5_sts16(a0r, __half_as_ushort(__hadd(__ushort_as_half(h1), __high2half(wa))));
                        f5_sts16(a1, __half_as_ushort(__hadd(__ushort_as_half(h2), __low2half(wb))));
                        f5_sts16(a1r, __half_as_ushort(__hadd(__ushort_as_half(h3), __high2half(wb))));
                    }
                    // chunks (16 cells) between the first and the last tapped cell of the row
                    kmask = (2u << ((uint32_t)(pmax + Sw + 1) >> 4)) - (1u << ((uint32_t)pmin >> 4));
                    tapped = true;
                }
                kmask = __reduce_or_sync(VER_FULL_MASK, kmask);
                tw.lap(4);                          // taps: arithmetic + read-modify-writes
            }
            if (!have_next) break;
            pending = true;
            p_kmask = kmask;
            p_J = J;
        }
        walk_foreign(Jbase);                        // the other group's batches after my last one
        tw.lap(7);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 512);
    if (*(volatile unsigned int*)&g_f5_abort && !(g_f5_timing_on & 2)) __trap();     // fail loudly (see f5_wait)
}

template <int DH, int NP>
int launch_fwd_tc5(const __half* vimg, const float* logits, int ld, const float* rpc, const int32_t* order,
                   const uint32_t* smask, const uint32_t* tile_union, __half* slots, int B, int Ncam, int Nq,
                   int Sh, int Sw, int SP, int NH, cudaStream_t st) {
    const F5Smem L(DH, SP);
    const int chunks_per_b = (Nq + kF5ChunkRows - 1) / kF5ChunkRows;
    const int n_items = B * NH * chunks_per_b;
    const int sms = ver_device_sm_count();
    const int grid = n_items < sms ? n_items : sms;
    if (Sw % 2 == 0) {
        auto kern = sca_fwd_tc5_kernel<DH, NP, true>;
        VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        kern<<<grid, kF5Threads, L.total, st>>>(vimg, logits, ld, rpc, order, smask, tile_union, slots, B, Ncam, Nq,
                                                Sh, Sw, SP, NH, chunks_per_b, n_items);
    } else {
        auto kern = sca_fwd_tc5_kernel<DH, NP, false>;
        VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        kern<<<grid, kF5Threads, L.total, st>>>(vimg, logits, ld, rpc, order, smask, tile_union, slots, B, Ncam, Nq,
                                                Sh, Sw, SP, NH, chunks_per_b, n_items);
    }
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

}  // namespace

extern "C" int ver_debug_tc5_timing(int enable, unsigned long long* host_out32) {
    if (host_out32) VER_CHECK_CUDA(cudaMemcpyFromSymbol(host_out32, g_f5_timing, sizeof(unsigned long long) * 32));
    unsigned long long zero[32] = {0};
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f5_timing, zero, sizeof(zero)));
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f5_timing_on, &enable, sizeof(int)));
    return VER_OK;
}

// record of the first timed-out wait of sca_fwd_tc5_kernel: code (1 built, 2 V, 3 accumulator free, 4/5 V buffer,
// 6 accumulator full, 7 operand retired), block, thread, two wait-specific words; all zero if none.  Resets it.
extern "C" int ver_debug_tc5_diag(unsigned int* host_out8) {
    unsigned int abort_flag = 0, zero[8] = {0};
    VER_CHECK_CUDA(cudaMemcpyFromSymbol(&abort_flag, g_f5_abort, sizeof(abort_flag)));
    if (host_out8) VER_CHECK_CUDA(cudaMemcpyFromSymbol(host_out8, g_f5_diag, sizeof(zero)));
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f5_diag, zero, sizeof(zero)));
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f5_abort, zero, sizeof(unsigned int)));
    return (int)abort_flag;
}

// shapes the three-operand kernel covers: the rest of ver_tc4_supported's shapes take sca_fwd_tc4_kernel
int ver_tc5_supported(int Ncam, int S, int Dh, int NP) {
    if (!(Ncam <= 32 && (NP == 4 || NP == 8) && S <= 256 && (Dh == 32 || Dh == 64 || Dh == 96))) return 0;
    const int SP = (S + 15) / 16 * 16;
    if (2 * Dh + kF5Ops * (SP / 2) > 512) return 0;    // TMEM columns: two accumulators + three A operands
    return F5Smem(Dh, SP).total + 1024 <= ver_device_max_smem_optin();
}

int ver_sca_forward_tc5(const void* vimg, const float* logits, int ld, const float* rpc, const int32_t* order,
                        const uint32_t* smask, const uint32_t* tile_union, void* slots, int B, int Ncam, int Nq,
                        int Sh, int Sw, int NH, int Dh, int NP, cudaStream_t st) {
    const int SP = (Sh * Sw + 15) / 16 * 16;
#define FWD5(D)                                                                                                  \
    (NP == 8 ? launch_fwd_tc5<D, 8>((const __half*)vimg, logits, ld, rpc, order, smask, tile_union, (__half*)slots, B, \
                                    Ncam, Nq, Sh, Sw, SP, NH, st)                                                 \
             : launch_fwd_tc5<D, 4>((const __half*)vimg, logits, ld, rpc, order, smask, tile_union, (__half*)slots, B, \
                                    Ncam, Nq, Sh, Sw, SP, NH, st))
    switch (Dh) {
        case 32: return FWD5(32);
        case 64: return FWD5(64);
        default: return FWD5(96);
    }
#undef FWD5
}

// Fused SCA sampler forward, sixth generation: rows of the interpolation matrix A are built STRAIGHT INTO the
// shared-memory tensor-core operand by two threads per row, in an image layout made for the builder.
//
//     slots[b, n, h, :] = 1/max(count,1) * sum_{cam sees n, ascending} A_cam[n, :] V_{b,cam,h}[:, :]
//
// Replaces SpatialCrossAttention.forward's rebatch / sampling / scatter-mean
// (M/spatial_cross_attention.py:138-173, MSDeformableAttention3D :340-374).  Same tiling as sca_tc4.cu
// (visibility-sorted 128-row tiles, persistent CTAs over (panorama, 256-row chunk, head) items, tcgen05 with fp32
// accumulators in TMEM).  What the measurements of generations 4 and 5 showed (profiles/r02b, r02i): the builder
// code alone runs a whole launch's rows in 129 us (tools/tap_bench.cu), the kernels took 410-440 us -- the time was
// in the couplings around the TMEM A operand (tcgen05.st queueing behind the other group's MMA batch, the wait for
// the group's own batch before the copy) and in latency-bound 16-bit read-modify-writes.  This generation:
//   * K index = 16 cells per image row: cell X = pixel x + 1, cells 0 and Sw + 1 are zero padding (the V image carries
//     zeros there, ver_value_image16_f16).  A sampling point is clamped to [0, Sw + 1] and its bilinear footprint
//     is the three cells e, e + 1, e + 2 of the ALIGNED pair base e = 2 floor(X / 2): weights max(0, 1 - u),
//     1 - |u - 1|, max(0, u - 1) with u = X - e -- no corner cases, no cell clamping, and every update is a 32-bit
//     read-modify-write of an aligned fp16 pair (half the shared-memory instructions, no pack / unpack);
//   * two threads per row, split by IMAGE-ROW PARITY: thread pi owns image rows 2 j + pi.  A point touches two
//     consecutive image rows, one of each parity, so the two threads never touch the same cell; the row
//     of parity pi that carries weight is j = floor((Y - pi) / 2), its weight the tent max(0, 1 - |Y - 1 - (2 j + pi)|)
//     which is 0 whenever j had to be clamped.  16 builder warps instead of 8 hide the read-modify-write latency;
//   * A is written in the canonical no-swizzle UMMA layout (core matrix = 8 rows x 16 B; image row y = K groups
//     2 y, 2 y + 1), the MMA reads it from shared memory: no scratch -> TMEM copy, no tcgen05.st, K chunk = image row;
//   * two accumulators per row group in TMEM (4 x Dh columns): the next item's first MMA never waits for the epilogue;
//   * every mbarrier wait is bounded (f6_wait): a protocol error fails the launch instead of hanging the GPU.
// MEASURED (profiles/r02i_tc6_check.txt, B200, 8 x 18 views, 16x40x40): correct on every test shape (3e-4 of
// sca_fwd_tc4_kernel), but 564 us per launch against 407 us -- NOT the default.  Two reasons, both inherent to a
// tensor-core-readable A tile: (1) in every UMMA shared-memory layout a row owns 16 contiguous bytes per K group, so
// the 32 rows of a warp map onto 8 bank groups and rows r, r + 8, r + 16, r + 24 collide whenever they update the
// same cell pair -- which neighbouring voxels do: 4-way conflicts on every read-modify-write (the lane-interleaved
// scratch of generation 4 is conflict free by construction); (2) with A read from shared memory an MMA of this
// size takes ~200 cycles under the builders' shared-memory traffic (phase timers: control "MMA issue" 512 k of 990 k
// cycles per CTA, builders 590 k cycles waiting for their batch to retire) against ~150 with A in TMEM.
// It stays selectable (ver_sca_forward_sorted16, ops.TC_FORWARD = 'sorted16') with its parity tests.
//
// Roles: warps 0-7 = builders of even image rows (warp w: rows 32 w .. 32 w + 31 of the chunk), warps 8-15 = builders of
// odd image rows (warp 8 + w: the same rows), warps 16-19 = epilogue (TMEM lane quarter = warp % 4), warp 20 = control
// (one lane): TMA of the value images, MMA issue.  Row group g = rows 128 g .. 128 g + 127 = UMMA M.
// Hand-offs (mbarriers):
//     bar_built[g]       group g wrote A_g for its next camera                         (8 arrivals, one per warp)
//     bar_mma[g]         tcgen05.commit: the MMAs reading A_g retired -> A_g may be rewritten
//     bar_full[g][a]     tcgen05.commit after the item's last camera: accumulator a of group g is complete
//     bar_free[g][a]     the epilogue warps drained accumulator a of group g           (4 arrivals)
//     bar_v[buf] / bar_vfree[buf]   value image landed (transaction bytes) / all MMAs reading it retired
//   named barrier 1 + w (64 threads): the two builder warps of rows 32 w .. share one prefetch slot (logits of the item)
#include "sampler.cuh"
#include "tap16.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int kF6Builders = 512;
constexpr int kF6Threads = kF6Builders + 128 + 32;
constexpr int kF6Rows = 128;                  // rows per group = UMMA M
constexpr int kF6ChunkRows = 2 * kF6Rows;
constexpr int kF6SlotUnits = 6;               // 16-byte units per row: 4 x offsets (8 points x 2), 2 x attention logits
constexpr int kF6SlotWarpBytes = kF6SlotUnits * 512;
constexpr float kF6Magic = 8388608.f;         // 2^23: floor() and float -> int through round-down adds

struct F6Smem {
    int v_bytes, tile, off_v[2], off_a, off_dummy, off_slots, total;
    __host__ __device__ F6Smem(int Dh, int Sh) {
        v_bytes = Dh * Sh * 16 * 2;           // [Dh / 8][2 Sh][8][8] halves
        tile = Sh * 4096;                     // [2 Sh K groups][16 row groups][8 rows][16 B]
        off_v[0] = 0;
        off_v[1] = v_bytes;
        off_a = 2 * v_bytes;
        off_dummy = off_a + 2 * tile;         // one word per builder thread: sink of the always-zero second pair at X = Sw + 1
        off_slots = off_dummy + kF6Builders * 4;
        total = off_slots + 8 * kF6SlotWarpBytes;
    }
};

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

struct F6Item {
    int b, chunk, h;
};
__device__ __forceinline__ F6Item f6_item(int item, int NH, int chunks_per_b) {
    F6Item it;
    it.h = item % NH;
    const int r = item / NH;
    it.chunk = r % chunks_per_b;
    it.b = r / chunks_per_b;
    return it;
}

// Bounded mbarrier wait (as sca_tc5.cu's): after kF6WaitLimit failed try_waits the waiter records what it was waiting
// for, raises g_f6_abort and returns; every other wait of the grid returns at its next wake-up, the kernel runs off its
// (garbage) end and traps there: the launch FAILS instead of hanging.  ver_debug_tc6(…) reads the record.
__device__ unsigned int g_f6_abort = 0;
__device__ unsigned int g_f6_diag[8];
__device__ int g_f6_flags = 0;                         // bit 0: phase timers, bit 1: do not trap on a failed wait
constexpr unsigned int kF6WaitLimit = 200000;
__device__ __noinline__ void f6_wait_failed(uint32_t code, uint32_t a, uint32_t b) {
    if (atomicExch(&g_f6_abort, 1u) == 0u) {
        g_f6_diag[0] = code;
        g_f6_diag[1] = blockIdx.x;
        g_f6_diag[2] = threadIdx.x;
        g_f6_diag[3] = a;
        g_f6_diag[4] = b;
        __threadfence();
    }
}
__device__ __forceinline__ void f6_wait(uint64_t* bar, uint32_t parity, uint32_t code, uint32_t a, uint32_t b) {
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
        if (tries >= 64 && *(volatile unsigned int*)&g_f6_abort) return;
        if (tries >= kF6WaitLimit) {
            f6_wait_failed(code, a, b);
            return;
        }
    }
}

// phase timers (debug; enabled through ver_debug_tc6, read by tools/tc_timing.py, never by the product)
__device__ unsigned long long g_f6_timing[32];
struct F6Timer {
    bool on;
    long long t;
    __device__ __forceinline__ F6Timer(bool active) : on(active && (g_f6_flags & 1)), t(0) {
        if (on) t = clock64();
    }
    __device__ __forceinline__ void lap(int slot) {
        if (on) {
            const long long n = clock64();
            atomicAdd(&g_f6_timing[slot], (unsigned long long)(n - t));
            t = n;
        }
    }
};

template <int DH, int NP>
__global__ void __launch_bounds__(kF6Threads, 1)
sca_fwd_tc6_kernel(const __half* __restrict__ vimg, const float* __restrict__ logits, int ld,
                   const float* __restrict__ rpc, const int32_t* __restrict__ order,
                   const uint32_t* __restrict__ smask, const uint32_t* __restrict__ tile_union,
                   __half* __restrict__ slots, int B, int Ncam, int Nq, int Sh, int Sw, int NH,
                   int chunks_per_b, int n_items) {
    extern __shared__ __align__(128) unsigned char smem[];
    const F6Smem L(DH, Sh);
    __shared__ __align__(8) uint64_t bar_built[2], bar_mma[2], bar_full[2][2], bar_free[2][2], bar_v[2], bar_vfree[2];
    __shared__ uint32_t s_tmem;
    __shared__ volatile uint32_t s_kmask[2][2][8];     // [group][batch parity][warp of the group]: image rows that hold taps

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_built[i], 8);
            mbar_init(&bar_mma[i], 1);
            mbar_init(&bar_v[i], 1);
            mbar_init(&bar_vfree[i], 1);
            for (int a = 0; a < 2; ++a) {
                mbar_init(&bar_full[i][a], 1);
                mbar_init(&bar_free[i][a], 4);
            }
        }
        mbar_fence_init();
    }
    if (warp == 20) tmem_alloc(&s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int tiles_per_b = (Nq + kF6Rows - 1) / kF6Rows;      // order.cu's tile_union row length
    const size_t v_elems = (size_t)DH * Sh * 16;               // halves per (view, head) image
    const int G = 2 * Sh;                                      // 8-cell K groups of an image

    if (warp == 20) {
        // ================================================================ control: TMA + MMA issue
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(128, DH, 0, 0);
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
                    const F6Item it = f6_item(nx_item, NH, chunks_per_b);
                    nx_b = it.b;
                    nx_h = it.h;
                    const uint32_t* tu = tile_union + (size_t)it.b * tiles_per_b + 2 * it.chunk;
                    nx_u0 = tu[0];
                    nx_u1 = (2 * it.chunk + 1 < tiles_per_b) ? tu[1] : 0u;
                    nx_rest = nx_u0 | nx_u1;
                }
            };
            auto load_v = [&](int buf) {
                mbar_expect_tx(&bar_v[buf], L.v_bytes);
                bulk_g2s(smem + L.off_v[buf], vimg + ((size_t)(nx_b * Ncam + nx_cam) * NH + nx_h) * v_elems,
                         L.v_bytes, &bar_v[buf]);
            };
            F6Timer tc(true);
            bool has_next = advance();
            if (has_next) load_v(0);
            uint32_t kk = 0, itg[2] = {0, 0}, acc_items[2] = {0, 0};
            // descriptors of image row 0: A tile (K stride 2048 B between the two core matrices of an image row,
            // M stride 128 B between 8-row groups), V image (K stride 128 B, N stride G * 128 B)
            const uint64_t da0[2] = {umma_desc(smem_u32(smem + L.off_a), 2048, 128),
                                     umma_desc(smem_u32(smem + L.off_a + L.tile), 2048, 128)};
            const uint64_t dv0[2] = {umma_desc(smem_u32(smem + L.off_v[0]), 128, G * 128),
                                     umma_desc(smem_u32(smem + L.off_v[1]), 128, G * 128)};
            while (has_next) {
                const int cam = nx_cam;
                const uint32_t u[2] = {nx_u0, nx_u1};
                has_next = advance();                  // nx_* now describe step kk + 1
                const int buf = kk & 1;
                bool first = true;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (!((u[g] >> cam) & 1u)) continue;
                    const bool first_cam = !(u[g] & ((1u << cam) - 1u));       // lowest camera of this tile overwrites
                    const bool last_cam = !(u[g] >> (cam + 1));
                    const uint32_t ab = acc_items[g] & 1u;                     // accumulator of this item
                    f6_wait(&bar_built[g], itg[g] & 1, 1, kk, g);
                    tc.lap(9);                         // control: wait for a built A
                    if (first) f6_wait(&bar_v[buf], (kk >> 1) & 1, 2, kk, 0);
                    tc.lap(10);                        // control: wait for the value image
                    if (first_cam && acc_items[g] >= 2) f6_wait(&bar_free[g][ab], ((acc_items[g] >> 1) - 1) & 1, 3, kk, g);
                    tc_fence_after();
                    tc.lap(13);                        // control: wait for a drained accumulator
                    const int par = itg[g] & 1;
                    uint32_t km = 0;
#pragma unroll
                    for (int w = 0; w < 8; ++w) km |= s_kmask[g][par][w];
                    uint32_t acc = first_cam ? 0u : 1u;
                    if (!acc && !km) km = 1u;          // (an all-zero image row zeroes the accumulator)
                    const uint32_t d_addr = tmem + (g * 2 + ab) * DH;
                    for (int y = 0; y < Sh; ++y) {
                        if ((km >> y) & 1u) {
                            umma_f16(d_addr, da0[g] + (uint64_t)(y * 256), dv0[buf] + (uint64_t)(y * 16), idesc, acc);
                            acc = 1u;
                        }
                    }
                    umma_commit(&bar_mma[g]);
                    if (last_cam) {
                        umma_commit(&bar_full[g][ab]);
                        ++acc_items[g];
                    }
                    tc.lap(11);                        // control: MMA issue
                    ++itg[g];
                    if (first && has_next) {
                        // value image of step kk + 1 -> the other buffer, once step kk - 1 stopped reading it
                        if (kk >= 1) f6_wait(&bar_vfree[(kk + 1) & 1], ((kk - 1) >> 1) & 1, 4, kk, 0);
                        load_v((kk + 1) & 1);
                        tc.lap(12);                    // control: wait for a free value buffer + TMA issue
                    }
                    first = false;
                }
                umma_commit(&bar_vfree[buf]);
                ++kk;
            }
            // drain: the last commits must have arrived before the CTA tears TMEM / smem down
            if (kk >= 1) f6_wait(&bar_vfree[(kk - 1) & 1], ((kk - 1) >> 1) & 1, 5, kk, 0);
        }
    } else if (warp >= 16) {
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
                const F6Item it = f6_item(item, NH, chunks_per_b);
                const int tile = 2 * it.chunk + g, i = tile * kF6Rows + r;
                if (tile < tiles_per_b) u_nx[g] = __ldg(tile_union + (size_t)it.b * tiles_per_b + tile);
                if (i < Nq) {
                    n_nx[g] = __ldg(order + (size_t)it.b * Nq + i);
                    m_nx[g] = __ldg(smask + (size_t)it.b * Nq + i);
                }
            }
        };
        F6Timer te(tid == kF6Builders);
        load_ids(blockIdx.x);
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const F6Item it = f6_item(item, NH, chunks_per_b);
            const int n[2] = {n_nx[0], n_nx[1]};
            const uint32_t m[2] = {m_nx[0], m_nx[1]}, u[2] = {u_nx[0], u_nx[1]};
            load_ids(item + gridDim.x);                 // in flight during this item's epilogue
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const float inv_cnt = 1.f / (float)max(__popc(m[g]), 1);
                __half* dst = (n[g] >= 0) ? slots + (((size_t)it.b * Nq + n[g]) * NH + it.h) * DH : nullptr;
                if (u[g]) {                             // warp-uniform (tile property)
                    const uint32_t ab = full_seen[g] & 1u;
                    f6_wait(&bar_full[g][ab], (full_seen[g] >> 1) & 1, 6, (uint32_t)item, g);
                    ++full_seen[g];
                    tc_fence_after();
                    te.lap(16);                         // epilogue: wait for a complete accumulator
#pragma unroll
                    for (int c0 = 0; c0 < DH; c0 += 32) {
                        float vv[32];
                        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (g * 2 + ab) * DH + c0, vv);
                        if (c0 + 32 >= DH) {                       // last read of this accumulator: hand it back
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&bar_free[g][ab]);
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
        const uint32_t a_row = smem_u32(smem + L.off_a) + (uint32_t)g * L.tile + (uint32_t)(rr >> 3) * 128u + (uint32_t)(rr & 7) * 16u;
        const uint32_t dummy = smem_u32(smem + L.off_dummy) + (uint32_t)tid * 4u;
        const uint32_t slot = smem_u32(smem + L.off_slots) + (uint32_t)rw * kF6SlotWarpBytes + (uint32_t)lane * 16u;
        const float fSw = (float)Sw, fSh = (float)Sh, fpi = (float)pi;
        const float xmax = (float)(Sw + 1);
        const float jtop = kF6Magic + (float)((Sh - pi + 1) / 2 - 1);       // last image row of my parity, magic domain
        const uint32_t a_par = a_row + (uint32_t)pi * 4096u;                // image row 2 j + pi at + j * 8192
        const float2* rp2 = reinterpret_cast<const float2*>(rpc);
        uint32_t it = 0, seen = 0;                // MMA batches handed over / observed retired (this group)
        bool tapped = false;                      // my cells hold the taps recorded in ua / ub
        uint32_t ua[8], ub[8];                    // addresses of the two pairs each point updated

        // my cells of the A tile start out all zero
        for (int y = pi; y < Sh; y += 2)
            for (int c = 0; c < 8; ++c) sts32(a_row + (uint32_t)(2 * y + (c >> 2)) * 2048u + (uint32_t)(c & 3) * 4u, 0u);
        sts32(dummy, 0u);

        auto load_ids = [&](int item, int& n, uint32_t& m, uint32_t& u) {
            n = -1;
            m = u = 0;
            if (item >= n_items) return;
            const F6Item q = f6_item(item, NH, chunks_per_b);
            const int tile = 2 * q.chunk + g, i = tile * kF6Rows + rr;
            if (tile < tiles_per_b) u = __ldg(tile_union + (size_t)q.b * tiles_per_b + tile);
            if (i < Nq) {
                n = __ldg(order + (size_t)q.b * Nq + i);
                m = __ldg(smask + (size_t)q.b * Nq + i);
            }
        };
        auto issue_row = [&](int item, int n) {                 // logits of (row, head) of `item` -> the pair's slot
            if (item >= n_items || n < 0) return;
            const F6Item q = f6_item(item, NH, chunks_per_b);
            const float* row = logits + ((size_t)q.b * Nq + n) * ld;
            const float* po = row + q.h * NP * 2;
            const float* pl = row + NH * NP * 2 + q.h * NP;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i * 2 < NP) cp_async16(slot + i * 512, po + i * 4);
            cp_async16(slot + 4 * 512, pl);
            if (NP > 4) cp_async16(slot + 5 * 512, pl + 4);
        };

        F6Timer tw(tid == 0);
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
            const F6Item q = f6_item(item, NH, chunks_per_b);
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
                        const F6Item qn = f6_item(item + gridDim.x, NH, chunks_per_b);
                        ref_nx = __ldg(rp2 + ((size_t)cam2 * B + qn.b) * Nq + n_nx);
                    }
                    ref_item = item + gridDim.x;
                }
                // ---- the MMAs of my group's previous batch retired -> A_g is ours again
                if (seen < it) {
                    f6_wait(&bar_mma[g], seen & 1, 7, it, (uint32_t)warp);
                    ++seen;
                    tc_fence_after();
                }
                tw.lap(2);                              // wait: my previous MMA batch retired
                if (tapped) {                           // un-tap: my cells are all zero again
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        sts32(ua[p], 0u);
                        sts32(ub[p], 0u);
                    }
                    tapped = false;
                }
                tw.lap(3);                              // un-tap
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
                        const uint32_t c = __float_as_uint(hh) & 15u, j = __float_as_uint(jm) & 15u;
                        const uint32_t rowoff = a_par + j * 8192u;
                        const uint32_t ad = rowoff + ((c & 4u) << 9) + ((c & 3u) << 2);
                        const uint32_t c1 = c + 1u;
                        const uint32_t ad2 = (c == 7u) ? dummy : rowoff + ((c1 & 4u) << 9) + ((c1 & 3u) << 2);
                        kmask |= 1u << (2u * j);
                        ua[p] = ad;
                        ub[p] = ad2;
                        const uint32_t v0 = lds32(ad), v1 = lds32(ad2);
                        sts32(ad, h2_bits(__hadd2(bits_h2(v0), h01)));
                        sts32(ad2, h2_bits(__hadd2(bits_h2(v1), h2)));
                    }
                    tapped = true;
                }
                tw.lap(4);                              // taps: arithmetic + read-modify-writes
                kmask = __reduce_or_sync(VER_FULL_MASK, kmask) << pi;          // bit y: image row y holds taps
                proxy_fence();                          // generic-proxy writes of A -> async proxy (tensor core)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    s_kmask[g][it & 1][(rw & 3) + 4 * pi] = kmask;
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
    if (*(volatile unsigned int*)&g_f6_abort && !(g_f6_flags & 2)) __trap();      // fail loudly (see f6_wait)
}

// value [Bv][S = Sh Sw][NH][Dh] fp16  ->  vimg [Bv][NH][Dh / 8][2 Sh][8 ch][8 cells] fp16 with cell k = y * 16 + x + 1,
// cells 0 and Sw + 1 .. 15 of every image row zero
__global__ void value_image16_kernel(const __half* __restrict__ value, __half* __restrict__ vimg, int Bv, int Sh, int Sw,
                                     int NH, int Dh) {
    const int G = 2 * Sh;
    const size_t chunks = (size_t)Bv * NH * (Dh / 8) * G * 8;        // 16-byte chunks: 8 cells of one channel
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < chunks; i += (size_t)gridDim.x * blockDim.x) {
        const int c8 = i & 7;
        size_t r = i >> 3;
        const int kg = r % G;
        r /= G;
        const int cg = r % (Dh / 8);
        r /= (Dh / 8);
        const int h = r % NH;
        const int bv = r / NH;
        const int ch = cg * 8 + c8, y = kg >> 1, x0 = (kg & 1) * 8 - 1;
        __half out[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int x = x0 + p;
            out[p] = (x >= 0 && x < Sw) ? value[(((size_t)bv * Sh * Sw + y * Sw + x) * NH + h) * Dh + ch] : __float2half(0.f);
        }
        *reinterpret_cast<uint4*>(vimg + i * 8) = *reinterpret_cast<const uint4*>(out);
    }
}

template <int DH, int NP>
int launch_fwd_tc6(const __half* vimg, const float* logits, int ld, const float* rpc, const int32_t* order,
                   const uint32_t* smask, const uint32_t* tile_union, __half* slots, int B, int Ncam, int Nq,
                   int Sh, int Sw, int NH, cudaStream_t st) {
    const F6Smem L(DH, Sh);
    auto kern = sca_fwd_tc6_kernel<DH, NP>;
    VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    const int chunks_per_b = (Nq + kF6ChunkRows - 1) / kF6ChunkRows;
    const int n_items = B * NH * chunks_per_b;
    const int sms = ver_device_sm_count();
    const int grid = n_items < sms ? n_items : sms;
    kern<<<grid, kF6Threads, L.total, st>>>(vimg, logits, ld, rpc, order, smask, tile_union, slots, B, Ncam, Nq, Sh,
                                            Sw, NH, chunks_per_b, n_items);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

}  // namespace

// debug hook of tools/ and tests/: flags bit 0 = phase timers on, bit 1 = report a failed wait instead of trapping;
// returns the phase timers and the record of the first failed wait (diag[0] != 0), and clears both
extern "C" int ver_debug_tc6(int flags, unsigned long long* timing32, unsigned int* diag8) {
    unsigned int abort_flag = 0;
    VER_CHECK_CUDA(cudaMemcpyFromSymbol(&abort_flag, g_f6_abort, sizeof(abort_flag)));
    if (timing32) VER_CHECK_CUDA(cudaMemcpyFromSymbol(timing32, g_f6_timing, sizeof(unsigned long long) * 32));
    if (diag8) {
        VER_CHECK_CUDA(cudaMemcpyFromSymbol(diag8, g_f6_diag, sizeof(unsigned int) * 8));
        if (!abort_flag) diag8[0] = 0;
    }
    unsigned long long zero[32] = {0};
    unsigned int zero8[8] = {0};
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f6_timing, zero, sizeof(zero)));
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f6_diag, zero8, sizeof(zero8)));
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f6_abort, zero8, sizeof(unsigned int)));
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f6_flags, &flags, sizeof(int)));
    return abort_flag ? 1 : 0;
}

// shapes the two-threads-per-row kernel covers: image rows of at most 14 pixels (16 cells with the zero padding)
extern "C" int ver_tc6_supported(int Ncam, int Sh, int Sw, int Dh, int NP) {
    if (!(Ncam <= 32 && (NP == 4 || NP == 8) && Sh >= 2 && Sh <= 14 && Sw >= 2 && Sw <= 14 &&
          (Dh == 32 || Dh == 64 || Dh == 96)))
        return 0;
    return F6Smem(Dh, Sh).total + 1024 <= ver_device_max_smem_optin();
}

extern "C" int ver_value_image16_f16(const void* value, void* vimg, int Bv, int Sh, int Sw, int NH, int Dh,
                                     ver_stream_t stream) {
    VER_CHECK_ARG(value && vimg, "null pointer");
    VER_CHECK_ARG(Bv > 0 && Sh > 0 && Sw > 0 && Sw <= 14 && NH > 0 && Dh > 0 && Dh % 8 == 0, "bad dims (Sw <= 14)");
    const size_t chunks = (size_t)Bv * NH * (Dh / 8) * (2 * Sh) * 8;
    const int blocks = (int)((chunks + 255) / 256 > 148 * 32 ? 148 * 32 : (chunks + 255) / 256);
    value_image16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)value, (__half*)vimg, Bv, Sh, Sw, NH, Dh);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

extern "C" int ver_sca_forward_sorted16(const void* vimg16, const float* logits, int ld_logits, const float* rpc,
                                        const int32_t* order, const uint32_t* smask, const uint32_t* tile_union,
                                        void* slots, int B, int Ncam, int Nq, int Sh, int Sw, int NH, int Dh, int NP,
                                        int variant, ver_stream_t stream) {
    VER_CHECK_ARG(vimg16 && logits && rpc && order && smask && tile_union && slots, "null pointer");
    VER_CHECK_ARG(variant == 0 || variant == 6 || variant == 7, "variant must be 0 (default), 6 or 7");
    VER_CHECK_ARG(B > 0 && Ncam > 0 && Nq > 0 && Sh > 0 && Sw > 0 && NH > 0, "non-positive dimension");
    VER_CHECK_ARG(ld_logits >= NH * NP * 3 && ld_logits % 4 == 0 && NP % 4 == 0,
                  "logits rows must be 16-byte aligned per head (NP %% 4 == 0, ld %% 4 == 0)");
    if (!ver_tc6_supported(Ncam, Sh, Sw, Dh, NP)) {
        ver_set_error("ver_sca_forward_sorted16 needs Ncam <= 32, 2 <= Sh, Sw <= 14, Dh in {32,64,96}, NP in {4,8}");
        return VER_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // variant 0 / 7: the TMEM-operand form (sca_fwd_tc7_kernel) where it applies; 6: A in the shared-memory operand
    if (variant != 6 && ver_tc7_supported(Ncam, Sh, Sw, Dh, NP))
        return ver_sca_forward_tc7(vimg16, logits, ld_logits, rpc, order, smask, tile_union, slots, B, Ncam, Nq, Sh, Sw, NH,
                                   Dh, NP, st);
#define FWD6(D)                                                                                                      \
    (NP == 8 ? launch_fwd_tc6<D, 8>((const __half*)vimg16, logits, ld_logits, rpc, order, smask, tile_union,         \
                                    (__half*)slots, B, Ncam, Nq, Sh, Sw, NH, st)                                     \
             : launch_fwd_tc6<D, 4>((const __half*)vimg16, logits, ld_logits, rpc, order, smask, tile_union,         \
                                    (__half*)slots, B, Ncam, Nq, Sh, Sw, NH, st))
    switch (Dh) {
        case 32: return FWD6(32);
        case 64: return FWD6(64);
        default: return FWD6(96);
    }
#undef FWD6
}

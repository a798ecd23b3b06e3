// Fused SCA sampler forward, third generation: tcgen05 + TMEM, visibility-sorted row tiles, one thread per row.
//
//     slots[b, n, h, :] = 1/max(count,1) * sum_{cam sees n, ascending} A_cam[n, :] V_{b,cam,h}[:, :]
//
// with A the 32-nonzero-per-row interpolation matrix (8 points x 4 bilinear corners, softmax weight folded in)
// and V the 14x14 value map of one (view, head) as a [SP = 208][Dh] fp16 operand image (sca_tc.cu, value_image).
// Replaces SpatialCrossAttention.forward's rebatch / sampling / scatter-mean
// (M/spatial_cross_attention.py:138-173, MSDeformableAttention3D :340-374).
//
// What changed against sca_fwd_tc_kernel (measured 1.06 ms per launch, 7.6 % of the HBM roofline; the builders
// were issue bound: 467 M warp instructions, 8 threads per row serialising their read-modify-writes):
//   * rows are voxels of ONE panorama sorted by their camera bit set (order.cu), so every (tile, camera)
//     product is dense in visible rows: ~2.9x fewer MMAs than the 4x8x8 voxel block with its camera union;
//   * ONE THREAD BUILDS ONE ROW of A: 24 registers hold the row's offsets / softmax weights for all cameras,
//     taps are plain shared-memory read-modify-writes without any intra-warp ordering, the row's 16-byte
//     chunks are zeroed with conflict-free 128-bit stores -> ~10x fewer issued instructions per row;
//   * persistent CTAs (one per SM) walk (panorama, 256-row chunk, head) items; two 128-row groups alternate on
//     the tensor pipe, each group does its own epilogue, and the next item's logits are in flight during it.
//
// Roles: warps 0-3 = group 0 (rows 0..127 of the chunk), warps 4-7 = group 1, warp 8 = control (one lane):
// streams value images in with cp.async.bulk (double buffered) and issues the MMAs.  Hand-offs (mbarriers):
//     bar_built[g]  group g finished A_g for its next camera            (4 arrivals: one per warp)
//     bar_mma[g]    tcgen05.commit: MMAs reading A_g retired -> A_g reusable, accumulator g readable
//     bar_v[buf]    value image landed (transaction bytes)
//     bar_vfree[buf] tcgen05.commit: every MMA that read V[buf] retired -> buffer may be overwritten
#include "sampler.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int kF3Workers = 256;
constexpr int kF3Threads = kF3Workers + 32;
constexpr int kF3Rows = 128;                 // rows per group = UMMA M
constexpr int kF3ChunkRows = 2 * kF3Rows;
constexpr int kF3LgBytes = 96;               // 8 points x (2 offsets + 1 attention logit) fp32 of one (row, head)

struct F3Smem {
    int a_bytes, v_bytes, off_a[2], off_v[2], off_stage, stage_stride, off_lg, total;
    // epilogue staging (per warp: 32 rows x pass_cols fp16, padded rows -> conflict-free 16-byte accesses) turns the
    // row-scattered 16-byte global stores of the accumulator into contiguous runs; Dh = 128 has no room for it
    static __host__ __device__ constexpr bool staged(int Dh) { return Dh <= 96; }
    static __host__ __device__ constexpr int pass_cols(int) { return 32; }
    __host__ __device__ F3Smem(int Dh, int SP) {
        a_bytes = kF3Rows * SP * 2;
        v_bytes = Dh * SP * 2;
        off_a[0] = 0;
        off_a[1] = a_bytes;
        off_v[0] = 2 * a_bytes;
        off_v[1] = 2 * a_bytes + v_bytes;
        off_stage = off_v[1] + v_bytes;
        stage_stride = pass_cols(Dh) * 2 + 16;
        off_lg = off_stage + (staged(Dh) ? (kF3Workers / 32) * 32 * stage_stride : 0);
        total = off_lg + kF3Workers * kF3LgBytes;     // per-worker landing slot of the next item's logits (cp.async)
    }
};

// byte offset of element k inside a row of the A image ([row/8][k/8][row%8][k%8] fp16)
__device__ __forceinline__ uint32_t koff(int k) { return (uint32_t)(k >> 3) * 128u + (uint32_t)(k & 7) * 2u; }
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// loads the compiler may not sink to their first use (register pressure would otherwise expose the DRAM latency)
__device__ __forceinline__ int ldg_pinned(const int32_t* p) {
    int v;
    asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float2 ldg_pinned(const float2* p) {
    float2 v;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
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
__device__ __forceinline__ uint16_t hadd16(uint16_t a, uint16_t b) {
    return __half_as_ushort(__hadd(__ushort_as_half(a), __ushort_as_half(b)));
}
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float (&v)[N]) {
    static_assert(N == 32 || N == 48, "pass width");
    uint32_t r[N];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    if constexpr (N == 48) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
              "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47])
            : "r"(taddr + 32));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = __uint_as_float(r[i]);
}

struct F3Item {
    int b, chunk, h;
};

// phase timers (debug; enabled through ver_debug_tc3_timing, read by tools/tc_timing.py, never by the product)
__device__ unsigned long long g_f3_timing[32];
__device__ int g_f3_timing_on = 0;
struct F3Timer {
    bool on;
    long long t;
    __device__ __forceinline__ F3Timer(bool active) : on(active && g_f3_timing_on), t(0) {
        if (on) t = clock64();
    }
    __device__ __forceinline__ void lap(int slot) {
        if (on) {
            const long long n = clock64();
            atomicAdd(&g_f3_timing[slot], (unsigned long long)(n - t));
            t = n;
        }
    }
};
__device__ __forceinline__ F3Item f3_item(int item, int NH, int chunks_per_b) {
    F3Item it;
    it.h = item % NH;
    const int r = item / NH;
    it.chunk = r % chunks_per_b;
    it.b = r / chunks_per_b;
    return it;
}

template <int DH>
__global__ void __launch_bounds__(kF3Threads, 1)
sca_fwd_tc3_kernel(const __half* __restrict__ vimg, const float* __restrict__ logits, int ld,
                   const float* __restrict__ rpc, const int32_t* __restrict__ order,
                   const uint32_t* __restrict__ smask, const uint32_t* __restrict__ tile_union,
                   __half* __restrict__ slots, int B, int Ncam, int Nq, int Sh, int Sw, int SP, int NH, int NP,
                   int chunks_per_b, int n_items) {
    const int G = SP >> 3;                       // 8-pixel groups per row
    extern __shared__ __align__(128) unsigned char smem[];
    const F3Smem L(DH, SP);
    __shared__ __align__(8) uint64_t bar_built[2], bar_mma[2], bar_v[2], bar_vfree[2];
    __shared__ uint32_t s_tmem;
    __shared__ volatile uint32_t s_kmask[2][2][4];     // [group][batch parity][warp]: K chunks (16 pixels) that hold taps

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_built[i], 4);
            mbar_init(&bar_mma[i], 1);
            mbar_init(&bar_v[i], 1);
            mbar_init(&bar_vfree[i], 1);
        }
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc(&s_tmem, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int tiles_per_b = (Nq + kF3Rows - 1) / kF3Rows;      // order.cu's tile_union row length
    const size_t v_elems = (size_t)DH * SP;               // halves per (view, head) image

    if (warp == 8) {
        // ================================================================ control: TMA + MMA issue
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(128, DH, 0, 0);
            // iterator over (item, camera) steps of this CTA, in the order the workers walk them
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
                    const F3Item it = f3_item(nx_item, NH, chunks_per_b);
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
            F3Timer tc(true);
            bool has_next = advance();
            if (has_next) load_v(0);
            uint32_t kk = 0, itg[2] = {0, 0};
            while (has_next) {
                const int cam = nx_cam;
                const uint32_t u[2] = {nx_u0, nx_u1};
                has_next = advance();                  // nx_* now describe step kk + 1
                const int buf = kk & 1;
                bool first = true;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (!((u[g] >> cam) & 1u)) continue;
                    mbar_wait_park(&bar_built[g], itg[g] & 1);
                    tc.lap(9);                         // control: wait for a built A
                    if (first) mbar_wait_park(&bar_v[buf], (kk >> 1) & 1);
                    tc_fence_after();
                    tc.lap(10);                        // control: wait for the value image
                    const uint32_t a_addr = smem_u32(smem + L.off_a[g]);
                    const uint32_t v_addr = smem_u32(smem + L.off_v[buf]);
                    // only the 16-pixel K chunks some row of this batch tapped (all other columns of A are zero)
                    const int par = itg[g] & 1;
                    uint32_t km = s_kmask[g][par][0] | s_kmask[g][par][1] | s_kmask[g][par][2] | s_kmask[g][par][3];
                    uint32_t acc = (u[g] & ((1u << cam) - 1u)) ? 1u : 0u;      // lowest camera of this tile overwrites
                    if (!acc && !km) km = 1u;                                 // (an all-zero chunk zeroes the accumulator)
                    for (; km; km &= km - 1) {
                        const int ks = __ffs(km) - 1;
                        umma_f16(tmem + g * 128, umma_desc(a_addr + ks * 256, 128, G * 128),
                                 umma_desc(v_addr + ks * 256, 128, G * 128), idesc, acc);
                        acc = 1u;
                    }
                    umma_commit(&bar_mma[g]);
                    tc.lap(11);                        // control: MMA issue
                    ++itg[g];
                    if (first && has_next) {
                        // value image of step kk + 1 -> the other buffer, once step kk - 1 stopped reading it
                        if (kk >= 1) mbar_wait_park(&bar_vfree[(kk + 1) & 1], ((kk - 1) >> 1) & 1);
                        load_v((kk + 1) & 1);
                        tc.lap(12);                    // control: wait for a free value buffer + TMA issue
                    }
                    first = false;
                }
                umma_commit(&bar_vfree[buf]);
                ++kk;
            }
            // drain: the last commits must have arrived before the CTA tears TMEM / smem down
            if (kk >= 1) mbar_wait_park(&bar_vfree[(kk - 1) & 1], ((kk - 1) >> 1) & 1);
        }
    } else {
        // ================================================================ workers: one thread = one row
        const int g = warp >> 2, r = (warp & 3) * 32 + lane;
        // shared-space byte address of element k = 0 of my row; element k lives at + (k >> 3) * 128 + (k & 7) * 2
        const uint32_t myrow = smem_u32(smem + L.off_a[g]) + (uint32_t)(r >> 3) * G * 128u + (uint32_t)(r & 7) * 16u;
        const uint32_t trash = koff(SP - 1);          // padded pixel column: V is zero there, any finite value is harmless
        const uint32_t tm_acc = tmem + ((uint32_t)((warp & 3) * 32) << 16) + g * 128;
        unsigned char* stage = smem + L.off_stage + warp * (32 * L.stage_stride);
        const float fSw = (float)Sw, fSh = (float)Sh;
        const float2* rp2 = reinterpret_cast<const float2*>(rpc);
        uint32_t it = 0, seen = 0;                // MMA batches handed over / observed retired (this group)
        bool tapped = false;                      // my row of A holds the taps recorded in off[]
        uint32_t off[16];                         // byte offsets of the 32 taps in my row (2 x 16 bit per word)

        // whole-buffer zero once (A must be finite / zero outside the taps from the first MMA on)
        for (int c = 0; c < G; ++c)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(myrow + c * 128), "r"(0) : "memory");

        // ---- software pipeline: ids of item i + 2 and logits / first reference point of item i + 1 are in flight
        // while item i is processed
        int n_cur = -1, n_nx = -1, n_n2 = -1;
        uint32_t m_cur = 0, u_cur = 0, m_nx = 0, u_nx = 0, m_n2 = 0, u_n2 = 0;
        float2 ref_nx = make_float2(0.f, 0.f);
        auto load_ids = [&](int item, int& n_o, uint32_t& m_o, uint32_t& u_o) {
            n_o = -1;
            m_o = 0;
            u_o = 0;
            if (item >= n_items) return;
            const F3Item q = f3_item(item, NH, chunks_per_b);
            const int tile = 2 * q.chunk + g;
            const int i = tile * kF3Rows + r;
            if (tile < tiles_per_b)
                u_o = (uint32_t)ldg_pinned(reinterpret_cast<const int32_t*>(tile_union) + (size_t)q.b * tiles_per_b + tile);
            if (i < Nq) {
                n_o = ldg_pinned(order + (size_t)q.b * Nq + i);
                m_o = (uint32_t)ldg_pinned(reinterpret_cast<const int32_t*>(smask) + (size_t)q.b * Nq + i);
            }
        };
        const uint32_t lg_slot = smem_u32(smem + L.off_lg) + (uint32_t)tid * kF3LgBytes;
        auto load_row = [&](int item) {            // uses n_nx / m_nx / u_nx (ids of `item`); logits land in my smem slot
            ref_nx = make_float2(0.f, 0.f);
            if (item >= n_items || n_nx < 0) return;
            const F3Item q = f3_item(item, NH, chunks_per_b);
            const float* row = logits + ((size_t)q.b * Nq + n_nx) * ld;
            const float* po = row + q.h * NP * 2;
            const float* pl = row + NH * NP * 2 + q.h * NP;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i * 2 < NP) cp_async16(lg_slot + i * 16, po + i * 4);
            cp_async16(lg_slot + 64, pl);
            if (NP > 4) cp_async16(lg_slot + 80, pl + 4);
            if (u_nx) {
                const int c0 = __ffs(u_nx) - 1;
                if ((m_nx >> c0) & 1u) ref_nx = ldg_pinned(rp2 + ((size_t)c0 * B + q.b) * Nq + n_nx);
            }
        };

        // ---- epilogue of one item: slots[row] = accumulator / max(count, 1), staged through shared memory so that
        // PPR consecutive lanes write one contiguous run of a row
        auto epilogue = [&](int n, uint32_t m, uint32_t u, int b, int h) {
            const float inv_cnt = 1.f / (float)max(__popc(m), 1);
            const size_t row0 = ((size_t)b * Nq) * NH * DH + (size_t)h * DH;     // + n * NH * DH
            if constexpr (F3Smem::staged(DH)) {
                constexpr int PC = F3Smem::pass_cols(DH), PPR = PC / 8;      // channels per pass, 16-B pieces per row
#pragma unroll
                for (int c0 = 0; c0 < DH; c0 += PC) {
                    float vv[PC];
                    if (u) {                            // warp-uniform (tile property)
                        tmem_ld_cols<PC>(tm_acc + c0, vv);
                    } else {
#pragma unroll
                        for (int i = 0; i < PC; ++i) vv[i] = 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < PC; ++i) vv[i] *= inv_cnt;
                    __syncwarp();                   // previous pass fully read
                    store_channels16<PC>(reinterpret_cast<__half*>(stage + lane * L.stage_stride), vv);
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < PPR; ++k) {
                        const int qi = k * 32 + lane, row = qi / PPR, piece = qi % PPR;
                        const int nr = __shfl_sync(VER_FULL_MASK, n, row);
                        const uint4 val = *reinterpret_cast<const uint4*>(stage + row * L.stage_stride + piece * 16);
                        if (nr >= 0)
                            *reinterpret_cast<uint4*>(slots + row0 + (size_t)nr * NH * DH + c0 + piece * 8) = val;
                    }
                }
            } else {
                __half* dst = (n >= 0) ? slots + row0 + (size_t)n * NH * DH : nullptr;
#pragma unroll
                for (int c0 = 0; c0 < DH; c0 += 32) {
                    float vv[32];
                    if (u) {
                        tmem_ld_cols<32>(tm_acc + c0, vv);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) vv[i] = 0.f;
                    }
                    if (dst) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) vv[i] *= inv_cnt;
                        store_channels16<32>(dst + c0, vv);
                    }
                }
            }
            tc_fence_before();                       // my tcgen05.ld's precede the next overwrite of the accumulator
        };

        F3Timer tw(tid == 0);
        int item = blockIdx.x;
        load_ids(item, n_nx, m_nx, u_nx);
        load_row(item);
        load_ids(item + gridDim.x, n_n2, m_n2, u_n2);
        // the epilogue of an item runs inside the first build of the NEXT item (between the MMA wait and the
        // read-modify-writes), so that the tensor-core round trip of its last batch is covered by the next item's
        // softmax and tap arithmetic
        bool pend = false;
        int pn = -1, pb = 0, ph = 0;
        uint32_t pm = 0, pu = 0;
        const int nchunks = SP >> 4;
        tw.lap(0);                                      // setup
        for (; item < n_items; item += gridDim.x) {
            const F3Item q = f3_item(item, NH, chunks_per_b);
            n_cur = n_nx;
            m_cur = m_nx;
            u_cur = u_nx;
            const int n = n_cur;
            const uint32_t m = m_cur, u = u_cur;
            // ---- offsets (pixel units) and softmax weights of my row, this head
            float ox[8], oy[8], aw[8];
            {
                cp_async_wait_all();                    // my slot holds this item's logits (issued one item ago)
                float4 raw_nx[6];
#pragma unroll
                for (int i = 0; i < 6; ++i)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(raw_nx[i].x), "=f"(raw_nx[i].y), "=f"(raw_nx[i].z), "=f"(raw_nx[i].w)
                                 : "r"(lg_slot + i * 16)
                                 : "memory");
                float mx = -INFINITY;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const float4 o4 = raw_nx[p >> 1];
                    ox[p] = (p & 1) ? o4.z : o4.x;
                    oy[p] = (p & 1) ? o4.w : o4.y;
                    const float4 l4 = raw_nx[4 + (p >> 2)];
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
            float2 ref = ref_nx;
            // ---- prefetch: logits of the next item (its ids are here), ids of the one after
            n_nx = n_n2;
            m_nx = m_n2;
            u_nx = u_n2;
            load_row(item + gridDim.x);
            load_ids(item + 2 * gridDim.x, n_n2, m_n2, u_n2);
            tw.lap(1);                                  // item top: softmax / offsets, prefetch issue

            // ---- cameras of my tile, ascending (= the reference's accumulation order, :166-168)
            for (uint32_t rest = u; rest; rest &= rest - 1) {
                const int cam = __ffs(rest) - 1;
                const bool vis = (m >> cam) & 1u;
                // reference point of the NEXT camera of this tile: in flight during this build
                float2 ref_next = make_float2(0.f, 0.f);
                {
                    const uint32_t nr = rest & (rest - 1);
                    if (nr) {
                        const int cn = __ffs(nr) - 1;
                        if ((m >> cn) & 1u) ref_next = ldg_pinned(rp2 + ((size_t)cn * B + q.b) * Nq + n);
                    }
                }
                // ---- the 32 taps of this (row, camera): byte offsets, fp16 coefficients, K-chunk mask -- registers
                // only, overlaps the tensor-core work on my previous batch.  floor() through the 2^23 trick
                // (add with round-down): no conversion-pipe instructions
                uint32_t noff[16], wts[16], kmask = 0;
                int pixv[8];
#pragma unroll
                for (int p = 0; p < 8; ++p) pixv[p] = -1000 - 64 * p;     // rows / points outside alias nothing
                if (vis) {
                    const float rx1 = fmaf(ref.x, fSw, 0.5f), ry1 = fmaf(ref.y, fSh, 0.5f);      // pixel coordinate + 1
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        const float tx = rx1 + ox[p], ty = ry1 + oy[p];
                        const float bx = __fadd_rd(tx, 8388608.f), by = __fadd_rd(ty, 8388608.f);
                        const int ix = __float_as_int(bx) - 0x4B000000, iy = __float_as_int(by) - 0x4B000000;   // x0 + 1, y0 + 1
                        const bool in = (p < NP) && (unsigned)ix <= (unsigned)Sw && (unsigned)iy <= (unsigned)Sh;
                        const float fx = tx - (bx - 8388608.f), fy = ty - (by - 8388608.f);
                        const bool vx0 = in && ix >= 1, vx1 = in && ix < Sw, vy0 = iy >= 1, vy1 = iy < Sh;
                        const int pix = (iy - 1) * Sw + (ix - 1);
                        const float a = aw[p];
                        const float gx = 1.f - fx, gy = 1.f - fy;
                        const bool v00 = vx0 && vy0, v01 = vx1 && vy0, v10 = vx0 && vy1, v11 = vx1 && vy1;
                        const uint32_t o00 = v00 ? koff(pix) : trash, o01 = v01 ? koff(pix + 1) : trash;
                        const uint32_t o10 = v10 ? koff(pix + Sw) : trash, o11 = v11 ? koff(pix + Sw + 1) : trash;
                        // chunks of the (up to) two pixel pairs; a neighbour that is out of the map only adds a chunk of zeros
                        const uint32_t top = (1u << (max(pix, 0) >> 4)) | (1u << ((pix + 1) >> 4));
                        const uint32_t bot = (1u << ((pix + Sw) >> 4)) | (1u << ((pix + Sw + 1) >> 4));
                        kmask |= ((in && vy0) ? top : 0u) | ((in && vy1) ? bot : 0u);
                        noff[2 * p] = o00 | (o01 << 16);
                        noff[2 * p + 1] = o10 | (o11 << 16);
                        wts[2 * p] = pack_half2(a * (gy * gx), a * (gy * fx));
                        wts[2 * p + 1] = pack_half2(a * (fy * gx), a * (fy * fx));
                        if (in) pixv[p] = pix;
                    }
                    kmask &= (1u << nchunks) - 1u;
                }
                // points p and p + 4 are read-modify-written together when their 2x2 cells are disjoint for every row
                // of the warp (the usual case: their offsets differ by 4 steps), else one after the other
                bool together[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const uint32_t d = (uint32_t)abs(pixv[p] - pixv[p + 4]);
                    const bool alias = d <= 1u || (d - (uint32_t)(Sw - 1)) <= 2u;
                    together[p] = !__any_sync(VER_FULL_MASK, alias);
                }
                tw.lap(4);                              // tap arithmetic
                if (seen < it) {                        // MMAs of my previous batch retired -> A_g is mine again
                    mbar_wait_park(&bar_mma[g], seen & 1);
                    ++seen;
                    tc_fence_after();
                }
                tw.lap(2);                              // wait: my previous MMA batch retired
                if (pend) {                             // ... which completed the previous item's accumulator
                    epilogue(pn, pm, pu, pb, ph);
                    pend = false;
                }
                tw.lap(8);                              // epilogue: TMEM -> slots
                if (tapped) {                           // un-tap: my row is all zero again
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        sts16(myrow + (off[i] & 0xffffu), 0);
                        sts16(myrow + (off[i] >> 16), 0);
                    }
                    tapped = false;
                }
                tw.lap(3);                              // un-tap
                if (vis) {
                    tapped = true;
                    // the LSU keeps program order, so consecutive rounds may touch the same element
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const int q2 = p + 4;
                        const uint32_t a0 = myrow + (noff[2 * p] & 0xffffu), a1 = myrow + (noff[2 * p] >> 16);
                        const uint32_t a2 = myrow + (noff[2 * p + 1] & 0xffffu), a3 = myrow + (noff[2 * p + 1] >> 16);
                        const uint32_t b0 = myrow + (noff[2 * q2] & 0xffffu), b1 = myrow + (noff[2 * q2] >> 16);
                        const uint32_t b2 = myrow + (noff[2 * q2 + 1] & 0xffffu), b3 = myrow + (noff[2 * q2 + 1] >> 16);
                        if (together[p]) {
                            const uint16_t h0 = lds16(a0), h1 = lds16(a1), h2 = lds16(a2), h3 = lds16(a3);
                            const uint16_t g0 = lds16(b0), g1 = lds16(b1), g2 = lds16(b2), g3 = lds16(b3);
                            sts16(a0, hadd16(h0, (uint16_t)(wts[2 * p] & 0xffffu)));
                            sts16(a1, hadd16(h1, (uint16_t)(wts[2 * p] >> 16)));
                            sts16(a2, hadd16(h2, (uint16_t)(wts[2 * p + 1] & 0xffffu)));
                            sts16(a3, hadd16(h3, (uint16_t)(wts[2 * p + 1] >> 16)));
                            sts16(b0, hadd16(g0, (uint16_t)(wts[2 * q2] & 0xffffu)));
                            sts16(b1, hadd16(g1, (uint16_t)(wts[2 * q2] >> 16)));
                            sts16(b2, hadd16(g2, (uint16_t)(wts[2 * q2 + 1] & 0xffffu)));
                            sts16(b3, hadd16(g3, (uint16_t)(wts[2 * q2 + 1] >> 16)));
                        } else {
                            const uint16_t h0 = lds16(a0), h1 = lds16(a1), h2 = lds16(a2), h3 = lds16(a3);
                            sts16(a0, hadd16(h0, (uint16_t)(wts[2 * p] & 0xffffu)));
                            sts16(a1, hadd16(h1, (uint16_t)(wts[2 * p] >> 16)));
                            sts16(a2, hadd16(h2, (uint16_t)(wts[2 * p + 1] & 0xffffu)));
                            sts16(a3, hadd16(h3, (uint16_t)(wts[2 * p + 1] >> 16)));
                            const uint16_t g0 = lds16(b0), g1 = lds16(b1), g2 = lds16(b2), g3 = lds16(b3);
                            sts16(b0, hadd16(g0, (uint16_t)(wts[2 * q2] & 0xffffu)));
                            sts16(b1, hadd16(g1, (uint16_t)(wts[2 * q2] >> 16)));
                            sts16(b2, hadd16(g2, (uint16_t)(wts[2 * q2 + 1] & 0xffffu)));
                            sts16(b3, hadd16(g3, (uint16_t)(wts[2 * q2 + 1] >> 16)));
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) off[i] = noff[i];
                }
                tw.lap(5);                              // taps (shared-memory read-modify-writes)
                kmask = __reduce_or_sync(VER_FULL_MASK, kmask);
                proxy_fence();                          // generic-proxy writes of A -> async proxy (tensor core)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    s_kmask[g][it & 1][warp & 3] = kmask;
                    mbar_arrive(&bar_built[g]);
                }
                ++it;
                ref = ref_next;
                tw.lap(6);                              // fences + arrive
            }
            if (u) {                                    // my last batch is in flight: epilogue deferred
                pend = true;
                pn = n;
                pm = m;
                pu = u;
                pb = q.b;
                ph = q.h;
            } else {                                    // no camera sees this tile: zeros (after any pending epilogue)
                if (pend) {
                    if (seen < it) {
                        mbar_wait_park(&bar_mma[g], seen & 1);
                        ++seen;
                        tc_fence_after();
                    }
                    epilogue(pn, pm, pu, pb, ph);
                    pend = false;
                }
                epilogue(n, m, 0u, q.b, q.h);
            }
            tw.lap(7);                                  // item end
        }
        if (pend) {
            if (seen < it) {
                mbar_wait_park(&bar_mma[g], seen & 1);
                ++seen;
                tc_fence_after();
            }
            epilogue(pn, pm, pu, pb, ph);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 256);
}

template <int DH>
int launch_fwd_tc3(const __half* vimg, const float* logits, int ld, const float* rpc, const int32_t* order,
                   const uint32_t* smask, const uint32_t* tile_union, __half* slots, int B, int Ncam, int Nq,
                   int Sh, int Sw, int SP, int NH, int NP, cudaStream_t st) {
    const F3Smem L(DH, SP);
    VER_CHECK_ARG(L.total + 512 <= ver_device_max_smem_optin(), "TC forward needs %d B of shared memory", L.total);
    auto kern = sca_fwd_tc3_kernel<DH>;
    VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    const int chunks_per_b = (Nq + kF3ChunkRows - 1) / kF3ChunkRows;
    const int n_items = B * NH * chunks_per_b;
    const int sms = ver_device_sm_count();
    const int grid = n_items < sms ? n_items : sms;
    kern<<<grid, kF3Threads, L.total, st>>>(vimg, logits, ld, rpc, order, smask, tile_union, slots, B, Ncam, Nq, Sh,
                                            Sw, SP, NH, NP, chunks_per_b, n_items);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

}  // namespace

extern "C" int ver_debug_tc3_timing(int enable, unsigned long long* host_out32) {
    if (host_out32) VER_CHECK_CUDA(cudaMemcpyFromSymbol(host_out32, g_f3_timing, sizeof(unsigned long long) * 32));
    unsigned long long zero[32] = {0};
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f3_timing, zero, sizeof(zero)));
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_f3_timing_on, &enable, sizeof(int)));
    return VER_OK;
}

int ver_tc3_supported(int Ncam, int S, int Dh, int NP) {
    // S % 16 != 0: the padded pixel columns of the operand images double as the sink of out-of-map corners
    if (!(Ncam <= 32 && NP >= 1 && NP <= 8 && S <= 256 && S % 16 != 0 && (Dh == 32 || Dh == 64 || Dh == 96 || Dh == 128)))
        return 0;
    const int SP = (S + 15) / 16 * 16;
    return F3Smem(Dh, SP).total + 512 <= ver_device_max_smem_optin();
}

/* Forward of the fused SCA sampler on visibility-sorted rows (see include/ver_b200.h). */
extern "C" int ver_sca_forward_sorted(const void* vimg, const float* logits, int ld_logits, const float* rpc,
                                      const int32_t* order, const uint32_t* smask, const uint32_t* tile_union,
                                      void* slots, int B, int Ncam, int Nq, int Sh, int Sw, int NH, int Dh, int NP,
                                      int variant, ver_stream_t stream) {
    VER_CHECK_ARG(vimg && logits && rpc && order && smask && tile_union && slots, "null pointer");
    VER_CHECK_ARG(variant == 0 || (variant >= 3 && variant <= 5), "variant must be 0 (newest) or 3, 4, 5");
    VER_CHECK_ARG(B > 0 && Ncam > 0 && Nq > 0 && Sh > 0 && Sw > 0 && NH > 0, "non-positive dimension");
    VER_CHECK_ARG(ld_logits >= NH * NP * 3 && ld_logits % 4 == 0 && (NP * 2) % 4 == 0 && NP % 4 == 0,
                  "logits rows must be 16-byte aligned per head (NP %% 4 == 0, ld %% 4 == 0)");
    if (!ver_tc3_supported(Ncam, Sh * Sw, Dh, NP)) {
        ver_set_error("sorted tensor-core sampler needs Ncam <= 32, S <= 256, Dh in {32,64,96,128}");
        return VER_ERR_UNSUPPORTED;
    }
    const int SP = (Sh * Sw + 15) / 16 * 16;
    cudaStream_t st = (cudaStream_t)stream;
    // variant 0 takes sca_fwd_tc4_kernel: the three-operand kernel measured 410-440 us against 420 us at the
    // benchmark shape (profiles/r02b), i.e. no gain -- it stays selectable for A/B runs and as the test bed of
    // the bounded-wait protocol
    if (variant == 5 && Sh >= 2 && Sw >= 2 && ver_tc5_supported(Ncam, Sh * Sw, Dh, NP))
        return ver_sca_forward_tc5(vimg, logits, ld_logits, rpc, order, smask, tile_union, slots, B, Ncam, Nq, Sh, Sw,
                                   NH, Dh, NP, st);
    if (variant != 3 && Sh >= 2 && Sw >= 2 && ver_tc4_supported(Ncam, Sh * Sw, Dh, NP))
        return ver_sca_forward_tc4(vimg, logits, ld_logits, rpc, order, smask, tile_union, slots, B, Ncam, Nq, Sh, Sw,
                                   NH, Dh, NP, st);
#define FWD3(D)                                                                                                   \
    launch_fwd_tc3<D>((const __half*)vimg, logits, ld_logits, rpc, order, smask, tile_union, (__half*)slots, B, \
                      Ncam, Nq, Sh, Sw, SP, NH, NP, st)
    switch (Dh) {
        case 32: return FWD3(32);
        case 64: return FWD3(64);
        case 96: return FWD3(96);
        default: return FWD3(128);
    }
#undef FWD3
}

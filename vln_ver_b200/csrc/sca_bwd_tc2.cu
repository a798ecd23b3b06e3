// Fused SCA sampler backward, second tensor-core generation (fp16 maps).
//
// Same contraction as sca_bwd_tc_kernel (sca_tc.cu): per (view, head) CTA and chunk of 128 hits
//     dV^T[ch, pix] += G^T[ch, hit] A'[hit, pix]          (accumulated over all chunks in TMEM, written once)
//     Dots[hit, pix] = G[hit, ch] V[pix, ch]^T            (read back at the 32 taps of each hit)
// replacing mmcv._ext.ms_deform_attn_backward + the SCA scatter (M/multi_scale_deformable_attn_function.py:128-163,
// M/spatial_cross_attention.py:166-173).  What changed (profiles/r01e: 17 k cycles per chunk, 46 % of them in the
// A' rows -- 8 threads per hit serialising 8 read-modify-write rounds -- and 25 % in the Dots read-back):
//   * two threads per hit (4 points each): bilinear weights are the branch-free tent max(0, 1 - |coordinate - cell|)
//     on a clamped 2x2 cell block (no per-corner validity tests, no trash cell), the two threads commit their
//     taps in two rounds; rows are un-tapped after use instead of re-zeroing the whole 53 KB image;
//   * Dots go TMEM -> registers -> a LANE-INTERLEAVED fp32 staging buffer of their own (word (quarter, col, lane)),
//     written and read without bank conflicts, in one pass over all 208 columns by all 8 warps;
//   * the Dots MMAs are committed before the dV^T MMAs, so the read-back overlaps the second half of the batch;
//   * tap geometry (cell, distances) stays in registers between the build and the read-back.
#include "sampler.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int kB2Threads = 256;
constexpr int kB2Hits = 128;                  // hits per chunk = UMMA M / K

struct B2Smem {
    int v_bytes, a_bytes, g_bytes, d_bytes, off_v, off_a, off_g, off_d, off_n, total;
    __host__ __device__ B2Smem(int Dh, int SP) {
        v_bytes = Dh * SP * 2;
        a_bytes = kB2Hits * SP * 2;
        g_bytes = kB2Hits * Dh * 2 + 1024;            // + slack: M = 128 over-reads past ch < Dh
        d_bytes = 4 * SP * 32 * 4;                    // Dots staging: [lane quarter][column][lane] fp32
        off_v = 0;
        off_a = off_v + v_bytes;
        off_g = off_a + a_bytes;
        off_d = off_g + g_bytes;
        off_n = off_d + d_bytes;
        total = off_n + 3 * kB2Hits * 4;
    }
};

__device__ unsigned long long g_b2_timing[32];
__device__ int g_b2_timing_on = 0;
struct B2Timer {
    bool on;
    long long t;
    __device__ __forceinline__ B2Timer(bool active) : on(active && g_b2_timing_on), t(0) {
        if (on) t = clock64();
    }
    __device__ __forceinline__ void lap(int slot) {
        if (on) {
            const long long n = clock64();
            atomicAdd(&g_b2_timing[slot], (unsigned long long)(n - t));
            t = n;
        }
    }
};

__device__ __forceinline__ uint32_t koffb(int k) { return ((uint32_t)(k >> 3) << 7) | ((uint32_t)(k & 7) << 1); }
__device__ __forceinline__ uint16_t b2_lds16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.b16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void b2_sts16(uint32_t a, uint16_t v) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(v) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// loads the compiler may not sink to their first use (the software pipeline relies on their issue position)
__device__ __forceinline__ float4 ldg_pin(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ldg_pin(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float2 ldg_pin(const float2* p) {
    float2 v;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ldg_pin(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// vector reductions into global memory (sm_90+): one L2 operation per 16 / 8 bytes instead of one per float
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// tent weight of a cell at signed distance d, and its derivative with respect to the coordinate
__device__ __forceinline__ float tent(float d) { return fmaxf(1.f - fabsf(d), 0.f); }
// (cell to the left of / at the coordinate: -1 on [0, 1); cell to the right: +1 on [-1, 0) -- the reference's
// per-corner derivative, including a valid corner whose weight is exactly 0)
__device__ __forceinline__ float tent_slope(float d) { return d >= 0.f ? (d < 1.f ? -1.f : 0.f) : (d >= -1.f ? 1.f : 0.f); }

template <int DH, int NP>
__global__ void __launch_bounds__(kB2Threads, 1)
sca_bwd_tc2_kernel(const __half* __restrict__ vimg, const float* __restrict__ logits, int ld,
                   const float* __restrict__ rpc, const uint32_t* __restrict__ vis_bits,
                   const int32_t* __restrict__ counts, const int32_t* __restrict__ index,
                   const __half* __restrict__ gslots, float* __restrict__ gvalue,
                   float* __restrict__ glogits, int B, int Ncam, int Nq, int Sh, int Sw, int SP, int NH) {
    static_assert(NP == 4 || NP == 8, "points per head");
    constexpr int PPT = NP / 2;                          // points per thread (two threads per hit)
    constexpr int CG = DH / 8;                           // channel groups
    const int S = Sh * Sw;
    const int G = SP >> 3;
    extern __shared__ __align__(1024) unsigned char smem[];
    const B2Smem L(DH, SP);
    __half* Gimg = reinterpret_cast<__half*>(smem + L.off_g);
    float* dstage = reinterpret_cast<float*>(smem + L.off_d);
    int* s_n = reinterpret_cast<int*>(smem + L.off_n);       // [3][128] hit -> voxel ids of chunks c, c + 1, c + 2 (c % 3)
    __shared__ __align__(8) uint64_t bar_v, bar_dots, bar_dv, bar_g;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = blockIdx.x;
    // longest-processing-time-first: CTA y takes the view with the y-th largest hit count, so that the last
    // (partial) wave of the 1-CTA-per-SM grid holds the cheapest views.  Every CTA ranks the views itself
    // (n^2 / 256 L1-resident loads per thread, n = B * Ncam); skipped for very large batches.
    __shared__ int s_bv;
    const int nviews = B * Ncam;
    if (nviews <= 512) {
        for (int v = tid; v < nviews; v += kB2Threads) {
            const int c = counts[v];
            int rank = 0;
            for (int u = 0; u < nviews; ++u) {
                const int cu = counts[u];
                rank += (cu > c || (cu == c && u < v)) ? 1 : 0;
            }
            if (rank == (int)blockIdx.y) s_bv = v;
        }
    } else if (tid == 0) {
        s_bv = blockIdx.y;
    }
    __syncthreads();
    const int bv = s_bv;
    const int b = bv / Ncam, cam = bv % Ncam;

    if (tid == 0) {
        mbar_init(&bar_v, 1);
        mbar_init(&bar_dots, 1);
        mbar_init(&bar_dv, 1);
        mbar_init(&bar_g, kB2Threads);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&s_tmem, 512);
    // the A' image starts out all zero (rows are un-tapped after every chunk)
    for (int i = tid; i < L.a_bytes / 16; i += kB2Threads)
        reinterpret_cast<uint4*>(smem + L.off_a)[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t tm_dv = tmem;                  // dV^T : lanes = channel, columns [0, SP) = pixel
    const uint32_t tm_dots = tmem + 256;          // Dots : lanes = hit,     columns [0, SP) = pixel
    B2Timer tb(tid == 0);
    if (tid == 0) {
        mbar_expect_tx(&bar_v, L.v_bytes);
        bulk_g2s(smem + L.off_v, vimg + ((size_t)bv * NH + h) * DH * SP, L.v_bytes, &bar_v);
    }
    const int nitems = counts[bv];
    const int32_t* idx = index + (size_t)bv * Nq;
    const float2* rp = reinterpret_cast<const float2*>(rpc) + ((size_t)cam * B + b) * Nq;
    constexpr uint32_t idesc_dv = umma_idesc(128, 0, 1, 1);       // N patched in at run time (SP)
    constexpr uint32_t idesc_dots = umma_idesc(128, 0, 0, 1);
    const uint32_t n_bits = (uint32_t)(SP >> 3) << 17;
    const float fSw = (float)Sw, fSh = (float)Sh;
    const float pix_bias = 8388608.f - (float)(Sw + 1);

    // my hit row and my half of its points
    const int r = tid >> 1, half = tid & 1;
    const uint32_t myrow = smem_u32(smem + L.off_a) + (uint32_t)(r >> 3) * G * 128u + (uint32_t)(r & 7) * 16u;
    const float* my_dots = dstage + ((size_t)(r >> 5) * SP) * 32 + (r & 31);       // column c at + c * 32
    int tap_pix[PPT];
    uint32_t tap_o01[PPT], tap_o23[PPT];          // byte offsets of the four cells of each tap, 16 bits each
    bool tapped = false;

    uint32_t phase = 0;
    const int nchunks = (nitems + kB2Hits - 1) / kB2Hits;
    // ---- software pipeline: the per-hit inputs (logits, reference point, camera count) of chunk i + 1, its
    // grad_slots rows and the hit ids of chunk i + 2 are in flight while chunk i is processed
    struct RowData {
        float4 lg4[NP / 4], off4[PPT / 2];
        float2 ref;
        uint32_t bits;
    };
    auto load_row = [&](int n, RowData& d) {
        if (n < 0) return;
        const float* row = logits + ((size_t)b * Nq + n) * ld;
#pragma unroll
        for (int i = 0; i < NP / 4; ++i) d.lg4[i] = ldg_pin(reinterpret_cast<const float4*>(row + NH * NP * 2 + h * NP + i * 4));
#pragma unroll
        for (int i = 0; i < PPT / 2; ++i)
            d.off4[i] = ldg_pin(reinterpret_cast<const float4*>(row + h * NP * 2 + half * PPT * 2 + i * 4));
        d.ref = ldg_pin(rp + n);
        d.bits = ldg_pin(vis_bits + (size_t)b * Nq + n);
    };
    constexpr int kG = (kB2Hits * CG + kB2Threads - 1) / kB2Threads;
    // 16-byte piece i of the G image -> (hit row, channel group): a warp instruction covers 8 rows x 4 channel groups,
    // lane = (channel group, row % 8).  The 8 lanes of a store phase then write the 8 rows of ONE core matrix (128
    // contiguous bytes, conflict free) and 4 lanes read 64 contiguous bytes of a row.  (Row-contiguous lanes -- 12 per
    // row -- stored at a 128-byte stride: 8-way bank conflicts, 1 344 of ~2 900 shared-memory wavefronts per chunk,
    // profiles/r02t_sca_bwd_tc2.txt.)
    static_assert(CG % 4 == 0, "channel groups come in fours");
    auto piece = [&](int i, int& rr, int& cg) {
        const int wi = i >> 5, l = i & 31;
        rr = (wi / (CG / 4)) * 8 + (l & 7);
        cg = (wi % (CG / 4)) * 4 + (l >> 3);
    };
    auto load_g = [&](const int* sn, uint4 (&gv)[kG]) {
#pragma unroll
        for (int u = 0; u < kG; ++u) {
            const int i = tid + u * kB2Threads;
            gv[u] = make_uint4(0, 0, 0, 0);
            if (i < kB2Hits * CG) {
                int rr, cg;
                piece(i, rr, cg);
                const int n = sn[rr];
                if (n >= 0)
                    gv[u] = ldg_pin(reinterpret_cast<const uint4*>(gslots + ((size_t)b * Nq + n) * NH * DH + h * DH + cg * 8));
            }
        }
    };
    int n_cur = r < nitems ? idx[r] : -1, n_nx = kB2Hits + r < nitems ? idx[kB2Hits + r] : -1;
    RowData d_cur, d_nx;
    uint4 g_cur[kG];
    int sn_pref = -1;                                                                       // row tid of chunk c + 2
    if (tid < kB2Hits) {
        s_n[tid] = tid < nitems ? idx[tid] : -1;
        s_n[kB2Hits + tid] = kB2Hits + tid < nitems ? idx[kB2Hits + tid] : -1;
        sn_pref = 2 * kB2Hits + tid < nitems ? idx[2 * kB2Hits + tid] : -1;
    }
    load_row(n_cur, d_cur);
    __syncthreads();
    load_g(s_n, g_cur);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int base = chunk * kB2Hits;
        // ---- un-tap my points of the previous chunk (its MMAs retired); prefetches
        if (tapped) {
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                b2_sts16(myrow + (tap_o01[p] & 0xffffu), 0);
                b2_sts16(myrow + (tap_o01[p] >> 16), 0);
                b2_sts16(myrow + (tap_o23[p] & 0xffffu), 0);
                b2_sts16(myrow + (tap_o23[p] >> 16), 0);
            }
            tapped = false;
        }
        // ---- prefetches for the next chunk.  Measured placements (profiles/r02m): here, 877 us per launch; after the
        // iteration's second proxy fence, 926 us; two chunks ahead in a second register set, 953 us -- the issue of
        // these per-hit gathers (16 cache lines per warp instruction) costs the same ~2 300 cycles wherever it stands,
        // and a proxy fence waits for the loads in flight
        if (tid < kB2Hits) {
            s_n[((chunk + 2) % 3) * kB2Hits + tid] = sn_pref;          // ids of chunk + 2, loaded one chunk ago
            sn_pref = base + 3 * kB2Hits + tid < nitems
                          ? (int)ldg_pin(reinterpret_cast<const uint32_t*>(idx + base + 3 * kB2Hits + tid))
                          : -1;
        }
        load_row(n_nx, d_nx);
        const int n_n2 = base + 2 * kB2Hits + r < nitems
                             ? (int)ldg_pin(reinterpret_cast<const uint32_t*>(idx + base + 2 * kB2Hits + r))
                             : -1;
        tb.lap(16);                                  // un-tap + prefetch issue
        // ---- this chunk's grad_slots rows -> G image (its last readers, the previous chunk's dV^T MMAs, retired).
        // Dots = G V^T does not depend on A': thread 0 issues those MMAs as soon as every thread stored its part,
        // and they run while the A' rows are built
#pragma unroll
        for (int u = 0; u < kG; ++u) {
            const int i = tid + u * kB2Threads;
            if (i < kB2Hits * CG) {
                int rr, cg;
                piece(i, rr, cg);
                *reinterpret_cast<uint4*>(Gimg + ((rr >> 3) * CG + cg) * 64 + (rr & 7) * 8) = g_cur[u];
            }
        }
        // the next chunk's grad_slots rows (ids published one chunk ago): in flight for a whole chunk
        load_g(s_n + ((chunk + 1) % 3) * kB2Hits, g_cur);
        proxy_fence();
        tc_fence_before();
        mbar_arrive(&bar_g);
        if (tid == 0) {
            mbar_wait(&bar_g, phase);
            tc_fence_after();
            if (chunk == 0) mbar_wait(&bar_v, 0);
            const uint32_t g_addr = smem_u32(Gimg), v_addr = smem_u32(smem + L.off_v);
            // Dots[hit, pix] = G[hit, ch] V[pix, ch]^T     (A: G image K-major; B: V image MN-major)
            for (int ks = 0; ks < DH / 16; ++ks)
                umma_f16(tm_dots, umma_desc(g_addr + ks * 256, 128, CG * 128),
                         umma_desc(v_addr + ks * 2 * G * 128, G * 128, 128), idesc_dots | n_bits, ks > 0 ? 1u : 0u);
            umma_commit(&bar_dots);
        }
        tb.lap(17);                                  // G image stores, Dots MMA issue
        // ---- my hit: softmax over all points (both threads of the pair compute it), then my PPT points
        const int n = n_cur;
        float aw[PPT], ddx[PPT], ddy[PPT], inv_cnt = 0.f;
        __half2 wa[PPT], wb[PPT];
        if (n >= 0) {
            float lg[NP];
#pragma unroll
            for (int i = 0; i < NP / 4; ++i) {
                const float4 t = d_cur.lg4[i];
                lg[4 * i] = t.x;
                lg[4 * i + 1] = t.y;
                lg[4 * i + 2] = t.z;
                lg[4 * i + 3] = t.w;
            }
            float2 off[PPT];
#pragma unroll
            for (int i = 0; i < PPT / 2; ++i) {
                const float4 t = d_cur.off4[i];
                off[2 * i] = make_float2(t.x, t.y);
                off[2 * i + 1] = make_float2(t.z, t.w);
            }
            const float2 ref = d_cur.ref;
            inv_cnt = 1.f / (float)max(__popc(d_cur.bits), 1);
            float mx = lg[0];
#pragma unroll
            for (int p = 1; p < NP; ++p) mx = fmaxf(mx, lg[p]);
            float sum = 0.f;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                lg[p] = __expf(lg[p] - mx);
                sum += lg[p];
            }
            const float inv = 1.f / sum;
            const float rx1 = fmaf(ref.x, fSw, 0.5f), ry1 = fmaf(ref.y, fSh, 0.5f);        // pixel coordinate + 1
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                aw[p] = (half ? lg[PPT + p] : lg[p]) * inv;
                const float tx = rx1 + off[p].x, ty = ry1 + off[p].y;
                const float flx = __fadd_rd(tx, 8388608.f) - 8388608.f, fly = __fadd_rd(ty, 8388608.f) - 8388608.f;
                const float cx = fminf(fmaxf(flx, 1.f), fSw - 1.f), cy = fminf(fmaxf(fly, 1.f), fSh - 1.f);
                ddx[p] = tx - cx;                        // signed distance to the left / upper cell of the 2x2 block
                ddy[p] = ty - cy;
                const float a = aw[p] * inv_cnt;
                const float wxa = tent(ddx[p]), wxb = tent(ddx[p] - 1.f);
                const float wya = a * tent(ddy[p]), wyb = a * tent(ddy[p] - 1.f);
                wa[p] = __floats2half2_rn(wya * wxa, wya * wxb);
                wb[p] = __floats2half2_rn(wyb * wxa, wyb * wxb);
                const int k = __float_as_int(fmaf(cy, fSw, cx) + pix_bias) - 0x4B000000;
                tap_pix[p] = k;
                tap_o01[p] = koffb(k) | (koffb(k + 1) << 16);
                tap_o23[p] = koffb(k + Sw) | (koffb(k + Sw + 1) << 16);
            }
            tapped = true;
        }
        // the two threads of a hit commit their taps one after the other (their cells may coincide)
#pragma unroll
        for (int round = 0; round < 2; ++round) {
            if (n >= 0 && half == round) {
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    const uint32_t a0 = myrow + (tap_o01[p] & 0xffffu), a1 = myrow + (tap_o01[p] >> 16);
                    const uint32_t a2 = myrow + (tap_o23[p] & 0xffffu), a3 = myrow + (tap_o23[p] >> 16);
                    const uint16_t h0 = b2_lds16(a0), h1 = b2_lds16(a1), h2 = b2_lds16(a2), h3 = b2_lds16(a3);
                    b2_sts16(a0, __half_as_ushort(__hadd(__ushort_as_half(h0), __low2half(wa[p]))));
                    b2_sts16(a1, __half_as_ushort(__hadd(__ushort_as_half(h1), __high2half(wa[p]))));
                    b2_sts16(a2, __half_as_ushort(__hadd(__ushort_as_half(h2), __low2half(wb[p]))));
                    b2_sts16(a3, __half_as_ushort(__hadd(__ushort_as_half(h3), __high2half(wb[p]))));
                }
            }
            __syncwarp();
        }
        tb.lap(18);                                  // A' rows
        proxy_fence();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + L.off_a), g_addr = smem_u32(Gimg);
            // dV^T[ch, pix] += G^T[ch, hit] A'[hit, pix]   (A: G image MN-major; B: A' image MN-major)
            for (int ks = 0; ks < kB2Hits / 16; ++ks)
                umma_f16(tm_dv, umma_desc(g_addr + ks * 2 * CG * 128, CG * 128, 128),
                         umma_desc(a_addr + ks * 2 * G * 128, G * 128, 128), idesc_dv | n_bits,
                         (chunk > 0 || ks > 0) ? 1u : 0u);
            umma_commit(&bar_dv);
        }
        mbar_wait(&bar_dots, phase);
        tc_fence_after();
        tb.lap(19);                                  // fences, dV^T MMA issue, wait for the Dots MMAs
        // ---- Dots: TMEM -> registers -> lane-interleaved staging; warp w reads lane quarter w & 3, column half w >> 2
        {
            const int c_beg = (warp >> 2) ? (SP / 16) * 8 : 0, c_end = (warp >> 2) ? SP : (SP / 16) * 8;
            float* dq = dstage + ((size_t)(warp & 3) * SP) * 32 + lane;
            const uint32_t tl = tm_dots + ((uint32_t)((warp & 3) * 32) << 16);
            int c = c_beg;
            for (; c + 16 <= c_end; c += 16) {
                float vv[16];
                tmem_ld16(tl + c, vv);
#pragma unroll
                for (int i = 0; i < 16; ++i) dq[(c + i) * 32] = vv[i];
            }
            if (c < c_end) {
                float vv[8];
                tmem_ld8(tl + c, vv);
#pragma unroll
                for (int i = 0; i < 8; ++i) dq[(c + i) * 32] = vv[i];
            }
        }
        tc_fence_before();
        __syncthreads();
        tb.lap(20);                                  // Dots dump
        // ---- my points: d/d(attention weight), d/d(offset); softmax backward over the pair; logit gradients
        float ga[PPT], gx[PPT], gy[PPT], t = 0.f;
        if (n >= 0) {
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                const int k = tap_pix[p];
                const float d00 = my_dots[k * 32], d01 = my_dots[(k + 1) * 32];
                const float d10 = my_dots[(k + Sw) * 32], d11 = my_dots[(k + Sw + 1) * 32];
                const float wxa = tent(ddx[p]), wxb = tent(ddx[p] - 1.f);
                const float wya = tent(ddy[p]), wyb = tent(ddy[p] - 1.f);
                const float sxa = tent_slope(ddx[p]), sxb = tent_slope(ddx[p] - 1.f);
                const float sya = tent_slope(ddy[p]), syb = tent_slope(ddy[p] - 1.f);
                const float top = wxa * d00 + wxb * d01, bot = wxa * d10 + wxb * d11;
                ga[p] = (wya * top + wyb * bot) * inv_cnt;
                gx[p] = wya * (sxa * d00 + sxb * d01) + wyb * (sxa * d10 + sxb * d11);
                gy[p] = sya * top + syb * bot;
                t += aw[p] * ga[p];
            }
        }
        t += __shfl_xor_sync(VER_FULL_MASK, t, 1);               // the other half of my hit's points
        if (n >= 0) {
            // my PPT attention-logit gradients and 2 PPT offset gradients are contiguous: vector reductions
            // (d loc = aw * size * sum(...), d offset = d loc / size  -> the size cancels)
            float* grow = glogits + ((size_t)b * Nq + n) * ld;
            float gw[PPT], go[2 * PPT];
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                gw[p] = aw[p] * (ga[p] - t);
                go[2 * p] = inv_cnt * aw[p] * gx[p];
                go[2 * p + 1] = inv_cnt * aw[p] * gy[p];
            }
            float* dw = grow + NH * NP * 2 + h * NP + half * PPT;
            float* dof = grow + h * NP * 2 + half * PPT * 2;
            if (PPT == 4) {
                red_add_v4(dw, gw[0], gw[1], gw[2], gw[3]);
                red_add_v4(dof, go[0], go[1], go[2], go[3]);
                red_add_v4(dof + 4, go[4], go[5], go[6], go[7]);
            } else {
                red_add_v2(dw, gw[0], gw[1]);
                red_add_v4(dof, go[0], go[1], go[2], go[3]);
            }
        }
        tb.lap(21);                                  // tap read-back, softmax backward, atomics
        mbar_wait(&bar_dv, phase);                   // A' / G may be rewritten, Dots may be overwritten
        phase ^= 1;
        tc_fence_after();
        __syncthreads();
        n_cur = n_nx;
        n_nx = n_n2;
        d_cur = d_nx;
        tb.lap(23);                                  // wait: dV^T MMAs retired
    }
    // ---- grad_value: dV^T (TMEM lanes = channel) -> [bv][pix][h][ch] fp32
    if (nchunks == 0 && tid == 0) mbar_wait(&bar_v, 0);       // never leave a bulk copy in flight
    if (warp < 4) {
        const int ch = warp * 32 + lane;
        float* gdst = gvalue + ((size_t)bv * S * NH + h) * DH + ch;
        for (int c0 = 0; c0 < SP; c0 += 16) {
            float vv[16];
            if (nchunks > 0) {
                tmem_ld16(tm_dv + ((uint32_t)(warp * 32) << 16) + c0, vv);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) vv[i] = 0.f;
            }
            if (ch < DH) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (c0 + i < S) gdst[(size_t)(c0 + i) * NH * DH] = vv[i];
            }
        }
    }
    tb.lap(22);                                  // dV^T -> grad_value
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int DH, int NP>
int launch_bwd_tc2(const __half* vimg, const float* logits, int ld, const float* rpc, const uint32_t* vis_bits,
                   const int32_t* counts, const int32_t* index, const __half* gslots, float* gvalue,
                   float* glogits, int B, int Ncam, int Nq, int Sh, int Sw, int SP, int NH, cudaStream_t st) {
    const B2Smem L(DH, SP);
    auto kern = sca_bwd_tc2_kernel<DH, NP>;
    VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    VER_CHECK_CUDA(cudaMemset2DAsync(glogits, (size_t)ld * sizeof(float), 0, (size_t)NH * NP * 3 * sizeof(float),
                                     (size_t)B * Nq, st));
    kern<<<dim3(NH, B * Ncam), kB2Threads, L.total, st>>>(vimg, logits, ld, rpc, vis_bits, counts, index, gslots,
                                                         gvalue, glogits, B, Ncam, Nq, Sh, Sw, SP, NH);
    VER_CHECK_LAUNCH();
    g_ver_launches += 2;
    return VER_OK;
}

}  // namespace

extern "C" int ver_debug_bwd2_timing(int enable, unsigned long long* host_out32) {
    if (host_out32) VER_CHECK_CUDA(cudaMemcpyFromSymbol(host_out32, g_b2_timing, sizeof(unsigned long long) * 32));
    unsigned long long zero[32] = {0};
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_b2_timing, zero, sizeof(zero)));
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_b2_timing_on, &enable, sizeof(int)));
    return VER_OK;
}

int ver_bwd_tc2_supported(int Ncam, int Sh, int Sw, int Dh, int NP, int ld) {
    const int S = Sh * Sw;
    if (!(Ncam <= 32 && (NP == 4 || NP == 8) && S <= 256 && Sh >= 2 && Sw >= 2 && ld % 4 == 0 &&
          (Dh == 32 || Dh == 64 || Dh == 96)))
        return 0;
    const int SP = (S + 15) / 16 * 16;
    return B2Smem(Dh, SP).total + 1024 <= ver_device_max_smem_optin();
}

int ver_sca_backward_tc2(const void* vimg, const float* logits, int ld, const float* rpc, const uint32_t* vis_bits,
                         const int32_t* counts, const int32_t* index, const void* gslots, float* gvalue,
                         float* glogits, int B, int Ncam, int Nq, int Sh, int Sw, int NH, int Dh, int NP,
                         cudaStream_t st) {
    const int SP = (Sh * Sw + 15) / 16 * 16;
#define BWD2(D)                                                                                                    \
    (NP == 8 ? launch_bwd_tc2<D, 8>((const __half*)vimg, logits, ld, rpc, vis_bits, counts, index,                \
                                    (const __half*)gslots, gvalue, glogits, B, Ncam, Nq, Sh, Sw, SP, NH, st)      \
             : launch_bwd_tc2<D, 4>((const __half*)vimg, logits, ld, rpc, vis_bits, counts, index,                \
                                    (const __half*)gslots, gvalue, glogits, B, Ncam, Nq, Sh, Sw, SP, NH, st))
    switch (Dh) {
        case 32: return BWD2(32);
        case 64: return BWD2(64);
        default: return BWD2(96);
    }
#undef BWD2
}

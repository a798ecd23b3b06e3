// Tensor-core (tcgen05 / TMEM) formulation of the fused SCA sampler, fp16 storage.
//
// Deformable sampling over ONE 14x14 map is a contraction with a sparse interpolation matrix:
//     out[row, ch] = sum_pix A[row, pix] * V[pix, ch],     A[row, pix] = sum_taps aw * bilinear
// with 32 taps (8 points x 4 corners) per (voxel|hit, head) row and only S = 196 "keys".
// Gathering the taps from shared memory costs ~300 issue slots per row (measured: the gather
// kernels are issue/LSU bound at <5% of the HBM roofline); building the row of A (32 scalar
// read-modify-writes) and letting the 5th-gen tensor cores do the dense K = 208 contraction
// costs ~15.  Forward:  O = A V  accumulated over cameras in TMEM (ascending camera order, fp32).
// Backward (per view, head, chunks of 128 hits):  dV^T = G^T A'   and   Dots = G V^T,
// from which d/d(attention logits) and d/d(offsets) are read back at the 32 taps of each row.
//
// Shared-memory operand images use the canonical no-swizzle core-matrix layout (8 rows x 16 B
// contiguous); one image serves as K-major or MN-major operand depending on which index is
// called "K" (validated on hardware with tools/tc_probe.cu):
//     V image  [ch/8][pix/8][ch%8][pix%8]   (built once per layer by value_image_kernel)
//     A image  [row/8][pix/8][row%8][pix%8]
//     G image  [hit/8][ch/8][hit%8][ch%8]
// Descriptor convention: LBO = byte stride between core matrices along K, SBO = along M/N.
#include "sca_bwd.cuh"
#include "tcgen05.cuh"

namespace {

// ---------------------------------------------------------------- value image
// value [Bv][S][NH][Dh] fp16  ->  vimg [Bv][NH][Dh/8][SP/8][8][8] fp16, pixels >= S zero
__global__ void value_image_kernel(const __half* __restrict__ value, __half* __restrict__ vimg, int Bv,
                                   int S, int NH, int Dh, int SP) {
    // one thread per 16-byte output chunk = (bv, h, ch/8, pix/8, ch%8): 8 pixels of one channel
    const size_t chunks = (size_t)Bv * NH * (Dh / 8) * (SP / 8) * 8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < chunks;
         i += (size_t)gridDim.x * blockDim.x) {
        const int c8 = i & 7;
        size_t r = i >> 3;
        const int pg = r % (SP / 8);
        r /= (SP / 8);
        const int cg = r % (Dh / 8);
        r /= (Dh / 8);
        const int h = r % NH;
        const int bv = r / NH;
        const int ch = cg * 8 + c8;
        __half out[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int pix = pg * 8 + p;
            out[p] = pix < S ? value[(((size_t)bv * S + pix) * NH + h) * Dh + ch] : __float2half(0.f);
        }
        *reinterpret_cast<uint4*>(vimg + i * 8) = *reinterpret_cast<const uint4*>(out);
    }
}

// The same through a shared-memory tile: one CTA per (view, head) stages the [S][Dh] slice with coalesced 16-byte loads
// (the 2 Dh contiguous bytes of every pixel) and writes the transposed image as full 128-byte core matrices.  The
// kernel above gathers eight 2-byte elements 2 NH Dh bytes apart per output chunk: 34 us for 43 MB at the benchmark
// shape, about half of what the traffic costs.
__global__ void __launch_bounds__(256)
value_image_tiled_kernel(const __half* __restrict__ value, __half* __restrict__ vimg, int S, int NH, int Dh, int SP) {
    extern __shared__ __align__(16) unsigned char vi_smem[];
    __half* tile = reinterpret_cast<__half*>(vi_smem);                 // [S][Dh]
    const int bv = blockIdx.x / NH, h = blockIdx.x % NH;
    const int vec_per_row = Dh / 8;
    const __half* src = value + ((size_t)bv * S * NH + h) * Dh;
    for (int i = threadIdx.x; i < S * vec_per_row; i += blockDim.x) {
        const int pix = i / vec_per_row, v = i % vec_per_row;
        reinterpret_cast<uint4*>(tile)[i] = *reinterpret_cast<const uint4*>(src + (size_t)pix * NH * Dh + v * 8);
    }
    __syncthreads();
    __half* dst = vimg + (size_t)blockIdx.x * Dh * SP;                  // [Dh / 8][SP / 8][8][8]
    const int PG = SP / 8;
    // chunk (pg, ch): 8 pixels of one channel; consecutive threads take consecutive channels (conflict-free tile reads,
    // 8 consecutive threads fill one 128-byte core matrix)
    for (int i = threadIdx.x; i < PG * Dh; i += blockDim.x) {
        const int pg = i / Dh, ch = i % Dh;
        __half out[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int pix = pg * 8 + p;
            out[p] = pix < S ? tile[pix * Dh + ch] : __float2half(0.f);
        }
        *reinterpret_cast<uint4*>(dst + ((size_t)(ch >> 3) * PG + pg) * 64 + (ch & 7) * 8) = *reinterpret_cast<const uint4*>(out);
    }
}

// ---------------------------------------------------------------- phase timers (debug)
// g_tc_timing[i] accumulates SM-clock deltas of phase i, written by ONE thread per role and CTA when
// g_tc_timing_on != 0 (set through ver_debug_tc_timing); read back by tests/tools, never by the product.
__device__ unsigned long long g_tc_timing[32];
__device__ int g_tc_timing_on = 0;
struct PhaseTimer {
    bool on;
    long long t;
    __device__ __forceinline__ PhaseTimer(bool active) : on(active && g_tc_timing_on), t(0) {
        if (on) t = clock64();
    }
    __device__ __forceinline__ void lap(int slot) {
        if (on) {
            const long long n = clock64();
            atomicAdd(&g_tc_timing[slot], (unsigned long long)(n - t));
            t = n;
        }
    }
};

// ---------------------------------------------------------------- forward
constexpr int kTcThreads = 512;
constexpr int kTcWarps = kTcThreads / 32;
constexpr int kBuildWarps = kTcWarps - 1;   // warps 0..14 build A, warp 15 feeds the tensor cores
constexpr int kTZ = 4, kTH = 8, kTW = 8;
constexpr int kTV = kTZ * kTH * kTW;        // 256 voxel rows = two M=128 halves
constexpr int kMaxCam = 32;

struct TcFwdSmem {
    // byte offsets inside dynamic shared memory
    int v_bytes, a_half_bytes;
    int off_v0, off_v1, off_a0, off_a1, off_soff, off_saw, off_n, off_bits, off_list, off_cnt, off_trash, total;
    __host__ __device__ TcFwdSmem(int Dh, int SP) {
        v_bytes = Dh * SP * 2;
        a_half_bytes = 128 * SP * 2;
        off_v0 = 0;
        off_v1 = off_v0 + v_bytes;
        off_a0 = off_v1 + v_bytes;
        off_a1 = off_a0 + a_half_bytes;
        off_soff = off_a1 + a_half_bytes;
        off_saw = off_soff + kTV * 16 * 4;
        off_n = off_saw + kTV * 8 * 4;
        off_bits = off_n + kTV * 4;
        off_list = off_bits + kTV * 4;
        off_cnt = off_list + kMaxCam * kTV;
        off_trash = off_cnt + 2 * kMaxCam * 4;     // one fp16 sink per thread for out-of-map taps
        total = off_trash + kTcThreads * 2;
    }
};

template <int DH>
__global__ void __launch_bounds__(kTcThreads, 1)
sca_fwd_tc_kernel(const __half* __restrict__ vimg, const float* __restrict__ logits, int ld,
                  const float* __restrict__ rpc, const uint32_t* __restrict__ vis_bits,
                  __half* __restrict__ slots, int B, int Ncam, int Z, int H, int W, int Sh, int Sw, int SP,
                  int NH, int NP) {
    const int Nq = Z * H * W;
    const int G = SP >> 3;                       // 8-pixel groups per row
    extern __shared__ __align__(1024) unsigned char smem[];
    const TcFwdSmem L(DH, SP);
    float* s_off = reinterpret_cast<float*>(smem + L.off_soff);
    float* s_aw = reinterpret_cast<float*>(smem + L.off_saw);
    int* s_n = reinterpret_cast<int*>(smem + L.off_n);
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(smem + L.off_bits);
    uint8_t* s_list = smem + L.off_list;
    int* s_cnt = reinterpret_cast<int*>(smem + L.off_cnt);
    __shared__ __align__(8) uint64_t bar_v[2], bar_mma[2], bar_built[2];
    __shared__ uint32_t s_union, s_tmem;

    const int b = blockIdx.z, h = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_w = (W + kTW - 1) / kTW, tiles_h = (H + kTH - 1) / kTH;
    const int tw = blockIdx.x % tiles_w, th = (blockIdx.x / tiles_w) % tiles_h,
              tz = blockIdx.x / (tiles_w * tiles_h);

    PhaseTimer tm(tid == 0), tm_mma(tid == kBuildWarps * 32);
    // The offset / attention-logit rows of "my" voxels do not depend on anything computed below: issue
    // the global loads first so that their latency overlaps barrier init, TMEM allocation and the
    // visibility-list construction.
    constexpr int kIter = kTV * 8 / kTcThreads;           // 4 voxels per 8-lane group
    float2 off[kIter];
    float lg[kIter];
#pragma unroll
    for (int it = 0; it < kIter; ++it) {
        const int v = (tid >> 3) + it * (kTcThreads / 8);
        const int w = tw * kTW + (v % kTW), hh = th * kTH + (v / kTW) % kTH, z = tz * kTZ + v / (kTW * kTH);
        off[it] = make_float2(0.f, 0.f);
        lg[it] = -INFINITY;
        if (w < W && hh < H && z < Z && (lane & 7) < NP) {
            const float* row = logits + ((size_t)b * Nq + (z * H + hh) * W + w) * ld;
            off[it] = reinterpret_cast<const float2*>(row + h * NP * 2)[lane & 7];
            lg[it] = row[NH * NP * 2 + h * NP + (lane & 7)];
        }
    }
    if (tid == 0) {
        mbar_init(&bar_v[0], 1);
        mbar_init(&bar_v[1], 1);
        mbar_init(&bar_mma[0], 1);
        mbar_init(&bar_mma[1], 1);
        mbar_init(&bar_built[0], kBuildWarps);
        mbar_init(&bar_built[1], kBuildWarps);
        mbar_fence_init();
        s_union = 0;
    }
    if (tid < 2 * kMaxCam) s_cnt[tid] = 0;
    if (warp == 1) tmem_alloc(&s_tmem, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    // ---- voxel ids, visibility, per-camera voxel lists
    if (tid < kTV) {
        const int v = tid;
        const int w = tw * kTW + (v % kTW), hh = th * kTH + (v / kTW) % kTH, z = tz * kTZ + v / (kTW * kTH);
        int n = -1;
        uint32_t bits = 0;
        if (w < W && hh < H && z < Z) {
            n = (z * H + hh) * W + w;
            bits = vis_bits[(size_t)b * Nq + n];
        }
        s_n[v] = n;
        s_bits[v] = bits;
        if (bits) atomicOr(&s_union, bits);
        for (uint32_t r = bits; r; r &= r - 1) {      // per (camera, 128-row half) lists of visible rows
            const int c = __ffs(r) - 1, hf = v >> 7;
            s_list[(c * 2 + hf) * 128 + atomicAdd(&s_cnt[c * 2 + hf], 1)] = (uint8_t)(v & 127);
        }
    }
    __syncthreads();
    const uint32_t cams = s_union;
    tm.lap(0);                                   // setup: barriers, TMEM alloc, voxel ids + lists
    const size_t v_stride_bv = (size_t)NH * DH * SP;         // halves per view
    const __half* vbase = vimg + (size_t)b * Ncam * v_stride_bv + (size_t)h * DH * SP;
    if (tid == 0 && cams) {                                  // first two camera maps start streaming in
        const int c0 = __ffs(cams) - 1;
        mbar_expect_tx(&bar_v[0], L.v_bytes);
        bulk_g2s(smem + L.off_v0, vbase + (size_t)c0 * v_stride_bv, L.v_bytes, &bar_v[0]);
        const uint32_t r1 = cams & (cams - 1);
        if (r1) {
            const int c1 = __ffs(r1) - 1;
            mbar_expect_tx(&bar_v[1], L.v_bytes);
            bulk_g2s(smem + L.off_v1, vbase + (size_t)c1 * v_stride_bv, L.v_bytes, &bar_v[1]);
        }
    }
    // ---- per-voxel offsets / softmax of this head (loads were issued at kernel entry)
    {
        const int p = lane & 7;
#pragma unroll
        for (int it = 0; it < kIter; ++it) {
            const int v = (tid >> 3) + it * (kTcThreads / 8);
            float m = lg[it];
            m = fmaxf(m, __shfl_xor_sync(VER_FULL_MASK, m, 1));
            m = fmaxf(m, __shfl_xor_sync(VER_FULL_MASK, m, 2));
            m = fmaxf(m, __shfl_xor_sync(VER_FULL_MASK, m, 4));
            const float e = (lg[it] > -INFINITY) ? expf(lg[it] - m) : 0.f;
            float s = e;
            s += __shfl_xor_sync(VER_FULL_MASK, s, 1);
            s += __shfl_xor_sync(VER_FULL_MASK, s, 2);
            s += __shfl_xor_sync(VER_FULL_MASK, s, 4);
            s_off[v * 16 + 2 * p] = off[it].x / (float)Sw;
            s_off[v * 16 + 2 * p + 1] = off[it].y / (float)Sh;
            s_aw[v * 8 + p] = s > 0.f ? e / s : 0.f;
        }
    }
    __syncthreads();

    tm.lap(1);                                   // logits -> offsets / softmax
    tm_mma.lap(8);
    constexpr uint32_t idesc = umma_idesc(128, DH, 0, 0);
    // ---- cameras in ascending order.  Warp roles: warps 0..14 build the interpolation matrix (two
    //      128-row halves, so building one half overlaps the tensor-core work on the other), warp 15
    //      streams value images in (TMA) and issues the MMAs.  Hand-offs are mbarriers only:
    //        bar_built[hf]  builders -> MMA warp   (one arrival per builder warp)
    //        bar_mma[hf]    MMA retired (tcgen05.commit) -> builders may overwrite A[hf]; V buffer reuse
    //        bar_v[buf]     value image landed
    int k = 0;
    if (warp == kBuildWarps) {
        if (lane == 0) {
            for (uint32_t rest = cams; rest; rest &= rest - 1, ++k) {
                for (int hf = 0; hf < 2; ++hf) {
                    if (k > 0) mbar_wait(&bar_mma[hf], (k - 1) & 1);   // safe: phase k of this half not issued yet
                    if (hf == 1 && k > 0) {
                        // both halves of camera k-1 retired -> V buffer (k+1)&1 is free: prefetch camera k+1
                        const uint32_t nxt = rest & (rest - 1);
                        if (nxt) {
                            const int cn = __ffs(nxt) - 1;
                            mbar_expect_tx(&bar_v[(k + 1) & 1], L.v_bytes);
                            bulk_g2s(smem + (((k + 1) & 1) ? L.off_v1 : L.off_v0),
                                     vbase + (size_t)cn * v_stride_bv, L.v_bytes, &bar_v[(k + 1) & 1]);
                        }
                    }
                    mbar_wait(&bar_built[hf], k & 1);
                    tm_mma.lap(9);                       // MMA warp: waiting for builders
                    tc_fence_after();
                    if (hf == 0) mbar_wait(&bar_v[k & 1], (k >> 1) & 1);
                    tm_mma.lap(10);                      // MMA warp: waiting for the value image
                    const uint32_t a_addr = smem_u32(smem + (hf ? L.off_a1 : L.off_a0));
                    const uint32_t v_addr = smem_u32(smem + ((k & 1) ? L.off_v1 : L.off_v0));
                    for (int ks = 0; ks < SP / 16; ++ks)
                        umma_f16(tmem + hf * DH, umma_desc(a_addr + ks * 256, 128, G * 128),
                                 umma_desc(v_addr + ks * 256, 128, G * 128), idesc, (k > 0 || ks > 0) ? 1u : 0u);
                    umma_commit(&bar_mma[hf]);
                    tm_mma.lap(11);                      // MMA warp: issuing
                }
            }
        }
        k = __popc(cams);
    } else {
        constexpr int kBT = kBuildWarps * 32;                  // 480 builder threads
        constexpr int kRounds = (128 * 8 + kBT - 1) / kBT;     // (row, point) pairs of a half / builders
        const int p = tid & 7;                                 // kBT % 8 == 0: my sampling point
        // my (row, reference point) pairs of one (camera, half); loaded ONE ITERATION AHEAD so the
        // L2/HBM latency of reference_points_cam hides behind the previous half's work
        auto prefetch = [&](int c, int hf, float2 (&rf)[kRounds], int (&rw)[kRounds]) {
            const float2* rp = reinterpret_cast<const float2*>(rpc) + ((size_t)c * B + b) * Nq;
            const int cnt = s_cnt[c * 2 + hf];
#pragma unroll
            for (int rr = 0; rr < kRounds; ++rr) {
                const int q = tid + rr * kBT;
                rw[rr] = -1;
                if (q < cnt * 8 && p < NP) {
                    rw[rr] = s_list[(c * 2 + hf) * 128 + (q >> 3)];
                    rf[rr] = rp[s_n[hf * 128 + rw[rr]]];
                }
            }
        };
        float2 ref[kRounds];
        int row[kRounds];
#pragma unroll
        for (int rr = 0; rr < kRounds; ++rr) row[rr] = -1;
        if (cams) prefetch(__ffs(cams) - 1, 0, ref, row);
        for (uint32_t rest = cams; rest; rest &= rest - 1, ++k) {
            const int c = __ffs(rest) - 1;
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                __half* A = reinterpret_cast<__half*>(smem + (hf ? L.off_a1 : L.off_a0));
                const int trash_rel = (int)(reinterpret_cast<__half*>(smem + L.off_trash) - A) + tid;
                float2 nref[kRounds];
                int nrow[kRounds];
#pragma unroll
                for (int rr = 0; rr < kRounds; ++rr) nrow[rr] = -1;
                if (hf == 0) {
                    prefetch(c, 1, nref, nrow);
                } else {
                    const uint32_t nxt = rest & (rest - 1);
                    if (nxt) prefetch(__ffs(nxt) - 1, 0, nref, nrow);
                }
                if (k > 0) mbar_wait(&bar_mma[hf], (k - 1) & 1);  // previous camera's MMAs on this half retired
                tm.lap(2);                               // builders: prefetch issue + wait for MMA retire
                for (int i = tid; i < L.a_half_bytes / 16; i += kBT)
                    reinterpret_cast<uint4*>(A)[i] = make_uint4(0, 0, 0, 0);
                asm volatile("bar.sync 1, %0;" ::"n"(kBT) : "memory");
                tm.lap(3);                               // builders: zero + barrier
                // ---- taps of my (row, point) pairs: 4 bilinear corners each
                int toff[kRounds][4];
                float twgt[kRounds][4];
#pragma unroll
                for (int rr = 0; rr < kRounds; ++rr) {
#pragma unroll
                    for (int cn = 0; cn < 4; ++cn) {           // out-of-map taps: weight 0 into my private sink
                        toff[rr][cn] = trash_rel;
                        twgt[rr][cn] = 0.f;
                    }
                    if (row[rr] >= 0) {
                        const int r = row[rr], v = hf * 128 + r;
                        const float aw = s_aw[v * 8 + p];
                        const float x = (ref[rr].x + s_off[v * 16 + 2 * p]) * (float)Sw - 0.5f;
                        const float y = (ref[rr].y + s_off[v * 16 + 2 * p + 1]) * (float)Sh - 0.5f;
                        if (x > -1.f && y > -1.f && x < (float)Sw && y < (float)Sh) {
                            const float xf = floorf(x), yf = floorf(y);
                            const float fx = x - xf, fy = y - yf;
                            const int x0 = (int)xf, y0 = (int)yf;
#pragma unroll
                            for (int cn = 0; cn < 4; ++cn) {
                                const int xi = x0 + (cn & 1), yi = y0 + (cn >> 1);
                                if (xi < 0 || xi >= Sw || yi < 0 || yi >= Sh) continue;
                                twgt[rr][cn] = aw * (((cn >> 1) ? fy : 1.f - fy) * ((cn & 1) ? fx : 1.f - fx));
                                toff[rr][cn] = img_off(r, yi * Sw + xi, G);
                            }
                        }
                    }
                }
                tm.lap(13);                              // builders: tap arithmetic (incl. exposed ref latency)
                // ---- commit.  Two points of a row may hit the same pixel, so the 8 points of a row (8 lanes
                // of one warp) write one after the other: plain read-modify-writes, deterministic, no atomics.
                // Pairs of different rounds belong to different rows and never collide.
#pragma unroll
                for (int rr = 0; rr < kRounds; ++rr) {
                    if (!__any_sync(VER_FULL_MASK, row[rr] >= 0)) continue;   // no pair of this warp in the round
#pragma unroll
                    for (int step = 0; step < 8; ++step) {
                        if (p == step) {
                            float old[4];                       // branch-free: 4 loads in flight, then 4 stores
#pragma unroll
                            for (int cn = 0; cn < 4; ++cn) old[cn] = __half2float(A[toff[rr][cn]]);
#pragma unroll
                            for (int cn = 0; cn < 4; ++cn) A[toff[rr][cn]] = __float2half_rn(old[cn] + twgt[rr][cn]);
                        }
                        __syncwarp();
                    }
                }
                tm.lap(4);                               // builders: commit
                proxy_fence();                   // generic-proxy writes of A -> async proxy (tensor core)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_built[hf]);
                tm.lap(5);                               // builders: fences + arrive
#pragma unroll
                for (int rr = 0; rr < kRounds; ++rr) {
                    ref[rr] = nref[rr];
                    row[rr] = nrow[rr];
                }
            }
        }
    }
    __syncwarp();
    tm.lap(6);
    // ---- epilogue: slots = sum / max(count, 1)
    if (cams) {
        mbar_wait(&bar_mma[0], (k - 1) & 1);
        mbar_wait(&bar_mma[1], (k - 1) & 1);
    }
    tc_fence_after();
    tm.lap(12);                                  // epilogue: wait for the last MMAs
    {
        const int q = warp & 3, grp = warp >> 2;          // TMEM lane quarter, column group
        const int hf = grp >> 1, cpart = grp & 1;         // 4 groups = 2 halves x 2 channel halves
        const int v = hf * 128 + q * 32 + lane;
        const int n = s_n[v];
        const float inv_cnt = 1.f / (float)max(__popc(s_bits[v]), 1);   // fp16 output: reciprocal multiply
        constexpr int CH = DH / 2;                        // channels per thread
        __half* dst = (n >= 0) ? slots + ((size_t)b * Nq + n) * NH * DH + h * DH + cpart * CH : nullptr;
#pragma unroll
        for (int c0 = 0; c0 < CH; c0 += 16) {
            float vv[16];
            if (cams) {
                tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + hf * DH + cpart * CH + c0, vv);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) vv[i] = 0.f;
            }
            if (dst) {
                float o[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = vv[i] * inv_cnt;
                store_channels16<16>(dst + c0, o);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tm.lap(7);                                   // epilogue: wait last MMAs, TMEM -> slots
    if (warp == 1) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------- backward
constexpr int kBtThreads = 256;
constexpr int kHitsPerChunk = 128;
constexpr int kRnd = kHitsPerChunk * 8 / kBtThreads;    // (hit, point) pairs per thread and chunk

struct TcBwdSmem {
    int v_bytes, a_bytes, g_bytes;
    int off_v, off_a, off_g, off_n, off_trash, total;
    __host__ __device__ TcBwdSmem(int Dh, int SP) {
        v_bytes = Dh * SP * 2;
        a_bytes = kHitsPerChunk * SP * 2;               // also reused as the fp32 Dots staging buffer
        g_bytes = kHitsPerChunk * Dh * 2 + 1024;        // + slack: M=128 over-reads past ch < Dh
        off_v = 0;
        off_a = off_v + v_bytes;
        off_g = off_a + a_bytes;
        // the fp32 Dots staging buffer (128 rows x dstride words) overlays A' and G once they are dead
        const int half_cols = ((SP / 2 + 15) / 16) * 16;
        const int dstride = half_cols + ((36 - (half_cols & 31)) & 31);
        const int dots_bytes = kHitsPerChunk * dstride * 4;
        const int ag = a_bytes + g_bytes;
        off_n = off_a + (ag > dots_bytes ? ag : dots_bytes);
        off_trash = off_n + kHitsPerChunk * 4;
        total = off_trash + kBtThreads * 2;
    }
};

struct HitTaps {          // per-point quantities a thread recomputes for "its" hit
    float aw, x, y;
};

template <int DH>
__global__ void __launch_bounds__(kBtThreads, 1)
sca_bwd_tc_kernel(const __half* __restrict__ vimg, const float* __restrict__ logits, int ld,
                  const float* __restrict__ rpc, const uint32_t* __restrict__ vis_bits,
                  const int32_t* __restrict__ counts, const int32_t* __restrict__ index,
                  const __half* __restrict__ gslots, float* __restrict__ gvalue,
                  float* __restrict__ glogits, int B, int Ncam, int Nq, int Sh, int Sw, int SP, int NH,
                  int NP) {
    const int S = Sh * Sw;
    const int G = SP >> 3;
    constexpr int CG = DH / 8;                           // channel groups
    extern __shared__ __align__(1024) unsigned char smem[];
    const TcBwdSmem L(DH, SP);
    __half* Aimg = reinterpret_cast<__half*>(smem + L.off_a);
    float* dots = reinterpret_cast<float*>(smem + L.off_a);      // reused after the MMAs retire
    __half* Gimg = reinterpret_cast<__half*>(smem + L.off_g);
    int* s_n = reinterpret_cast<int*>(smem + L.off_n);
    __shared__ __align__(8) uint64_t bar_v, bar_mma;
    __shared__ uint32_t s_tmem;

    const int bv = blockIdx.y, h = blockIdx.x;
    const int b = bv / Ncam, cam = bv % Ncam;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        mbar_init(&bar_v, 1);
        mbar_init(&bar_mma, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t tm_dv = tmem;                  // dV^T : lanes = channel, columns [0, SP) = pixel
    const uint32_t tm_dots = tmem + 256;          // Dots : lanes = hit,     columns [0, SP) = pixel
    PhaseTimer tb(tid == 0);
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

    uint32_t phase = 0;
    const int nchunks = (nitems + kHitsPerChunk - 1) / kHitsPerChunk;
    int next_n = (tid < kHitsPerChunk && tid < nitems) ? idx[tid] : -1;
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int base = chunk * kHitsPerChunk;
        const int rows = min(kHitsPerChunk, nitems - base);
        // ---- zero A', publish the chunk's voxel ids
        for (int i = tid; i < L.a_bytes / 16; i += kBtThreads)
            reinterpret_cast<uint4*>(Aimg)[i] = make_uint4(0, 0, 0, 0);
        if (tid < kHitsPerChunk) s_n[tid] = tid < rows ? next_n : -1;
        // hit ids of the NEXT chunk: in flight during this chunk's work
        if (tid < kHitsPerChunk && base + kHitsPerChunk + tid < nitems) next_n = idx[base + kHitsPerChunk + tid];
        __syncthreads();
        tb.lap(16);                                  // zero A' + hit ids
        // ---- build.  G image: every thread first puts its 16-byte gathers of the hits' grad_slots rows in
        //      flight; A' rows: one thread per (hit, point), the 8 lanes of a hit share the softmax via
        //      shuffles; the G stores come last, when the loads have landed behind the A' work.
        constexpr int kG = (kHitsPerChunk * CG + kBtThreads - 1) / kBtThreads;
        uint4 gval[kG];
#pragma unroll
        for (int u = 0; u < kG; ++u) {
            const int i = tid + u * kBtThreads;
            gval[u] = make_uint4(0, 0, 0, 0);
            if (i < kHitsPerChunk * CG) {
                const int n = s_n[i / CG];
                if (n >= 0)
                    gval[u] = *reinterpret_cast<const uint4*>(gslots + ((size_t)b * Nq + n) * NH * DH + h * DH +
                                                              (i % CG) * 8);
            }
        }
        float q_aw[kRnd], q_x[kRnd], q_y[kRnd], q_ic[kRnd];
        int q_n[kRnd];
        {
            const int p = tid & 7;
            float2 off[kRnd], ref[kRnd];
            float lg[kRnd];
            uint32_t vb[kRnd];
#pragma unroll
            for (int rr = 0; rr < kRnd; ++rr) {           // all global loads first
                const int n = s_n[(tid >> 3) + rr * (kBtThreads / 8)];
                q_n[rr] = n;
                off[rr] = ref[rr] = make_float2(0.f, 0.f);
                lg[rr] = -INFINITY;
                vb[rr] = 0;
                if (n >= 0) {
                    const float* row = logits + ((size_t)b * Nq + n) * ld;
                    if (p < NP) {
                        off[rr] = reinterpret_cast<const float2*>(row + h * NP * 2)[p];
                        lg[rr] = row[NH * NP * 2 + h * NP + p];
                    }
                    ref[rr] = rp[n];
                    vb[rr] = vis_bits[(size_t)b * Nq + n];
                }
            }
            int toff[kRnd][4];
            float twgt[kRnd][4];
            const int trash_rel = (int)(reinterpret_cast<__half*>(smem + L.off_trash) - Aimg) + tid;
#pragma unroll
            for (int rr = 0; rr < kRnd; ++rr) {
                const int r = (tid >> 3) + rr * (kBtThreads / 8);
                float m = lg[rr];
                m = fmaxf(m, __shfl_xor_sync(VER_FULL_MASK, m, 1));
                m = fmaxf(m, __shfl_xor_sync(VER_FULL_MASK, m, 2));
                m = fmaxf(m, __shfl_xor_sync(VER_FULL_MASK, m, 4));
                const float e = lg[rr] > -INFINITY ? expf(lg[rr] - m) : 0.f;
                float sm = e;
                sm += __shfl_xor_sync(VER_FULL_MASK, sm, 1);
                sm += __shfl_xor_sync(VER_FULL_MASK, sm, 2);
                sm += __shfl_xor_sync(VER_FULL_MASK, sm, 4);
                q_aw[rr] = sm > 0.f ? e / sm : 0.f;
                q_ic[rr] = 1.f / (float)max(__popc(vb[rr]), 1);
                q_x[rr] = (ref[rr].x + off[rr].x / (float)Sw) * (float)Sw - 0.5f;
                q_y[rr] = (ref[rr].y + off[rr].y / (float)Sh) * (float)Sh - 0.5f;
                const float x = q_x[rr], y = q_y[rr];
#pragma unroll
                for (int cn = 0; cn < 4; ++cn) {
                    toff[rr][cn] = trash_rel;
                    twgt[rr][cn] = 0.f;
                }
                if (q_n[rr] >= 0 && p < NP && x > -1.f && y > -1.f && x < (float)Sw && y < (float)Sh) {
                    const float xf = floorf(x), yf = floorf(y);
                    const float fx = x - xf, fy = y - yf;
                    const int x0 = (int)xf, y0 = (int)yf;
#pragma unroll
                    for (int cn = 0; cn < 4; ++cn) {
                        const int xi = x0 + (cn & 1), yi = y0 + (cn >> 1);
                        if (xi < 0 || xi >= Sw || yi < 0 || yi >= Sh) continue;
                        twgt[rr][cn] = q_ic[rr] * q_aw[rr] * (((cn >> 1) ? fy : 1.f - fy) * ((cn & 1) ? fx : 1.f - fx));
                        toff[rr][cn] = img_off(r, yi * Sw + xi, G);
                    }
                }
            }
            // the 8 points of a hit (8 lanes of one warp) commit their taps one after the other; the
            // pairs a thread holds from different rounds belong to different hits and never collide
#pragma unroll
            for (int step = 0; step < 8; ++step) {
                if (p == step) {
                    float old[kRnd][4];
#pragma unroll
                    for (int rr = 0; rr < kRnd; ++rr)
#pragma unroll
                        for (int cn = 0; cn < 4; ++cn) old[rr][cn] = __half2float(Aimg[toff[rr][cn]]);
#pragma unroll
                    for (int rr = 0; rr < kRnd; ++rr)
#pragma unroll
                        for (int cn = 0; cn < 4; ++cn) Aimg[toff[rr][cn]] = __float2half_rn(old[rr][cn] + twgt[rr][cn]);
                }
                __syncwarp();
            }
        }
        tb.lap(18);                                  // A' rows (+ latency of the G gathers)
#pragma unroll
        for (int u = 0; u < kG; ++u) {
            const int i = tid + u * kBtThreads;
            if (i < kHitsPerChunk * CG) {
                const int r = i / CG, cg = i % CG;
                *reinterpret_cast<uint4*>(Gimg + ((r >> 3) * CG + cg) * 64 + (r & 7) * 8) = gval[u];
            }
        }
        tb.lap(17);                                  // G image stores
        proxy_fence();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            if (chunk == 0) mbar_wait(&bar_v, 0);
            const uint32_t a_addr = smem_u32(Aimg), g_addr = smem_u32(Gimg), v_addr = smem_u32(smem + L.off_v);
            // dV^T[ch, pix] += G^T[ch, hit] A'[hit, pix]   (A: G image MN-major; B: A' image MN-major)
            for (int ks = 0; ks < kHitsPerChunk / 16; ++ks)
                umma_f16(tm_dv, umma_desc(g_addr + ks * 2 * CG * 128, CG * 128, 128),
                         umma_desc(a_addr + ks * 2 * G * 128, G * 128, 128), idesc_dv | n_bits,
                         (chunk > 0 || ks > 0) ? 1u : 0u);
            // Dots[hit, pix] = G[hit, ch] V[pix, ch]^T     (A: G image K-major; B: V image MN-major)
            for (int ks = 0; ks < DH / 16; ++ks)
                umma_f16(tm_dots, umma_desc(g_addr + ks * 256, 128, CG * 128),
                         umma_desc(v_addr + ks * 2 * G * 128, G * 128, 128), idesc_dots | n_bits, ks > 0 ? 1u : 0u);
            umma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, phase);
        phase ^= 1;
        tc_fence_after();
        tb.lap(19);                                  // fences, MMA issue, MMA execution
        // ---- Dots: TMEM -> shared (fp32, two column halves through the retired A'/G buffers), then
        //      every hit thread picks the 32 taps of its row
        float ga[kRnd], gx[kRnd], gy[kRnd];
#pragma unroll
        for (int rr = 0; rr < kRnd; ++rr) ga[rr] = gx[rr] = gy[rr] = 0.f;
        const int half_cols = ((SP / 2 + 15) / 16) * 16;         // 112 for SP = 208
        // row stride of the fp32 staging buffer: == 4 (mod 32) words, so the 16-byte row stores of 8
        // consecutive hits fall in 8 different bank groups (112 would be a 16-way conflict)
        const int dstride = half_cols + ((36 - (half_cols & 31)) & 31);
        for (int part = 0; part < 2; ++part) {
            const int col0 = part * half_cols;
            const int ncols = min(half_cols, SP - col0);
            {   // all 8 warps dump: warp w reads TMEM lane quarter w & 3, column chunks split by w >> 2
                const int hit = (warp & 3) * 32 + lane;
                const int nch = ncols / 16, cbeg = (warp >> 2) ? (nch + 1) / 2 : 0,
                          cend = (warp >> 2) ? nch : (nch + 1) / 2;
                for (int ci = cbeg; ci < cend; ++ci) {
                    float vv[16];
                    tmem_ld16(tm_dots + ((uint32_t)((warp & 3) * 32) << 16) + col0 + ci * 16, vv);
                    float4* d = reinterpret_cast<float4*>(dots + hit * dstride + ci * 16);
#pragma unroll
                    for (int i = 0; i < 4; ++i) d[i] = make_float4(vv[4 * i], vv[4 * i + 1], vv[4 * i + 2], vv[4 * i + 3]);
                }
            }
            __syncthreads();
#pragma unroll
            for (int rr = 0; rr < kRnd; ++rr) {
                const float x = q_x[rr], y = q_y[rr];
                if (q_n[rr] < 0 || (tid & 7) >= NP || !(x > -1.f && y > -1.f && x < (float)Sw && y < (float)Sh))
                    continue;
                const int r = (tid >> 3) + rr * (kBtThreads / 8);
                const float xf = floorf(x), yf = floorf(y);
                const float fx = x - xf, fy = y - yf;
                const int x0 = (int)xf, y0 = (int)yf;
#pragma unroll
                for (int cn = 0; cn < 4; ++cn) {
                    const int xi = x0 + (cn & 1), yi = y0 + (cn >> 1);
                    if (xi < 0 || xi >= Sw || yi < 0 || yi >= Sh) continue;
                    const int pix = yi * Sw + xi;
                    if (pix < col0 || pix >= col0 + ncols) continue;
                    const float d = dots[r * dstride + pix - col0];
                    const float wx = (cn & 1) ? fx : 1.f - fx, wy = (cn >> 1) ? fy : 1.f - fy;
                    ga[rr] += wy * wx * d;
                    gx[rr] += ((cn & 1) ? wy : -wy) * d;
                    gy[rr] += ((cn >> 1) ? wx : -wx) * d;
                }
            }
            __syncthreads();
        }
        tb.lap(20);                                  // Dots: TMEM -> smem -> taps
        // ---- softmax backward (8-lane groups) + accumulate into the per-voxel logit gradients
#pragma unroll
        for (int rr = 0; rr < kRnd; ++rr) {
            const float g = ga[rr] * q_ic[rr];
            float t = q_aw[rr] * g;
            t += __shfl_xor_sync(VER_FULL_MASK, t, 1);
            t += __shfl_xor_sync(VER_FULL_MASK, t, 2);
            t += __shfl_xor_sync(VER_FULL_MASK, t, 4);
            const int p = tid & 7;
            if (q_n[rr] >= 0 && p < NP) {
                float* grow = glogits + ((size_t)b * Nq + q_n[rr]) * ld;
                atomicAdd(grow + NH * NP * 2 + h * NP + p, q_aw[rr] * (g - t));
                // d loc = aw * size * sum(...), d offset = d loc / size  -> the size cancels
                atomicAdd(grow + h * NP * 2 + 2 * p, q_ic[rr] * q_aw[rr] * gx[rr]);
                atomicAdd(grow + h * NP * 2 + 2 * p + 1, q_ic[rr] * q_aw[rr] * gy[rr]);
            }
        }
        tc_fence_before();
        __syncthreads();           // A'/dots and G buffers are free for the next chunk
        tc_fence_after();
        tb.lap(21);                                  // softmax backward + atomics
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

template <int DH>
int launch_fwd_tc(const __half* vimg, const float* logits, int ld, const float* rpc, const uint32_t* vis_bits,
                  __half* slots, int B, int Ncam, int Z, int H, int W, int Sh, int Sw, int SP, int NH, int NP,
                  cudaStream_t st) {
    const TcFwdSmem L(DH, SP);
    VER_CHECK_ARG(L.total + 2048 <= ver_device_max_smem_optin(), "TC forward needs %d B of shared memory", L.total);
    auto kern = sca_fwd_tc_kernel<DH>;
    VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    const int tiles = ((W + kTW - 1) / kTW) * ((H + kTH - 1) / kTH) * ((Z + kTZ - 1) / kTZ);
    kern<<<dim3(tiles, NH, B), kTcThreads, L.total, st>>>(vimg, logits, ld, rpc, vis_bits, slots, B, Ncam, Z, H, W,
                                                         Sh, Sw, SP, NH, NP);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

template <int DH>
int launch_bwd_tc(const __half* vimg, const float* logits, int ld, const float* rpc, const uint32_t* vis_bits,
                  const int32_t* counts, const int32_t* index, const __half* gslots, float* gvalue,
                  float* glogits, int B, int Ncam, int Nq, int Sh, int Sw, int SP, int NH, int NP,
                  cudaStream_t st) {
    const TcBwdSmem L(DH, SP);
    VER_CHECK_ARG(L.total + 2048 <= ver_device_max_smem_optin(), "TC backward needs %d B of shared memory", L.total);
    auto kern = sca_bwd_tc_kernel<DH>;
    VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    VER_CHECK_CUDA(cudaMemset2DAsync(glogits, (size_t)ld * sizeof(float), 0, (size_t)NH * NP * 3 * sizeof(float),
                                     (size_t)B * Nq, st));
    kern<<<dim3(NH, B * Ncam), kBtThreads, L.total, st>>>(vimg, logits, ld, rpc, vis_bits, counts, index, gslots,
                                                         gvalue, glogits, B, Ncam, Nq, Sh, Sw, SP, NH, NP);
    VER_CHECK_LAUNCH();
    g_ver_launches += 2;
    return VER_OK;
}

}  // namespace

// ---- entry points used by sca.cu's dispatch (value_layout == VER_LAYOUT_TC_IMAGE) and the C ABI
extern "C" int ver_debug_tc_timing(int enable, unsigned long long* host_out32) {
    if (host_out32) VER_CHECK_CUDA(cudaMemcpyFromSymbol(host_out32, g_tc_timing, sizeof(unsigned long long) * 32));
    unsigned long long zero[32] = {0};
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_tc_timing, zero, sizeof(zero)));
    VER_CHECK_CUDA(cudaMemcpyToSymbol(g_tc_timing_on, &enable, sizeof(int)));
    return VER_OK;
}

int ver_tc_supported(int Ncam, int S, int Dh, int NP) {
    if (!(Ncam <= kMaxCam && NP >= 1 && NP <= 8 && S <= 256 && (Dh == 32 || Dh == 64 || Dh == 96 || Dh == 128)))
        return 0;
    const int SP = (S + 15) / 16 * 16;
    return TcFwdSmem(Dh, SP).total + 2048 <= ver_device_max_smem_optin() &&
           TcBwdSmem(Dh, SP).total + 2048 <= ver_device_max_smem_optin();
}

extern "C" int ver_value_image_f16(const void* value, void* vimg, int Bv, int S, int NH, int Dh,
                                   ver_stream_t stream) {
    VER_CHECK_ARG(value && vimg, "null pointer");
    VER_CHECK_ARG(Bv > 0 && S > 0 && NH > 0 && Dh > 0 && Dh % 8 == 0, "bad dims");
    const int SP = (S + 15) / 16 * 16;
    const size_t tile_bytes = (size_t)S * Dh * 2;
    if (tile_bytes <= 96 * 1024 && ((uintptr_t)value & 15) == 0) {          // staged transpose (one CTA per view and head)
        auto kern = value_image_tiled_kernel;
        if (tile_bytes > 48 * 1024)
            VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes));
        kern<<<Bv * NH, 256, tile_bytes, (cudaStream_t)stream>>>((const __half*)value, (__half*)vimg, S, NH, Dh, SP);
    } else {
        const size_t chunks = (size_t)Bv * NH * (Dh / 8) * (SP / 8) * 8;
        const int blocks = (int)((chunks + 255) / 256 > 148 * 32 ? 148 * 32 : (chunks + 255) / 256);
        value_image_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)value, (__half*)vimg, Bv, S, NH, Dh, SP);
    }
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

int ver_sca_forward_tc(const void* vimg, const float* logits, int ld, const float* rpc, const uint32_t* vis_bits,
                       void* slots, int B, int Ncam, int Z, int H, int W, int Sh, int Sw, int NH, int Dh, int NP,
                       cudaStream_t st) {
    const int SP = (Sh * Sw + 15) / 16 * 16;
#define FWD(D) launch_fwd_tc<D>((const __half*)vimg, logits, ld, rpc, vis_bits, (__half*)slots, B, Ncam, Z, H, W, Sh, Sw, SP, NH, NP, st)
    switch (Dh) {
        case 32: return FWD(32);
        case 64: return FWD(64);
        case 96: return FWD(96);
        default: return FWD(128);
    }
#undef FWD
}

// debug: 0 = newest backward kernel that covers the shape (default), 1 = force sca_bwd_tc_kernel
static int g_bwd_variant = 0;
extern "C" int ver_debug_bwd_variant(int v) {
    g_bwd_variant = v;
    return VER_OK;
}

int ver_sca_backward_tc(const void* vimg, const float* logits, int ld, const float* rpc, const uint32_t* vis_bits,
                        const int32_t* counts, const int32_t* index, const void* gslots, float* gvalue,
                        float* glogits, int B, int Ncam, int Nq, int Sh, int Sw, int NH, int Dh, int NP,
                        cudaStream_t st) {
    if (g_bwd_variant != 1 && ver_bwd_tc2_supported(Ncam, Sh, Sw, Dh, NP, ld))
        return ver_sca_backward_tc2(vimg, logits, ld, rpc, vis_bits, counts, index, gslots, gvalue, glogits, B, Ncam,
                                    Nq, Sh, Sw, NH, Dh, NP, st);
    const int SP = (Sh * Sw + 15) / 16 * 16;
#define BWD(D) launch_bwd_tc<D>((const __half*)vimg, logits, ld, rpc, vis_bits, counts, index, (const __half*)gslots, gvalue, glogits, B, Ncam, Nq, Sh, Sw, SP, NH, NP, st)
    switch (Dh) {
        case 32: return BWD(32);
        case 64: return BWD(64);
        case 96: return BWD(96);
        default: return BWD(128);
    }
#undef BWD
}

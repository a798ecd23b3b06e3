// A3 (+) A4 (+) A5 fused: the sampling half of SpatialCrossAttention.forward
// (M/spatial_cross_attention.py:138-173) with MSDeformableAttention3D's softmax and
// location arithmetic (:340-374) folded in.
//
// Forward is voxel-tile centric: a CTA (16 warps) owns a compact 4 x 8 x 8 block of
// voxels and one head.  It walks the cameras that see any voxel of the block in
// ascending order (the reference's accumulation order), stages that
// camera's [S][Dh] value map in shared memory with the bulk-copy (TMA) engine --
// double buffered for fp16 maps, so the next camera streams in while the current
// one is sampled -- lets its warps sample the visible voxels (compacted per camera),
// and keeps the per-voxel sums in shared memory; the count-normalised slots are
// written exactly once.  No padded rebatch, no per-hit intermediate, no atomics, no
// host sync.
//
// Backward is camera centric (one CTA per (view, head), see sca_bwd.cuh) so that
// grad_value never needs atomics; the per-voxel logit gradients (<= Ncam addends,
// typically 1-3) are accumulated with fp32 atomics.
#include "sca_bwd.cuh"

namespace {

constexpr int kFwdThreads = 512;
constexpr int kFwdWarps = kFwdThreads / 32;
constexpr int kTZ = 4, kTH = 8, kTW = 8;
constexpr int kTV = kTZ * kTH * kTW;   // 256 voxels per CTA
constexpr int kMaxCam = 32;

template <typename T, int CPL>
struct FwdSmem {
    static constexpr int Dh = CPL * 8;
    static constexpr int NBUF = sizeof(T) == 2 ? 2 : 1;
    __host__ __device__ static size_t tile_bytes(int S) { return ((size_t)S * Dh * sizeof(T) + 127) / 128 * 128; }
    __host__ __device__ static size_t bytes(int S) {
        return NBUF * tile_bytes(S) + (size_t)kTV * Dh * sizeof(float)   // acc
               + (size_t)kTV * 16 * sizeof(float)                        // offsets (8 x float2)
               + (size_t)kTV * 8 * sizeof(float)                         // softmaxed weights
               + (size_t)kTV * 2 * sizeof(int)                           // voxel id, visibility bits
               + (size_t)kMaxCam * kTV                                   // per-camera voxel lists (u8)
               + (size_t)kMaxCam * sizeof(int);                          // per-camera counts
    }
};

template <typename T, int CPL>
__global__ void __launch_bounds__(kFwdThreads, 1)
sca_fwd_kernel(const T* __restrict__ value, MapLayout L, const float* __restrict__ logits, int ld,
               const float* __restrict__ rpc, const uint32_t* __restrict__ vis_bits,
               T* __restrict__ slots, int B, int Ncam, int Z, int H, int W, int Sh, int Sw, int NH,
               int NP) {
    constexpr int Dh = CPL * 8;
    constexpr int NBUF = FwdSmem<T, CPL>::NBUF;
    const int S = Sh * Sw;
    const int Nq = Z * H * W;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t tb = FwdSmem<T, CPL>::tile_bytes(S);
    float* s_acc = reinterpret_cast<float*>(smem_raw + NBUF * tb);
    float* s_off = s_acc + kTV * Dh;
    float* s_aw = s_off + kTV * 16;
    int* s_n = reinterpret_cast<int*>(s_aw + kTV * 8);
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(s_n + kTV);
    uint8_t* s_list = reinterpret_cast<uint8_t*>(s_bits + kTV);
    int* s_cnt = reinterpret_cast<int*>(s_list + kMaxCam * kTV);
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t s_union;

    const int b = blockIdx.z, h = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles_w = (W + kTW - 1) / kTW, tiles_h = (H + kTH - 1) / kTH;
    const int tw = blockIdx.x % tiles_w, th = (blockIdx.x / tiles_w) % tiles_h,
              tz = blockIdx.x / (tiles_w * tiles_h);

    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
        s_union = 0;
    }
    if (threadIdx.x < kMaxCam) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    // ---- phase 0: voxel ids, visibility, per-camera voxel lists of the block
    if (threadIdx.x < kTV) {
        const int v = threadIdx.x;
        const int w = tw * kTW + (v % kTW), hh = th * kTH + (v / kTW) % kTH,
                  z = tz * kTZ + v / (kTW * kTH);
        int n = -1;
        uint32_t bits = 0;
        if (w < W && hh < H && z < Z) {
            n = (z * H + hh) * W + w;
            bits = vis_bits[(size_t)b * Nq + n];
        }
        s_n[v] = n;
        s_bits[v] = bits;
        if (bits) atomicOr(&s_union, bits);
        for (uint32_t r = bits; r; r &= r - 1) {
            const int c = __ffs(r) - 1;
            s_list[c * kTV + atomicAdd(&s_cnt[c], 1)] = (uint8_t)v;   // order inside a camera is free
        }
    }
    for (int i = threadIdx.x; i < kTV * Dh / 4; i += kFwdThreads)
        reinterpret_cast<float4*>(s_acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const uint32_t cams = s_union;
    const T* vbase = value + (size_t)b * Ncam * L.s_bv + (size_t)h * L.s_h;
    // first camera's map starts streaming in while the prologue computes the softmaxes
    if (cams && warp == 0)
        stage_map(reinterpret_cast<T*>(smem_raw), vbase + (size_t)(__ffs(cams) - 1) * L.s_bv, S, Dh,
                  L.s_row, &bar[0], lane);
    // ---- phase 1: per-voxel offsets and softmax(attention logits) of this head
    {
        const int p = lane & 7;
        for (int v = threadIdx.x >> 3; v < kTV; v += kFwdThreads / 8) {
            const int n = s_n[v];
            if (n < 0 || s_bits[v] == 0) continue;      // uniform over the 8-lane group
            const float* row = logits + ((size_t)b * Nq + n) * ld;
            float2 off = make_float2(0.f, 0.f);
            float lg = -INFINITY;
            if (p < NP) {
                off = reinterpret_cast<const float2*>(row + h * NP * 2)[p];
                lg = row[NH * NP * 2 + h * NP + p];
            }
            const unsigned gm = 0xffu << (lane & 24);
            float m = lg;
            m = fmaxf(m, __shfl_xor_sync(gm, m, 1));
            m = fmaxf(m, __shfl_xor_sync(gm, m, 2));
            m = fmaxf(m, __shfl_xor_sync(gm, m, 4));
            const float e = (p < NP) ? expf(lg - m) : 0.f;
            float s = e;
            s += __shfl_xor_sync(gm, s, 1);
            s += __shfl_xor_sync(gm, s, 2);
            s += __shfl_xor_sync(gm, s, 4);
            // offsets / (W, H)  (spatial_cross_attention.py:359-365)
            s_off[v * 16 + 2 * p] = off.x / (float)Sw;
            s_off[v * 16 + 2 * p + 1] = off.y / (float)Sh;
            s_aw[v * 8 + p] = e / s;
        }
    }
    __syncthreads();
    // ---- phase 2: cameras in ascending order
    const int corner = lane >> 3, g = lane & 7;
    int k = 0;
    for (uint32_t rest = cams; rest; rest &= rest - 1, ++k) {
        const int c = __ffs(rest) - 1;
        const int buf = (NBUF == 2) ? (k & 1) : 0;
        const T* tile = reinterpret_cast<const T*>(smem_raw + buf * tb);
        if (NBUF == 2) {
            const uint32_t nxt = rest & (rest - 1);
            if (nxt && warp == 0)     // buffer (k+1)&1 was released by the barrier ending iteration k-1
                stage_map(reinterpret_cast<T*>(smem_raw + ((k + 1) & 1) * tb),
                          vbase + (size_t)(__ffs(nxt) - 1) * L.s_bv, S, Dh, L.s_row, &bar[(k + 1) & 1], lane);
            mbar_wait(&bar[buf], (k >> 1) & 1);
        } else {
            if (k > 0 && warp == 0)
                stage_map(reinterpret_cast<T*>(smem_raw), vbase + (size_t)c * L.s_bv, S, Dh, L.s_row, &bar[0], lane);
            mbar_wait(&bar[0], k & 1);
        }
        const float2* rp = reinterpret_cast<const float2*>(rpc) + ((size_t)c * B + b) * Nq;
        const int cnt = s_cnt[c];
        for (int j = warp; j < cnt; j += kFwdWarps) {
            const int v = s_list[c * kTV + j];
            const float2 ref = rp[s_n[v]];
            const float lx = ref.x + s_off[v * 16 + 2 * g];
            const float ly = ref.y + s_off[v * 16 + 2 * g + 1];
            const float aw = (g < NP) ? s_aw[v * 8 + g] : 0.f;
            const Tap tap = make_tap(lx, ly, aw, corner, Sh, Sw, Dh);
            float acc[CPL];
#pragma unroll
            for (int q = 0; q < CPL; ++q) acc[q] = 0.f;
            gather8<CPL>(tile, tap, lane, NP, acc);
            reduce_corners<CPL>(acc);
            if (lane < 8) {
                float4* dst = reinterpret_cast<float4*>(s_acc + v * Dh + g * CPL);
#pragma unroll
                for (int q = 0; q < CPL / 4; ++q) {
                    float4 o = dst[q];
                    o.x += acc[4 * q];
                    o.y += acc[4 * q + 1];
                    o.z += acc[4 * q + 2];
                    o.w += acc[4 * q + 3];
                    dst[q] = o;
                }
            }
        }
        __syncthreads();                 // this camera's map and accumulator updates are retired
    }
    // ---- phase 3: slots = sum / max(count, 1)   (spatial_cross_attention.py:170-173)
    for (int i = threadIdx.x; i < kTV * (Dh / 4); i += kFwdThreads) {
        const int v = i / (Dh / 4), c4 = (i % (Dh / 4)) * 4;
        const int n = s_n[v];
        if (n < 0) continue;
        const float cnt = (float)max(__popc(s_bits[v]), 1);
        const float4 a = *reinterpret_cast<const float4*>(s_acc + v * Dh + c4);
        float r[4] = {a.x / cnt, a.y / cnt, a.z / cnt, a.w / cnt};
        store_channels<4>(slots + ((size_t)b * Nq + n) * NH * Dh + h * Dh + c4, r);
    }
}

// ------------------------------------------------------------------ backward
template <typename T, int CPL>
__global__ void __launch_bounds__(kBwdThreads, 2)
sca_bwd_kernel(const T* __restrict__ value, MapLayout L, const float* __restrict__ logits, int ld,
               const float* __restrict__ rpc, const uint32_t* __restrict__ vis_bits,
               const int32_t* __restrict__ counts, const int32_t* __restrict__ index,
               const T* __restrict__ gslots, float* __restrict__ gvalue,
               float* __restrict__ glogits, int B, int Ncam, int Nq, int Sh, int Sw, int NH, int NP) {
    constexpr int Dh = CPL * 8;
    const int S = Sh * Sw;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BwdSmem<T, CPL> sm(smem_raw, S);
    const int bv = blockIdx.y, h = blockIdx.x;
    const int b = bv / Ncam, cam = bv % Ncam;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bwd_prologue<T, CPL>(sm, value + (size_t)bv * L.s_bv + (size_t)h * L.s_h, S, L.s_row, lane, warp);

    const int corner = lane >> 3, g = lane & 7;
    const int nitems = counts[bv];
    const int32_t* idx = index + (size_t)bv * Nq;
    const float2* rp = reinterpret_cast<const float2*>(rpc) + ((size_t)cam * B + b) * Nq;
    for (int k = 0; k < nitems; ++k) {
        const int n = idx[k];
        const float* row = logits + ((size_t)b * Nq + n) * ld;
        float2 off = make_float2(0.f, 0.f);
        float lg = -INFINITY;
        if (g < NP) {
            off = reinterpret_cast<const float2*>(row + h * NP * 2)[g];
            lg = row[NH * NP * 2 + h * NP + g];
        }
        float m = lg;
        m = fmaxf(m, __shfl_xor_sync(VER_FULL_MASK, m, 1));
        m = fmaxf(m, __shfl_xor_sync(VER_FULL_MASK, m, 2));
        m = fmaxf(m, __shfl_xor_sync(VER_FULL_MASK, m, 4));
        const float e = (g < NP) ? expf(lg - m) : 0.f;
        float s = e;
        s += __shfl_xor_sync(VER_FULL_MASK, s, 1);
        s += __shfl_xor_sync(VER_FULL_MASK, s, 2);
        s += __shfl_xor_sync(VER_FULL_MASK, s, 4);
        const float aw = e / s;
        const float2 ref = rp[n];
        const TapB tap = make_tap_bwd(ref.x + off.x / (float)Sw, ref.y + off.y / (float)Sh, aw,
                                      corner, Sh, Sw);
        const float cnt = (float)max(__popc(vis_bits[(size_t)b * Nq + n]), 1);
        const float3 gr = bwd_process_item<T, CPL>(
            sm, tap, gslots + ((size_t)b * Nq + n) * NH * Dh + h * Dh, 1.f / cnt, Sh, Sw, NP, lane,
            warp, k);
        if (warp == (k & (kBwdWarps - 1))) {
            // softmax backward over the 8 points: d logit_p = aw_p (ga_p - sum_j aw_j ga_j)
            float t = aw * gr.x;
            t += __shfl_xor_sync(VER_FULL_MASK, t, 1);
            t += __shfl_xor_sync(VER_FULL_MASK, t, 2);
            t += __shfl_xor_sync(VER_FULL_MASK, t, 4);
            if (lane < 8 && g < NP) {
                float* grow = glogits + ((size_t)b * Nq + n) * ld;
                atomicAdd(grow + NH * NP * 2 + h * NP + g, aw * (gr.x - t));
                atomicAdd(grow + h * NP * 2 + 2 * g, gr.y / (float)Sw);
                atomicAdd(grow + h * NP * 2 + 2 * g + 1, gr.z / (float)Sh);
            }
        }
    }
    bwd_epilogue<T, CPL>(sm, gvalue + (size_t)bv * L.s_bv + (size_t)h * L.s_h, S, L.s_row, Sw, lane, warp);
}

template <typename T, int CPL>
int launch_sca_fwd(const T* value, int layout, const float* logits, int ld, const float* rpc,
                   const uint32_t* vis_bits, T* slots, int B, int Ncam, int Z, int H, int W, int Sh,
                   int Sw, int NH, int NP, cudaStream_t st) {
    const size_t smem = FwdSmem<T, CPL>::bytes(Sh * Sw);
    VER_CHECK_ARG(smem + 1024 <= (size_t)ver_device_max_smem_optin(),
                  "feature map %dx%dx%d does not fit in shared memory", Sh, Sw, CPL * 8);
    auto kern = sca_fwd_kernel<T, CPL>;
    VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = ((W + kTW - 1) / kTW) * ((H + kTH - 1) / kTH) * ((Z + kTZ - 1) / kTZ);
    dim3 grid(tiles, NH, B);
    kern<<<grid, kFwdThreads, smem, st>>>(value, make_layout(layout, Sh * Sw, NH, CPL * 8), logits, ld,
                                          rpc, vis_bits, slots, B, Ncam, Z, H, W, Sh, Sw, NH, NP);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

template <typename T, int CPL>
int launch_sca_bwd(const T* value, int layout, const float* logits, int ld, const float* rpc,
                   const uint32_t* vis_bits, const int32_t* counts, const int32_t* index,
                   const T* gslots, float* gvalue, float* glogits, int B, int Ncam, int Nq, int Sh,
                   int Sw, int NH, int NP, cudaStream_t st) {
    const size_t smem = BwdSmem<T, CPL>::bytes(Sh * Sw);
    VER_CHECK_ARG(smem + 1024 <= (size_t)ver_device_max_smem_optin(),
                  "feature map %dx%dx%d does not fit in shared memory", Sh, Sw, CPL * 8);
    auto kern = sca_bwd_kernel<T, CPL>;
    VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VER_CHECK_CUDA(cudaMemset2DAsync(glogits, (size_t)ld * sizeof(float), 0,
                                     (size_t)NH * NP * 3 * sizeof(float), (size_t)B * Nq, st));
    dim3 grid(NH, B * Ncam);
    kern<<<grid, kBwdThreads, smem, st>>>(value, make_layout(layout, Sh * Sw, NH, CPL * 8), logits, ld, rpc,
                                          vis_bits, counts, index, gslots, gvalue, glogits, B, Ncam, Nq,
                                          Sh, Sw, NH, NP);
    VER_CHECK_LAUNCH();
    g_ver_launches += 2;
    return VER_OK;
}

int check_sca(int dtype, int layout, const void* value, int B, int Ncam, int Z, int H, int W, int Sh,
              int Sw, int NH, int Dh, int NP, int ld) {
    VER_CHECK_ARG(dtype == VER_F32 || dtype == VER_F16, "bad dtype %d", dtype);
    VER_CHECK_ARG(layout == VER_LAYOUT_MMCV || layout == VER_LAYOUT_HEAD_MAJOR, "bad value layout %d", layout);
    VER_CHECK_ARG(B > 0 && Ncam > 0 && Z > 0 && H > 0 && W > 0 && Sh > 0 && Sw > 0 && NH > 0,
                  "non-positive dimension");
    VER_CHECK_ARG(ld >= NH * NP * 3, "ld_logits %d < NH*NP*3 = %d", ld, NH * NP * 3);
    VER_CHECK_ARG(((uintptr_t)value & 15) == 0, "value must be 16-byte aligned");
    if (Ncam > kMaxCam || NP < 1 || NP > 8 || !(Dh == 32 || Dh == 64 || Dh == 96 || Dh == 128) ||
        Sh * Sw > 65535) {
        ver_set_error("fused SCA supports Ncam<=32, NP<=8, Dh in {32,64,96,128}; got Ncam=%d NP=%d Dh=%d",
                      Ncam, NP, Dh);
        return VER_ERR_UNSUPPORTED;
    }
    return VER_OK;
}

}  // namespace

#define DISPATCH_CPL(Dh, CALL)                  \
    switch (Dh) {                               \
        case 32: { constexpr int CPL = 4; return CALL; }  \
        case 64: { constexpr int CPL = 8; return CALL; }  \
        case 96: { constexpr int CPL = 12; return CALL; } \
        default: { constexpr int CPL = 16; return CALL; } \
    }

extern "C" int ver_sca_forward(int dtype, const void* value, int value_layout, const float* logits,
                               int ld_logits, const float* rpc, const uint32_t* vis_bits, void* slots,
                               int B, int Ncam, int Z, int H, int W, int Sh, int Sw, int NH, int Dh,
                               int NP, ver_stream_t stream) {
    VER_CHECK_ARG(value && logits && rpc && vis_bits && slots, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (value_layout == VER_LAYOUT_TC_IMAGE) {
        int rc = check_sca(dtype, VER_LAYOUT_MMCV, value, B, Ncam, Z, H, W, Sh, Sw, NH, Dh, NP, ld_logits);
        if (rc) return rc;
        if (dtype != VER_F16 || !ver_tc_supported(Ncam, Sh * Sw, Dh, NP)) {
            ver_set_error("tensor-core sampler needs fp16 maps, S <= 256, Dh in {32,64,96,128}");
            return VER_ERR_UNSUPPORTED;
        }
        return ver_sca_forward_tc(value, logits, ld_logits, rpc, vis_bits, slots, B, Ncam, Z, H, W, Sh, Sw, NH, Dh,
                                  NP, st);
    }
    int rc = check_sca(dtype, value_layout, value, B, Ncam, Z, H, W, Sh, Sw, NH, Dh, NP, ld_logits);
    if (rc) return rc;
    if (dtype == VER_F32) {
        DISPATCH_CPL(Dh, (launch_sca_fwd<float, CPL>((const float*)value, value_layout, logits, ld_logits, rpc,
                                                     vis_bits, (float*)slots, B, Ncam, Z, H, W, Sh, Sw, NH, NP, st)));
    }
    DISPATCH_CPL(Dh, (launch_sca_fwd<__half, CPL>((const __half*)value, value_layout, logits, ld_logits, rpc,
                                                  vis_bits, (__half*)slots, B, Ncam, Z, H, W, Sh, Sw, NH, NP, st)));
}

extern "C" int ver_sca_backward(int dtype, const void* value, int value_layout, const float* logits,
                                int ld_logits, const float* rpc, const uint32_t* vis_bits,
                                const int32_t* counts, const int32_t* index, const void* grad_slots,
                                float* grad_value, float* grad_logits, int B, int Ncam, int Z, int H,
                                int W, int Sh, int Sw, int NH, int Dh, int NP, ver_stream_t stream) {
    VER_CHECK_ARG(value && logits && rpc && vis_bits && counts && index && grad_slots && grad_value &&
                      grad_logits, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int Nq = Z * H * W;
    if (value_layout == VER_LAYOUT_TC_IMAGE) {
        int rc = check_sca(dtype, VER_LAYOUT_MMCV, value, B, Ncam, Z, H, W, Sh, Sw, NH, Dh, NP, ld_logits);
        if (rc) return rc;
        if (dtype != VER_F16 || !ver_tc_supported(Ncam, Sh * Sw, Dh, NP)) {
            ver_set_error("tensor-core sampler needs fp16 maps, S <= 256, Dh in {32,64,96,128}");
            return VER_ERR_UNSUPPORTED;
        }
        return ver_sca_backward_tc(value, logits, ld_logits, rpc, vis_bits, counts, index, grad_slots, grad_value,
                                   grad_logits, B, Ncam, Nq, Sh, Sw, NH, Dh, NP, st);
    }
    int rc = check_sca(dtype, value_layout, value, B, Ncam, Z, H, W, Sh, Sw, NH, Dh, NP, ld_logits);
    if (rc) return rc;
    if (dtype == VER_F32) {
        DISPATCH_CPL(Dh, (launch_sca_bwd<float, CPL>((const float*)value, value_layout, logits, ld_logits, rpc,
                                                     vis_bits, counts, index, (const float*)grad_slots,
                                                     grad_value, grad_logits, B, Ncam, Nq, Sh, Sw, NH, NP, st)));
    }
    DISPATCH_CPL(Dh, (launch_sca_bwd<__half, CPL>((const __half*)value, value_layout, logits, ld_logits, rpc,
                                                  vis_bits, counts, index, (const __half*)grad_slots,
                                                  grad_value, grad_logits, B, Ncam, Nq, Sh, Sw, NH, NP, st)));
}

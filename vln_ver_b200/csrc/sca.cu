// A3 (+) A4 (+) A5 fused: the sampling half of SpatialCrossAttention.forward
// (M/spatial_cross_attention.py:138-173) with MSDeformableAttention3D's softmax and
// location arithmetic (:340-374) folded in.
//
// Forward is voxel-tile centric: a CTA owns a compact TZ x TH x TW block of voxels
// and one head.  It walks the cameras that see any voxel of the block in ascending
// order (the reference's accumulation order), stages that camera's
// [S][Dh] value map in shared memory with bulk async copies, lets its warps sample
// the visible voxels, and keeps the per-voxel sums in shared memory; the
// count-normalised slots are written exactly once.  No padded rebatch, no per-hit
// intermediate, no atomics, no host sync.
//
// Backward is camera centric (one CTA per (view, head), see sca_bwd.cuh) so that
// grad_value never needs atomics; the per-voxel logit gradients (<= Ncam addends,
// typically 1-3) are accumulated with fp32 atomics.
#include "sca_bwd.cuh"

namespace {

constexpr int kFwdThreads = 256;
constexpr int kFwdWarps = kFwdThreads / 32;
constexpr int kTZ = 4, kTH = 4, kTW = 8;
constexpr int kTV = kTZ * kTH * kTW;   // voxels per CTA

template <typename T, int CPL>
struct FwdSmem {
    static constexpr int Dh = CPL * 8;
    __host__ __device__ static size_t tile_bytes(int S) { return ((size_t)S * Dh * sizeof(T) + 127) / 128 * 128; }
    __host__ __device__ static size_t bytes(int S) {
        return tile_bytes(S) + (size_t)kTV * Dh * sizeof(float)   // acc
               + (size_t)kTV * 16 * sizeof(float)                 // offsets (8 x float2)
               + (size_t)kTV * 8 * sizeof(float)                  // softmaxed weights
               + (size_t)kTV * 2 * sizeof(int);                   // voxel id, visibility bits
    }
};

template <typename T, int CPL>
__global__ void __launch_bounds__(kFwdThreads)
sca_fwd_kernel(const T* __restrict__ value, const float* __restrict__ logits, int ld,
               const float* __restrict__ rpc, const uint32_t* __restrict__ vis_bits,
               T* __restrict__ slots, int B, int Ncam, int Z, int H, int W, int Sh, int Sw, int NH,
               int NP) {
    constexpr int Dh = CPL * 8;
    const int S = Sh * Sw;
    const int Nq = Z * H * W;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    float* s_acc = reinterpret_cast<float*>(smem_raw + FwdSmem<T, CPL>::tile_bytes(S));
    float* s_off = s_acc + kTV * Dh;
    float* s_aw = s_off + kTV * 16;
    int* s_n = reinterpret_cast<int*>(s_aw + kTV * 8);
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(s_n + kTV);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_union;

    const int b = blockIdx.z, h = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles_w = (W + kTW - 1) / kTW, tiles_h = (H + kTH - 1) / kTH;
    const int tw = blockIdx.x % tiles_w, th = (blockIdx.x / tiles_w) % tiles_h,
              tz = blockIdx.x / (tiles_w * tiles_h);

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        s_union = 0;
    }
    __syncthreads();
    // ---- phase 0: voxel ids + visibility of the block
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
    }
    for (int i = threadIdx.x; i < kTV * Dh / 4; i += kFwdThreads)
        reinterpret_cast<float4*>(s_acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const uint32_t cams = s_union;
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
    // ---- phase 2: cameras in ascending order
    const int corner = lane >> 3, g = lane & 7;
    uint32_t phase = 0;
    for (uint32_t rest = cams; rest; rest &= rest - 1) {
        const int c = __ffs(rest) - 1;
        __syncthreads();                 // tile is free, phase-1 results visible
        if (warp == 0)
            stage_tile_rows(tile, value + (((size_t)(b * Ncam + c) * S) * NH + h) * Dh, S, Dh,
                            (size_t)NH * Dh, Dh, &bar, lane);
        mbar_wait(&bar, phase);
        phase ^= 1;
        const float2* rp = reinterpret_cast<const float2*>(rpc) + ((size_t)c * B + b) * Nq;
        for (int v = warp; v < kTV; v += kFwdWarps) {
            if (!((s_bits[v] >> c) & 1u)) continue;
            const float2 ref = rp[s_n[v]];
            const float lx = ref.x + s_off[v * 16 + 2 * g];
            const float ly = ref.y + s_off[v * 16 + 2 * g + 1];
            const float aw = (g < NP) ? s_aw[v * 8 + g] : 0.f;
            const Tap tap = make_tap(lx, ly, aw, corner, Sh, Sw, Dh);
            float acc[CPL];
#pragma unroll
            for (int k = 0; k < CPL; ++k) acc[k] = 0.f;
            gather8<CPL>(tile, tap, lane, NP, acc);
            reduce_corners<CPL>(acc);
            if (lane < 8) {
                float4* dst = reinterpret_cast<float4*>(s_acc + v * Dh + g * CPL);
#pragma unroll
                for (int k = 0; k < CPL / 4; ++k) {
                    float4 o = dst[k];
                    o.x += acc[4 * k];
                    o.y += acc[4 * k + 1];
                    o.z += acc[4 * k + 2];
                    o.w += acc[4 * k + 3];
                    dst[k] = o;
                }
            }
        }
    }
    __syncthreads();
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
sca_bwd_kernel(const T* __restrict__ value, const float* __restrict__ logits, int ld,
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
    bwd_prologue<T, CPL>(sm, value + ((size_t)bv * S * NH + h) * Dh, S, NH, lane, warp);

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
    bwd_epilogue<T, CPL>(sm, gvalue + ((size_t)bv * S * NH + h) * Dh, S, NH, Sw, lane, warp);
}

template <typename T, int CPL>
int launch_sca_fwd(const T* value, const float* logits, int ld, const float* rpc,
                   const uint32_t* vis_bits, T* slots, int B, int Ncam, int Z, int H, int W, int Sh,
                   int Sw, int NH, int NP, cudaStream_t st) {
    const size_t smem = FwdSmem<T, CPL>::bytes(Sh * Sw);
    VER_CHECK_ARG(smem + 1024 <= (size_t)ver_device_max_smem_optin(),
                  "feature map %dx%dx%d does not fit in shared memory", Sh, Sw, CPL * 8);
    auto kern = sca_fwd_kernel<T, CPL>;
    VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = ((W + kTW - 1) / kTW) * ((H + kTH - 1) / kTH) * ((Z + kTZ - 1) / kTZ);
    dim3 grid(tiles, NH, B);
    kern<<<grid, kFwdThreads, smem, st>>>(value, logits, ld, rpc, vis_bits, slots, B, Ncam, Z, H, W, Sh,
                                          Sw, NH, NP);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

template <typename T, int CPL>
int launch_sca_bwd(const T* value, const float* logits, int ld, const float* rpc,
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
    kern<<<grid, kBwdThreads, smem, st>>>(value, logits, ld, rpc, vis_bits, counts, index, gslots,
                                          gvalue, glogits, B, Ncam, Nq, Sh, Sw, NH, NP);
    VER_CHECK_LAUNCH();
    g_ver_launches += 2;
    return VER_OK;
}

int check_sca(int dtype, int B, int Ncam, int Z, int H, int W, int Sh, int Sw, int NH, int Dh, int NP,
              int ld) {
    VER_CHECK_ARG(dtype == VER_F32 || dtype == VER_F16, "bad dtype %d", dtype);
    VER_CHECK_ARG(B > 0 && Ncam > 0 && Z > 0 && H > 0 && W > 0 && Sh > 0 && Sw > 0 && NH > 0,
                  "non-positive dimension");
    VER_CHECK_ARG(ld >= NH * NP * 3, "ld_logits %d < NH*NP*3 = %d", ld, NH * NP * 3);
    if (Ncam > 32 || NP < 1 || NP > 8 || !(Dh == 32 || Dh == 64 || Dh == 96 || Dh == 128) ||
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

extern "C" int ver_sca_forward(int dtype, const void* value, const float* logits, int ld_logits,
                               const float* rpc, const uint32_t* vis_bits, void* slots, int B,
                               int Ncam, int Z, int H, int W, int Sh, int Sw, int NH, int Dh, int NP,
                               ver_stream_t stream) {
    VER_CHECK_ARG(value && logits && rpc && vis_bits && slots, "null pointer");
    int rc = check_sca(dtype, B, Ncam, Z, H, W, Sh, Sw, NH, Dh, NP, ld_logits);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == VER_F32) {
        DISPATCH_CPL(Dh, (launch_sca_fwd<float, CPL>((const float*)value, logits, ld_logits, rpc, vis_bits,
                                                     (float*)slots, B, Ncam, Z, H, W, Sh, Sw, NH, NP, st)));
    }
    DISPATCH_CPL(Dh, (launch_sca_fwd<__half, CPL>((const __half*)value, logits, ld_logits, rpc, vis_bits,
                                                  (__half*)slots, B, Ncam, Z, H, W, Sh, Sw, NH, NP, st)));
}

extern "C" int ver_sca_backward(int dtype, const void* value, const float* logits, int ld_logits,
                                const float* rpc, const uint32_t* vis_bits, const int32_t* counts,
                                const int32_t* index, const void* grad_slots, float* grad_value,
                                float* grad_logits, int B, int Ncam, int Z, int H, int W, int Sh,
                                int Sw, int NH, int Dh, int NP, ver_stream_t stream) {
    VER_CHECK_ARG(value && logits && rpc && vis_bits && counts && index && grad_slots && grad_value &&
                      grad_logits, "null pointer");
    int rc = check_sca(dtype, B, Ncam, Z, H, W, Sh, Sw, NH, Dh, NP, ld_logits);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int Nq = Z * H * W;
    if (dtype == VER_F32) {
        DISPATCH_CPL(Dh, (launch_sca_bwd<float, CPL>((const float*)value, logits, ld_logits, rpc, vis_bits,
                                                     counts, index, (const float*)grad_slots, grad_value,
                                                     grad_logits, B, Ncam, Nq, Sh, Sw, NH, NP, st)));
    }
    DISPATCH_CPL(Dh, (launch_sca_bwd<__half, CPL>((const __half*)value, logits, ld_logits, rpc, vis_bits,
                                                  counts, index, (const __half*)grad_slots, grad_value,
                                                  grad_logits, B, Ncam, Nq, Sh, Sw, NH, NP, st)));
}

// N2 / N3 -- 3-D (voxel-volume) multi-scale deformable attention: the sm_100a replacement of
// voxel_multi_scale_deformable_attn_pytorch (M/voxel_temporal_self_attention.py:275-335), the
// sampler behind VoxelCustomMSDeformableAttention (M/voxel_decoder.py:135-337, 100 box queries
// reading the encoded voxel volume) and VoxelTemporalSelfAttention
// (M/voxel_temporal_self_attention.py:26-273, every voxel reading the volume).
//
// Shape of the work: one warp owns one (view-batch b, query q, head h).  The 8 trilinear corners of
// a sampling point are 8 runs of Dh contiguous channels in the [Bv][S][NH][Dh] volume (Dh = 96 at
// embed_dims 768: 384 B fp32 per run), so lane = channel gives fully coalesced 128-B requests and
// the tap arithmetic (trilinear.cuh) is warp-uniform.  The volume of one panorama (Z*H*W*768*4 B =
// 78.6 MB at 16x40x40) fits the 126 MB L2, so the gather is served from L2 after first touch: the
// bound is L2 request rate, not HBM.  Backward accumulates grad_value with coalesced fp32 RED
// (one 128-B atomic request per corner per 32 channels); grad_loc / grad_w need one 4-value warp
// reduction per sampling point and are written without atomics (each (b,q,h,l,p) has one owner).
#include "common.cuh"
#include "trilinear.cuh"

namespace {

constexpr int kThreads3 = 256;
constexpr int kWarps3 = kThreads3 / 32;
constexpr int kMaxLevels3 = 16;

struct LevelTable3 {
    int dhw[3 * kMaxLevels3];
    int start[kMaxLevels3];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(VER_FULL_MASK, v, o);
    return v;
}

// ------------------------------------------------------------------ forward
template <typename T, int CPW>
__global__ void __launch_bounds__(kThreads3, 2)
msda3d_fwd_kernel(const T* __restrict__ value, LevelTable3 tab, int NL, const float* __restrict__ loc,
                  const float* __restrict__ w, T* __restrict__ out, long long rows, int S, int NH, int Dh,
                  int Nq, int NP) {
    __shared__ int s_dhw[3 * kMaxLevels3];
    __shared__ int s_start[kMaxLevels3];
    if (threadIdx.x < 3 * kMaxLevels3) s_dhw[threadIdx.x] = tab.dhw[threadIdx.x];
    if (threadIdx.x < kMaxLevels3) s_start[threadIdx.x] = tab.start[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kWarps3 + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kWarps3;
    const size_t pstride = (size_t)NH * Dh;
    for (long long qh = warp0; qh < rows; qh += nwarps) {
        const int h = (int)(qh % NH);
        const long long bq = qh / NH;
        const int bv = (int)(bq / Nq);
        float acc[CPW];
#pragma unroll
        for (int j = 0; j < CPW; ++j) acc[j] = 0.f;
        for (int l = 0; l < NL; ++l) {
            const int D = s_dhw[3 * l], H = s_dhw[3 * l + 1], W = s_dhw[3 * l + 2];
            const T* vbase = value + (((size_t)bv * S + s_start[l]) * NH + h) * Dh;
            for (int p = 0; p < NP; ++p) {
                const size_t li = ((size_t)qh * NL + l) * NP + p;
                const Tap3 tap = make_tap3(loc[3 * li], loc[3 * li + 1], loc[3 * li + 2], D, H, W);
                if (!tap.any) continue;
                const float aw = w[li];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (tap.off[k] < 0) continue;
                    const T* src = vbase + (size_t)tap.off[k] * pstride;
                    const float cw = aw * tap.wgt[k];
#pragma unroll
                    for (int j = 0; j < CPW; ++j) {
                        const int c = lane + 32 * j;
                        if (c < Dh) acc[j] = fmaf(cw, to_f32(src[c]), acc[j]);
                    }
                }
            }
        }
        T* dst = out + (size_t)qh * Dh;
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
            const int c = lane + 32 * j;
            if (c < Dh) from_f32(dst[c], acc[j]);
        }
    }
}

// ------------------------------------------------------------------ backward
template <typename T, int CPW>
__global__ void __launch_bounds__(kThreads3, 2)
msda3d_bwd_kernel(const T* __restrict__ value, LevelTable3 tab, int NL, const float* __restrict__ loc,
                  const float* __restrict__ w, const T* __restrict__ gout, float* __restrict__ gvalue,
                  float* __restrict__ gloc, float* __restrict__ gw, long long rows, int S, int NH, int Dh,
                  int Nq, int NP) {
    __shared__ int s_dhw[3 * kMaxLevels3];
    __shared__ int s_start[kMaxLevels3];
    if (threadIdx.x < 3 * kMaxLevels3) s_dhw[threadIdx.x] = tab.dhw[threadIdx.x];
    if (threadIdx.x < kMaxLevels3) s_start[threadIdx.x] = tab.start[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kWarps3 + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kWarps3;
    const size_t pstride = (size_t)NH * Dh;
    for (long long qh = warp0; qh < rows; qh += nwarps) {
        const int h = (int)(qh % NH);
        const long long bq = qh / NH;
        const int bv = (int)(bq / Nq);
        float go[CPW];
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
            const int c = lane + 32 * j;
            go[j] = c < Dh ? to_f32(gout[(size_t)qh * Dh + c]) : 0.f;
        }
        for (int l = 0; l < NL; ++l) {
            const int D = s_dhw[3 * l], H = s_dhw[3 * l + 1], W = s_dhw[3 * l + 2];
            const size_t base = (((size_t)bv * S + s_start[l]) * NH + h) * Dh;
            for (int p = 0; p < NP; ++p) {
                const size_t li = ((size_t)qh * NL + l) * NP + p;
                const Tap3 tap = make_tap3(loc[3 * li], loc[3 * li + 1], loc[3 * li + 2], D, H, W);
                const float aw = w[li];
                float pw = 0.f, px = 0.f, py = 0.f, pz = 0.f;   // per-lane partial dots
                if (tap.any) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (tap.off[k] < 0) continue;
                        const size_t o = base + (size_t)tap.off[k] * pstride;
                        const float cw = aw * tap.wgt[k];
                        float part = 0.f;
#pragma unroll
                        for (int j = 0; j < CPW; ++j) {
                            const int c = lane + 32 * j;
                            if (c < Dh) {
                                part = fmaf(go[j], to_f32(value[o + c]), part);
                                atomicAdd(gvalue + o + c, cw * go[j]);
                            }
                        }
                        pw = fmaf(tap.wgt[k], part, pw);
                        px = fmaf(tap.gx[k], part, px);
                        py = fmaf(tap.gy[k], part, py);
                        pz = fmaf(tap.gz[k], part, pz);
                    }
                    pw = warp_sum(pw);
                    px = warp_sum(px);
                    py = warp_sum(py);
                    pz = warp_sum(pz);
                }
                if (lane == 0) {
                    gw[li] = pw;
                    gloc[3 * li] = aw * (float)W * px;
                    gloc[3 * li + 1] = aw * (float)H * py;
                    gloc[3 * li + 2] = aw * (float)D * pz;
                }
            }
        }
    }
}

int fill_table3(LevelTable3& tab, const int32_t* shapes_dhw, int NL) {
    int start = 0;
    memset(&tab, 0, sizeof(tab));
    for (int l = 0; l < NL; ++l) {
        for (int a = 0; a < 3; ++a) tab.dhw[3 * l + a] = shapes_dhw[3 * l + a];
        tab.start[l] = start;
        start += shapes_dhw[3 * l] * shapes_dhw[3 * l + 1] * shapes_dhw[3 * l + 2];
    }
    return start;
}

int grid_for(long long rows) {
    const long long want = (rows + kWarps3 - 1) / kWarps3;
    const long long cap = (long long)ver_device_sm_count() * 8;   // 8 resident CTAs of 256 threads per SM
    return (int)(want < cap ? want : cap);
}

template <typename T, int CPW>
int launch_fwd3(const T* value, const LevelTable3& tab, int NL, const float* loc, const float* w, T* out,
                int Bv, int S, int NH, int Dh, int Nq, int NP, cudaStream_t st) {
    const long long rows = (long long)Bv * Nq * NH;
    msda3d_fwd_kernel<T, CPW><<<grid_for(rows), kThreads3, 0, st>>>(value, tab, NL, loc, w, out, rows, S, NH,
                                                                  Dh, Nq, NP);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

template <typename T, int CPW>
int launch_bwd3(const T* value, const LevelTable3& tab, int NL, const float* loc, const float* w,
                const T* gout, float* gvalue, float* gloc, float* gw, int Bv, int S, int NH, int Dh, int Nq,
                int NP, cudaStream_t st) {
    const long long rows = (long long)Bv * Nq * NH;
    VER_CHECK_CUDA(cudaMemsetAsync(gvalue, 0, (size_t)Bv * S * NH * Dh * sizeof(float), st));
    msda3d_bwd_kernel<T, CPW><<<grid_for(rows), kThreads3, 0, st>>>(value, tab, NL, loc, w, gout, gvalue, gloc,
                                                                  gw, rows, S, NH, Dh, Nq, NP);
    VER_CHECK_LAUNCH();
    g_ver_launches += 2;
    return VER_OK;
}

template <typename T>
int msda3d_forward_t(const T* value, const LevelTable3& tab, int NL, const float* loc, const float* w, T* out,
                     int Bv, int S, int NH, int Dh, int Nq, int NP, cudaStream_t st) {
    switch ((Dh + 31) / 32) {
        case 1: return launch_fwd3<T, 1>(value, tab, NL, loc, w, out, Bv, S, NH, Dh, Nq, NP, st);
        case 2: return launch_fwd3<T, 2>(value, tab, NL, loc, w, out, Bv, S, NH, Dh, Nq, NP, st);
        case 3: return launch_fwd3<T, 3>(value, tab, NL, loc, w, out, Bv, S, NH, Dh, Nq, NP, st);
        case 4: return launch_fwd3<T, 4>(value, tab, NL, loc, w, out, Bv, S, NH, Dh, Nq, NP, st);
        default: return launch_fwd3<T, 8>(value, tab, NL, loc, w, out, Bv, S, NH, Dh, Nq, NP, st);
    }
}

template <typename T>
int msda3d_backward_t(const T* value, const LevelTable3& tab, int NL, const float* loc, const float* w,
                      const T* gout, float* gvalue, float* gloc, float* gw, int Bv, int S, int NH, int Dh,
                      int Nq, int NP, cudaStream_t st) {
    switch ((Dh + 31) / 32) {
        case 1: return launch_bwd3<T, 1>(value, tab, NL, loc, w, gout, gvalue, gloc, gw, Bv, S, NH, Dh, Nq, NP, st);
        case 2: return launch_bwd3<T, 2>(value, tab, NL, loc, w, gout, gvalue, gloc, gw, Bv, S, NH, Dh, Nq, NP, st);
        case 3: return launch_bwd3<T, 3>(value, tab, NL, loc, w, gout, gvalue, gloc, gw, Bv, S, NH, Dh, Nq, NP, st);
        case 4: return launch_bwd3<T, 4>(value, tab, NL, loc, w, gout, gvalue, gloc, gw, Bv, S, NH, Dh, Nq, NP, st);
        default: return launch_bwd3<T, 8>(value, tab, NL, loc, w, gout, gvalue, gloc, gw, Bv, S, NH, Dh, Nq, NP, st);
    }
}

int check_msda3d_args(int dtype, const void* value, const int32_t* shapes_dhw, int NL, const float* loc,
                      const float* w, int Bv, int S, int NH, int Dh, int Nq, int NP) {
    VER_CHECK_ARG(dtype == VER_F32 || dtype == VER_F16, "bad dtype %d", dtype);
    VER_CHECK_ARG(value && shapes_dhw && loc && w, "null pointer");
    VER_CHECK_ARG(Bv > 0 && S > 0 && NH > 0 && Dh > 0 && Nq > 0 && NP > 0, "non-positive dimension");
    VER_CHECK_ARG(Dh <= 256, "channels per head %d > 256", Dh);
    VER_CHECK_ARG(NL >= 1 && NL <= kMaxLevels3, "num_levels %d out of range [1,%d]", NL, kMaxLevels3);
    long long tot = 0;
    for (int l = 0; l < NL; ++l) {
        VER_CHECK_ARG(shapes_dhw[3 * l] > 0 && shapes_dhw[3 * l + 1] > 0 && shapes_dhw[3 * l + 2] > 0,
                      "bad level shape");
        tot += (long long)shapes_dhw[3 * l] * shapes_dhw[3 * l + 1] * shapes_dhw[3 * l + 2];
    }
    // mirrors `assert (spatial_shapes[:,0]*spatial_shapes[:,1]*spatial_shapes[:,2]).sum() == num_value`
    // (M/voxel_decoder.py:283, M/voxel_temporal_self_attention.py:194)
    VER_CHECK_ARG(tot == S, "sum(d*h*w)=%lld != num_value=%d", tot, S);
    return VER_OK;
}

}  // namespace

extern "C" int ver_msda3d_forward(int dtype, const void* value, const int32_t* shapes_dhw, int NL,
                                  const float* loc, const float* w, void* out, int Bv, int S, int NH, int Dh,
                                  int Nq, int NP, ver_stream_t stream) {
    int rc = check_msda3d_args(dtype, value, shapes_dhw, NL, loc, w, Bv, S, NH, Dh, Nq, NP);
    if (rc) return rc;
    VER_CHECK_ARG(out, "null pointer");
    LevelTable3 tab;
    fill_table3(tab, shapes_dhw, NL);
    if (dtype == VER_F32)
        return msda3d_forward_t<float>((const float*)value, tab, NL, loc, w, (float*)out, Bv, S, NH, Dh, Nq, NP,
                                       (cudaStream_t)stream);
    return msda3d_forward_t<__half>((const __half*)value, tab, NL, loc, w, (__half*)out, Bv, S, NH, Dh, Nq, NP,
                                    (cudaStream_t)stream);
}

extern "C" int ver_msda3d_backward(int dtype, const void* value, const int32_t* shapes_dhw, int NL,
                                   const float* loc, const float* w, const void* grad_out, float* grad_value,
                                   float* grad_loc, float* grad_w, int Bv, int S, int NH, int Dh, int Nq,
                                   int NP, ver_stream_t stream) {
    int rc = check_msda3d_args(dtype, value, shapes_dhw, NL, loc, w, Bv, S, NH, Dh, Nq, NP);
    if (rc) return rc;
    VER_CHECK_ARG(grad_out && grad_value && grad_loc && grad_w, "null pointer");
    LevelTable3 tab;
    fill_table3(tab, shapes_dhw, NL);
    if (dtype == VER_F32)
        return msda3d_backward_t<float>((const float*)value, tab, NL, loc, w, (const float*)grad_out, grad_value,
                                        grad_loc, grad_w, Bv, S, NH, Dh, Nq, NP, (cudaStream_t)stream);
    return msda3d_backward_t<__half>((const __half*)value, tab, NL, loc, w, (const __half*)grad_out,
                                     grad_value, grad_loc, grad_w, Bv, S, NH, Dh, Nq, NP,
                                     (cudaStream_t)stream);
}

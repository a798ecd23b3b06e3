// A5 -- operator-level multi-scale deformable attention (the mmcv._ext
// ms_deform_attn_forward/backward boundary, M/multi_scale_deformable_attn_function.py:118-160).
//
// Fast path (num_levels == 1, NP <= 8, Dh in {32,64,96,128}): one CTA stages the
// [S][Dh] feature map of one (view, head) in shared memory with bulk async copies
// (TMA engine) and its 8 warps stream over a chunk of queries; see sampler.cuh for
// the lane mapping.  Everything else takes the generic global-memory kernels.
#include "sampler.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ------------------------------------------------------------------ staged forward
template <typename T, int CPL>
__global__ void __launch_bounds__(kThreads)
msda_fwd_staged(const T* __restrict__ value, const float* __restrict__ loc,
                const float* __restrict__ w, T* __restrict__ out, int S, int Sh, int Sw, int NH,
                int Nq, int NP, int q_per_cta) {
    constexpr int Dh = CPL * 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    __shared__ __align__(8) uint64_t bar;

    const int bv = blockIdx.z, h = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == 0)
        stage_tile_rows(tile, value + ((size_t)bv * S * NH + h) * Dh, S, Dh, (size_t)NH * Dh, Dh,
                        &bar, lane);
    mbar_wait(&bar, 0);

    const int corner = lane >> 3, g = lane & 7;
    const int q0 = blockIdx.x * q_per_cta;
    const int q1 = min(q0 + q_per_cta, Nq);
    for (int q = q0 + warp; q < q1; q += kWarps) {
        const size_t qh = ((size_t)bv * Nq + q) * NH + h;
        float lx = 0.f, ly = 0.f, aw = 0.f;
        if (g < NP) {
            const float2 l = reinterpret_cast<const float2*>(loc)[qh * NP + g];
            lx = l.x;
            ly = l.y;
            aw = w[qh * NP + g];
        }
        const Tap tap = make_tap(lx, ly, aw, corner, Sh, Sw, Dh);
        float acc[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) acc[k] = 0.f;
        gather8<CPL>(tile, tap, lane, NP, acc);
        reduce_corners<CPL>(acc);
        if (lane < 8) store_channels<CPL>(out + qh * Dh + g * CPL, acc);
    }
}

}  // namespace

#include "sca_bwd.cuh"

namespace {

template <typename T, int CPL>
__global__ void __launch_bounds__(kBwdThreads, 2)
msda_bwd_staged(const T* __restrict__ value, const float* __restrict__ loc,
                const float* __restrict__ w, const T* __restrict__ gout,
                float* __restrict__ gvalue, float* __restrict__ gloc, float* __restrict__ gw, int S,
                int Sh, int Sw, int NH, int Nq, int NP) {
    constexpr int Dh = CPL * 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BwdSmem<T, CPL> sm(smem_raw, S);
    const int bv = blockIdx.y, h = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bwd_prologue<T, CPL>(sm, value + ((size_t)bv * S * NH + h) * Dh, S, (size_t)NH * Dh, lane, warp);

    const int corner = lane >> 3, g = lane & 7;
    for (int q = 0; q < Nq; ++q) {
        const size_t qh = ((size_t)bv * Nq + q) * NH + h;
        float lx = 0.f, ly = 0.f, aw = 0.f;
        if (g < NP) {
            const float2 l = reinterpret_cast<const float2*>(loc)[qh * NP + g];
            lx = l.x;
            ly = l.y;
            aw = w[qh * NP + g];
        }
        const TapB tap = make_tap_bwd(lx, ly, aw, corner, Sh, Sw);
        float3 gr = bwd_process_item<T, CPL>(sm, tap, gout + qh * Dh, 1.f, Sh, Sw, NP, lane, warp, q);
        // gr = (d/d aw, d/d loc.x, d/d loc.y) of tap point p = g, valid in the finalizer warp
        if (warp == (q & (kBwdWarps - 1)) && lane < 8 && g < NP) {
            gw[qh * NP + g] = gr.x;
            reinterpret_cast<float2*>(gloc)[qh * NP + g] = make_float2(gr.y, gr.z);
        }
    }
    bwd_epilogue<T, CPL>(sm, gvalue + ((size_t)bv * S * NH + h) * Dh, S, (size_t)NH * Dh, Sw, lane, warp);
}

struct LevelTable {
    int hw[32];
    int start[16];
};

// ------------------------------------------------------------------ generic kernels
// any number of levels / points / channels; the level table travels as a kernel parameter
template <typename T>
__global__ void msda_fwd_generic_tab(const T* value, LevelTable tab, int NL, const float* loc,
                                     const float* w, T* out, int Bv, int S, int NH, int Dh, int Nq,
                                     int NP) {
    __shared__ int s_hw[32];
    __shared__ int s_start[16];
    if (threadIdx.x < 32) s_hw[threadIdx.x] = tab.hw[threadIdx.x];
    if (threadIdx.x < 16) s_start[threadIdx.x] = tab.start[threadIdx.x];
    __syncthreads();
    // inline body (same as msda_fwd_generic, reading the table from shared memory)
    const size_t total = (size_t)Bv * Nq * NH * Dh;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int c = i % Dh;
        const int h = (i / Dh) % NH;
        const size_t bq = i / ((size_t)Dh * NH);
        const int bv = bq / Nq;
        const size_t qh = bq * NH + h;
        const size_t pstride = (size_t)NH * Dh;
        float acc = 0.f;
        for (int l = 0; l < NL; ++l) {
            const int Hh = s_hw[2 * l], Ww = s_hw[2 * l + 1];
            const T* vbase = value + (((size_t)bv * S + s_start[l]) * NH + h) * Dh + c;
            for (int p = 0; p < NP; ++p) {
                const size_t li = (qh * NL + l) * NP + p;
                const float x = loc[2 * li] * Ww - 0.5f, y = loc[2 * li + 1] * Hh - 0.5f;
                if (!(x > -1.f && y > -1.f && x < (float)Ww && y < (float)Hh)) continue;
                const float xf = floorf(x), yf = floorf(y);
                const float fx = x - xf, fy = y - yf;
                const int x0 = (int)xf, y0 = (int)yf;
                float s = 0.f;
                if (y0 >= 0) {
                    if (x0 >= 0) s += (1.f - fy) * (1.f - fx) * to_f32(vbase[((size_t)y0 * Ww + x0) * pstride]);
                    if (x0 + 1 < Ww) s += (1.f - fy) * fx * to_f32(vbase[((size_t)y0 * Ww + x0 + 1) * pstride]);
                }
                if (y0 + 1 < Hh) {
                    if (x0 >= 0) s += fy * (1.f - fx) * to_f32(vbase[((size_t)(y0 + 1) * Ww + x0) * pstride]);
                    if (x0 + 1 < Ww) s += fy * fx * to_f32(vbase[((size_t)(y0 + 1) * Ww + x0 + 1) * pstride]);
                }
                acc += w[li] * s;
            }
        }
        from_f32(out[i], acc);
    }
}

template <typename T>
__global__ void msda_bwd_generic_tab(const T* value, LevelTable tab, int NL, const float* loc,
                                     const float* w, const T* gout, float* gvalue, float* gloc,
                                     float* gw, int Bv, int S, int NH, int Dh, int Nq, int NP) {
    __shared__ int s_hw[32];
    __shared__ int s_start[16];
    if (threadIdx.x < 32) s_hw[threadIdx.x] = tab.hw[threadIdx.x];
    if (threadIdx.x < 16) s_start[threadIdx.x] = tab.start[threadIdx.x];
    __syncthreads();
    const size_t total = (size_t)Bv * Nq * NH * NL * NP;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int l = (i / NP) % NL;
        const size_t qh = i / ((size_t)NP * NL);
        const int h = qh % NH;
        const size_t bq = qh / NH;
        const int bv = bq / Nq;
        const int Hh = s_hw[2 * l], Ww = s_hw[2 * l + 1];
        const float x = loc[2 * i] * Ww - 0.5f, y = loc[2 * i + 1] * Hh - 0.5f;
        float g_w = 0.f, g_x = 0.f, g_y = 0.f;
        if (x > -1.f && y > -1.f && x < (float)Ww && y < (float)Hh) {
            const float xf = floorf(x), yf = floorf(y);
            const float fx = x - xf, fy = y - yf;
            const int x0 = (int)xf, y0 = (int)yf;
            const float aw = w[i];
            const T* go = gout + qh * Dh;
            const size_t base = (((size_t)bv * S + s_start[l]) * NH + h) * Dh;
            const size_t pstride = (size_t)NH * Dh;
            for (int cy = 0; cy < 2; ++cy) {
                const int yi = y0 + cy;
                if (yi < 0 || yi >= Hh) continue;
                const float wy = cy ? fy : 1.f - fy;
                for (int cx = 0; cx < 2; ++cx) {
                    const int xi = x0 + cx;
                    if (xi < 0 || xi >= Ww) continue;
                    const float wx = cx ? fx : 1.f - fx;
                    const size_t o = base + ((size_t)yi * Ww + xi) * pstride;
                    float dot = 0.f;
                    for (int c = 0; c < Dh; ++c) {
                        const float gg = to_f32(go[c]);
                        dot += gg * to_f32(value[o + c]);
                        atomicAdd(gvalue + o + c, aw * wy * wx * gg);
                    }
                    g_w += wy * wx * dot;
                    g_x += (cx ? 1.f : -1.f) * wy * dot;
                    g_y += (cy ? 1.f : -1.f) * wx * dot;
                }
            }
            g_x *= aw * Ww;
            g_y *= aw * Hh;
        }
        gw[i] = g_w;
        gloc[2 * i] = g_x;
        gloc[2 * i + 1] = g_y;
    }
}

int fill_table(LevelTable& tab, const int32_t* shapes_hw, int NL, int S) {
    int start = 0;
    for (int l = 0; l < NL; ++l) {
        tab.hw[2 * l] = shapes_hw[2 * l];
        tab.hw[2 * l + 1] = shapes_hw[2 * l + 1];
        tab.start[l] = start;
        start += shapes_hw[2 * l] * shapes_hw[2 * l + 1];
    }
    return start == S;
}

template <typename T, int CPL>
int launch_fwd_staged(const T* value, int Sh, int Sw, const float* loc, const float* w, T* out, int Bv,
                      int NH, int Nq, int NP, cudaStream_t st) {
    const int S = Sh * Sw;
    const size_t smem = (size_t)S * CPL * 8 * sizeof(T);
    auto kern = msda_fwd_staged<T, CPL>;
    VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int q_per_cta = 256;
    dim3 grid((Nq + q_per_cta - 1) / q_per_cta, NH, Bv);
    kern<<<grid, kThreads, smem, st>>>(value, loc, w, out, S, Sh, Sw, NH, Nq, NP, q_per_cta);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

template <typename T, int CPL>
int launch_bwd_staged(const T* value, int Sh, int Sw, const float* loc, const float* w, const T* gout,
                      float* gvalue, float* gloc, float* gw, int Bv, int NH, int Nq, int NP,
                      cudaStream_t st) {
    const int S = Sh * Sw;
    const size_t smem = BwdSmem<T, CPL>::bytes(S);
    auto kern = msda_bwd_staged<T, CPL>;
    VER_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(NH, Bv);
    kern<<<grid, kBwdThreads, smem, st>>>(value, loc, w, gout, gvalue, gloc, gw, S, Sh, Sw, NH, Nq, NP);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

bool staged_ok(int NL, int NP, int Dh, int S, size_t esize, size_t extra_per_elem) {
    if (NL != 1 || NP > 8 || NP < 1) return false;
    if (!(Dh == 32 || Dh == 64 || Dh == 96 || Dh == 128)) return false;
    const size_t need = (size_t)S * Dh * (esize + extra_per_elem) + 4096;
    return need <= (size_t)ver_device_max_smem_optin();
}

}  // namespace

template <typename T>
static int msda_forward_t(const T* value, const int32_t* shapes_hw, int NL, const float* loc,
                          const float* w, T* out, int Bv, int S, int NH, int Dh, int Nq, int NP,
                          cudaStream_t st) {
    if (staged_ok(NL, NP, Dh, S, sizeof(T), 0)) {
        const int Sh = shapes_hw[0], Sw = shapes_hw[1];
        switch (Dh) {
            case 32: return launch_fwd_staged<T, 4>(value, Sh, Sw, loc, w, out, Bv, NH, Nq, NP, st);
            case 64: return launch_fwd_staged<T, 8>(value, Sh, Sw, loc, w, out, Bv, NH, Nq, NP, st);
            case 96: return launch_fwd_staged<T, 12>(value, Sh, Sw, loc, w, out, Bv, NH, Nq, NP, st);
            case 128: return launch_fwd_staged<T, 16>(value, Sh, Sw, loc, w, out, Bv, NH, Nq, NP, st);
        }
    }
    LevelTable tab;
    fill_table(tab, shapes_hw, NL, S);
    const size_t total = (size_t)Bv * Nq * NH * Dh;
    const int blocks = (int)((total + 255) / 256 > 1048576 ? 1048576 : (total + 255) / 256);
    msda_fwd_generic_tab<T><<<blocks, 256, 0, st>>>(value, tab, NL, loc, w, out, Bv, S, NH, Dh, Nq, NP);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

template <typename T>
static int msda_backward_t(const T* value, const int32_t* shapes_hw, int NL, const float* loc,
                           const float* w, const T* gout, float* gvalue, float* gloc, float* gw,
                           int Bv, int S, int NH, int Dh, int Nq, int NP, cudaStream_t st) {
    if (staged_ok(NL, NP, Dh, S, sizeof(T), sizeof(float))) {
        const int Sh = shapes_hw[0], Sw = shapes_hw[1];
        switch (Dh) {
            case 32: return launch_bwd_staged<T, 4>(value, Sh, Sw, loc, w, gout, gvalue, gloc, gw, Bv, NH, Nq, NP, st);
            case 64: return launch_bwd_staged<T, 8>(value, Sh, Sw, loc, w, gout, gvalue, gloc, gw, Bv, NH, Nq, NP, st);
            case 96: return launch_bwd_staged<T, 12>(value, Sh, Sw, loc, w, gout, gvalue, gloc, gw, Bv, NH, Nq, NP, st);
            case 128: return launch_bwd_staged<T, 16>(value, Sh, Sw, loc, w, gout, gvalue, gloc, gw, Bv, NH, Nq, NP, st);
        }
    }
    LevelTable tab;
    fill_table(tab, shapes_hw, NL, S);
    VER_CHECK_CUDA(cudaMemsetAsync(gvalue, 0, (size_t)Bv * S * NH * Dh * sizeof(float), st));
    const size_t total = (size_t)Bv * Nq * NH * NL * NP;
    const int blocks = (int)((total + 127) / 128 > 1048576 ? 1048576 : (total + 127) / 128);
    msda_bwd_generic_tab<T><<<blocks, 128, 0, st>>>(value, tab, NL, loc, w, gout, gvalue, gloc, gw, Bv,
                                                   S, NH, Dh, Nq, NP);
    VER_CHECK_LAUNCH();
    g_ver_launches += 2;
    return VER_OK;
}

static int check_msda_args(int dtype, const void* value, const int32_t* shapes_hw, int NL,
                           const float* loc, const float* w, int Bv, int S, int NH, int Dh, int Nq,
                           int NP) {
    VER_CHECK_ARG(dtype == VER_F32 || dtype == VER_F16, "bad dtype %d", dtype);
    VER_CHECK_ARG(value && shapes_hw && loc && w, "null pointer");
    VER_CHECK_ARG(Bv > 0 && S > 0 && NH > 0 && Dh > 0 && Nq > 0 && NP > 0, "non-positive dimension");
    VER_CHECK_ARG(NL >= 1 && NL <= 16, "num_levels %d out of range [1,16]", NL);
    int tot = 0;
    for (int l = 0; l < NL; ++l) {
        VER_CHECK_ARG(shapes_hw[2 * l] > 0 && shapes_hw[2 * l + 1] > 0, "bad level shape");
        tot += shapes_hw[2 * l] * shapes_hw[2 * l + 1];
    }
    // mirrors `assert (spatial_shapes[:,0]*spatial_shapes[:,1]).sum() == num_value`
    // (M/spatial_cross_attention.py:334)
    VER_CHECK_ARG(tot == S, "sum(h*w)=%d != num_value=%d", tot, S);
    return VER_OK;
}

extern "C" int ver_msda_forward(int dtype, const void* value, const int32_t* shapes_hw, int NL,
                                const float* loc, const float* w, void* out, int Bv, int S, int NH,
                                int Dh, int Nq, int NP, ver_stream_t stream) {
    int rc = check_msda_args(dtype, value, shapes_hw, NL, loc, w, Bv, S, NH, Dh, Nq, NP);
    if (rc) return rc;
    VER_CHECK_ARG(out, "null pointer");
    if (dtype == VER_F32)
        return msda_forward_t<float>((const float*)value, shapes_hw, NL, loc, w, (float*)out, Bv, S,
                                     NH, Dh, Nq, NP, (cudaStream_t)stream);
    return msda_forward_t<__half>((const __half*)value, shapes_hw, NL, loc, w, (__half*)out, Bv, S, NH,
                                  Dh, Nq, NP, (cudaStream_t)stream);
}

extern "C" int ver_msda_backward(int dtype, const void* value, const int32_t* shapes_hw, int NL,
                                 const float* loc, const float* w, const void* grad_out,
                                 float* grad_value, float* grad_loc, float* grad_w, int Bv, int S,
                                 int NH, int Dh, int Nq, int NP, ver_stream_t stream) {
    int rc = check_msda_args(dtype, value, shapes_hw, NL, loc, w, Bv, S, NH, Dh, Nq, NP);
    if (rc) return rc;
    VER_CHECK_ARG(grad_out && grad_value && grad_loc && grad_w, "null pointer");
    if (dtype == VER_F32)
        return msda_backward_t<float>((const float*)value, shapes_hw, NL, loc, w,
                                      (const float*)grad_out, grad_value, grad_loc, grad_w, Bv, S, NH,
                                      Dh, Nq, NP, (cudaStream_t)stream);
    return msda_backward_t<__half>((const __half*)value, shapes_hw, NL, loc, w,
                                   (const __half*)grad_out, grad_value, grad_loc, grad_w, Bv, S, NH,
                                   Dh, Nq, NP, (cudaStream_t)stream);
}

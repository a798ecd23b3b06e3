// HBM-bound elementwise / row-wise kernels around the sampler:
//   A8 prologue (camera + level embedding add and view reorder), A6 epilogue
//   (residual + LayerNorm), A11 (sigmoid focal loss fwd+bwd with on-the-fly dense
//   target), A12 (occupancy decode: argmax + ordered compaction).
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ A8 prologue
// feats [Ncam, B, S, C] fp32 -> out [B*Ncam, S, C]; 4 channels per thread
template <typename T>
__global__ void feat_embed_kernel(const float4* __restrict__ feats, const float4* __restrict__ cams,
                                  const float4* __restrict__ level, T* __restrict__ out, int Ncam,
                                  int B, int S, int C4) {
    const size_t total = (size_t)Ncam * B * S * C4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = i % C4;
        const size_t r = i / C4;           // (b*Ncam + cam)*S + s
        const int s = r % S;
        const size_t bc = r / S;
        const int cam = bc % Ncam, b = bc / Ncam;
        float4 v = feats[(((size_t)cam * B + b) * S + s) * C4 + c4];
        // feat + cams_embeds, then + level_embeds[0]  (voxel_transformer.py:154-157)
        if (cams) {
            const float4 e = cams[(size_t)cam * C4 + c4];
            v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
        }
        const float4 l = level[c4];
        v.x += l.x; v.y += l.y; v.z += l.z; v.w += l.w;
        T* o = out + i * 4;
        from_f32(o[0], v.x); from_f32(o[1], v.y); from_f32(o[2], v.z); from_f32(o[3], v.w);
    }
}

// ------------------------------------------------------------------ A6 epilogue
// one warp per row; two-pass mean/variance in registers (C <= 32*4*MAXV)
template <typename T, int MAXV>
__global__ void __launch_bounds__(256)
add_layernorm_kernel(const T* __restrict__ x, const T* __restrict__ res,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     T* __restrict__ y, int64_t rows, int C, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T* xr = x + row * C;
    const T* rr = res ? res + row * C : nullptr;
    float v[MAXV * 4];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int c = (k * 32 + lane) * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float t = 0.f;
            if (c + e < C) {
                t = to_f32(xr[c + e]);
                if (rr) t += to_f32(rr[c + e]);
            }
            v[4 * k + e] = t;
            sum += t;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(VER_FULL_MASK, sum, o);
    const float mean = sum / (float)C;
    float var = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int c = (k * 32 + lane) * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (c + e < C) {
                const float d = v[4 * k + e] - mean;
                var += d * d;
            }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) var += __shfl_xor_sync(VER_FULL_MASK, var, o);
    const float rstd = rsqrtf(var / (float)C + eps);
    T* yr = y + row * C;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int c = (k * 32 + lane) * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (c + e < C) from_f32(yr[c + e], (v[4 * k + e] - mean) * rstd * gamma[c + e] + beta[c + e]);
    }
}

// ------------------------------------------------------------------ A11
__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        p[i] = v;
}
// gt[0][occ_gt[:,0]] = occ_gt[:,1]   (HEAD:1330 / :1409)
__global__ void scatter_gt_kernel(const int64_t* __restrict__ occ_gt, int n_gt,
                                  int32_t* __restrict__ dense, int64_t N, int Ccls,
                                  int32_t* __restrict__ num_pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_gt) return;
    const int64_t idx = occ_gt[2 * i];
    const int cls = (int)occ_gt[2 * i + 1];
    if (idx >= 0 && idx < N) dense[idx] = cls;
}
__global__ void count_pos_kernel(const int32_t* __restrict__ dense, int64_t N, int Ccls,
                                 int32_t* __restrict__ num_pos) {
    int local = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N;
         i += (int64_t)gridDim.x * blockDim.x)
        local += dense[i] < Ccls;
#pragma unroll
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(VER_FULL_MASK, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(num_pos, local);
}
// element loss (mmdet py_sigmoid_focal_loss): t in {0,1}
//   p = sigmoid(x); pt = t ? 1-p : p; fw = (t ? alpha : 1-alpha) * pt^gamma
//   loss = fw * bce(x, t),  bce = max(x,0) - x t + log1p(exp(-|x|))
__global__ void __launch_bounds__(256)
focal_kernel(const float* __restrict__ logits, const int32_t* __restrict__ dense,
             float* __restrict__ loss_sum, float* __restrict__ grad, int64_t N, int Ccls,
             float gamma, float alpha) {
    float local = 0.f;
    const int64_t total = N * Ccls;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i / Ccls;
        const int c = (int)(i % Ccls);
        const float x = logits[i];
        const bool t = dense[n] == c;
        const float p = 1.f / (1.f + expf(-x));
        const float pt = t ? 1.f - p : p;
        const float a = t ? alpha : 1.f - alpha;
        const float bce = fmaxf(x, 0.f) - (t ? x : 0.f) + log1pf(expf(-fabsf(x)));
        const float ptg = powf(pt, gamma);
        local += a * ptg * bce;
        if (grad) {
            // d/dx [a pt^g bce]: dpt/dx = +-p(1-p) (minus for t=1); dbce/dx = p - t
            const float dpt = t ? -p * (1.f - p) : p * (1.f - p);
            const float dptg = (pt > 0.f) ? gamma * powf(pt, gamma - 1.f) * dpt : 0.f;
            grad[i] = a * (dptg * bce + ptg * (p - (t ? 1.f : 0.f)));
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(VER_FULL_MASK, local, o);
    __shared__ float s_part[8];
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += s_part[i];
        atomicAdd(loss_sum, t);
    }
}

// ------------------------------------------------------------------ A12
// class of a row = argmax over [sigmoid(logit_0..C-1), thr] (first max wins, like torch)
__device__ __forceinline__ int decode_row(const float* __restrict__ row, int Ccls, float thr) {
    int best = 0;
    float bv = 1.f / (1.f + expf(-row[0]));
    for (int c = 1; c < Ccls; ++c) {
        const float p = 1.f / (1.f + expf(-row[c]));
        if (p > bv) { bv = p; best = c; }
    }
    if (thr > bv) best = Ccls;
    return best;
}
__global__ void __launch_bounds__(1024)
decode_count_kernel(const float* __restrict__ logits, int64_t N, int Ccls, float thr,
                    int32_t* __restrict__ block_counts) {
    const int64_t n = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const bool occ = n < N && decode_row(logits + n * Ccls, Ccls, thr) < Ccls;
    const int c = __syncthreads_count(occ);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}
__global__ void __launch_bounds__(1024)
decode_scan_kernel(int32_t* __restrict__ block_counts, int nblocks, int32_t* __restrict__ out_count) {
    // single CTA exclusive scan of the per-block counts (in place)
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int start = 0; start < nblocks; start += 1024) {
        const int i = start + threadIdx.x;
        const int v = i < nblocks ? block_counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(VER_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        int wpre = 0;
        for (int k = 0; k < wid; ++k) wpre += s_warp[k];
        const int excl = s_carry + wpre + incl - v;
        if (i < nblocks) block_counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_count = s_carry;
}
__global__ void __launch_bounds__(1024)
decode_emit_kernel(const float* __restrict__ logits, int64_t N, int Ccls, float thr,
                   const int32_t* __restrict__ block_offsets, int64_t* __restrict__ out_pairs) {
    __shared__ int s_warp[32];
    const int64_t n = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int cls = n < N ? decode_row(logits + n * Ccls, Ccls, thr) : Ccls;
    const bool occ = cls < Ccls;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(VER_FULL_MASK, occ);
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    int pre = block_offsets[blockIdx.x];
    for (int k = 0; k < wid; ++k) pre += s_warp[k];
    if (occ) {
        const int64_t o = pre + __popc(bal & ((1u << lane) - 1));
        out_pairs[2 * o] = n;
        out_pairs[2 * o + 1] = cls;
    }
}

int grid_for(size_t total, int threads, int cap = 148 * 16) {
    size_t b = (total + threads - 1) / threads;
    return (int)(b < (size_t)cap ? (b ? b : 1) : cap);
}

}  // namespace

extern "C" int ver_feat_embed(int dtype, const float* feats, const float* cams_embeds,
                              const float* level_embed, void* out, int Ncam, int B, int S, int C,
                              ver_stream_t stream) {
    VER_CHECK_ARG(dtype == VER_F32 || dtype == VER_F16, "bad dtype %d", dtype);
    VER_CHECK_ARG(feats && level_embed && out, "null pointer");
    VER_CHECK_ARG(Ncam > 0 && B > 0 && S > 0 && C > 0 && C % 4 == 0, "bad dims (C %% 4 != 0?)");
    const size_t total = (size_t)Ncam * B * S * (C / 4);
    const int blocks = grid_for(total, 256, 148 * 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == VER_F32)
        feat_embed_kernel<float><<<blocks, 256, 0, st>>>((const float4*)feats, (const float4*)cams_embeds,
                                                         (const float4*)level_embed, (float*)out, Ncam, B, S, C / 4);
    else
        feat_embed_kernel<__half><<<blocks, 256, 0, st>>>((const float4*)feats, (const float4*)cams_embeds,
                                                          (const float4*)level_embed, (__half*)out, Ncam, B, S, C / 4);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

extern "C" int ver_add_layernorm(int dtype, const void* x, const void* residual, const float* gamma,
                                 const float* beta, void* y, int64_t rows, int C, float eps,
                                 ver_stream_t stream) {
    VER_CHECK_ARG(dtype == VER_F32 || dtype == VER_F16, "bad dtype %d", dtype);
    VER_CHECK_ARG(x && gamma && beta && y, "null pointer");
    VER_CHECK_ARG(rows > 0 && C > 0, "bad dims");
    if (C > 32 * 4 * 8) {
        ver_set_error("add_layernorm supports C <= 1024 (got %d)", C);
        return VER_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (int)((rows + 7) / 8);
#define LN_LAUNCH(T, MAXV)                                                                         \
    add_layernorm_kernel<T, MAXV><<<blocks, 256, 0, st>>>((const T*)x, (const T*)residual, gamma, beta, \
                                                         (T*)y, rows, C, eps)
    const int maxv = (C + 127) / 128;
    if (dtype == VER_F32) {
        if (maxv <= 1) LN_LAUNCH(float, 1); else if (maxv <= 2) LN_LAUNCH(float, 2);
        else if (maxv <= 6) LN_LAUNCH(float, 6); else LN_LAUNCH(float, 8);
    } else {
        if (maxv <= 1) LN_LAUNCH(__half, 1); else if (maxv <= 2) LN_LAUNCH(__half, 2);
        else if (maxv <= 6) LN_LAUNCH(__half, 6); else LN_LAUNCH(__half, 8);
    }
#undef LN_LAUNCH
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

extern "C" int ver_focal_loss(const float* logits, const int64_t* occ_gt, int n_gt, int32_t* dense_gt,
                              float* loss_sum, int32_t* num_pos, float* grad_logits, int64_t N,
                              int Ccls, float gamma, float alpha, ver_stream_t stream) {
    VER_CHECK_ARG(logits && dense_gt && loss_sum && num_pos, "null pointer");
    VER_CHECK_ARG(n_gt <= 0 || occ_gt, "null occ_gt with n_gt > 0");
    VER_CHECK_ARG(N > 0 && Ccls > 0, "bad dims");
    cudaStream_t st = (cudaStream_t)stream;
    // n_gt < 0: dense_gt is an INPUT (already-dense class targets), nothing to build
    if (n_gt >= 0) fill_i32_kernel<<<grid_for(N, 256), 256, 0, st>>>(dense_gt, N, Ccls);
    if (n_gt > 0) scatter_gt_kernel<<<(n_gt + 255) / 256, 256, 0, st>>>(occ_gt, n_gt, dense_gt, N, Ccls, num_pos);
    VER_CHECK_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(float), st));
    VER_CHECK_CUDA(cudaMemsetAsync(num_pos, 0, sizeof(int32_t), st));
    count_pos_kernel<<<grid_for(N, 256), 256, 0, st>>>(dense_gt, N, Ccls, num_pos);
    focal_kernel<<<grid_for((size_t)N * Ccls, 256, 148 * 8), 256, 0, st>>>(logits, dense_gt, loss_sum,
                                                                          grad_logits, N, Ccls, gamma, alpha);
    VER_CHECK_LAUNCH();
    g_ver_launches += n_gt > 0 ? 4 : (n_gt == 0 ? 3 : 2);
    return VER_OK;
}

extern "C" int ver_occupancy_decode(const float* logits, int64_t N, int Ccls, float threshold,
                                    int64_t* out_pairs, int32_t* out_count, int32_t* scratch,
                                    ver_stream_t stream) {
    VER_CHECK_ARG(logits && out_pairs && out_count && scratch, "null pointer");
    VER_CHECK_ARG(N > 0 && Ccls > 0, "bad dims");
    cudaStream_t st = (cudaStream_t)stream;
    const int nblocks = (int)((N + 1023) / 1024);
    decode_count_kernel<<<nblocks, 1024, 0, st>>>(logits, N, Ccls, threshold, scratch);
    decode_scan_kernel<<<1, 1024, 0, st>>>(scratch, nblocks, out_count);
    decode_emit_kernel<<<nblocks, 1024, 0, st>>>(logits, N, Ccls, threshold, scratch, out_pairs);
    VER_CHECK_LAUNCH();
    g_ver_launches += 3;
    return VER_OK;
}

// Training-path epilogues of VoxelFormerLayer (A6): HBM-bound row kernels that replace chains of
// separate dropout / add / cast / LayerNorm / ReLU passes (measured: 28 ms of a 139 ms step were
// such passes).  One warp per row, 16-byte vector accesses, fp32 statistics, counter-based
// (Philox4x32-10) dropout so the mask is regenerated in backward instead of stored.
//
//   z = residual + dropout(x) ; y = LayerNorm(z) * gamma + beta          (SCA / FFN -> 'norm')
//   h = dropout(relu(a))                                                  (FFN inner activation)
#include "common.cuh"
#include "philox.cuh"

namespace {

template <typename T>
struct Vec8;
template <>
struct Vec8<__half> {
    uint4 raw;
    __device__ __forceinline__ void load(const __half* p) { raw = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void store(__half* p) const { *reinterpret_cast<uint4*>(p) = raw; }
    __device__ __forceinline__ void get(float (&f)[8]) const {
        const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = __half22float2(h[i]);
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        }
    }
    __device__ __forceinline__ void set(const float (&f)[8]) {
        __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    }
};
template <>
struct Vec8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) {
        a = reinterpret_cast<const float4*>(p)[0];
        b = reinterpret_cast<const float4*>(p)[1];
    }
    __device__ __forceinline__ void store(float* p) const {
        reinterpret_cast<float4*>(p)[0] = a;
        reinterpret_cast<float4*>(p)[1] = b;
    }
    __device__ __forceinline__ void get(float (&f)[8]) const {
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    __device__ __forceinline__ void set(const float (&f)[8]) {
        a = make_float4(f[0], f[1], f[2], f[3]);
        b = make_float4(f[4], f[5], f[6], f[7]);
    }
};

constexpr int kRowsPerBlock = 8;       // 8 warps, one row each per iteration

// ---------------------------------------------------------------- forward
// NV = 8-element vectors per lane (C = 256 * NV ... handled as C <= 32*8*NV with bounds checks)
template <typename T, int NV>
__global__ void __launch_bounds__(256)
dropout_add_ln_fwd(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ gamma,
                   const float* __restrict__ beta, T* __restrict__ y, T* __restrict__ z_out,
                   float2* __restrict__ stats, int64_t rows, int C, float eps, float p, uint64_t seed,
                   const unsigned long long* __restrict__ seed_epoch, uint8_t* __restrict__ keep_bits) {
    if (seed_epoch) seed += *seed_epoch;        // device-side word: fresh masks on every replay of a captured graph
    // gamma / beta staged in shared memory as [lo | hi][C / 8] float4: a lane's 8 columns are two conflict-free
    // 16-byte reads.  (Reading them from global cost 32 L1 sectors per warp request -- lane stride 32 B -- and made
    // the L1 path, not HBM, the bound of this kernel: 10.7 GB of L1 traffic for 1.2 GB of DRAM traffic.)
    __shared__ float4 s_gb[4][128];
    for (int i = threadIdx.x; i < C / 8; i += blockDim.x) {
        s_gb[0][i] = reinterpret_cast<const float4*>(gamma)[2 * i];
        s_gb[1][i] = reinterpret_cast<const float4*>(gamma)[2 * i + 1];
        s_gb[2][i] = reinterpret_cast<const float4*>(beta)[2 * i];
        s_gb[3][i] = reinterpret_cast<const float4*>(beta)[2 * i + 1];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t thr16 = (uint32_t)(p * 65536.f);
    const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
    for (int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + warp; row < rows;
         row += (int64_t)gridDim.x * kRowsPerBlock) {
        float v[NV][8];
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = (k * 32 + lane) * 8;
            if (c < C) {
                Vec8<T> vx;
                vx.load(x + row * C + c);
                vx.get(v[k]);
                if (p > 0.f) {
                    const uint32_t m = keep8((uint64_t)row * C + c, seed, thr16);
                    if (keep_bits) keep_bits[((uint64_t)row * C + c) >> 3] = (uint8_t)m;     // one byte per 8 elements
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[k][e] = ((m >> e) & 1) ? v[k][e] * scale : 0.f;
                }
                if (res) {
                    Vec8<T> vr;
                    float r[8];
                    vr.load(res + row * C + c);
                    vr.get(r);
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[k][e] += r[e];
                }
                if (z_out) {      // the LN input as stored (rounded to T): backward recomputes from it
                    Vec8<T> vz;
                    vz.set(v[k]);
                    vz.store(z_out + row * C + c);
                    vz.get(v[k]);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) sum += v[k][e];
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[k][e] = 0.f;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(VER_FULL_MASK, sum, o);
        const float mean = sum / (float)C;
        float var = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = (k * 32 + lane) * 8;
            if (c < C) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float d = v[k][e] - mean;
                    var += d * d;
                }
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) var += __shfl_xor_sync(VER_FULL_MASK, var, o);
        const float rstd = rsqrtf(var / (float)C + eps);
        if (stats && lane == 0) stats[row] = make_float2(mean, rstd);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = (k * 32 + lane) * 8;
            if (c < C) {
                float o[8];
                const float4 g0 = s_gb[0][c >> 3], g1 = s_gb[1][c >> 3], b0 = s_gb[2][c >> 3], b1 = s_gb[3][c >> 3];
                const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = (v[k][e] - mean) * rstd * gm[e] + bt[e];
                Vec8<T> vy;
                vy.set(o);
                vy.store(y + row * C + c);
            }
        }
    }
}

// ---------------------------------------------------------------- backward
// dz = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
// d residual = dz ; dx = dz * keep / (1-p) ; per-block partial column sums of dgamma, dbeta and (optionally) dx --
// the latter is the bias gradient of the Linear that produced x (output_proj / the FFN's last Linear).
// Register budget: the packed dy / z vectors are kept between the statistics pass and the dz pass and unpacked
// twice, so that the three accumulator sets (9 x NV x 8 floats) fit in 128 registers -> 2 CTAs (16 warps) per SM.
template <typename T, int NV>
__global__ void __launch_bounds__(256, 2)
dropout_add_ln_bwd(const T* __restrict__ dy, const T* __restrict__ z, const float2* __restrict__ stats,
                   const float* __restrict__ gamma, T* __restrict__ dx, T* __restrict__ dres,
                   float* __restrict__ dgamma_part, float* __restrict__ dbeta_part, float* __restrict__ dxsum_part,
                   int64_t rows, int C, float p, uint64_t seed, const unsigned long long* __restrict__ seed_epoch,
                   const uint8_t* __restrict__ keep_bits) {
    if (seed_epoch) seed += *seed_epoch;
    extern __shared__ float s_part[];       // [8 warps][C], reused for the three reductions
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t thr16 = (uint32_t)(p * 65536.f);
    const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
    float dg[NV][8], db[NV][8], ds[NV][8];
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) dg[k][e] = db[k][e] = ds[k][e] = 0.f;
    // gamma staged in shared memory as [lo | hi][C / 8] float4 (conflict-free per-lane reads; see the forward kernel)
    __shared__ float4 s_g[2][128];
    for (int i = threadIdx.x; i < C / 8; i += blockDim.x) {
        s_g[0][i] = reinterpret_cast<const float4*>(gamma)[2 * i];
        s_g[1][i] = reinterpret_cast<const float4*>(gamma)[2 * i + 1];
    }
    __syncthreads();
    auto load_gamma = [&](int c, float (&gm)[8]) {
        const float4 a = s_g[0][c >> 3], b = s_g[1][c >> 3];
        gm[0] = a.x; gm[1] = a.y; gm[2] = a.z; gm[3] = a.w; gm[4] = b.x; gm[5] = b.y; gm[6] = b.z; gm[7] = b.w;
    };

    for (int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + warp; row < rows;
         row += (int64_t)gridDim.x * kRowsPerBlock) {
        const float2 st = stats[row];
        Vec8<T> vy[NV], vz[NV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = (k * 32 + lane) * 8;
            if (c < C) {
                vy[k].load(dy + row * C + c);
                vz[k].load(z + row * C + c);
            }
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = (k * 32 + lane) * 8;
            if (c < C) {
                float fy[8], fz[8], gm[8];
                vy[k].get(fy);
                vz[k].get(fz);
                load_gamma(c, gm);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float xh = (fz[e] - st.x) * st.y, g = fy[e] * gm[e];
                    s1 += g;
                    s2 += g * xh;
                    dg[k][e] += fy[e] * xh;
                    db[k][e] += fy[e];
                }
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            s1 += __shfl_xor_sync(VER_FULL_MASK, s1, o);
            s2 += __shfl_xor_sync(VER_FULL_MASK, s2, o);
        }
        const float m1 = s1 / (float)C, m2 = s2 / (float)C;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = (k * 32 + lane) * 8;
            if (c < C) {
                float fy[8], fz[8], dz[8], gm[8];
                vy[k].get(fy);
                vz[k].get(fz);
                load_gamma(c, gm);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float xh = (fz[e] - st.x) * st.y;
                    dz[e] = st.y * (fy[e] * gm[e] - m1 - xh * m2);
                }
                Vec8<T> o;
                if (dres) {
                    o.set(dz);
                    o.store(dres + row * C + c);
                }
                if (p > 0.f) {
                    const uint32_t m = keep_bits ? (uint32_t)keep_bits[((uint64_t)row * C + c) >> 3]
                                                 : keep8((uint64_t)row * C + c, seed, thr16);
#pragma unroll
                    for (int e = 0; e < 8; ++e) dz[e] = ((m >> e) & 1) ? dz[e] * scale : 0.f;
                }
                o.set(dz);
                o.store(dx + row * C + c);
                if (dxsum_part) {          // what was stored (rounded to T) is what the bias gradient sums
                    o.get(dz);
#pragma unroll
                    for (int e = 0; e < 8; ++e) ds[k][e] += dz[e];
                }
            }
        }
    }
    // block-level reduction of the parameter-gradient partials, one quantity at a time through [8][C] floats
    auto reduce = [&](float (&acc)[NV][8], float* out) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = (k * 32 + lane) * 8;
            if (c < C) {
#pragma unroll
                for (int e = 0; e < 8; ++e) s_part[warp * C + c + e] = acc[k][e];
            }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float a = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) a += s_part[w * C + c];
            out[(size_t)blockIdx.x * C + c] = a;
        }
        __syncthreads();
    };
    reduce(dg, dgamma_part);
    reduce(db, dbeta_part);
    if (dxsum_part) reduce(ds, dxsum_part);
}

// ---------------------------------------------------------------- backward, rows fed by the TMA engine (fp16)
// Same arithmetic as dropout_add_ln_bwd.  What changed: the dy / z rows reach the warp through a ring of kLnStages
// shared-memory row slots filled by cp.async.bulk (one elected lane, completion on an mbarrier per slot) instead of
// register loads.  The register version keeps 72 column accumulators per thread and therefore runs 16 warps per SM
// with 3 KB in flight each (49 KB per SM: 4.0 TB/s, profiles/r02n); the ring keeps kLnStages rows of both tensors
// in flight per warp without costing a register, and the second pass re-reads shared memory instead of holding the
// packed rows in 24 registers.
constexpr int kLnStages = 4;
template <int NV>
__global__ void __launch_bounds__(256, 2)
dropout_add_ln_bwd_tma(const __half* __restrict__ dy, const __half* __restrict__ z, const float2* __restrict__ stats,
                       const float* __restrict__ gamma, __half* __restrict__ dx, __half* __restrict__ dres,
                       float* __restrict__ dgamma_part, float* __restrict__ dbeta_part, float* __restrict__ dxsum_part,
                       int64_t rows, int C, float p, uint64_t seed, const unsigned long long* __restrict__ seed_epoch,
                       const uint8_t* __restrict__ keep_bits) {
    if (seed_epoch) seed += *seed_epoch;
    // [8 warps][kLnStages][dy row | z row]; the first 8 * C floats double as the reduction buffer at the end
    extern __shared__ __align__(128) unsigned char s_ring[];
    __shared__ __align__(8) uint64_t s_full[8][kLnStages];
    __shared__ float4 s_g[2][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t row_bytes = (uint32_t)C * 2u;
    unsigned char* my_ring = s_ring + (size_t)warp * kLnStages * 2 * row_bytes;
    const uint32_t thr16 = (uint32_t)(p * 65536.f);
    const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
    float dg[NV][8], db[NV][8], ds[NV][8];
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) dg[k][e] = db[k][e] = ds[k][e] = 0.f;
    for (int i = threadIdx.x; i < C / 8; i += blockDim.x) {
        s_g[0][i] = reinterpret_cast<const float4*>(gamma)[2 * i];
        s_g[1][i] = reinterpret_cast<const float4*>(gamma)[2 * i + 1];
    }
    if (lane == 0) {
        for (int st = 0; st < kLnStages; ++st) mbar_init(&s_full[warp][st], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto load_gamma = [&](int c, float (&gm)[8]) {
        const float4 a = s_g[0][c >> 3], b = s_g[1][c >> 3];
        gm[0] = a.x; gm[1] = a.y; gm[2] = a.z; gm[3] = a.w; gm[4] = b.x; gm[5] = b.y; gm[6] = b.z; gm[7] = b.w;
    };
    const int64_t row0 = (int64_t)blockIdx.x * kRowsPerBlock + warp, stride = (int64_t)gridDim.x * kRowsPerBlock;
    auto issue = [&](int64_t it) {                      // row of iteration `it` -> slot it % kLnStages (lane 0)
        const int64_t row = row0 + it * stride;
        if (row >= rows) return;
        const int st = (int)(it % kLnStages);
        unsigned char* slot = my_ring + (size_t)st * 2 * row_bytes;
        mbar_expect_tx(&s_full[warp][st], 2 * row_bytes);
        bulk_g2s(slot, dy + row * C, row_bytes, &s_full[warp][st]);
        bulk_g2s(slot + row_bytes, z + row * C, row_bytes, &s_full[warp][st]);
    };
    if (lane == 0)
        for (int it = 0; it < kLnStages; ++it) issue(it);
    int64_t it = 0;
    for (int64_t row = row0; row < rows; row += stride, ++it) {
        const int st = (int)(it % kLnStages);
        const float2 stt = stats[row];
        uint32_t mk[NV];                                // the forward's keep bits (in flight during the wait)
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = (k * 32 + lane) * 8;
            mk[k] = (keep_bits && p > 0.f && c < C) ? (uint32_t)keep_bits[((uint64_t)row * C + c) >> 3] : 0u;
        }
        mbar_wait(&s_full[warp][st], (uint32_t)((it / kLnStages) & 1));
        const unsigned char* slot = my_ring + (size_t)st * 2 * row_bytes;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = (k * 32 + lane) * 8;
            if (c < C) {
                Vec8<__half> vy, vz;
                vy.load(reinterpret_cast<const __half*>(slot) + c);
                vz.load(reinterpret_cast<const __half*>(slot + row_bytes) + c);
                float fy[8], fz[8], gm[8];
                vy.get(fy);
                vz.get(fz);
                load_gamma(c, gm);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float xh = (fz[e] - stt.x) * stt.y, g = fy[e] * gm[e];
                    s1 += g;
                    s2 += g * xh;
                    dg[k][e] += fy[e] * xh;
                    db[k][e] += fy[e];
                }
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            s1 += __shfl_xor_sync(VER_FULL_MASK, s1, o);
            s2 += __shfl_xor_sync(VER_FULL_MASK, s2, o);
        }
        const float m1 = s1 / (float)C, m2 = s2 / (float)C;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = (k * 32 + lane) * 8;
            if (c < C) {
                Vec8<__half> vy, vz;
                vy.load(reinterpret_cast<const __half*>(slot) + c);
                vz.load(reinterpret_cast<const __half*>(slot + row_bytes) + c);
                float fy[8], fz[8], dz[8], gm[8];
                vy.get(fy);
                vz.get(fz);
                load_gamma(c, gm);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float xh = (fz[e] - stt.x) * stt.y;
                    dz[e] = stt.y * (fy[e] * gm[e] - m1 - xh * m2);
                }
                Vec8<__half> o;
                if (dres) {
                    o.set(dz);
                    o.store(dres + row * C + c);
                }
                if (p > 0.f) {
                    const uint32_t m = keep_bits ? mk[k] : keep8((uint64_t)row * C + c, seed, thr16);
#pragma unroll
                    for (int e = 0; e < 8; ++e) dz[e] = ((m >> e) & 1) ? dz[e] * scale : 0.f;
                }
                o.set(dz);
                o.store(dx + row * C + c);
                if (dxsum_part) {          // what was stored (rounded to fp16) is what the bias gradient sums
                    o.get(dz);
#pragma unroll
                    for (int e = 0; e < 8; ++e) ds[k][e] += dz[e];
                }
            }
        }
        __syncwarp();                                   // every lane is done with the slot: refill it
        if (lane == 0) issue(it + kLnStages);
    }
    // block-level reduction of the parameter-gradient partials through [8][C] floats (the ring is idle now: every
    // issued copy has been waited for)
    __syncthreads();
    float* s_part = reinterpret_cast<float*>(s_ring);
    auto reduce = [&](float (&acc)[NV][8], float* out) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = (k * 32 + lane) * 8;
            if (c < C) {
#pragma unroll
                for (int e = 0; e < 8; ++e) s_part[warp * C + c + e] = acc[k][e];
            }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float a = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) a += s_part[w * C + c];
            out[(size_t)blockIdx.x * C + c] = a;
        }
        __syncthreads();
    };
    reduce(dg, dgamma_part);
    reduce(db, dbeta_part);
    if (dxsum_part) reduce(ds, dxsum_part);
}

// ---------------------------------------------------------------- h = dropout(relu(a)), in place capable
template <typename T>
__global__ void __launch_bounds__(256)
relu_dropout_fwd(const T* __restrict__ a, T* __restrict__ h, int64_t n8, float p, uint64_t seed,
                 const unsigned long long* __restrict__ seed_epoch) {
    if (seed_epoch) seed += *seed_epoch;
    const uint32_t thr16 = (uint32_t)(p * 65536.f);
    const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        Vec8<T> v;
        float f[8];
        v.load(a + i * 8);
        v.get(f);
        const uint32_t m = p > 0.f ? keep8((uint64_t)i * 8, seed, thr16) : 0xffu;
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = (((m >> e) & 1) && f[e] > 0.f) ? f[e] * scale : 0.f;
        v.set(f);
        v.store(h + i * 8);
    }
}
// da = dh * [h > 0] / (1 - p)   (a kept element with relu(a) == 0 has zero gradient either way)
// Optional column sums of da (= bias gradient of the FFN's first Linear): the launch makes gridDim * blockDim a
// multiple of C / 8, so a thread meets the same 8 columns in every iteration and keeps 8 partial sums; thread t of
// the grid writes them to part[t][8] and the caller sums the rows with equal t % (C / 8).
template <typename T>
__global__ void __launch_bounds__(256)
relu_dropout_bwd(const T* __restrict__ dh, const T* __restrict__ h, T* __restrict__ da, int64_t n8, float p,
                 float* __restrict__ part) {
    const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        Vec8<T> vg, vh;
        float g[8], f[8];
        vg.load(dh + i * 8);
        vg.get(g);
        vh.load(h + i * 8);
        vh.get(f);
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] = f[e] > 0.f ? g[e] * scale : 0.f;
        vg.set(g);
        vg.store(da + i * 8);
        if (part) {
            vg.get(g);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += g[e];
        }
    }
    if (part) {
        float4* o = reinterpret_cast<float4*>(part + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8);
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

// y = T(x) for fp32 x, plus the column sums of x (same partial-sum scheme): the fp32 gradients the sampler's backward
// produces (grad_value, grad_logits) become GEMM operands and bias gradients in one pass
template <typename T>
__global__ void __launch_bounds__(256)
cast_colsum_kernel(const float* __restrict__ x, T* __restrict__ y, int64_t n8, float* __restrict__ part) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        Vec8<float> v;
        float f[8];
        v.load(x + i * 8);
        v.get(f);
        Vec8<T> o;
        o.set(f);
        o.store(y + i * 8);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
    float4* o = reinterpret_cast<float4*>(part + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8);
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

// column sums of an fp16 matrix (bias gradient of a Linear whose output gradient is already fp16)
__global__ void __launch_bounds__(256)
colsum_f16_kernel(const __half* __restrict__ x, int64_t n8, float* __restrict__ part) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        Vec8<__half> v;
        float f[8];
        v.load(x + i * 8);
        v.get(f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
    float4* o = reinterpret_cast<float4*>(part + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8);
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

constexpr int kColsumGrid = 148 * 12;      // x 256 threads: a multiple of 3 * 256, i.e. of C / 8 for C | 6144

int ln_grid(int64_t rows) {
    const int64_t need = (rows + kRowsPerBlock - 1) / kRowsPerBlock;
    const int cap = 148 * 4;
    return (int)(need < cap ? (need ? need : 1) : cap);
}

}  // namespace

#define LN_DISPATCH(KERN, T, nv, ...)                                   \
    do {                                                                \
        if (nv <= 1) KERN<T, 1> __VA_ARGS__;                            \
        else if (nv <= 2) KERN<T, 2> __VA_ARGS__;                       \
        else if (nv <= 3) KERN<T, 3> __VA_ARGS__;                       \
        else KERN<T, 4> __VA_ARGS__;                                    \
    } while (0)

extern "C" int ver_dropout_add_layernorm_fwd(int dtype, const void* x, const void* residual,
                                             const float* gamma, const float* beta, void* y, void* z_out,
                                             float* stats, int64_t rows, int C, float eps, float p_drop,
                                             uint64_t seed, const uint64_t* seed_epoch, ver_stream_t stream) {
    return ver_dropout_add_layernorm_fwd_bits(dtype, x, residual, gamma, beta, y, z_out, stats, nullptr, rows, C, eps,
                                              p_drop, seed, seed_epoch, stream);
}

extern "C" int ver_dropout_add_layernorm_fwd_bits(int dtype, const void* x, const void* residual,
                                                  const float* gamma, const float* beta, void* y, void* z_out,
                                                  float* stats, uint8_t* keep_bits, int64_t rows, int C, float eps,
                                                  float p_drop, uint64_t seed, const uint64_t* seed_epoch,
                                                  ver_stream_t stream) {
    const unsigned long long* ep = (const unsigned long long*)seed_epoch;
    VER_CHECK_ARG(dtype == VER_F32 || dtype == VER_F16, "bad dtype %d", dtype);
    VER_CHECK_ARG(x && gamma && beta && y, "null pointer");
    VER_CHECK_ARG(rows > 0 && C > 0 && C % 8 == 0 && C <= 1024, "C must be a multiple of 8, <= 1024 (got %d)", C);
    VER_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "bad dropout probability");
    cudaStream_t st = (cudaStream_t)stream;
    const int nv = (C + 255) / 256;
    const int grid = ln_grid(rows);
    if (dtype == VER_F16)
        LN_DISPATCH(dropout_add_ln_fwd, __half, nv, <<<grid, 256, 0, st>>>((const __half*)x, (const __half*)residual, gamma, beta, (__half*)y, (__half*)z_out, (float2*)stats, rows, C, eps, p_drop, seed, ep, keep_bits));
    else
        LN_DISPATCH(dropout_add_ln_fwd, float, nv, <<<grid, 256, 0, st>>>((const float*)x, (const float*)residual, gamma, beta, (float*)y, (float*)z_out, (float2*)stats, rows, C, eps, p_drop, seed, ep, keep_bits));
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

extern "C" int ver_dropout_add_layernorm_bwd_blocks(int64_t rows) { return ln_grid(rows); }

// A/B switch of tools/ln_bench.py and the tests: 1 (default) = TMA-fed fp16 backward where it applies, 0 = register loads
static int g_ln_bwd_tma = 1;
extern "C" int ver_debug_ln_bwd_tma(int on) {
    g_ln_bwd_tma = on;
    return VER_OK;
}

extern "C" int ver_dropout_add_layernorm_bwd(int dtype, const void* dy, const void* z, const float* stats,
                                             const float* gamma, void* dx, void* dresidual,
                                             float* dgamma_part, float* dbeta_part, float* dxsum_part,
                                             int64_t rows, int C, float p_drop, uint64_t seed,
                                             const uint64_t* seed_epoch, ver_stream_t stream) {
    return ver_dropout_add_layernorm_bwd_bits(dtype, dy, z, stats, gamma, nullptr, dx, dresidual, dgamma_part, dbeta_part,
                                              dxsum_part, rows, C, p_drop, seed, seed_epoch, stream);
}

extern "C" int ver_dropout_add_layernorm_bwd_bits(int dtype, const void* dy, const void* z, const float* stats,
                                                  const float* gamma, const uint8_t* keep_bits, void* dx,
                                                  void* dresidual, float* dgamma_part, float* dbeta_part,
                                                  float* dxsum_part, int64_t rows, int C, float p_drop, uint64_t seed,
                                                  const uint64_t* seed_epoch, ver_stream_t stream) {
    const unsigned long long* ep = (const unsigned long long*)seed_epoch;
    VER_CHECK_ARG(dtype == VER_F32 || dtype == VER_F16, "bad dtype %d", dtype);
    VER_CHECK_ARG(dy && z && stats && gamma && dx && dgamma_part && dbeta_part, "null pointer");
    VER_CHECK_ARG(rows > 0 && C > 0 && C % 8 == 0 && C <= 1024, "C must be a multiple of 8, <= 1024 (got %d)", C);
    cudaStream_t st = (cudaStream_t)stream;
    const int nv = (C + 255) / 256;
    const int grid = ln_grid(rows);
    const size_t smem = (size_t)8 * C * sizeof(float);
    // fp16 rows of a whole number of 16-byte units, ring + partials within half an SM's shared memory: TMA-fed kernel
    const size_t ring = (size_t)8 * kLnStages * 2 * C * 2;
    if (dtype == VER_F16 && g_ln_bwd_tma && ring >= smem && ring <= 100 * 1024) {
        cudaFuncSetAttribute(dropout_add_ln_bwd_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring);
        cudaFuncSetAttribute(dropout_add_ln_bwd_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring);
        cudaFuncSetAttribute(dropout_add_ln_bwd_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring);
        cudaFuncSetAttribute(dropout_add_ln_bwd_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring);
#define LN_TMA(NVV) dropout_add_ln_bwd_tma<NVV><<<grid, 256, ring, st>>>((const __half*)dy, (const __half*)z, (const float2*)stats, gamma, (__half*)dx, (__half*)dresidual, dgamma_part, dbeta_part, dxsum_part, rows, C, p_drop, seed, ep, keep_bits)
        if (nv <= 1) LN_TMA(1);
        else if (nv <= 2) LN_TMA(2);
        else if (nv <= 3) LN_TMA(3);
        else LN_TMA(4);
#undef LN_TMA
    } else if (dtype == VER_F16) {
        if (smem > 48 * 1024) {
            cudaFuncSetAttribute(dropout_add_ln_bwd<__half, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(dropout_add_ln_bwd<__half, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        LN_DISPATCH(dropout_add_ln_bwd, __half, nv, <<<grid, 256, smem, st>>>((const __half*)dy, (const __half*)z, (const float2*)stats, gamma, (__half*)dx, (__half*)dresidual, dgamma_part, dbeta_part, dxsum_part, rows, C, p_drop, seed, ep, keep_bits));
    } else {
        if (smem > 48 * 1024) {
            cudaFuncSetAttribute(dropout_add_ln_bwd<float, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(dropout_add_ln_bwd<float, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        LN_DISPATCH(dropout_add_ln_bwd, float, nv, <<<grid, 256, smem, st>>>((const float*)dy, (const float*)z, (const float2*)stats, gamma, (float*)dx, (float*)dresidual, dgamma_part, dbeta_part, dxsum_part, rows, C, p_drop, seed, ep, keep_bits));
    }
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

extern "C" int ver_relu_dropout_fwd(int dtype, const void* a, void* h, int64_t n, float p_drop, uint64_t seed,
                                    const uint64_t* seed_epoch, ver_stream_t stream) {
    const unsigned long long* ep = (const unsigned long long*)seed_epoch;
    VER_CHECK_ARG(dtype == VER_F32 || dtype == VER_F16, "bad dtype %d", dtype);
    VER_CHECK_ARG(a && h && n > 0 && n % 8 == 0, "bad arguments (n %% 8 != 0?)");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n8 = n / 8;
    const int grid = (int)((n8 + 255) / 256 < 148 * 16 ? (n8 + 255) / 256 : 148 * 16);
    if (dtype == VER_F16) relu_dropout_fwd<__half><<<grid, 256, 0, st>>>((const __half*)a, (__half*)h, n8, p_drop, seed, ep);
    else relu_dropout_fwd<float><<<grid, 256, 0, st>>>((const float*)a, (float*)h, n8, p_drop, seed, ep);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

extern "C" int ver_colsum_partial_rows(void) { return kColsumGrid * 256; }

extern "C" int ver_relu_dropout_bwd(int dtype, const void* dh, const void* h, void* da, int64_t n, float p_drop,
                                    int C, float* colsum_part, ver_stream_t stream) {
    VER_CHECK_ARG(dtype == VER_F32 || dtype == VER_F16, "bad dtype %d", dtype);
    VER_CHECK_ARG(dh && h && da && n > 0 && n % 8 == 0, "bad arguments (n %% 8 != 0?)");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n8 = n / 8;
    int grid = (int)((n8 + 255) / 256 < 148 * 16 ? (n8 + 255) / 256 : 148 * 16);
    if (colsum_part) {
        VER_CHECK_ARG(C > 0 && C % 8 == 0 && ((int64_t)kColsumGrid * 256) % (C / 8) == 0 && n % C == 0,
                      "column sums need C / 8 to divide %d (got C = %d)", kColsumGrid * 256, C);
        grid = kColsumGrid;
    }
    if (dtype == VER_F16) relu_dropout_bwd<__half><<<grid, 256, 0, st>>>((const __half*)dh, (const __half*)h, (__half*)da, n8, p_drop, colsum_part);
    else relu_dropout_bwd<float><<<grid, 256, 0, st>>>((const float*)dh, (const float*)h, (float*)da, n8, p_drop, colsum_part);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

// ---------------------------------------------------------------- out[m, C] = sum over the P rows of part[m, P, C]
// Folds the partial sums the kernels above emit (per-thread 8-float groups viewed as rows of C floats, or per-block
// rows).  One block of 32 warps per 128-column slab and matrix: warp w adds rows w, w + 32, ... (coalesced 512-byte
// row segments, four independent loads in flight per lane), the 32 warp partials are added in a fixed order --
// deterministic, one pass over the data.  (Round-2 profile r02n: the previous version let the LAST block add 296
// block partials serially -- 54 us per call, 31 calls per step = 1.7 ms of a 21 ms step for a few megabytes.)
__global__ void __launch_bounds__(1024)
colsum_fold_kernel(const float* __restrict__ part, int64_t P, int C, float* __restrict__ out) {
    __shared__ float4 s_acc[32][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 128 + lane * 4;
    const float* src = part + (size_t)blockIdx.y * P * C + c;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C) {
        int64_t r = warp;
        for (; r + 96 < P; r += 128) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(src + r * C));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(src + (r + 32) * C));
            const float4 v2 = __ldg(reinterpret_cast<const float4*>(src + (r + 64) * C));
            const float4 v3 = __ldg(reinterpret_cast<const float4*>(src + (r + 96) * C));
            a.x += (v0.x + v1.x) + (v2.x + v3.x);
            a.y += (v0.y + v1.y) + (v2.y + v3.y);
            a.z += (v0.z + v1.z) + (v2.z + v3.z);
            a.w += (v0.w + v1.w) + (v2.w + v3.w);
        }
        for (; r < P; r += 32) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + r * C));
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
    }
    s_acc[warp][lane] = a;
    __syncthreads();
    if (warp == 0 && c < C) {
        float4 t = s_acc[0][lane];
#pragma unroll
        for (int w = 1; w < 32; ++w) {
            const float4 v = s_acc[w][lane];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        *reinterpret_cast<float4*>(out + (size_t)blockIdx.y * C + c) = t;
    }
}

extern "C" int ver_colsum_fold_scratch_floats(int C) { return 4; }     // (kept for ABI stability: no scratch needed)

extern "C" int ver_colsum_fold_batched(const float* part, int n_mat, int64_t P, int C, float* out, ver_stream_t stream) {
    VER_CHECK_ARG(part && out && n_mat > 0 && P > 0 && C > 0 && C % 4 == 0, "bad arguments");
    colsum_fold_kernel<<<dim3((C + 127) / 128, n_mat), 1024, 0, (cudaStream_t)stream>>>(part, P, C, out);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

extern "C" int ver_colsum_fold(const float* part, int64_t P, int C, float* out, float* scratch, ver_stream_t stream) {
    (void)scratch;
    return ver_colsum_fold_batched(part, 1, P, C, out, stream);
}

extern "C" int ver_colsum_f16(const void* x, int64_t rows, int C, float* colsum_part, ver_stream_t stream) {
    VER_CHECK_ARG(x && colsum_part && rows > 0, "bad arguments");
    VER_CHECK_ARG(C > 0 && C % 8 == 0 && ((int64_t)kColsumGrid * 256) % (C / 8) == 0,
                  "column sums need C / 8 to divide %d (got C = %d)", kColsumGrid * 256, C);
    colsum_f16_kernel<<<kColsumGrid, 256, 0, (cudaStream_t)stream>>>((const __half*)x, rows * C / 8, colsum_part);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

extern "C" int ver_cast_colsum(int dtype, const float* x, void* y, int64_t rows, int C, float* colsum_part,
                               ver_stream_t stream) {
    VER_CHECK_ARG(dtype == VER_F32 || dtype == VER_F16, "bad dtype %d", dtype);
    VER_CHECK_ARG(x && y && colsum_part && rows > 0, "bad arguments");
    VER_CHECK_ARG(C > 0 && C % 8 == 0 && ((int64_t)kColsumGrid * 256) % (C / 8) == 0,
                  "column sums need C / 8 to divide %d (got C = %d)", kColsumGrid * 256, C);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n8 = rows * C / 8;
    if (dtype == VER_F16) cast_colsum_kernel<__half><<<kColsumGrid, 256, 0, st>>>(x, (__half*)y, n8, colsum_part);
    else cast_colsum_kernel<float><<<kColsumGrid, 256, 0, st>>>(x, (float*)y, n8, colsum_part);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

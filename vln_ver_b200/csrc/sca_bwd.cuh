// Backward building blocks shared by the operator-level (msda.cu) and fused (sca.cu)
// samplers.
//
// One CTA (4 warps) owns one (view, head): its [S][Dh] value map AND its fp32
// grad_value map both live in shared memory, so grad_value needs no atomics at
// all -- it is written to HBM once, coalesced, when the CTA retires.  Race
// freedom comes from ownership: warp w owns the channels [w*Dh/4, (w+1)*Dh/4) of
// every pixel, all four warps walk the same item list, and inside a warp the 8
// sampling points are processed one after the other (lane = corner*8 + j, lane j
// owns channels w*Dh/4 + j + 8k).  The grad map is XOR-swizzled at 8-word
// granularity with the pixel parity so the four bilinear corners of a tap fall
// in four different bank groups.
#pragma once
#include "sampler.cuh"

constexpr int kBwdThreads = 128;
constexpr int kBwdWarps = kBwdThreads / 32;

struct TapB {
    float coef;  // aw * bilinear weight (0 if the corner is outside)
    float bil;   // bilinear weight of this corner (0 if outside)
    float dx;    // d bil / d x  (+-wy, 0 if outside)
    float dy;    // d bil / d y  (+-wx, 0 if outside)
    float aw;
    int pixsw;   // pixel index | swizzle << 16, or -1 if outside
};

__device__ __forceinline__ TapB make_tap_bwd(float lx, float ly, float aw, int corner, int Sh,
                                             int Sw) {
    TapB t;
    t.coef = t.bil = t.dx = t.dy = 0.f;
    t.aw = aw;
    t.pixsw = -1;
    const float x = lx * (float)Sw - 0.5f, y = ly * (float)Sh - 0.5f;
    if (x > -1.f && y > -1.f && x < (float)Sw && y < (float)Sh) {
        const float xf = floorf(x), yf = floorf(y);
        const float fx = x - xf, fy = y - yf;
        const int xi = (int)xf + (corner & 1), yi = (int)yf + (corner >> 1);
        const float wx = (corner & 1) ? fx : 1.f - fx;
        const float wy = (corner >> 1) ? fy : 1.f - fy;
        if (xi >= 0 && xi < Sw && yi >= 0 && yi < Sh) {
            t.bil = wy * wx;
            t.coef = aw * t.bil;
            t.dx = (corner & 1) ? wy : -wy;
            t.dy = (corner >> 1) ? wx : -wx;
            t.pixsw = (yi * Sw + xi) | (((xi & 1) | ((yi & 1) << 1)) << 16);
        }
    }
    return t;
}

template <typename T, int CPL>
struct BwdSmem {
    static constexpr int Dh = CPL * 8;
    T* val;         // [S][Dh]
    float* grad;    // [S][Dh] swizzled
    float* part;    // [2][kBwdWarps][32]
    uint64_t* bar;
    __device__ BwdSmem(unsigned char* raw, int S) {
        grad = reinterpret_cast<float*>(raw);
        val = reinterpret_cast<T*>(raw + (size_t)S * Dh * sizeof(float));
        part = reinterpret_cast<float*>(raw + (size_t)S * Dh * (sizeof(float) + sizeof(T)));
        bar = reinterpret_cast<uint64_t*>(part + 2 * kBwdWarps * 32);
    }
    static size_t bytes(int S) {
        return (size_t)S * Dh * (sizeof(float) + sizeof(T)) + 2 * kBwdWarps * 32 * sizeof(float) + 16;
    }
};

__device__ __forceinline__ int swz_word(int pix, int sw, int c, int Dh) {
    const int grp = c >> 3;
    return pix * Dh + ((((grp & ~3) | ((grp & 3) ^ sw))) << 3) + (c & 7);
}

template <typename T, int CPL>
__device__ __forceinline__ void bwd_prologue(BwdSmem<T, CPL>& sm, const T* gsrc, int S, size_t s_row,
                                             int lane, int warp) {
    constexpr int Dh = CPL * 8;
    if (threadIdx.x == 0) {
        mbar_init(sm.bar, 1);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < S * Dh / 4; i += kBwdThreads)
        reinterpret_cast<float4*>(sm.grad)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (warp == 0) stage_map(sm.val, gsrc, S, Dh, s_row, sm.bar, lane);
    mbar_wait(sm.bar, 0);
}

// Processes one (item, head).  `gptr` points at the Dh grad_out channels of the item.
// Returns, in ALL lanes of warp (item % 4), (dL/d aw_p, dL/d loc_x_p, dL/d loc_y_p) for
// p = lane & 7; other warps return garbage.
template <typename T, int CPL>
__device__ __forceinline__ float3 bwd_process_item(BwdSmem<T, CPL>& sm, const TapB& tap,
                                                   const T* __restrict__ gptr, float gscale, int Sh,
                                                   int Sw, int NP, int lane, int warp, int item) {
    constexpr int Dh = CPL * 8;
    constexpr int KPL = CPL / 4;           // channels per lane (stride 8)
    constexpr int CPW = Dh / kBwdWarps;    // channels per warp
    const int cbase = lane & 24, j = lane & 7;
    const int c0 = warp * CPW + j;
    float gk[KPL];
#pragma unroll
    for (int k = 0; k < KPL; ++k) gk[k] = gscale * to_f32(gptr[c0 + 8 * k]);

    float mydot = 0.f;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        if (p >= NP) break;
        const float coef = __shfl_sync(VER_FULL_MASK, tap.coef, cbase | p);
        const int pixsw = __shfl_sync(VER_FULL_MASK, tap.pixsw, cbase | p);
        const bool valid = pixsw >= 0;
        if (__any_sync(VER_FULL_MASK, valid)) {
            float partial = 0.f;
            if (valid) {
                const int pix = pixsw & 0xffff, sw = pixsw >> 16;
                const T* vrow = sm.val + pix * Dh + c0;
#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    partial = fmaf(gk[k], to_f32(vrow[8 * k]), partial);
                    float* gp = sm.grad + swz_word(pix, sw, c0 + 8 * k, Dh);
                    *gp = fmaf(coef, gk[k], *gp);
                }
            }
            partial += __shfl_xor_sync(VER_FULL_MASK, partial, 1);
            partial += __shfl_xor_sync(VER_FULL_MASK, partial, 2);
            partial += __shfl_xor_sync(VER_FULL_MASK, partial, 4);
            if (j == p) mydot = partial;
            __syncwarp();
        }
    }
    float* part = sm.part + (item & 1) * kBwdWarps * 32;
    part[warp * 32 + lane] = mydot;
    __syncthreads();
    float3 r = make_float3(0.f, 0.f, 0.f);
    if (warp == (item & (kBwdWarps - 1))) {
        const float d = part[lane] + part[32 + lane] + part[64 + lane] + part[96 + lane];
        float ga = tap.bil * d, gx = tap.dx * d, gy = tap.dy * d;
        ga += __shfl_xor_sync(VER_FULL_MASK, ga, 8);
        ga += __shfl_xor_sync(VER_FULL_MASK, ga, 16);
        gx += __shfl_xor_sync(VER_FULL_MASK, gx, 8);
        gx += __shfl_xor_sync(VER_FULL_MASK, gx, 16);
        gy += __shfl_xor_sync(VER_FULL_MASK, gy, 8);
        gy += __shfl_xor_sync(VER_FULL_MASK, gy, 16);
        r = make_float3(ga, gx * tap.aw * (float)Sw, gy * tap.aw * (float)Sh);
    }
    return r;
}

// un-swizzle and write the finished grad_value map: gdst[pix*s_row + c]
template <typename T, int CPL>
__device__ __forceinline__ void bwd_epilogue(BwdSmem<T, CPL>& sm, float* gdst, int S, size_t s_row,
                                             int Sw, int lane, int warp) {
    constexpr int Dh = CPL * 8;
    __syncthreads();
    const int vec_per_row = Dh / 4;
    for (int i = threadIdx.x; i < S * vec_per_row; i += kBwdThreads) {
        const int pix = i / vec_per_row, c = (i % vec_per_row) * 4;
        const int sw = ((pix % Sw) & 1) | (((pix / Sw) & 1) << 1);
        const float4 v = *reinterpret_cast<const float4*>(sm.grad + swz_word(pix, sw, c, Dh));
        *reinterpret_cast<float4*>(gdst + (size_t)pix * s_row + c) = v;
    }
}

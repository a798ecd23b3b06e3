// Device-side building blocks shared by the op-level (msda.cu) and fused (sca.cu)
// deformable samplers.
//
// Warp mapping ("corner x channel-group"): a warp processes one (query|hit, head)
// at a time.  lane = corner*8 + g, corner bit0 = +1 in x, bit1 = +1 in y, and lane
// g owns the CPL = Dh/8 contiguous channels [g*CPL, (g+1)*CPL).  The same 32 lanes
// also enumerate the 8 points x 4 corners "taps" of that (query, head): lane
// (corner, p) sets up tap (p, corner) once and the warp broadcasts it with
// shuffles while looping over p.  With Dh = 96 and a [pixel][channel] tile this
// makes every shared-memory wavefront conflict free: the 8 lanes of a corner read
// 8*CPL contiguous elements, x and x+1 are 48 (fp16) / 96 (fp32) words apart and
// the two image rows are served by different half-warps.
#pragma once
#include "common.cuh"

struct Tap {
    float coef;  // attention weight * bilinear weight of this corner (0 if outside)
    int off;     // element offset of the corner pixel inside the tile (0 if outside)
};

// mmcv's ms_deform_attn im2col convention: pixel = loc * size - 0.5, zeros outside.
__device__ __forceinline__ Tap make_tap(float lx, float ly, float aw, int corner, int Sh, int Sw,
                                        int row_elems) {
    Tap t;
    const float x = lx * (float)Sw - 0.5f, y = ly * (float)Sh - 0.5f;
    t.coef = 0.f;
    t.off = 0;
    if (x > -1.f && y > -1.f && x < (float)Sw && y < (float)Sh) {
        const float xf = floorf(x), yf = floorf(y);
        const float fx = x - xf, fy = y - yf;
        const int xi = (int)xf + (corner & 1), yi = (int)yf + (corner >> 1);
        const float wx = (corner & 1) ? fx : 1.f - fx;
        const float wy = (corner >> 1) ? fy : 1.f - fy;
        if (xi >= 0 && xi < Sw && yi >= 0 && yi < Sh) {
            t.coef = aw * (wy * wx);
            t.off = (yi * Sw + xi) * row_elems;
        }
    }
    return t;
}

// acc[k] += coef * tile[off + g*CPL + k] over the 8 points (fp32 storage).
template <int CPL>
__device__ __forceinline__ void gather8(const float* __restrict__ tile, Tap tap, int lane, int NP,
                                        float (&acc)[CPL]) {
    const int cbase = lane & 24;
    const int chan = (lane & 7) * CPL;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        if (p >= NP) break;
        const float c = __shfl_sync(VER_FULL_MASK, tap.coef, cbase | p);
        const int o = __shfl_sync(VER_FULL_MASK, tap.off, cbase | p);
        if (__any_sync(VER_FULL_MASK, c != 0.f)) {
            const float4* src = reinterpret_cast<const float4*>(tile + o + chan);
#pragma unroll
            for (int k = 0; k < CPL / 4; ++k) {
                const float4 v = src[k];
                acc[4 * k + 0] = fmaf(c, v.x, acc[4 * k + 0]);
                acc[4 * k + 1] = fmaf(c, v.y, acc[4 * k + 1]);
                acc[4 * k + 2] = fmaf(c, v.z, acc[4 * k + 2]);
                acc[4 * k + 3] = fmaf(c, v.w, acc[4 * k + 3]);
            }
        }
    }
}

// fp16 storage: the tap coefficient is rounded to fp16 and multiplied exactly into
// the fp32 accumulator by FHFMA (fma.rn.f32.f16) -- no cvt in the inner loop.
template <int CPL>
__device__ __forceinline__ void gather8(const __half* __restrict__ tile, Tap tap, int lane, int NP,
                                        float (&acc)[CPL]) {
    const int cbase = lane & 24;
    const int chan = (lane & 7) * CPL;
    const uint32_t ch16 = __half_as_ushort(__float2half_rn(tap.coef));
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        if (p >= NP) break;
        const uint16_t c = (uint16_t)__shfl_sync(VER_FULL_MASK, ch16, cbase | p);
        const int o = __shfl_sync(VER_FULL_MASK, tap.off, cbase | p);
        if (__any_sync(VER_FULL_MASK, (c & 0x7fff) != 0)) {
            const uint2* src = reinterpret_cast<const uint2*>(tile + o + chan);
#pragma unroll
            for (int k = 0; k < CPL / 4; ++k) {
                const uint2 v = src[k];
                acc[4 * k + 0] = fhfma((uint16_t)(v.x & 0xffff), c, acc[4 * k + 0]);
                acc[4 * k + 1] = fhfma((uint16_t)(v.x >> 16), c, acc[4 * k + 1]);
                acc[4 * k + 2] = fhfma((uint16_t)(v.y & 0xffff), c, acc[4 * k + 2]);
                acc[4 * k + 3] = fhfma((uint16_t)(v.y >> 16), c, acc[4 * k + 3]);
            }
        }
    }
}

// sum the four corner partials; afterwards every lane holds the total of its channels
template <int CPL>
__device__ __forceinline__ void reduce_corners(float (&acc)[CPL]) {
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        acc[k] += __shfl_xor_sync(VER_FULL_MASK, acc[k], 8);
        acc[k] += __shfl_xor_sync(VER_FULL_MASK, acc[k], 16);
    }
}

template <int CPL>
__device__ __forceinline__ void store_channels(float* dst, const float (&acc)[CPL]) {
#pragma unroll
    for (int k = 0; k < CPL / 4; ++k)
        reinterpret_cast<float4*>(dst)[k] =
            make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
}
// 16-byte stores; dst must be 16-byte aligned and CPL a multiple of 8
template <int CPL>
__device__ __forceinline__ void store_channels16(__half* dst, const float (&acc)[CPL]) {
#pragma unroll
    for (int k = 0; k < CPL / 8; ++k) {
        uint4 u;
        __half2 h;
        h = __floats2half2_rn(acc[8 * k], acc[8 * k + 1]);
        u.x = *reinterpret_cast<const uint32_t*>(&h);
        h = __floats2half2_rn(acc[8 * k + 2], acc[8 * k + 3]);
        u.y = *reinterpret_cast<const uint32_t*>(&h);
        h = __floats2half2_rn(acc[8 * k + 4], acc[8 * k + 5]);
        u.z = *reinterpret_cast<const uint32_t*>(&h);
        h = __floats2half2_rn(acc[8 * k + 6], acc[8 * k + 7]);
        u.w = *reinterpret_cast<const uint32_t*>(&h);
        reinterpret_cast<uint4*>(dst)[k] = u;
    }
}
template <int CPL>
__device__ __forceinline__ void store_channels(__half* dst, const float (&acc)[CPL]) {
#pragma unroll
    for (int k = 0; k < CPL / 4; ++k) {
        const __half2 a = __floats2half2_rn(acc[4 * k], acc[4 * k + 1]);
        const __half2 b = __floats2half2_rn(acc[4 * k + 2], acc[4 * k + 3]);
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t*>(&a);
        u.y = *reinterpret_cast<const uint32_t*>(&b);
        reinterpret_cast<uint2*>(dst)[k] = u;
    }
}

// Stage one [S][Dh] (bv, head) tile into shared memory with the bulk-copy (TMA)
// engine: one cp.async.bulk per pixel row (row stride NH*Dh in global memory,
// `dst_row_elems` in shared memory), all completing on `bar`.  Called by one warp.
template <typename T>
__device__ __forceinline__ void stage_tile_rows(T* tile, const T* gsrc, int S, int Dh,
                                                size_t src_row_elems, int dst_row_elems,
                                                uint64_t* bar, int lane) {
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)(S * Dh * sizeof(T)));
    __syncwarp();
    for (int s = lane; s < S; s += 32)
        bulk_g2s(tile + (size_t)s * dst_row_elems, gsrc + (size_t)s * src_row_elems,
                 (uint32_t)(Dh * sizeof(T)), bar);
}

// where the [S][Dh] map of (view bv, head h) lives: base + bv*s_bv + h*s_h, rows s_row apart
struct MapLayout {
    size_t s_bv, s_h, s_row;
};
inline MapLayout make_layout(int layout, int S, int NH, int Dh) {
    MapLayout L;
    if (layout == VER_LAYOUT_HEAD_MAJOR) {       // [Bv][NH][S][Dh]
        L.s_bv = (size_t)NH * S * Dh; L.s_h = (size_t)S * Dh; L.s_row = Dh;
    } else {                                     // [Bv][S][NH][Dh]  (mmcv)
        L.s_bv = (size_t)S * NH * Dh; L.s_h = Dh; L.s_row = (size_t)NH * Dh;
    }
    return L;
}

// one bulk copy when the map is contiguous in global memory, else one per pixel row
template <typename T>
__device__ __forceinline__ void stage_map(T* tile, const T* gsrc, int S, int Dh, size_t s_row,
                                          uint64_t* bar, int lane) {
    if (s_row == (size_t)Dh) {
        if (lane == 0) {
            mbar_expect_tx(bar, (uint32_t)(S * Dh * sizeof(T)));
            bulk_g2s(tile, gsrc, (uint32_t)(S * Dh * sizeof(T)), bar);
        }
    } else {
        stage_tile_rows(tile, gsrc, S, Dh, s_row, Dh, bar, lane);
    }
}

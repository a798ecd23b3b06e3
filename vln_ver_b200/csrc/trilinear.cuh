// Trilinear tap of the 3-D (voxel-volume) deformable sampler: the arithmetic of
// F.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=False) on a 5-D input as
// voxel_multi_scale_deformable_attn_pytorch calls it
// (M/voxel_temporal_self_attention.py:275-335: grid = 2*loc - 1, so voxel coordinate =
// loc*size - 0.5; last dim of loc is (x, y, z) <-> (W, H, D)).
//
// Plain C++ on purpose: msda3d.cu uses it on the device, and tests/host_harness/msda3d_host.cpp
// compiles the very same header with g++ so the tap arithmetic is checked against the oracle on
// machines without a GPU.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define VER_HD __host__ __device__ __forceinline__
#define VER_UNROLL _Pragma("unroll")
#else
#define VER_HD inline
#define VER_UNROLL
#endif

struct Tap3 {
    bool any;      // false: the point is outside the zero-padded volume (all 8 corners contribute 0)
    int off[8];    // voxel index (d*H + h)*W + w of corner k (bit0 = +x, bit1 = +y, bit2 = +z), -1 = outside
    float wgt[8];  // trilinear weight of corner k
    float gx[8];   // d wgt / d x_voxel  (sign_x * wy * wz), likewise gy, gz
    float gy[8];
    float gz[8];
};

VER_HD Tap3 make_tap3(float lx, float ly, float lz, int D, int H, int W) {
    Tap3 t;
    const float x = lx * (float)W - 0.5f, y = ly * (float)H - 0.5f, z = lz * (float)D - 0.5f;
    // written so that NaN / inf locations fall out (every comparison is false)
    t.any = (x > -1.f) && (y > -1.f) && (z > -1.f) && (x < (float)W) && (y < (float)H) && (z < (float)D);
    if (!t.any) {
        VER_UNROLL
        for (int k = 0; k < 8; ++k) {
            t.off[k] = -1;
            t.wgt[k] = t.gx[k] = t.gy[k] = t.gz[k] = 0.f;
        }
        return t;
    }
    const float xf = floorf(x), yf = floorf(y), zf = floorf(z);
    const float fx = x - xf, fy = y - yf, fz = z - zf;
    const int x0 = (int)xf, y0 = (int)yf, z0 = (int)zf;
    VER_UNROLL
    for (int k = 0; k < 8; ++k) {
        const int bx = k & 1, by = (k >> 1) & 1, bz = k >> 2;
        const int xi = x0 + bx, yi = y0 + by, zi = z0 + bz;
        const bool ok = xi >= 0 && xi < W && yi >= 0 && yi < H && zi >= 0 && zi < D;
        const float wx = bx ? fx : 1.f - fx, wy = by ? fy : 1.f - fy, wz = bz ? fz : 1.f - fz;
        t.off[k] = ok ? (zi * H + yi) * W + xi : -1;
        t.wgt[k] = wx * wy * wz;
        t.gx[k] = (bx ? wy : -wy) * wz;
        t.gy[k] = (by ? wx : -wx) * wz;
        t.gz[k] = (bz ? wx : -wx) * wy;
    }
    return t;
}

// Bilinear footprint of one sampling point on 16-cell image rows (sca_tc6.cu, sca_tc7.cu): the arithmetic of
// F.grid_sample(bilinear, zeros, align_corners=False) as MSDeformableAttention3D uses it
// (M/multi_scale_deformable_attn_function.py:29-53; pixel coordinate = loc * size - 0.5, corners outside the map
// contribute 0), written for a row layout in which cell X = pixel x + 1 and cells 0 and Sw + 1 are zero padding.
//
//   x: the coordinate is clamped to [0, Sw + 1]; e = 2 floor(X / 2) is the ALIGNED pair base and the footprint is the
//      three cells e, e + 1, e + 2 with weights max(0, 1 - u), 1 - |u - 1|, max(0, u - 1), u = X - e.  (A point
//      outside the map puts its weight on a padding cell.)
//   y: a thread owns the image rows of one parity pi; the row of that parity that carries weight is 2 j + pi with
//      j = floor((Y - pi) / 2), Y = pixel y + 1, clamped to the rows that exist; its weight is the tent
//      max(0, 1 - |Y - 1 - (2 j + pi)|), which is 0 whenever j had to be clamped.
//
// Plain C++ on purpose: the kernels use it on the device (floor through a round-down add of 2^23), and
// tests/host_harness/tap16_host.cpp compiles the very same header with g++ so the arithmetic is checked against the
// oracle's bilinear weights on machines without a GPU (tests/test_tap16_host_math.py).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define TAP16_HD __host__ __device__ __forceinline__
#else
#define TAP16_HD inline
#endif

#define TAP16_MAGIC 8388608.f        // 2^23

struct Tap16 {
    float w0, w1, w2;   // x weights of cells e, e + 1, e + 2
    float wy;           // attention weight x y weight of the owned image row
    float hh;           // 2^23 + e / 2     (low mantissa bits = pair index 0..7)
    float jm;           // 2^23 + j         (low mantissa bits = row index of the owned parity)
};

// 2^23 + floor(v / 2) for |v| < 2^23 (beyond that the callers clamp)
TAP16_HD float tap16_half_floor(float v) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rd(v, 0.5f, TAP16_MAGIC);
#else
    return floorf(v * 0.5f) + TAP16_MAGIC;          // v * 0.5 is exact, the sum is exact below 2^23
#endif
}

// X = pixel x + 1 (unclamped); t = Y - pi with Y = pixel y + 1; a = attention weight;
// xmax = Sw + 1; jtop = 2^23 + (number of image rows of parity pi) - 1
TAP16_HD Tap16 tap16(float X, float t, float a, float xmax, float jtop) {
    Tap16 r;
    const float Xc = fminf(fmaxf(X, 0.f), xmax);
    r.hh = tap16_half_floor(Xc);
    const float v = fmaf(r.hh - TAP16_MAGIC, -2.f, Xc) - 1.f;             // u - 1
    r.w0 = fmaxf(-v, 0.f);
    r.w1 = 1.f - fabsf(v);
    r.w2 = fmaxf(v, 0.f);
    r.jm = fminf(fmaxf(tap16_half_floor(t), TAP16_MAGIC), jtop);
    const float d = fmaf(r.jm - TAP16_MAGIC, -2.f, t) - 1.f;              // Y - 1 - (2 j + pi)
    r.wy = a * fmaxf(1.f - fabsf(d), 0.f);
    return r;
}

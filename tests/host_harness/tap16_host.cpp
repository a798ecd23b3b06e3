// Test-only host harness: builds ONE row of the interpolation matrix A exactly as the two parity threads of
// sca_fwd_tc6_kernel / sca_fwd_tc7_kernel do (vln_ver_b200/csrc/tap16.cuh, compiled here with g++), in fp32 and without
// the fp16 rounding of the cells, into a dense [Sh][16] row.  tests/test_tap16_host_math.py compares it with the bilinear
// weights grid_sample defines.  NOT a product path.
#include "../../vln_ver_b200/csrc/tap16.cuh"

static inline uint32_t bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

// X1[p], Y1[p]: pixel coordinate + 1 of point p; aw[p]: attention weight; row: [Sh * 16] floats, zeroed by the caller.
// returns the mask of image rows that received taps (the kernels' K-chunk mask)
extern "C" unsigned tap16_row(const float* X1, const float* Y1, const float* aw, int np, int Sh, int Sw, float* row) {
    unsigned kmask = 0;
    float sink = 0.f;
    for (int pi = 0; pi < 2; ++pi) {                                   // the two threads of the row
        const float xmax = (float)(Sw + 1);
        const float jtop = TAP16_MAGIC + (float)((Sh - pi + 1) / 2 - 1);
        for (int p = 0; p < np; ++p) {
            const Tap16 tp = tap16(X1[p], Y1[p] - (float)pi, aw[p], xmax, jtop);
            const unsigned c = bits(tp.hh) & 15u, j = bits(tp.jm) & 15u;
            const int y = 2 * (int)j + pi;
            float* r = row + y * 16;
            r[2 * c] += tp.wy * tp.w0;
            r[2 * c + 1] += tp.wy * tp.w1;
            if (c == 7u) sink += tp.wy * tp.w2;                        // the kernels' sink word: always zero
            else r[2 * c + 2] += tp.wy * tp.w2;
            kmask |= 1u << y;
        }
    }
    return sink != 0.f ? 0x80000000u | kmask : kmask;
}

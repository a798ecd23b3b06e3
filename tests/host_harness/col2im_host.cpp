// TEST INFRASTRUCTURE ONLY: the tap/index relation of csrc/col2im.cu (the SAME header, convt_index.cuh) run with
// plain host loops shaped like convt_col2im_kernel / convt_im2col_kernel, fp64 buffers, so the index arithmetic
// is checked against torch's ConvTranspose3d on a machine without a GPU (tests/test_upsample_lattice.py).
#include <stddef.h>

#include "../../vln_ver_b200/csrc/convt_index.cuh"

extern "C" void col2im_host(const double* cols, double* out, int B, int Z, int Hi, int Wi, int s, int C) {
    const int Ho = s * Hi, Wo = s * Wi;
    const long long pos_in = (long long)Z * Hi * Wi, pos_out = (long long)Z * Ho * Wo;
    for (long long row = 0; row < B * pos_out; ++row) {
        const long long b = row / pos_out;
        const int o = (int)(row % pos_out);
        const int ox = o % Wo, oy = (o / Wo) % Ho, oz = o / (Wo * Ho);
        const double* src_b = cols + (size_t)b * pos_in * 75 * C;
        for (int c = 0; c < C; ++c) out[(size_t)row * C + c] = 0.0;
        for (int kz = 0; kz < 3; ++kz) {
            const int iz = convt_src_depth(oz, kz, Z);
            if (iz < 0) continue;
            for (int ky = 0; ky < 5; ++ky) {
                const int iy = convt_src_lateral(oy, ky, s, Hi);
                if (iy < 0) continue;
                for (int kx = 0; kx < 5; ++kx) {
                    const int ix = convt_src_lateral(ox, kx, s, Wi);
                    if (ix < 0) continue;
                    const size_t i = ((size_t)iz * Hi + iy) * Wi + ix;
                    const int k = (kz * 5 + ky) * 5 + kx;
                    for (int c = 0; c < C; ++c) out[(size_t)row * C + c] += src_b[(i * 75 + k) * C + c];
                }
            }
        }
    }
}

extern "C" void im2col_host(const double* gout, double* gcols, int B, int Z, int Hi, int Wi, int s, int C) {
    const int Ho = s * Hi, Wo = s * Wi;
    const long long pos_in = (long long)Z * Hi * Wi, pos_out = (long long)Z * Ho * Wo;
    for (long long row = 0; row < B * pos_in; ++row) {
        const long long b = row / pos_in;
        const int i = (int)(row % pos_in);
        const int ix = i % Wi, iy = (i / Wi) % Hi, iz = i / (Wi * Hi);
        const double* src_b = gout + (size_t)b * pos_out * C;
        double* dst = gcols + (size_t)row * 75 * C;
        for (int kz = 0; kz < 3; ++kz) {
            const int oz = convt_dst_depth(iz, kz, Z);
            for (int ky = 0; ky < 5; ++ky) {
                const int oy = convt_dst_lateral(iy, ky, s, Ho);
                for (int kx = 0; kx < 5; ++kx) {
                    const int ox = convt_dst_lateral(ix, kx, s, Wo);
                    const int k = (kz * 5 + ky) * 5 + kx;
                    const bool ok = oz >= 0 && oy >= 0 && ox >= 0;
                    const size_t o = ok ? ((size_t)oz * Ho + oy) * Wo + ox : 0;
                    for (int c = 0; c < C; ++c) dst[(size_t)k * C + c] = ok ? src_b[o * C + c] : 0.0;
                }
            }
        }
    }
}

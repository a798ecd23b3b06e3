// TEST INFRASTRUCTURE ONLY (never linked into libver_b200.so): runs the tap arithmetic of the 3-D
// sampler kernels -- the SAME header the device code includes, vln_ver_b200/csrc/trilinear.cuh --
// with plain host loops shaped like msda3d_fwd_kernel / msda3d_bwd_kernel (one (b,q,h) "warp" at a
// time, channels innermost), so that the arithmetic is checked against the oracle on a machine
// without a GPU (tests/test_msda3d_host_math.py).  The GPU parity tests remain the proof for the kernels.
#include <stddef.h>
#include <string.h>

#include "../../vln_ver_b200/csrc/trilinear.cuh"

extern "C" void msda3d_host_forward(const float* value, const int* shapes_dhw, int NL, const float* loc,
                                    const float* w, float* out, int Bv, int S, int NH, int Dh, int Nq, int NP) {
    int start[16];
    for (int l = 0, s = 0; l < NL; ++l) {
        start[l] = s;
        s += shapes_dhw[3 * l] * shapes_dhw[3 * l + 1] * shapes_dhw[3 * l + 2];
    }
    const size_t pstride = (size_t)NH * Dh;
    const long long rows = (long long)Bv * Nq * NH;
    for (long long qh = 0; qh < rows; ++qh) {
        const int h = (int)(qh % NH);
        const int bv = (int)((qh / NH) / Nq);
        float* dst = out + (size_t)qh * Dh;
        for (int c = 0; c < Dh; ++c) dst[c] = 0.f;
        for (int l = 0; l < NL; ++l) {
            const int D = shapes_dhw[3 * l], H = shapes_dhw[3 * l + 1], W = shapes_dhw[3 * l + 2];
            const float* vbase = value + (((size_t)bv * S + start[l]) * NH + h) * Dh;
            for (int p = 0; p < NP; ++p) {
                const size_t li = ((size_t)qh * NL + l) * NP + p;
                const Tap3 tap = make_tap3(loc[3 * li], loc[3 * li + 1], loc[3 * li + 2], D, H, W);
                if (!tap.any) continue;
                const float aw = w[li];
                for (int k = 0; k < 8; ++k) {
                    if (tap.off[k] < 0) continue;
                    const float* src = vbase + (size_t)tap.off[k] * pstride;
                    const float cw = aw * tap.wgt[k];
                    for (int c = 0; c < Dh; ++c) dst[c] = fmaf(cw, src[c], dst[c]);
                }
            }
        }
    }
}

extern "C" void msda3d_host_backward(const float* value, const int* shapes_dhw, int NL, const float* loc,
                                     const float* w, const float* gout, float* gvalue, float* gloc, float* gw,
                                     int Bv, int S, int NH, int Dh, int Nq, int NP) {
    int start[16];
    for (int l = 0, s = 0; l < NL; ++l) {
        start[l] = s;
        s += shapes_dhw[3 * l] * shapes_dhw[3 * l + 1] * shapes_dhw[3 * l + 2];
    }
    memset(gvalue, 0, (size_t)Bv * S * NH * Dh * sizeof(float));
    const size_t pstride = (size_t)NH * Dh;
    const long long rows = (long long)Bv * Nq * NH;
    for (long long qh = 0; qh < rows; ++qh) {
        const int h = (int)(qh % NH);
        const int bv = (int)((qh / NH) / Nq);
        const float* go = gout + (size_t)qh * Dh;
        for (int l = 0; l < NL; ++l) {
            const int D = shapes_dhw[3 * l], H = shapes_dhw[3 * l + 1], W = shapes_dhw[3 * l + 2];
            const size_t base = (((size_t)bv * S + start[l]) * NH + h) * Dh;
            for (int p = 0; p < NP; ++p) {
                const size_t li = ((size_t)qh * NL + l) * NP + p;
                const Tap3 tap = make_tap3(loc[3 * li], loc[3 * li + 1], loc[3 * li + 2], D, H, W);
                const float aw = w[li];
                float pw = 0.f, px = 0.f, py = 0.f, pz = 0.f;
                if (tap.any) {
                    for (int k = 0; k < 8; ++k) {
                        if (tap.off[k] < 0) continue;
                        const size_t o = base + (size_t)tap.off[k] * pstride;
                        const float cw = aw * tap.wgt[k];
                        float part = 0.f;
                        for (int c = 0; c < Dh; ++c) {
                            part = fmaf(go[c], value[o + c], part);
                            gvalue[o + c] += cw * go[c];
                        }
                        pw = fmaf(tap.wgt[k], part, pw);
                        px = fmaf(tap.gx[k], part, px);
                        py = fmaf(tap.gy[k], part, py);
                        pz = fmaf(tap.gz[k], part, pz);
                    }
                }
                gw[li] = pw;
                gloc[3 * li] = aw * (float)W * px;
                gloc[3 * li + 1] = aw * (float)H * py;
                gloc[3 * li + 2] = aw * (float)D * pz;
            }
        }
    }
}

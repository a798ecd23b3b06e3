// N1 -- col2im / im2col of the lattice-form transposed convolutions of the occupancy head's up_sample stack
// (HEAD:254-258 applied at HEAD:557-560; derivation in vln_ver_b200/upsample.py).  A layer is
//     cols[b, i, k, :] = e_in[b, i, :] @ W[:, k, :]          (library GEMM, M = positions, N = 75 * C)
//     e_out[b, o, :]   = sum_{(i, k) -> o} cols[b, i, k, :]   (this file, gather form: no atomics)
// Both kernels are pure data movement, channels innermost: one CTA per output row of C channels, every thread
// a 16-byte vector of channels, so each of the <= 75 taps of a position is one fully coalesced row read.
// Bound: HBM (cols is written once by the GEMM and read once here: 75 * C * sizeof(T) bytes per input position).
#include "common.cuh"
#include "convt_index.cuh"

namespace {

constexpr int kColThreads = 128;

template <typename T>
struct Vec16;
template <>
struct Vec16<__half> {
    static constexpr int kElems = 8;
    __device__ static void load(const __half* p, float (&f)[8]) {
        const uint4 v = *reinterpret_cast<const uint4*>(p);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 t = __half22float2(h[j]);
            f[2 * j] = t.x;
            f[2 * j + 1] = t.y;
        }
    }
    __device__ static void store(__half* p, const float (&f)[8]) {
        uint4 v;
        __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
        *reinterpret_cast<uint4*>(p) = v;
    }
};
template <>
struct Vec16<float> {
    static constexpr int kElems = 4;
    __device__ static void load(const float* p, float (&f)[4]) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        f[0] = v.x;
        f[1] = v.y;
        f[2] = v.z;
        f[3] = v.w;
    }
    __device__ static void store(float* p, const float (&f)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    }
};

// out[b, (oz, oy, ox), :] = sum over taps of cols[b, (iz, iy, ix), (kz, ky, kx), :]
template <typename T>
__global__ void __launch_bounds__(kColThreads)
convt_col2im_kernel(const T* __restrict__ cols, T* __restrict__ out, int Z, int Hi, int Wi, int s, int C,
                    long long rows_out) {
    constexpr int E = Vec16<T>::kElems;
    const int Ho = s * Hi, Wo = s * Wi;
    const long long pos_in = (long long)Z * Hi * Wi, pos_out = (long long)Z * Ho * Wo;
    for (long long row = blockIdx.x; row < rows_out; row += gridDim.x) {
        const long long b = row / pos_out;
        const int o = (int)(row % pos_out);
        const int ox = o % Wo, oy = (o / Wo) % Ho, oz = o / (Wo * Ho);
        const T* src_b = cols + (size_t)b * pos_in * 75 * C;
        for (int c0 = threadIdx.x * E; c0 < C; c0 += kColThreads * E) {
            float acc[E];
#pragma unroll
            for (int j = 0; j < E; ++j) acc[j] = 0.f;
            for (int kz = 0; kz < 3; ++kz) {
                const int iz = convt_src_depth(oz, kz, Z);
                if (iz < 0) continue;
                for (int ky = 0; ky < 5; ++ky) {
                    const int iy = convt_src_lateral(oy, ky, s, Hi);
                    if (iy < 0) continue;
#pragma unroll
                    for (int kx = 0; kx < 5; ++kx) {
                        const int ix = convt_src_lateral(ox, kx, s, Wi);
                        if (ix < 0) continue;
                        const size_t i = ((size_t)iz * Hi + iy) * Wi + ix;
                        const int k = (kz * 5 + ky) * 5 + kx;
                        float v[E];
                        Vec16<T>::load(src_b + (i * 75 + k) * C + c0, v);
#pragma unroll
                        for (int j = 0; j < E; ++j) acc[j] += v[j];
                    }
                }
            }
            Vec16<T>::store(out + (size_t)row * C + c0, acc);
        }
    }
}

// adjoint: grad_cols[b, i, k, :] = grad_out[b, o(i, k), :], zero where the tap falls outside
template <typename T>
__global__ void __launch_bounds__(kColThreads)
convt_im2col_kernel(const T* __restrict__ gout, T* __restrict__ gcols, int Z, int Hi, int Wi, int s, int C,
                    long long rows_in) {
    constexpr int E = Vec16<T>::kElems;
    const int Ho = s * Hi, Wo = s * Wi;
    const long long pos_in = (long long)Z * Hi * Wi, pos_out = (long long)Z * Ho * Wo;
    for (long long row = blockIdx.x; row < rows_in; row += gridDim.x) {
        const long long b = row / pos_in;
        const int i = (int)(row % pos_in);
        const int ix = i % Wi, iy = (i / Wi) % Hi, iz = i / (Wi * Hi);
        const T* src_b = gout + (size_t)b * pos_out * C;
        T* dst = gcols + (size_t)row * 75 * C;
        for (int kz = 0; kz < 3; ++kz) {
            const int oz = convt_dst_depth(iz, kz, Z);
            for (int ky = 0; ky < 5; ++ky) {
                const int oy = convt_dst_lateral(iy, ky, s, Ho);
                for (int kx = 0; kx < 5; ++kx) {
                    const int ox = convt_dst_lateral(ix, kx, s, Wo);
                    const int k = (kz * 5 + ky) * 5 + kx;
                    const bool ok = oz >= 0 && oy >= 0 && ox >= 0;
                    const size_t o = ok ? ((size_t)oz * Ho + oy) * Wo + ox : 0;
                    for (int c0 = threadIdx.x * E; c0 < C; c0 += kColThreads * E) {
                        float v[E];
#pragma unroll
                        for (int j = 0; j < E; ++j) v[j] = 0.f;
                        if (ok) Vec16<T>::load(src_b + o * C + c0, v);
                        Vec16<T>::store(dst + (size_t)k * C + c0, v);
                    }
                }
            }
        }
    }
}

int check_col_args(int dtype, const void* a, const void* b, int B, int Z, int Hi, int Wi, int s, int C) {
    VER_CHECK_ARG(dtype == VER_F32 || dtype == VER_F16, "bad dtype %d", dtype);
    VER_CHECK_ARG(a && b, "null pointer");
    VER_CHECK_ARG(B > 0 && Z > 0 && Hi > 0 && Wi > 0 && C > 0, "non-positive dimension");
    VER_CHECK_ARG(s == 1 || s == 2, "lateral stride %d (1 for the first layer, 2 after)", s);
    VER_CHECK_ARG(C % (dtype == VER_F16 ? 8 : 4) == 0, "channels %d must be a multiple of a 16-byte vector", C);
    VER_CHECK_ARG((((uintptr_t)a | (uintptr_t)b) & 15) == 0, "buffers must be 16-byte aligned");
    return VER_OK;
}

int col_grid(long long rows) {
    const long long cap = (long long)ver_device_sm_count() * 16;
    return (int)(rows < cap ? rows : cap);
}

}  // namespace

extern "C" int ver_convt_col2im(int dtype, const void* cols, void* out, int B, int Z, int Hi, int Wi, int s, int C,
                                ver_stream_t stream) {
    int rc = check_col_args(dtype, cols, out, B, Z, Hi, Wi, s, C);
    if (rc) return rc;
    const long long rows = (long long)B * Z * (s * Hi) * (s * Wi);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == VER_F16)
        convt_col2im_kernel<__half><<<col_grid(rows), kColThreads, 0, st>>>((const __half*)cols, (__half*)out, Z, Hi, Wi,
                                                                            s, C, rows);
    else
        convt_col2im_kernel<float><<<col_grid(rows), kColThreads, 0, st>>>((const float*)cols, (float*)out, Z, Hi, Wi, s,
                                                                           C, rows);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

extern "C" int ver_convt_im2col(int dtype, const void* grad_out, void* grad_cols, int B, int Z, int Hi, int Wi, int s,
                                int C, ver_stream_t stream) {
    int rc = check_col_args(dtype, grad_out, grad_cols, B, Z, Hi, Wi, s, C);
    if (rc) return rc;
    const long long rows = (long long)B * Z * Hi * Wi;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == VER_F16)
        convt_im2col_kernel<__half><<<col_grid(rows), kColThreads, 0, st>>>((const __half*)grad_out, (__half*)grad_cols,
                                                                            Z, Hi, Wi, s, C, rows);
    else
        convt_im2col_kernel<float><<<col_grid(rows), kColThreads, 0, st>>>((const float*)grad_out, (float*)grad_cols, Z,
                                                                           Hi, Wi, s, C, rows);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

// A1 + A2 (voxel reference points -> per-camera image coordinates + visibility) and
// K8 (per-camera visible-voxel index lists).  fp32 arithmetic with explicit
// round-to-nearest intrinsics so nvcc cannot contract mul+add into FMA: the
// results are bit-identical to the reference's CPU ops
// (M/voxel_encoder.py:53-83, :136-195).
#include "common.cuh"

namespace {

struct PcRange {
    float mn[3];    // (float)pc_range[0..2]
    float ext[3];   // (float)(pc_range[3+i] - pc_range[i]), subtraction in double like Python
};

// One thread per (b, n): the world point is shared by all cameras.
__global__ void __launch_bounds__(256)
point_sampling_kernel(const float* __restrict__ lidar2img, const float* __restrict__ originshift,
                      PcRange pc, int B, int Ncam, int Z, int H, int W, float img_w, float img_h,
                      float* __restrict__ rpc, uint8_t* __restrict__ mask,
                      uint32_t* __restrict__ vis_bits, int32_t* __restrict__ count) {
    extern __shared__ float s_mat[];   // [Ncam][16] for this b
    const int b = blockIdx.y;
    const int Nq = Z * H * W;
    for (int i = threadIdx.x; i < Ncam * 16; i += blockDim.x)
        s_mat[i] = lidar2img[(size_t)b * Ncam * 16 + i];
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Nq) return;

    const int w = n % W, h = (n / W) % H, z = n / (W * H);
    // linspace(0.5, n-0.5, n) has step exactly 1 -> (i + 0.5) / n   (voxel_encoder.py:70-75)
    const float rx = __fdiv_rn((float)w + 0.5f, (float)W);
    const float ry = __fdiv_rn((float)h + 0.5f, (float)H);
    const float rz = __fdiv_rn((float)z + 0.5f, (float)Z);
    // ref * (max - min) + min + shift, evaluated left to right (voxel_encoder.py:146-151);
    // (max - min) is a Python double rounded to fp32 when it meets the fp32 tensor
    const float sx = originshift[b * 3 + 0], sy = originshift[b * 3 + 1], sz = originshift[b * 3 + 2];
    const float px = __fadd_rn(__fadd_rn(__fmul_rn(rx, pc.ext[0]), pc.mn[0]), sx);
    const float py = __fadd_rn(__fadd_rn(__fmul_rn(ry, pc.ext[1]), pc.mn[1]), sy);
    const float pz = __fadd_rn(__fadd_rn(__fmul_rn(rz, pc.ext[2]), pc.mn[2]), sz);

    uint32_t bits = 0;
    int cnt = 0;
    const float eps = 1e-5f;
    for (int c = 0; c < Ncam; ++c) {
        const float* m = s_mat + c * 16;
        // row . [px py pz 1], sequential k = 0..3, no FMA (== torch CPU matmul here)
        float cx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], px), __fmul_rn(m[1], py)),
                                       __fmul_rn(m[2], pz)), m[3]);
        float cy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[4], px), __fmul_rn(m[5], py)),
                                       __fmul_rn(m[6], pz)), m[7]);
        float cz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[8], px), __fmul_rn(m[9], py)),
                                       __fmul_rn(m[10], pz)), m[11]);
        bool vis = cz > eps;
        const float zc = fmaxf(cz, eps);
        float u = __fdiv_rn(__fdiv_rn(cx, zc), img_w);
        float v = __fdiv_rn(__fdiv_rn(cy, zc), img_h);
        vis = vis && (v > 0.0f) && (v < 1.0f) && (u < 1.0f) && (u > 0.0f);
        const size_t o = ((size_t)c * B + b) * Nq + n;
        reinterpret_cast<float2*>(rpc)[o] = make_float2(u, v);
        mask[o] = vis ? 1 : 0;
        if (vis) {
            ++cnt;
            if (c < 32) bits |= (1u << c);
        }
    }
    if (vis_bits) vis_bits[(size_t)b * Nq + n] = bits;
    if (count) count[(size_t)b * Nq + n] = cnt;
}

// One CTA per (b, cam): ordered stream compaction of the visible voxel ids.
__global__ void __launch_bounds__(1024)
visible_index_kernel(const uint8_t* __restrict__ mask, int B, int Ncam, int Nq,
                     int32_t* __restrict__ counts, int32_t* __restrict__ index) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int cam = blockIdx.x % Ncam, b = blockIdx.x / Ncam;
    const uint8_t* m = mask + ((size_t)cam * B + b) * Nq;
    int32_t* out = index + ((size_t)b * Ncam + cam) * Nq;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int start = 0; start < Nq; start += blockDim.x) {
        const int n = start + threadIdx.x;
        const bool v = (n < Nq) && m[n];
        const unsigned bal = __ballot_sync(VER_FULL_MASK, v);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        int prefix = s_base;
        for (int i = 0; i < wid; ++i) prefix += s_warp[i];
        if (v) out[prefix + __popc(bal & ((1u << lane) - 1))] = n;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int i = 0; i < nw; ++i) t += s_warp[i];
            s_base += t;
        }
        __syncthreads();
    }
    const int total = s_base;
    for (int i = total + threadIdx.x; i < Nq; i += blockDim.x) out[i] = -1;
    if (threadIdx.x == 0) counts[b * Ncam + cam] = total;
}

}  // namespace

extern "C" int ver_point_sampling_f32(const float* lidar2img, const float* originshift,
                                      const double* pc_range, int B, int Ncam, int Z, int H, int W,
                                      float img_w, float img_h, float* rpc, uint8_t* mask,
                                      uint32_t* vis_bits, int32_t* count, ver_stream_t stream) {
    VER_CHECK_ARG(lidar2img && originshift && pc_range && rpc && mask, "null pointer");
    VER_CHECK_ARG(B > 0 && Ncam > 0 && Z > 0 && H > 0 && W > 0, "non-positive dimension");
    VER_CHECK_ARG(vis_bits == nullptr || Ncam <= 32, "vis_bits needs Ncam <= 32 (got %d)", Ncam);
    VER_CHECK_ARG(Ncam * 16 * sizeof(float) <= 48 * 1024, "Ncam too large (%d)", Ncam);
    PcRange pc;
    for (int i = 0; i < 3; ++i) {
        pc.mn[i] = (float)pc_range[i];
        pc.ext[i] = (float)(pc_range[3 + i] - pc_range[i]);
    }
    const int Nq = Z * H * W;
    dim3 grid((Nq + 255) / 256, B);
    point_sampling_kernel<<<grid, 256, Ncam * 16 * sizeof(float), (cudaStream_t)stream>>>(
        lidar2img, originshift, pc, B, Ncam, Z, H, W, img_w, img_h, rpc, mask, vis_bits, count);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

extern "C" int ver_visible_index(const uint8_t* mask, int B, int Ncam, int Nq, int32_t* counts,
                                 int32_t* index, ver_stream_t stream) {
    VER_CHECK_ARG(mask && counts && index, "null pointer");
    VER_CHECK_ARG(B > 0 && Ncam > 0 && Nq > 0, "non-positive dimension");
    visible_index_kernel<<<B * Ncam, 1024, 0, (cudaStream_t)stream>>>(mask, B, Ncam, Nq, counts, index);
    VER_CHECK_LAUNCH();
    g_ver_launches += 1;
    return VER_OK;
}

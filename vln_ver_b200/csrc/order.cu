// Visibility-sorted voxel order for the tensor-core sampler (sca_tc3.cu).
//
// The reference rebatches the visible voxels of every camera and pads to max_len
// (M/spatial_cross_attention.py:138-154).  The tcgen05 sampler instead wants M = 128-row tiles whose
// rows are seen by the SAME cameras, so that (tile, camera) products carry no invisible rows and the sum
// over cameras accumulates in TMEM.  Per panorama b the voxels are therefore sorted (stable) by their camera
// bit set; tiles are 128 consecutive sorted rows, and the union of the bit sets of a tile tells the kernel
// which cameras it has to walk.  The order only affects performance: every row is computed independently.
//
// The sort is CUB's device radix sort (CUDA toolkit header library) on 64-bit keys (b << 32 | bits); it runs
// once per forward (the geometry is shared by the three encoder layers and all heads).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace {

constexpr int kTileRows = 128;

__global__ void order_keys_kernel(const uint32_t* __restrict__ vis_bits, int B, int Nq,
                                  unsigned long long* __restrict__ keys, int32_t* __restrict__ vals) {
    const size_t total = (size_t)B * Nq;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long b = i / Nq;
        keys[i] = (b << 32) | vis_bits[i];
        vals[i] = (int32_t)(i - b * Nq);
    }
}

// one warp per tile: sorted bit sets out, union of the tile's bit sets
__global__ void order_tiles_kernel(const unsigned long long* __restrict__ keys, int B, int Nq, int tiles_per_b,
                                   uint32_t* __restrict__ smask, uint32_t* __restrict__ tile_union) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * tiles_per_b) return;
    const int b = warp / tiles_per_b, t = warp % tiles_per_b;
    uint32_t u = 0;
    for (int r = lane; r < kTileRows; r += 32) {
        const int i = t * kTileRows + r;
        if (i < Nq) {
            const uint32_t m = (uint32_t)keys[(size_t)b * Nq + i];
            smask[(size_t)b * Nq + i] = m;
            u |= m;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) u |= __shfl_xor_sync(VER_FULL_MASK, u, o);
    if (lane == 0) tile_union[warp] = u;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int key_bits(int B) {
    int bits = 32;
    while ((1 << (bits - 32)) < B) ++bits;
    return bits;
}

}  // namespace

extern "C" int ver_visibility_order_workspace(int B, int Nq, size_t* bytes) {
    VER_CHECK_ARG(bytes && B > 0 && Nq > 0, "bad arguments");
    const size_t n = (size_t)B * Nq;
    size_t temp = 0;
    VER_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp, (const unsigned long long*)nullptr,
                                                   (unsigned long long*)nullptr, (const int32_t*)nullptr,
                                                   (int32_t*)nullptr, (int)n, 0, key_bits(B)));
    *bytes = 2 * align256(n * 8) + align256(n * 4) + align256(temp);
    return VER_OK;
}

extern "C" int ver_visibility_order(const uint32_t* vis_bits, int B, int Nq, int32_t* order, uint32_t* smask,
                                    uint32_t* tile_union, void* workspace, size_t workspace_bytes,
                                    ver_stream_t stream) {
    VER_CHECK_ARG(vis_bits && order && smask && tile_union && workspace, "null pointer");
    VER_CHECK_ARG(B > 0 && Nq > 0 && (size_t)B * Nq < (1u << 31), "bad dimensions");
    size_t need = 0;
    int rc = ver_visibility_order_workspace(B, Nq, &need);
    if (rc) return rc;
    VER_CHECK_ARG(workspace_bytes >= need, "workspace too small (%zu < %zu)", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)B * Nq;
    char* ws = (char*)workspace;
    unsigned long long* keys_in = (unsigned long long*)ws;
    unsigned long long* keys_out = (unsigned long long*)(ws + align256(n * 8));
    int32_t* vals_in = (int32_t*)(ws + 2 * align256(n * 8));
    void* temp = ws + 2 * align256(n * 8) + align256(n * 4);
    size_t temp_bytes = workspace_bytes - (2 * align256(n * 8) + align256(n * 4));
    const int blocks = (int)((n + 255) / 256 > 148 * 8 ? 148 * 8 : (n + 255) / 256);
    order_keys_kernel<<<blocks, 256, 0, st>>>(vis_bits, B, Nq, keys_in, vals_in);
    VER_CHECK_LAUNCH();
    VER_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, order, (int)n, 0,
                                                   key_bits(B), st));
    const int tiles_per_b = (Nq + kTileRows - 1) / kTileRows;
    const int warps = B * tiles_per_b;
    order_tiles_kernel<<<(warps * 32 + 255) / 256, 256, 0, st>>>(keys_out, B, Nq, tiles_per_b, smask, tile_union);
    VER_CHECK_LAUNCH();
    g_ver_launches += 3;
    return VER_OK;
}

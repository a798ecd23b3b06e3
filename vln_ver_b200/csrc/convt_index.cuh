// Index relation of the lattice form of the head's transposed convolutions (vln_ver_b200/upsample.py,
// HEAD:254-258): kernel (3,5,5); along H and W  o = s*i - 2 + k (k = 0..4, s = 1 for the first layer, 2 after);
// along Z  o = i - 2 + 2*k (k = 0..2).  Plain C++: col2im.cu uses it on the device and
// tests/host_harness/col2im_host.cpp compiles the same header with g++ (checked against torch on CPU).
#pragma once

#if defined(__CUDACC__)
#define VER_IDX_HD __host__ __device__ __forceinline__
#else
#define VER_IDX_HD inline
#endif

// input index that tap k of a 5-tap lateral kernel connects to output index o, or -1
VER_IDX_HD int convt_src_lateral(int o, int k, int s, int n_in) {
    const int t = o + 2 - k;
    if (t < 0) return -1;
    if (s == 2 && (t & 1)) return -1;
    const int i = (s == 2) ? (t >> 1) : t;
    return i < n_in ? i : -1;
}

// input index that tap k of the 3-tap, dilation-2 depth kernel connects to output index o, or -1
VER_IDX_HD int convt_src_depth(int o, int k, int n_in) {
    const int i = o + 2 - 2 * k;
    return (i >= 0 && i < n_in) ? i : -1;
}

// output index written by input index i through tap k, or -1 (the adjoint view)
VER_IDX_HD int convt_dst_lateral(int i, int k, int s, int n_out) {
    const int o = s * i - 2 + k;
    return (o >= 0 && o < n_out) ? o : -1;
}
VER_IDX_HD int convt_dst_depth(int i, int k, int n_out) {
    const int o = i - 2 + 2 * k;
    return (o >= 0 && o < n_out) ? o : -1;
}

/* libver_b200.so -- C ABI of the B200-native (sm_100a) implementation of VER's
 * 2D->3D volumetric lifting hot path.
 *
 * Conventions
 *   - every pointer is DEVICE memory unless the parameter is documented "host";
 *   - the caller owns every buffer; the library never allocates, frees or
 *     synchronises; all work is enqueued on `stream` (a cudaStream_t);
 *   - return value: 0 = VER_OK, negative = error; ver_last_error() returns a
 *     thread-local message for the last failing call on this thread;
 *   - functions are re-entrant; one call per stream at a time is the caller's
 *     responsibility;
 *   - dtype: VER_F32 (=0) or VER_F16 (=1) selects the storage type of `value`
 *     feature maps and of sampled outputs; sampling locations, attention
 *     weights/logits, camera geometry and all accumulation are always fp32
 *     (the reference force-casts to fp32 at
 *     projects/mmdet3d_plugin/bevformer/modules/multi_scale_deformable_attn_function.py:93
 *     and runs SCA under @force_fp32, .../spatial_cross_attention.py:76).
 *
 * Path shorthands used in the citations below (under the reference root):
 *   M/   = projects/mmdet3d_plugin/bevformer/modules/
 *   HEAD = projects/mmdet3d_plugin/bevformer/dense_heads/voxelformer_occupancy_head.py
 */
#ifndef VER_B200_H_
#define VER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VER_OK 0
#define VER_ERR_INVALID_ARG (-1)
#define VER_ERR_CUDA (-2)
#define VER_ERR_UNSUPPORTED (-3)

#define VER_F32 0
#define VER_F16 1

/* layout of per-view feature maps handed to the fused SCA entry points */
#define VER_LAYOUT_MMCV 0       /* [Bv][S][NH][Dh]  (what mmcv's op takes) */
#define VER_LAYOUT_HEAD_MAJOR 1 /* [Bv][NH][S][Dh]  (one contiguous map per (view, head): single bulk copy) */
#define VER_LAYOUT_TC_IMAGE 2   /* output of ver_value_image_f16: tensor-core operand image, selects the
                                   tcgen05 sampler (fp16 only); with this layout ver_sca_backward writes
                                   grad_value in VER_LAYOUT_MMCV order */

typedef void* ver_stream_t; /* cudaStream_t */

/* ABI version of this header (bumped on any signature change). */
int ver_abi_version(void);
/* Thread-local message of the last failing call ("" if none). */
const char* ver_last_error(void);
/* Number of kernels this library has launched in this process (all threads);
 * bench.py reports the delta over the timed region as `gpu_launches`. */
int64_t ver_launch_count(void);

/* ---------------------------------------------------------------- A1 + A2
 * Replaces VoxelFormerEncoder.get_reference_points(dim='3d') (M/voxel_encoder.py:53-83)
 * and the arithmetic of VoxelFormerEncoder.point_sampling (M/voxel_encoder.py:136-195)
 * for B panoramas at once.  Voxel n = z*(H*W) + h*W + w, reference point
 * ((w+.5)/W, (h+.5)/H, (z+.5)/Z); world = ref*range + pc_min + originshift;
 * cam = lidar2img @ [world,1] with the fp32 summation order ((m0*x+m1*y)+m2*z)+m3
 * and no FMA contraction (bit-exact to torch's CPU matmul); mask = z>1e-5 &
 * 0<u<1 & 0<v<1 with u = x/max(z,1e-5)/img_w, v = y/max(z,1e-5)/img_h.
 *   lidar2img  [B, Ncam, 4, 4] fp32 row-major      originshift [B, 3] fp32
 *   pc_range   host, 6 doubles (x0,y0,z0,x1,y1,z1) -- doubles because the reference subtracts
 *              the Python floats before rounding to fp32 (M/voxel_encoder.py:146-151)
 *   rpc        [Ncam, B, Nq, 1, 2] fp32 (reference_points_cam)     (out)
 *   mask       [Ncam, B, Nq, 1] uint8 0/1 (bev_mask, torch.bool)   (out)
 *   vis_bits   [B, Nq] uint32, bit c = mask[c,b,n]; may be NULL; requires Ncam<=32 (out)
 *   count      [B, Nq] int32 = number of cameras seeing the voxel; may be NULL (out) */
int ver_point_sampling_f32(const float* lidar2img, const float* originshift, const double* pc_range,
                           int B, int Ncam, int Z, int H, int W, float img_w, float img_h,
                           float* rpc, uint8_t* mask, uint32_t* vis_bits, int32_t* count,
                           ver_stream_t stream);

/* ---------------------------------------------------------------- K8 (index tensors)
 * Replaces the per-camera `mask.sum(-1).nonzero()` + Python max(len) of
 * SpatialCrossAttention.forward (M/spatial_cross_attention.py:138-142) without a
 * host sync: for every (b, cam) the ascending list of visible voxel ids.
 *   mask    [Ncam, B, Nq] uint8
 *   counts  [B, Ncam] int32 (out)      index [B, Ncam, Nq] int32, entries >= count are -1 (out) */
int ver_visible_index(const uint8_t* mask, int B, int Ncam, int Nq, int32_t* counts,
                      int32_t* index, ver_stream_t stream);

/* ---------------------------------------------------------------- A5 (operator boundary B2)
 * Replaces mmcv._ext.ms_deform_attn_forward as called at
 * M/multi_scale_deformable_attn_function.py:118-124.
 *   value  [Bv, S, NH, Dh] dtype, S = sum_l h_l*w_l        shapes_hw host int32 [NL,2] = (h,w)
 *   loc    [Bv, Nq, NH, NL, NP, 2] fp32 (x,y) in [0,1]     w [Bv, Nq, NH, NL, NP] fp32
 *   out    [Bv, Nq, NH*Dh] dtype
 * zero padding, align_corners=False (pixel = loc*size - 0.5). */
int ver_msda_forward(int dtype, const void* value, const int32_t* shapes_hw, int NL, const float* loc,
                     const float* w, void* out, int Bv, int S, int NH, int Dh, int Nq, int NP,
                     ver_stream_t stream);

/* Replaces mmcv._ext.ms_deform_attn_backward (M/multi_scale_deformable_attn_function.py:150-160).
 *   grad_out [Bv, Nq, NH*Dh] dtype
 *   grad_value [Bv, S, NH, Dh] fp32, grad_loc like loc, grad_w like w: all three are
 *   OVERWRITTEN (the reference passes zero-filled buffers, :146-148; no pre-zeroing needed). */
int ver_msda_backward(int dtype, const void* value, const int32_t* shapes_hw, int NL, const float* loc,
                      const float* w, const void* grad_out, float* grad_value, float* grad_loc,
                      float* grad_w, int Bv, int S, int NH, int Dh, int Nq, int NP,
                      ver_stream_t stream);

/* ---------------------------------------------------------------- N2 / N3: 3-D sampler
 * Replaces voxel_multi_scale_deformable_attn_pytorch (M/voxel_temporal_self_attention.py:275-335)
 * as called by VoxelCustomMSDeformableAttention.forward (M/voxel_decoder.py:315-316) and
 * VoxelTemporalSelfAttention.forward (M/voxel_temporal_self_attention.py:256).
 *   value  [Bv, S, NH, Dh] dtype, S = sum_l d_l*h_l*w_l, voxel index (d*h_l + y)*w_l + x
 *   shapes_dhw host int32 [NL,3] = (d,h,w)
 *   loc    [Bv, Nq, NH, NL, NP, 3] fp32 (x,y,z) in [0,1]   w [Bv, Nq, NH, NL, NP] fp32
 *   out    [Bv, Nq, NH*Dh] dtype
 * trilinear, zero padding, align_corners=False (voxel coordinate = loc*size - 0.5); Dh <= 256. */
int ver_msda3d_forward(int dtype, const void* value, const int32_t* shapes_dhw, int NL, const float* loc,
                       const float* w, void* out, int Bv, int S, int NH, int Dh, int Nq, int NP,
                       ver_stream_t stream);

/* Gradient of ver_msda3d_forward (the reference differentiates through F.grid_sample).
 *   grad_out [Bv, Nq, NH*Dh] dtype
 *   grad_value [Bv, S, NH, Dh] fp32, grad_loc like loc, grad_w like w: all three are OVERWRITTEN. */
int ver_msda3d_backward(int dtype, const void* value, const int32_t* shapes_dhw, int NL, const float* loc,
                        const float* w, const void* grad_out, float* grad_value, float* grad_loc,
                        float* grad_w, int Bv, int S, int NH, int Dh, int Nq, int NP,
                        ver_stream_t stream);

/* ---------------------------------------------------------------- N1: up_sample stack
 * The three ConvTranspose3d(768, 768, (3,5,5), stride (1,2,2), padding (2,4,4), dilation (2,2,2),
 * output_padding (0,1,1)) of HEAD:254-258 (applied at HEAD:557-560) in lattice form: a layer is a library
 * GEMM  cols[b, i, k, :] = e_in[b, i, :] @ W[:, k, :]  (k = (kz*5 + ky)*5 + kx, 75 taps) followed by
 *   e_out[b, (oz,oy,ox), :] = sum of cols[b, (iz,iy,ix), k, :] over the taps with
 *   oz = iz - 2 + 2 kz,  oy = s iy - 2 + ky,  ox = s ix - 2 + kx      (s = 1 first layer, 2 after)
 * which is ver_convt_col2im (gather form, no atomics).  ver_convt_im2col is its adjoint (backward).
 *   cols / grad_cols  [B, Z*Hi*Wi, 75, C] dtype        out / grad_out  [B, Z*(s Hi)*(s Wi), C] dtype
 * C must be a multiple of 8 (fp16) / 4 (fp32); buffers 16-byte aligned. */
int ver_convt_col2im(int dtype, const void* cols, void* out, int B, int Z, int Hi, int Wi, int s, int C,
                     ver_stream_t stream);
int ver_convt_im2col(int dtype, const void* grad_out, void* grad_cols, int B, int Z, int Hi, int Wi, int s, int C,
                     ver_stream_t stream);

/* ---------------------------------------------------------------- A3 (+) A4 (+) A5 fused
 * The sampling part of SpatialCrossAttention.forward (M/spatial_cross_attention.py:138-173)
 * with MSDeformableAttention3D's softmax / location arithmetic (:340-374) fused in, for
 * num_levels = 1 and num_Z_anchors = 1:
 *   slots[b,n,:] = 1/max(count,1) * sum_{cam visible, ascending} sum_p softmax(aw)[p] *
 *                  bilinear(value[b*Ncam+cam], rpc[cam,b,n] + offs[p]/(Sw,Sh))
 * i.e. the tensor handed to output_proj (:174).  Never materialises the padded rebatch.
 *   value   dtype, already value_proj'ed; value_layout VER_LAYOUT_MMCV [B*Ncam, Sh*Sw, NH, Dh] or
 *           VER_LAYOUT_HEAD_MAJOR [B*Ncam, NH, Sh*Sw, Dh]; 16-byte aligned
 *   logits  [B*Nq, ld] fp32: columns [0, NH*NP*2) = sampling_offsets Linear output
 *           (layout (h, p, xy), :340-341), columns [NH*NP*2, NH*NP*3) = attention_weights
 *           Linear output BEFORE softmax (layout (h, p), :342-343)
 *   rpc, vis_bits from ver_point_sampling_f32          slots [B, Nq, NH*Dh] dtype (out)
 * Requires Ncam <= 32, NP == 8, Dh % 8 == 0 (else VER_ERR_UNSUPPORTED). */
int ver_sca_forward(int dtype, const void* value, int value_layout, const float* logits, int ld_logits,
                    const float* rpc, const uint32_t* vis_bits, void* slots, int B, int Ncam, int Z,
                    int H, int W, int Sh, int Sw, int NH, int Dh, int NP, ver_stream_t stream);

/* Re-lays fp16 value maps out as tcgen05 operand images (one per (view, head)):
 *   value [Bv, S, NH, Dh] fp16 (VER_LAYOUT_MMCV)  ->  vimg [Bv, NH, Dh/8, SP/8, 8, 8] fp16,
 *   SP = S rounded up to 16, padded pixels zero.  Dh % 8 == 0. */
int ver_value_image_f16(const void* value, void* vimg, int Bv, int S, int NH, int Dh, ver_stream_t stream);

/* Visibility-sorted voxel order for the tensor-core sampler (ver_sca_forward_sorted).  The device-side
 * counterpart of the reference's per-camera rebatch (M/spatial_cross_attention.py:138-154: nonzero() per
 * camera + pad to max_len): per panorama, voxels are stably sorted by their camera bit set, so that 128-row
 * tiles are seen by (nearly) the same cameras.
 *   vis_bits   [B, Nq] from ver_point_sampling_f32
 *   order      [B, Nq] int32 (out): voxel index n of sorted row i
 *   smask      [B, Nq] uint32 (out): camera bit set of sorted row i
 *   tile_union [B, ceil(Nq/128)] uint32 (out): OR of the bit sets of rows [128 t, 128 t + 128)
 *   workspace  device scratch of at least ver_visibility_order_workspace() bytes */
int ver_visibility_order_workspace(int B, int Nq, size_t* bytes);
int ver_visibility_order(const uint32_t* vis_bits, int B, int Nq, int32_t* order, uint32_t* smask,
                         uint32_t* tile_union, void* workspace, size_t workspace_bytes, ver_stream_t stream);

/* ver_sca_forward on visibility-sorted rows (fp16 tcgen05 operand images from ver_value_image_f16):
 * same result, slots written at their voxel positions.  Requires Ncam <= 32, NP in {4, 8}, S <= 256,
 * Dh in {32, 64, 96, 128}.
 *   variant    0 = the fastest measured kernel generation that covers the shape (what the product passes:
 *              sca_fwd_tc4_kernel, else sca_fwd_tc3_kernel);
 *              3 / 4 / 5 = force sca_fwd_tc3 / tc4 / tc5_kernel (A/B timing and cross-checks in tests/, tools/;
 *              a forced generation that does not cover the shape falls through to the next older one) */
int ver_sca_forward_sorted(const void* vimg, const float* logits, int ld_logits, const float* rpc,
                           const int32_t* order, const uint32_t* smask, const uint32_t* tile_union,
                           void* slots, int B, int Ncam, int Nq, int Sh, int Sw, int NH, int Dh, int NP,
                           int variant, ver_stream_t stream);

/* The same forward on operand images with 16-cell image rows (sca_fwd_tc6_kernel, csrc/sca_tc6.cu): the interpolation
 * rows are built by two threads per row straight into the shared-memory tensor-core operand.
 *   ver_value_image16_f16: value [Bv, Sh*Sw, NH*Dh] fp16 -> vimg16 [Bv, NH, Dh/8, 2*Sh, 8, 8] fp16, cell k = y*16 + x + 1
 *   of pixel (y, x), the other cells of an image row zero.  Requires Sw <= 14.
 *   ver_sca_forward_sorted16: arguments as ver_sca_forward_sorted with vimg16 in place of vimg; variant 0 = default,
 *   7 = sca_fwd_tc7_kernel (rows in a lane-interleaved scratch, A operand in TMEM), 6 = sca_fwd_tc6_kernel.  Requires
 *   ver_tc6_supported(): Ncam <= 32, 2 <= Sh, Sw <= 14, Dh in {32, 64, 96}, NP in {4, 8}. */
int ver_tc6_supported(int Ncam, int Sh, int Sw, int Dh, int NP);
int ver_value_image16_f16(const void* value, void* vimg16, int Bv, int Sh, int Sw, int NH, int Dh, ver_stream_t stream);
int ver_sca_forward_sorted16(const void* vimg16, const float* logits, int ld_logits, const float* rpc,
                             const int32_t* order, const uint32_t* smask, const uint32_t* tile_union,
                             void* slots, int B, int Ncam, int Nq, int Sh, int Sw, int NH, int Dh, int NP,
                             int variant, ver_stream_t stream);

/* Backward of ver_sca_forward.
 *   grad_slots  [B, Nq, NH*Dh] dtype
 *   grad_value  fp32, same layout as value (overwritten)
 *   grad_logits [B*Nq, ld] fp32, columns [0, NH*NP*3) overwritten
 *   counts/index from ver_visible_index (camera-major hit lists) */
int ver_sca_backward(int dtype, const void* value, int value_layout, const float* logits, int ld_logits,
                     const float* rpc, const uint32_t* vis_bits, const int32_t* counts,
                     const int32_t* index, const void* grad_slots, float* grad_value,
                     float* grad_logits, int B, int Ncam, int Z, int H, int W, int Sh, int Sw,
                     int NH, int Dh, int NP, ver_stream_t stream);

/* ---------------------------------------------------------------- A8 prologue
 * feat + cams_embeds[cam] + level_embeds[0] and the (Ncam,B,S,C) -> (B*Ncam,S,C) reorder of
 * VoxelPerceptionTransformer.get_voxel_features (M/voxel_transformer.py:146-168) followed by
 * SpatialCrossAttention's value.permute(2,0,1,3).reshape (M/spatial_cross_attention.py:158-161).
 *   feats [Ncam, B, S, C] fp32   cams_embeds [Ncam, C] fp32 or NULL   level_embed [C] fp32
 *   out   [B*Ncam, S, C] dtype */
int ver_feat_embed(int dtype, const float* feats, const float* cams_embeds, const float* level_embed,
                   void* out, int Ncam, int B, int S, int C, ver_stream_t stream);

/* ---------------------------------------------------------------- A6 epilogue
 * y = LayerNorm(x + residual) * gamma + beta, eps inside sqrt, over the last dim C
 * (the 'norm' steps of VoxelFormerLayer.forward, M/voxel_encoder.py:435-437, fused with the
 * residual adds of SCA (:176 of spatial_cross_attention.py) and of mmcv FFN).
 *   x, residual (may be NULL), y: [rows, C] dtype; gamma, beta fp32 [C] */
int ver_add_layernorm(int dtype, const void* x, const void* residual, const float* gamma,
                      const float* beta, void* y, int64_t rows, int C, float eps,
                      ver_stream_t stream);

/* Training forms of the same epilogue (dropout p = 0.1 in vocc.py:135 / spatial_cross_attention.py:62):
 *   z = residual + dropout(x, p) ;  y = LayerNorm(z) * gamma + beta
 * with a counter-based Philox dropout mask that backward regenerates from (seed, element index).
 *   seed_epoch (device, may be NULL): one 64-bit word that the kernels ADD to `seed` when they run -- a step
 *   captured in a CUDA graph bumps that word inside the graph and so draws fresh masks on every replay, while
 *   forward and backward of one step still see the same value.
 *   z_out (dtype) and stats (float2 mean, rstd per row) are saved for backward; both may be NULL.
 * Backward: dx = dz * keep/(1-p), dresidual = dz (may be NULL), and per-block partial sums of
 *   dgamma / dbeta and, if dxsum_part != NULL, of the columns of dx (= the bias gradient of the Linear that
 *   produced x): [ver_dropout_add_layernorm_bwd_blocks(rows), C] fp32 each (caller sums over dim 0). */
int ver_dropout_add_layernorm_fwd(int dtype, const void* x, const void* residual, const float* gamma,
                                  const float* beta, void* y, void* z_out, float* stats, int64_t rows,
                                  int C, float eps, float p_drop, uint64_t seed, const uint64_t* seed_epoch,
                                  ver_stream_t stream);
int ver_dropout_add_layernorm_bwd_blocks(int64_t rows);
int ver_dropout_add_layernorm_bwd(int dtype, const void* dy, const void* z, const float* stats,
                                  const float* gamma, void* dx, void* dresidual, float* dgamma_part,
                                  float* dbeta_part, float* dxsum_part, int64_t rows, int C, float p_drop,
                                  uint64_t seed, const uint64_t* seed_epoch, ver_stream_t stream);
/* The same pair with the forward's dropout keep bits handed to the backward instead of regenerated there
 * (one byte per 8 consecutive elements, bit e = element e kept: rows * C / 8 bytes; 1.6 % of one fp16 tensor).
 * keep_bits may be NULL: forward then stores nothing, backward regenerates the mask from (seed, element index). */
int ver_dropout_add_layernorm_fwd_bits(int dtype, const void* x, const void* residual, const float* gamma,
                                       const float* beta, void* y, void* z_out, float* stats, uint8_t* keep_bits,
                                       int64_t rows, int C, float eps, float p_drop, uint64_t seed,
                                       const uint64_t* seed_epoch, ver_stream_t stream);
int ver_dropout_add_layernorm_bwd_bits(int dtype, const void* dy, const void* z, const float* stats,
                                       const float* gamma, const uint8_t* keep_bits, void* dx, void* dresidual,
                                       float* dgamma_part, float* dbeta_part, float* dxsum_part, int64_t rows, int C,
                                       float p_drop, uint64_t seed, const uint64_t* seed_epoch, ver_stream_t stream);
/* ---------------------------------------------------------------- K5: the dense projections
 * nn.Linear as a hand-written tcgen05 GEMM with a fused epilogue (csrc/gemm_tc.cu):
 *     out[M, N] = epilogue(a[M, K] @ w[N, K]^T + bias[N])
 * for value_proj (M/spatial_cross_attention.py:336), sampling_offsets (+) attention_weights (:340-343),
 * output_proj (:174) and the two FFN Linears (M/custom_base_transformer_layer.py:157-158, vocc.py:134-135).
 *   epilogue   0: bias -> fp16;  1: bias -> fp32 (the offset / weight logits);
 *              2: bias + ReLU + dropout(p_drop) -> fp16 (= ver_relu_dropout_fwd applied to the Linear's output,
 *                 same Philox mask for the same seed / seed_epoch: element index = row * N + column)
 *   a, w       fp16 row-major, leading dimensions lda / ldw (elements, multiples of 8), 16-byte aligned
 *   bias       fp32 [N] or NULL;   out: fp16 / fp32 row-major, leading dimension ldo
 *   requires   K % 64 == 0 and N % 128 == 0 or N % 192 == 0 (ver_linear_supported); any M
 * fp32 accumulation in tensor memory; one persistent CTA per SM. */
int ver_linear_supported(int M, int N, int K);
int ver_linear_f16(int epilogue, const void* a, int lda, const void* w, int ldw, const float* bias, void* out, int ldo,
                   int M, int N, int K, float p_drop, uint64_t seed, const uint64_t* seed_epoch, ver_stream_t stream);

/* Backward of the FFN's second Linear fused with the backward of Dropout and ReLU and with the bias gradient of the
 * first Linear (mmcv FFN, M/custom_base_transformer_layer.py:157-158; autograd of Linear -> ReLU -> Dropout -> Linear):
 *     da[M, N] = (dy[M, K] @ w[N, K]^T) * [h > 0] / (1 - p_drop)          (two-CTA tcgen05 GEMM, csrc/gemm_tc.cu)
 *     colsum_part[r, :] = column sums of da over the 128 rows of row block r   (fold with ver_colsum_fold -> bias gradient)
 *   dy      fp16 gradient of the second Linear's output;   w = W2^T as [N = hidden, K = embed] fp16 row-major
 *   h       fp16 saved output of the Dropout (> 0 exactly where the unit was positive and kept), leading dimension ld
 *   da      fp16, leading dimension ld;   colsum_part fp32 [ver_linear_bwd_colsum_rows(M), N]
 *   requires K % 64 == 0, N % 256 == 0. */
int ver_linear_bwd_colsum_rows(int M);
int ver_linear_relu_dropout_bwd_f16(const void* dy, int lddy, const void* w, int ldw, const void* h, void* da, int ld,
                                    float* colsum_part, int M, int N, int K, float p_drop, ver_stream_t stream);

/* FFN inner activation (mmcv FFN: Linear -> ReLU -> Dropout): h = dropout(relu(a)), in place allowed;
 * backward da = dh * [h > 0] / (1 - p).  n % 8 == 0.
 * Column sums (bias gradients) come as partial sums: colsum_part is [ver_colsum_partial_rows(), 8] fp32 and row t
 * holds partial sums of columns 8 * (t % (C / 8)) ... + 7, i.e. colsum = part.view(-1, C / 8, 8).sum(0).view(C);
 * C / 8 must divide ver_colsum_partial_rows().  colsum_part may be NULL for ver_relu_dropout_bwd (C is then unused). */
int ver_relu_dropout_fwd(int dtype, const void* a, void* h, int64_t n, float p_drop, uint64_t seed,
                         const uint64_t* seed_epoch, ver_stream_t stream);
int ver_colsum_partial_rows(void);
int ver_relu_dropout_bwd(int dtype, const void* dh, const void* h, void* da, int64_t n, float p_drop, int C,
                         float* colsum_part, ver_stream_t stream);
/* y[rows, C] (dtype) = x[rows, C] (fp32) and the column sums of x as partial sums (see above): turns the fp32
 * gradients of the sampler's backward into GEMM operands + bias gradients in one pass. */
int ver_cast_colsum(int dtype, const float* x, void* y, int64_t rows, int C, float* colsum_part,
                    ver_stream_t stream);
/* out[C] = sum over the P rows of part[P, C] (fp32): folds the partial sums above (view the per-thread 8-float groups
 * as rows of C floats) in one deterministic two-level launch.  scratch: ver_colsum_fold_scratch_floats(C) floats whose
 * LAST word is a counter that must be zero on entry (the kernel leaves it zero); one launch per scratch at a time. */
int ver_colsum_fold_scratch_floats(int C);
int ver_colsum_fold(const float* part, int64_t P, int C, float* out, float* scratch, ver_stream_t stream);
/* the same for n_mat matrices at once: part [n_mat, P, C] -> out [n_mat, C] (one launch; scratch-free) */
int ver_colsum_fold_batched(const float* part, int n_mat, int64_t P, int C, float* out, ver_stream_t stream);
/* Column sums of an fp16 matrix x[rows, C] as partial sums (layout as above): the bias gradient of a Linear whose
 * output gradient is already fp16 (the occupancy head's occ_proj / occ_branches, HEAD:236-248). */
int ver_colsum_f16(const void* x, int64_t rows, int C, float* colsum_part, ver_stream_t stream);

/* ---------------------------------------------------------------- A11
 * Sigmoid focal loss of mmdet FocalLoss(use_sigmoid=True) (vocc.py:190-195; calls HEAD:981,
 * HEAD:1425) with the dense target built on the fly from the sparse GT
 * (HEAD:1326-1330): target[n] = classes ("empty") unless listed in occ_gt.
 *   logits [N, Ccls] fp32      dense_gt [N] int32 scratch (out; filled by this call)
 *   occ_gt [n_gt, 2] int64 (flat index, class); n_gt < 0 means "dense_gt is an INPUT holding
 *          ready-made class targets in [0, Ccls]" (the generic FocalLoss.forward(pred, target) call)
 *   loss_sum fp32[1] = sum of element losses (out, overwritten); num_pos int32[1] (out)
 *   grad_logits [N, Ccls] fp32 = d(loss_sum)/d(logits) or NULL */
int ver_focal_loss(const float* logits, const int64_t* occ_gt, int n_gt, int32_t* dense_gt,
                   float* loss_sum, int32_t* num_pos, float* grad_logits, int64_t N, int Ccls,
                   float gamma, float alpha, ver_stream_t stream);

/* ---------------------------------------------------------------- A12
 * get_occupancy_prediction (HEAD:1505-1524): class = argmax([sigmoid(logits), thr]);
 * rows with class < Ccls are emitted in ascending row order.
 *   logits [N, Ccls] fp32     out_pairs [N, 2] int64 capacity (index, class) (out)
 *   out_count int32[1] (out)  scratch: int32 [ceil(N/1024) + 1] */
int ver_occupancy_decode(const float* logits, int64_t N, int Ccls, float threshold,
                         int64_t* out_pairs, int32_t* out_count, int32_t* scratch,
                         ver_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VER_B200_H_ */

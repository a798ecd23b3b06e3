"""One VoxelFormerLayer ('cross_attn', 'norm', 'ffn', 'norm'; vocc.py:136-137) as ONE autograd node with a
hand-written backward, for fp16 storage on CUDA.

Same arithmetic as the module-by-module path (M/voxel_encoder.py:344-464 -> M/spatial_cross_attention.py:76-176,
mmcv FFN, LayerNorm); what changes is the plumbing around the sm_100a kernels and the library GEMMs:
  * fp16 copies of the fp32 master weights are made once per optimizer step (cache keyed on the parameter
    version), not once per Linear call;
  * the offset / attention-weight logits come out of ONE fp16 x fp16 -> fp32 GEMM with the fp32 bias as the
    beta term (no fp16 round trip, no separate cast and bias passes over the (B Nq, 192) tensor);
  * bias gradients are column sums produced by the kernels that already stream the gradient
    (ver_dropout_add_layernorm_bwd's dx sums, ver_relu_dropout_bwd, ver_cast_colsum) instead of separate
    reduction passes; weight gradients are fp16 x fp16 -> fp32 GEMMs written straight in the master precision;
  * residual-branch gradients are accumulated by the GEMM that produces the other branch (beta = 1) instead
    of a separate add.
torch is used for memory, streams and the plain library GEMMs; there is no CPU path.
"""
import torch
from torch.autograd.function import Function, once_differentiable

from . import ops
from ._lib import VER_F16, check, lib
from .ops import _ptr, _stream

# fp16 / concatenated copies of the fp32 master weights, made once per weight update instead of once per Linear
# call.  An entry is valid while (a) every source parameter is still alive (weak references: a rebuilt model
# cannot alias an id, and dropping the model frees the copies), (b) its autograd version counter and storage are
# unchanged, and (c) the global generation is unchanged.  Updates that bypass the version counter (`p.data.copy_`,
# apex / multi-tensor optimizers writing through .data, load_state_dict on some paths, manual surgery) MUST be
# followed by invalidate_weight_cache(); optimizer steps through torch.optim and load_state_dict are hooked by
# install_cache_hooks().  Inside a CUDA-graph capture a cache hit would freeze stale copies into the graph, so
# graph.capture_step() invalidates first.
_HALF_CACHE = {}
_GENERATION = [0]

# Which Linear layers of the layer's FORWARD run on the hand-written tcgen05 GEMM (csrc/gemm_tc.cu) instead of the
# library GEMM.  Measured on B200 at the benchmark shape (profiles/r02f_gemm_check.txt): the offset / weight logits
# (N = 192, fp32 out) 95 us against 194 us, FFN1 with bias + ReLU + dropout fused 492 us against 644 us for library
# GEMM + ver_relu_dropout_fwd; the three plain N = 768 projections are still 0.78-0.86x of the library (a single-CTA
# 128 x 256 tile is shared-memory-bandwidth bound at ~70 % of the tensor peak), so they stay on the library.
TC_GEMM = {'logits': True, 'ffn1': True, 'value_proj': False, 'output_proj': False, 'ffn2': False,
           'ffn2_bwd': True}      # backward of FFN2 fused with ReLU / dropout backward and the FFN1 bias gradient


def invalidate_weight_cache():
    """Drop every cached low-precision weight copy (call after any weight update that does not go through
    torch.optim / load_state_dict, e.g. writes through `.data`)."""
    _GENERATION[0] += 1
    _HALF_CACHE.clear()


def install_cache_hooks(module, optimizer=None):
    """Invalidate the copies after optimizer.step() and load_state_dict() of `module`."""
    handles = []
    if optimizer is not None and hasattr(optimizer, 'register_step_post_hook'):
        handles.append(optimizer.register_step_post_hook(lambda *a, **k: invalidate_weight_cache()))
    if hasattr(module, 'register_load_state_dict_post_hook'):
        handles.append(module.register_load_state_dict_post_hook(lambda *a, **k: invalidate_weight_cache()))
    return handles


def _cached(kind, params, make):
    import weakref
    key = (kind,) + tuple(id(p) for p in params)
    ver = (_GENERATION[0],) + tuple((p._version, p.data_ptr()) for p in params)
    hit = _HALF_CACHE.get(key)
    if hit is not None and hit[0] == ver and all(r() is p for r, p in zip(hit[2], params)):
        return hit[1]
    with torch.no_grad():
        t = make()
    refs = tuple(weakref.ref(p, lambda _r, k=key: _HALF_CACHE.pop(k, None)) for p in params)
    _HALF_CACHE[key] = (ver, t, refs)
    return t


# fp16 copies of SINGLE parameters are refreshed together: the first stale lookup after a weight update casts every
# registered copy that is out of date with one multi-tensor kernel (a training step used to launch ~35 separate 5 us
# cast kernels, inside the captured graph as well).  Entries: id(param) -> [weak reference, fp16 copy, version]; the
# validity rule is _cached()'s.
_SINGLES = {}


def _refresh_singles():
    dsts, srcs, done = [], [], []
    for key, ent in list(_SINGLES.items()):
        p = ent[0]()
        if p is None:
            _SINGLES.pop(key, None)
            continue
        ver = (_GENERATION[0], p._version, p.data_ptr())
        if ent[2] != ver:
            # a NEW tensor per refresh: a copy that an earlier forward saved for its backward is never overwritten
            ent[1] = torch.empty(p.shape, dtype=torch.float16, device=p.device)
            dsts.append(ent[1])
            srcs.append(p.detach())
            done.append((ent, ver))
    if dsts:
        with torch.no_grad():
            torch._foreach_copy_(dsts, srcs)
        for ent, ver in done:
            ent[2] = ver


def half_of(*params):
    """fp16 copy of a parameter (or the row-concatenation of several), refreshed when a parameter changes."""
    if len(params) == 1:
        import weakref
        p = params[0]
        ent = _SINGLES.get(id(p))
        if ent is None or ent[0]() is not p:
            ent = _SINGLES[id(p)] = [weakref.ref(p, lambda _r, k=id(p): _SINGLES.pop(k, None)), None, None]
        if ent[2] != (_GENERATION[0], p._version, p.data_ptr()):
            _refresh_singles()
        return ent[1]
    return _cached('f16', params, lambda: torch.cat([p.detach() for p in params], 0).to(torch.float16))


def half_t_of(param):
    """fp16 TRANSPOSED copy ([in, out] of an [out, in] Linear weight): the K-major W^T operand of a dX GEMM."""
    return _cached('f16t', (param,), lambda: param.detach().to(torch.float16).t().contiguous())


def f32_cat(*params):
    return _cached('f32', params, lambda: torch.cat([p.detach().float() for p in params], 0).contiguous())


# ------------------------------------------------------------------ thin kernel wrappers (no autograd)
def _ln_fwd(x, res, gamma32, beta32, p, eps, seed, save):
    """-> y, z (the LayerNorm input), stats (mean, rstd per row), keep bits of the dropout mask (one byte per 8
    elements; None without dropout): the last three are what backward needs."""
    rows, C = x.shape
    y = torch.empty_like(x)
    z = torch.empty_like(x) if save else None
    stats = torch.empty((rows, 2), dtype=torch.float32, device=x.device) if save else None
    bits = torch.empty(rows * C // 8, dtype=torch.uint8, device=x.device) if (save and p > 0) else None
    check(lib.ver_dropout_add_layernorm_fwd_bits(VER_F16, _ptr(x), _ptr(res), _ptr(gamma32), _ptr(beta32), _ptr(y),
                                                 _ptr(z), _ptr(stats), _ptr(bits), rows, C, float(eps), float(p), seed,
                                                 _ptr(ops._seed_epoch(x.device)), _stream()))
    return y, z, stats, bits


def _ln_bwd(dy, z, stats, gamma32, p, seed, bits=None):
    """-> dx (gradient of the dropout input), dres (gradient of the residual input), dgamma, dbeta, colsum(dx).
    `bits`: the forward's keep bits; None regenerates the mask from (seed, element index)."""
    rows, C = z.shape
    dx, dres = torch.empty_like(z), torch.empty_like(z)
    nb = lib.ver_dropout_add_layernorm_bwd_blocks(rows)
    part = torch.empty((3, nb, C), dtype=torch.float32, device=z.device)
    check(lib.ver_dropout_add_layernorm_bwd_bits(VER_F16, _ptr(dy), _ptr(z), _ptr(stats), _ptr(gamma32), _ptr(bits),
                                                 _ptr(dx), _ptr(dres), _ptr(part[0]), _ptr(part[1]), _ptr(part[2]),
                                                 rows, C, float(p), seed, _ptr(ops._seed_epoch(z.device)), _stream()))
    folded = _fold_mats(part)                       # dgamma, dbeta, colsum(dx) in one launch
    return dx, dres, folded[0], folded[1], folded[2]


def _colsum_part(device):
    return torch.empty((lib.ver_colsum_partial_rows(), 8), dtype=torch.float32, device=device)


def _fold_rows(part2d):
    """(P, C) fp32 partial sums -> (C,): ver_colsum_fold_batched (one deterministic launch)."""
    return _fold_mats(part2d[None])[0]


def _fold_mats(part3d):
    """(m, P, C) fp32 partial sums -> (m, C) in one launch."""
    m, P, C = part3d.shape
    assert part3d.is_contiguous()
    out = torch.empty((m, C), dtype=torch.float32, device=part3d.device)
    check(lib.ver_colsum_fold_batched(_ptr(part3d), m, P, C, _ptr(out), _stream()))
    return out


def _fold(part, C):
    """per-thread partials (n_threads, 8), thread t holding column group t % (C / 8) -> (C,)"""
    return _fold_rows(part.view(-1, C))


def _relu_dropout_bwd_(dh, h, p):
    """in place on dh -> da; returns colsum(da)."""
    C = h.shape[-1]
    part = _colsum_part(h.device)
    check(lib.ver_relu_dropout_bwd(VER_F16, _ptr(dh), _ptr(h), _ptr(dh), h.numel(), float(p), C, _ptr(part),
                                   _stream()))
    return _fold(part, C)


def _cast_colsum(x32):
    rows, C = x32.shape
    y = torch.empty((rows, C), dtype=torch.float16, device=x32.device)
    part = _colsum_part(x32.device)
    check(lib.ver_cast_colsum(VER_F16, _ptr(x32), _ptr(y), rows, C, _ptr(part), _stream()))
    return y, _fold(part, C)


class LinearF16Function(Function):
    """y = x @ W^T + b with fp16 storage for the occupancy head's Linear layers (HEAD:236-248): cached fp16 weight
    copies, and a hand-written backward whose bias gradient is a column-sum kernel over the fp16 output gradient
    (torch's fp16 sum reduction took 0.13 ms per Linear at 204 800 rows, 0.58 ms per step) and whose weight gradient
    is written straight in fp32."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        """x (rows, K) -> (rows, N): 2-D in, 2-D out (a view made inside a Function may not be modified in place by
        the caller; the FFN's ReLU is in place)."""
        w16 = half_of(weight)
        b16 = half_of(bias) if bias is not None else None
        x2 = x if x.dtype == torch.float16 else x.to(torch.float16)
        x2 = x2.contiguous()
        y = torch.addmm(b16, x2, w16.t()) if b16 is not None else torch.mm(x2, w16.t())
        ctx.save_for_backward(x2, w16)
        ctx.meta = (x.shape, x.dtype, weight.dtype, bias.dtype if bias is not None else None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, w16 = ctx.saved_tensors
        xshape, xdtype, wdtype, bdtype = ctx.meta
        dy2 = dy.reshape(-1, dy.shape[-1])
        if dy2.dtype != torch.float16:
            dy2 = dy2.to(torch.float16)
        dy2 = dy2.contiguous()
        N = dy2.shape[1]
        dx = torch.mm(dy2, w16).view(xshape).to(xdtype) if ctx.needs_input_grad[0] else None
        dw = torch.mm(dy2.t(), x2, out_dtype=torch.float32).to(wdtype) if ctx.needs_input_grad[1] else None
        db = None
        if bdtype is not None and ctx.needs_input_grad[2]:
            n_part = lib.ver_colsum_partial_rows()
            if N % 8 == 0 and n_part % (N // 8) == 0:
                part = _colsum_part(dy2.device)
                check(lib.ver_colsum_f16(_ptr(dy2), dy2.shape[0], N, _ptr(part), _stream()))
                db = _fold(part, N).to(bdtype)
            else:
                db = dy2.float().sum(0).to(bdtype)
        return dx, dw, db


def linear_f16(x, layer):
    y = LinearF16Function.apply(x.reshape(-1, x.shape[-1]), layer.weight, layer.bias)
    return y.view(*x.shape[:-1], layer.weight.shape[0])


def supported(q, feat, vis, NH, NP, S, F):
    """F: hidden width of the FFN.  The column-sum kernels need C / 8, F / 8 and NH * NP * 3 / 8 to divide their
    thread count (ver_colsum_partial_rows)."""
    C = q.shape[-1]
    n_part = lib.ver_colsum_partial_rows()
    L = NH * NP * 3
    return (q.is_cuda and q.dtype == torch.float16 and feat.dtype == torch.float16 and vis is not None
            and vis.bits is not None and C % 8 == 0 and C <= 1024 and NP % 4 == 0 and L % 8 == 0 and F % 8 == 0
            and ops.tc_supported(torch.float16, vis.rpc.shape[0], S, C // NH, NP)
            and n_part % (C // 8) == 0 and n_part % (F // 8) == 0 and n_part % (L // 8) == 0)


class VoxelLayerFunction(Function):
    """y2 = layer(q): q (B*Nq, C) fp16, feat (Bv*S, C) fp16 (view tokens + camera / level embeddings)."""

    @staticmethod
    def forward(ctx, q, feat, vis, cfg, Wv, bv, Wso, bso, Waw, baw, Wo, bo, g1, be1, W1, b1, W2, b2, g2, be2):
        NH, NP, Sh, Sw = cfg['NH'], cfg['NP'], cfg['Sh'], cfg['Sw']
        training = cfg['training']
        p_attn = cfg['p_attn'] if training else 0.0
        p_ffn = cfg['p_ffn'] if training else 0.0
        p_out = cfg['p_out'] if training else 0.0
        eps1, eps2 = cfg['eps1'], cfg['eps2']
        Ncam, B = vis.rpc.shape[:2]
        Z, H, W = vis.grid
        Nq, S = Z * H * W, Sh * Sw
        C = q.shape[1]
        Dh = C // NH
        Bv = B * Ncam
        assert q.shape[0] == B * Nq and feat.shape == (Bv * S, C)
        q = q.contiguous()
        feat = feat.contiguous()
        need_bwd = any(ctx.needs_input_grad)

        Wv16, Wo16, W116, W216 = half_of(Wv), half_of(Wo), half_of(W1), half_of(W2)
        Wcat16 = half_of(Wso, Waw)
        bcat32 = f32_cat(bso, baw)
        # value_proj (M/spatial_cross_attention.py:336) and its tcgen05 operand image
        with ops.nvtx_range('layer.value_proj+logits'):
            if TC_GEMM['value_proj'] and ops.linear_tc_supported(feat, Wv16):
                v = ops.linear_tc(feat, Wv16, bv, ops.LINEAR_BIAS_F16)
            else:
                v = torch.addmm(half_of(bv), feat, Wv16.t())
            vimg = ops.value_image(v.view(Bv, S, C), NH)
            vimg16 = (ops.value_image16(v.view(Bv, S, C), NH, Sh, Sw)
                      if ops.TC_FORWARD in ops._SORTED16_VARIANT and lib.ver_tc6_supported(Ncam, Sh, Sw, Dh, NP) else None)
            del v
            # sampling_offsets (+) attention_weights once per voxel (:340-343), fp32 out
            if TC_GEMM['logits'] and ops.linear_tc_supported(q, Wcat16):
                logits = ops.linear_tc(q, Wcat16, bcat32, ops.LINEAR_BIAS_F32)
            else:
                logits = torch.addmm(bcat32, q, Wcat16.t(), out_dtype=torch.float32)
        ops.nvtx_push('layer.sampler_fwd')
        slots = torch.empty((B * Nq, C), dtype=torch.float16, device=q.device)
        order, smask, tile_union = vis.order
        prof = ops.PROFILE_EVENTS
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        v16 = ops._SORTED16_VARIANT.get(ops.TC_FORWARD)
        if v16 is not None and lib.ver_tc6_supported(Ncam, Sh, Sw, Dh, NP):
            # the generations on 16-cell image rows need their own value image (built above, outside the timed span)
            check(lib.ver_sca_forward_sorted16(_ptr(vimg16), _ptr(logits), logits.shape[1], _ptr(vis.rpc), _ptr(order),
                                               _ptr(smask), _ptr(tile_union), _ptr(slots), B, Ncam, Nq, Sh, Sw, NH, Dh,
                                               NP, v16, _stream()))
        else:
            check(lib.ver_sca_forward_sorted(_ptr(vimg), _ptr(logits), logits.shape[1], _ptr(vis.rpc), _ptr(order),
                                             _ptr(smask), _ptr(tile_union), _ptr(slots), B, Ncam, Nq, Sh, Sw, NH, Dh, NP,
                                             ops._SORTED_VARIANT.get(ops.TC_FORWARD, 0), _stream()))
        if prof is not None:
            e1.record()
            prof.append((e0, e1))
        ops.nvtx_pop()
        # output_proj, dropout + residual + LayerNorm (:174-176, 'norm')
        ops.nvtx_push('layer.output_proj+norm')
        if TC_GEMM['output_proj'] and ops.linear_tc_supported(slots, Wo16):
            proj = ops.linear_tc(slots, Wo16, bo, ops.LINEAR_BIAS_F16)
        else:
            proj = torch.addmm(half_of(bo), slots, Wo16.t())
        seed1, seed2, seed3 = ops._next_seed(), ops._next_seed(), ops._next_seed()
        g1f, be1f, g2f, be2f = (t.detach().float().contiguous() for t in (g1, be1, g2, be2))
        y1, z1, st1, kb1 = _ln_fwd(proj, q, g1f, be1f, p_attn, eps1, seed1, need_bwd)
        del proj
        ops.nvtx_pop()
        # FFN: Linear -> ReLU -> Dropout -> Linear -> Dropout, + identity, LayerNorm
        ops.nvtx_push('layer.ffn+norm')
        if TC_GEMM['ffn1'] and ops.linear_tc_supported(y1, W116):
            # Linear + bias + ReLU + dropout in the GEMM epilogue: no separate pass over the (rows, 1536) tensor
            h = ops.linear_tc(y1, W116, b1, ops.LINEAR_BIAS_RELU_DROPOUT_F16, p=float(p_ffn), seed=seed2)
        else:
            h = torch.addmm(half_of(b1), y1, W116.t())
            check(lib.ver_relu_dropout_fwd(VER_F16, _ptr(h), _ptr(h), h.numel(), float(p_ffn), seed2,
                                           _ptr(ops._seed_epoch(h.device)), _stream()))
        if TC_GEMM['ffn2'] and ops.linear_tc_supported(h, W216):
            f = ops.linear_tc(h, W216, b2, ops.LINEAR_BIAS_F16)
        else:
            f = torch.addmm(half_of(b2), h, W216.t())
        y2, z2, st2, kb2 = _ln_fwd(f, y1, g2f, be2f, p_out, eps2, seed3, need_bwd)
        del f
        ops.nvtx_pop()
        if need_bwd:
            ctx.save_for_backward(q, feat, vimg, logits, slots, z1, st1, y1, h, z2, st2, Wv16, Wcat16, Wo16, W116,
                                  W216, g1f, g2f, *(t for t in (kb1, kb2) if t is not None))
            ctx.has_bits = (kb1 is not None, kb2 is not None)
            ctx.vis, ctx.dims = vis, (B, Ncam, Z, H, W, Sh, Sw, NH, Dh, NP)
            ctx.drop = (p_attn, seed1, p_ffn, p_out, seed3)
            ctx.n_so = Wso.shape[0]
            ctx.W2 = W2
            ctx.pdtypes = [t.dtype for t in (Wv, bv, Wso, bso, Waw, baw, Wo, bo, g1, be1, W1, b1, W2, b2, g2, be2)]
        return y2

    @staticmethod
    @once_differentiable
    def backward(ctx, dy2):
        (q, feat, vimg, logits, slots, z1, st1, y1, h, z2, st2, Wv16, Wcat16, Wo16, W116, W216, g1f,
         g2f) = ctx.saved_tensors[:18]
        extra = list(ctx.saved_tensors[18:])
        kb1 = extra.pop(0) if ctx.has_bits[0] else None           # dropout keep bits of the two LayerNorm inputs
        kb2 = extra.pop(0) if ctx.has_bits[1] else None
        vis = ctx.vis
        B, Ncam, Z, H, W, Sh, Sw, NH, Dh, NP = ctx.dims
        p_attn, seed1, p_ffn, p_out, seed3 = ctx.drop
        f32 = torch.float32
        dy2 = dy2.contiguous()
        if dy2.dtype != torch.float16:
            dy2 = dy2.to(torch.float16)
        # ---- norm 2 / FFN
        ops.nvtx_push('layer.bwd.ffn+norm')
        df, dy1, dg2, dbe2, db2 = _ln_bwd(dy2, z2, st2, g2f, p_out, seed3, kb2)
        dW2 = torch.mm(df.t(), h, out_dtype=f32)
        if (TC_GEMM['ffn2_bwd'] and ctx.W2 is not None and h.shape[1] % 256 == 0 and df.shape[1] % 64 == 0):
            # dX of FFN2, ReLU / dropout backward and the column sums (= FFN1 bias gradient) in one tcgen05 GEMM
            dh, part = ops.linear_relu_dropout_bwd(df, half_t_of(ctx.W2), h, p_ffn)
            db1 = _fold_rows(part)
        else:
            dh = torch.mm(df, W216)
            db1 = _relu_dropout_bwd_(dh, h, p_ffn)          # dh -> da in place
        del df
        dW1 = torch.mm(dh.t(), y1, out_dtype=f32)
        dy1.addmm_(dh, W116)                                # + residual branch, accumulated by the GEMM
        del dh
        ops.nvtx_pop()
        # ---- norm 1 / output_proj
        ops.nvtx_push('layer.bwd.output_proj+norm')
        dproj, dq, dg1, dbe1, dbo = _ln_bwd(dy1, z1, st1, g1f, p_attn, seed1, kb1)
        del dy1
        dWo = torch.mm(dproj.t(), slots, out_dtype=f32)
        dslots = torch.mm(dproj, Wo16)
        del dproj
        ops.nvtx_pop()
        # ---- fused sampler backward (A5 backward + SCA scatter)
        ops.nvtx_push('layer.bwd.sampler')
        counts, index = vis.index
        Bv, S, C = B * Ncam, Sh * Sw, NH * Dh
        gvalue = torch.empty((Bv * S, C), dtype=f32, device=q.device)
        glogits = torch.empty(logits.shape, dtype=f32, device=q.device)
        prof = ops.PROFILE_EVENTS_BWD
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        check(lib.ver_sca_backward(VER_F16, _ptr(vimg), ops.VER_LAYOUT_TC_IMAGE, _ptr(logits), logits.shape[1],
                                   _ptr(vis.rpc), _ptr(vis.bits), _ptr(counts), _ptr(index), _ptr(dslots),
                                   _ptr(gvalue), _ptr(glogits), B, Ncam, Z, H, W, Sh, Sw, NH, Dh, NP, _stream()))
        if prof is not None:
            e1.record()
            prof.append((e0, e1))
        del dslots
        ops.nvtx_pop()
        # ---- logits Linear (once per voxel): dW, db, and the query gradient joins the residual branch's
        ops.nvtx_push('layer.bwd.logits+value_proj')
        gl16, dbcat = _cast_colsum(glogits)
        del glogits
        dWcat = torch.mm(gl16.t(), q, out_dtype=f32)
        dq.addmm_(gl16, Wcat16)
        del gl16
        # ---- value_proj
        gv16, dbv = _cast_colsum(gvalue)
        del gvalue
        dWv = torch.mm(gv16.t(), feat, out_dtype=f32)
        dfeat = torch.mm(gv16, Wv16) if ctx.needs_input_grad[1] else None
        ops.nvtx_pop()
        n_so = ctx.n_so
        grads = [dWv, dbv, dWcat[:n_so], dbcat[:n_so], dWcat[n_so:], dbcat[n_so:], dWo, dbo, dg1, dbe1, dW1, db1,
                 dW2, db2, dg2, dbe2]
        grads = [g if g.dtype == dt else g.to(dt) for g, dt in zip(grads, ctx.pdtypes)]
        return (dq if ctx.needs_input_grad[0] else None, dfeat, None, None, *grads)


def voxel_layer(q, feat, vis, cfg, params):
    return VoxelLayerFunction.apply(q, feat, vis, cfg, *params)

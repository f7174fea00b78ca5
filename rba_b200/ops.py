"""Thin torch-tensor wrappers over the C ABI.  torch supplies device memory and the current stream only;
all compute happens in librba_b200.so.  Every function raises RbaError on failure — nothing falls back."""
import ctypes

import torch

from . import _lib
from ._lib import RBA_ACT_GELU, RBA_ACT_NONE, RBA_ACT_RELU, RBA_GEMM_FFMA, RBA_GEMM_TC, RbaError, RbaGemmArgs  # noqa: F401


_last_dev = [None]


def _stream():
    """Current stream of the device the last checked tensors live on (not of the thread's current device)."""
    return ctypes.c_void_p(torch.cuda.current_stream(_last_dev[0]).cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _chk_cuda(*ts):
    for t in ts:
        if t is not None:
            if not t.is_cuda:
                raise RbaError("rba_b200 ops need CUDA tensors (no CPU path)")
            if not t.is_contiguous():
                raise RbaError("rba_b200 ops need contiguous tensors")
            _last_dev[0] = t.device


def split_planes(x):
    """fp32 [rows, cols] -> (hi, lo) bf16 planes (as torch.bfloat16 tensors)."""
    _chk_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().rba_k_split(_p(x), x.shape[0], x.shape[1], x.shape[1], _p(hi), _p(lo), x.shape[1], _stream()))
    return hi, lo


def planes_to_float(hi, lo):
    return hi.float() + lo.float()


def score_fused(pred_masks, pred_logits, out_hw, want_sem_seg=False):
    """maskformer_model.py:294-299,381-386 + evaluate_ood.py:148-150 fused.
    pred_masks (B,Q,h,w), pred_logits (B,Q,K+1) -> rba (B,H,W) [, sem_seg (B,K,H,W)]."""
    _chk_cuda(pred_masks, pred_logits)
    B, Q, h, w = pred_masks.shape
    K = pred_logits.shape[-1] - 1
    H, W = out_hw
    rba = torch.empty((B, H, W), dtype=torch.float32, device=pred_masks.device)
    sem = torch.empty((B, K, H, W), dtype=torch.float32, device=pred_masks.device) if want_sem_seg else None
    _lib.check(_lib.lib().rba_score_fused(_p(pred_masks), _p(pred_logits), B, Q, K, h, w, H, W, _p(rba), _p(sem), _stream()))
    return (rba, sem) if want_sem_seg else rba


SCORE_FUNCS = {"rba": _lib.RBA_SCORE_RBA, "energy": _lib.RBA_SCORE_ENERGY, "pebal": _lib.RBA_SCORE_ENERGY,
               "densehybrid": _lib.RBA_SCORE_DENSEHYBRID, "dense_hybrid": _lib.RBA_SCORE_DENSEHYBRID}
KERNEL_SCORE_FUNCS = ("rba", "energy", "pebal")     # what the fused kernel itself implements


def einsum_score_fused(mask_embed, features, pred_logits, out_hw, bias=None, want_sem_seg=False, score_func="rba",
                       include_void=False):
    """einsum("bqc,bchw->bqhw") (mask2former_transformer_decoder.py:479) + score_fused in ONE kernel: pred_masks is
    never materialised.  mask_embed: (hi, lo) planes (B,Q,D); features: (hi, lo) planes (B,h,w,D) NHWC;
    pred_logits (B,Q,K+1) fp32; bias (B,Q) fp32 or None -> score (B,H,W) [, sem_seg (B,K|K+1,H,W)].
    score_func: "rba" (evaluate_ood.get_RbA) or "energy" (get_energy); include_void: semantic_inference_with_void."""
    e_hi, e_lo = mask_embed
    f_hi, f_lo = features
    _chk_cuda(e_hi, e_lo, f_hi, f_lo, pred_logits, bias)
    B, Q, D = e_hi.shape
    _, h, w, _ = f_hi.shape
    K = pred_logits.shape[-1] - 1
    H, W = out_hw
    Kc = K + 1 if include_void else K
    rba = torch.empty((B, H, W), dtype=torch.float32, device=f_hi.device)
    sem = torch.empty((B, Kc, H, W), dtype=torch.float32, device=f_hi.device) if want_sem_seg else None
    _lib.check(_lib.lib().rba_einsum_score_fused(_p(e_hi), _p(e_lo), _p(bias), _p(f_hi), _p(f_lo), _p(pred_logits), B, Q, K, D,
                                                h, w, H, W, SCORE_FUNCS[score_func], int(bool(include_void)), _p(rba), _p(sem),
                                                _stream()))
    return (rba, sem) if want_sem_seg else rba


def set_fused_score_variant(variant):
    """Test / profiling hook for RbA-only launches: 3 (default) = runs in registers (score_fused3.cu), 2 = tcgen05 score
    phase (score_fused2.cu), 1 = the mma.sync cell kernel (score_fused.cu, which serves every sem_seg / energy launch);
    0 restores the default."""
    _lib.check(_lib.lib().rba_k_set_fused_score_variant(int(variant)))


def _dev_i64(t, like):
    return t.is_cuda and t.device == like.device and t.dtype == torch.int64 and t.is_contiguous()


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, im2col_step):
    """Same signature and result as the reference pybind op (ops/src/vision.cpp:18-21)."""
    _chk_cuda(value, sampling_locations, attention_weights)
    if value.dtype != torch.float32:
        raise RbaError("ms_deform_attn_forward: only float32 is built (the pixel decoder forces fp32, msdeformattn.py:323,329)")
    B, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    out = torch.empty((B, Lq, M * D), dtype=torch.float32, device=value.device)
    if _dev_i64(spatial_shapes, value) and _dev_i64(level_start_index, value):
        # the reference passes CUDA int64 tensors: hand them through (no host copy, no sync, graph-capturable)
        _lib.check(_lib.lib().rba_msda_forward_dev(
            _p(value), _p(spatial_shapes), _p(level_start_index), _p(sampling_locations), _p(attention_weights),
            B, S, M, D, Lq, L, P, int(im2col_step), _p(out), _stream()))
        return out
    ss = spatial_shapes.detach().to("cpu", torch.int64).contiguous()
    ls = level_start_index.detach().to("cpu", torch.int64).contiguous()
    _lib.check(_lib.lib().rba_msda_forward(
        _p(value), ctypes.cast(ss.data_ptr(), ctypes.POINTER(ctypes.c_int64)),
        ctypes.cast(ls.data_ptr(), ctypes.POINTER(ctypes.c_int64)), _p(sampling_locations), _p(attention_weights),
        B, S, M, D, Lq, L, P, int(im2col_step), _p(out), _stream()))
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, grad_output,
                            im2col_step):
    """Same signature and result as the reference pybind op (ops/src/vision.cpp:20, called from
    ops/functions/ms_deform_attn_func.py:43-47): returns (grad_value, grad_sampling_loc, grad_attn_weight)."""
    _chk_cuda(value, sampling_locations, attention_weights, grad_output)
    if value.dtype != torch.float32 or grad_output.dtype != torch.float32:
        raise RbaError("ms_deform_attn_backward: only float32 is built")
    B, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    if tuple(grad_output.shape) != (B, Lq, M * D):
        raise RbaError(f"ms_deform_attn_backward: grad_output {tuple(grad_output.shape)} != {(B, Lq, M * D)}")
    g_value = torch.empty_like(value)
    g_loc = torch.empty_like(sampling_locations)
    g_aw = torch.empty_like(attention_weights)
    if _dev_i64(spatial_shapes, value) and _dev_i64(level_start_index, value):
        _lib.check(_lib.lib().rba_msda_backward_dev(
            _p(value), _p(spatial_shapes), _p(level_start_index), _p(sampling_locations), _p(attention_weights),
            _p(grad_output), B, S, M, D, Lq, L, P, int(im2col_step), _p(g_value), _p(g_loc), _p(g_aw), _stream()))
        return g_value, g_loc, g_aw
    ss = spatial_shapes.detach().to("cpu", torch.int64).contiguous()
    ls = level_start_index.detach().to("cpu", torch.int64).contiguous()
    _lib.check(_lib.lib().rba_msda_backward(
        _p(value), ctypes.cast(ss.data_ptr(), ctypes.POINTER(ctypes.c_int64)),
        ctypes.cast(ls.data_ptr(), ctypes.POINTER(ctypes.c_int64)), _p(sampling_locations), _p(attention_weights),
        _p(grad_output), B, S, M, D, Lq, L, P, int(im2col_step), _p(g_value), _p(g_loc), _p(g_aw), _stream()))
    return g_value, g_loc, g_aw


class MSDeformAttnFunction(torch.autograd.Function):
    """Autograd wrapper with the reference's name and call contract
    (ops/functions/ms_deform_attn_func.py:32-49): `MSDeformAttnFunction.apply(value, spatial_shapes, level_start_index,
    sampling_locations, attention_weights, im2col_step)`."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        out = ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                     attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, aw = ctx.saved_tensors
        gv, gl, ga = ms_deform_attn_backward(value, shapes, lsi, loc, aw, grad_output.contiguous(), ctx.im2col_step)
        return gv, None, None, gl, ga, None


OUTLIER_SCORE_MODES = {("nls", "none"): 0, ("nls", "tanh"): 1, ("nls", "sigmoid"): 2, ("energy", None): 3}


class _OutlierLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_masks, pred_logits, outlier_masks, mode, t_in, t_out):
        B, Q, h, w = pred_masks.shape
        K = pred_logits.shape[-1] - 1
        H, W = outlier_masks.shape[-2:]
        need_grad = pred_masks.requires_grad or pred_logits.requires_grad
        dev = pred_masks.device
        ws = torch.empty(int(_lib.lib().rba_outlier_loss_workspace_floats(B, Q, K, h, w)), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        d_masks = torch.empty_like(pred_masks) if need_grad else None
        d_logits = torch.empty_like(pred_logits) if need_grad else None
        _lib.check(_lib.lib().rba_outlier_loss(_p(pred_masks), _p(pred_logits), _p(outlier_masks), outlier_masks.element_size(),
                                               B, Q, K, h, w, H, W, mode, float(t_in), float(t_out), _p(loss), _p(d_masks),
                                               _p(d_logits), _p(ws), _stream()))
        if need_grad:
            ctx.save_for_backward(d_masks, d_logits)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        d_masks, d_logits = ctx.saved_tensors
        return d_masks * grad_out, d_logits * grad_out, None, None, None, None


def outlier_loss(pred_masks, pred_logits, outlier_masks, outlier_loss_target="nls", score_norm="tanh",
                 outlier_loss_func="squared_hinge", inlier_upper_threshold=-1.0, outlier_lower_threshold=-0.1):
    """SetCriterion.outlier_loss (mask2former/modeling/criterion.py:435-553) as one fused forward + backward launch
    sequence; differentiable w.r.t. pred_masks (B,Q,h,w) and pred_logits (B,Q,K+1).  outlier_masks: (B,H,W) uint8/int64
    (1 outlier, 0 inlier, other values ignored).  Argument names and defaults are the reference's config keys
    (MODEL.MASK_FORMER.OUTLIER_LOSS_TARGET / SCORE_NORM / OUTLIER_LOSS_FUNC / INLIER_UPPER_THRESHOLD / OUTLIER_LOWER_THRESHOLD,
    mask2former/config.py:188-227).  Returns the scalar loss tensor."""
    _chk_cuda(pred_masks, pred_logits, outlier_masks)
    if outlier_loss_func != "squared_hinge":
        raise RbaError(f"outlier_loss: OUTLIER_LOSS_FUNC {outlier_loss_func!r} is not built (squared_hinge only)")
    key = (outlier_loss_target, None if outlier_loss_target == "energy" else (score_norm or "none"))
    if key not in OUTLIER_SCORE_MODES:
        raise RbaError(f"outlier_loss: target {outlier_loss_target!r} / score_norm {score_norm!r} is not built "
                       "(nls with none|tanh|sigmoid, or energy)")
    if pred_masks.dtype != torch.float32 or pred_logits.dtype != torch.float32:
        raise RbaError("outlier_loss: float32 inputs expected")
    if outlier_masks.dtype not in (torch.uint8, torch.int64):
        outlier_masks = outlier_masks.to(torch.int64)
    if pred_masks.dim() != 4 or pred_logits.dim() != 3 or outlier_masks.dim() != 3 or \
            pred_masks.shape[:2] != pred_logits.shape[:2] or outlier_masks.shape[0] != pred_masks.shape[0]:
        raise RbaError(f"outlier_loss: shapes {tuple(pred_masks.shape)}, {tuple(pred_logits.shape)}, {tuple(outlier_masks.shape)}")
    return _OutlierLoss.apply(pred_masks, pred_logits, outlier_masks.contiguous(), OUTLIER_SCORE_MODES[key],
                              inlier_upper_threshold, outlier_lower_threshold)


def gemm(a, w, bias=None, act=RBA_ACT_NONE, residual=None, out_planes=False, out_f32=True, bias_per_row=False,
         swin=None, backend=RBA_GEMM_FFMA, out=None, qkv_tile_heads=0):
    """C = act(A W^T + bias) (+ residual).  a: (hi, lo) planes [M,K] or [batch,M,K]; w: planes [N,K] or [batch,N,K].
    swin = (B, H, W, ws, shift) scatters windowed rows to token rows (output has B*H*W rows)."""
    a_hi, a_lo = a
    w_hi, w_lo = w
    _chk_cuda(a_hi, a_lo, w_hi, w_lo, bias, residual, out)
    batched = a_hi.dim() == 3
    batch = a_hi.shape[0] if batched else 1
    M, K = a_hi.shape[-2:]
    N = w_hi.shape[-2]
    dev = a_hi.device
    rows_out = M
    if swin is not None:
        rows_out = swin[0] * swin[1] * swin[2]
    shape = (batch, rows_out, N) if batched else (rows_out, N)
    c = out
    if c is None and out_f32:
        c = residual.clone() if (swin is not None and residual is not None) else torch.empty(shape, dtype=torch.float32, device=dev)
    c_hi = c_lo = None
    if out_planes:
        c_hi = torch.zeros(shape, dtype=torch.bfloat16, device=dev)
        c_lo = torch.zeros(shape, dtype=torch.bfloat16, device=dev)
    g = RbaGemmArgs()
    g.a_hi, g.a_lo, g.lda = a_hi.data_ptr(), a_lo.data_ptr(), K
    g.w_hi, g.w_lo, g.ldw = w_hi.data_ptr(), w_lo.data_ptr(), K
    g.M, g.N, g.K, g.batch = M, N, K, batch
    g.a_bstride = M * K if batched else 0
    g.w_bstride = N * K if (batched and w_hi.dim() == 3) else 0
    if bias is not None:
        g.bias = bias.data_ptr()
        g.bias_per_row = 1 if bias_per_row else 0
        g.bias_bstride = bias.shape[-1] if (batched and bias.dim() == 2) else 0
    g.act = act
    if residual is not None:
        g.residual = residual.data_ptr()
    if c is not None:
        g.c, g.ldc, g.c_bstride = c.data_ptr(), N, rows_out * N
    if c_hi is not None:
        g.c_hi, g.c_lo, g.ldcp, g.cp_bstride = c_hi.data_ptr(), c_lo.data_ptr(), N, rows_out * N
    if swin is not None:
        g.swin_map, g.sw_H, g.sw_W, g.sw_ws, g.sw_shift = 1, swin[1], swin[2], swin[3], swin[4]
    g.backend = backend
    g.qkv_tile_heads = int(qkv_tile_heads)        # planes in the (window, part, head) tiled layout (tensor-core backend only)
    _lib.check(_lib.lib().rba_k_gemm(ctypes.byref(g), _stream()))
    if out_planes and c is not None:
        return c, (c_hi, c_lo)
    return (c_hi, c_lo) if out_planes else c


def conv3x3(x, w, backend=RBA_GEMM_FFMA):
    """x: planes (B,H,W,Cin) NHWC; w: planes [Cout, 9*Cin] (k = (ky*3+kx)*Cin + ci) -> fp32 (B,H,W,Cout)."""
    x_hi, x_lo = x
    w_hi, w_lo = w
    _chk_cuda(x_hi, x_lo, w_hi, w_lo)
    B, H, W, Cin = x_hi.shape
    Cout = w_hi.shape[0]
    y = torch.empty((B, H, W, Cout), dtype=torch.float32, device=x_hi.device)
    _lib.check(_lib.lib().rba_k_conv3x3(_p(x_hi), _p(x_lo), _p(w_hi), _p(w_lo), B, H, W, Cin, Cout, _p(y), backend, _stream()))
    return y


def layernorm(x, gamma, beta, mode=0, B=1, H=1, W=None, ws=0, shift=0, eps=1e-5, want_f32=True, want_planes=False):
    """x fp32 [B*H*W, C].  mode 0 plain / 1 Swin window gather / 2 PatchMerging gather (see rba_b200.h)."""
    _chk_cuda(x, gamma, beta)
    C = x.shape[-1]
    if W is None:
        W = x.numel() // C
    if mode == 0:
        rows, CO = B * H * W, C
    elif mode == 1:
        nWh, nWw = -(-H // ws), -(-W // ws)
        rows, CO = B * nWh * nWw * ws * ws, C
    else:
        rows, CO = B * (H // 2) * (W // 2), 4 * C
    y = torch.empty((rows, CO), dtype=torch.float32, device=x.device) if want_f32 else None
    hi = lo = None
    if want_planes:
        hi = torch.empty((rows, CO), dtype=torch.bfloat16, device=x.device)
        lo = torch.empty((rows, CO), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().rba_k_layernorm(_p(x), _p(gamma), _p(beta), mode, B, H, W, C, ws, shift, eps, _p(y), _p(hi), _p(lo), _stream()))
    if want_f32 and want_planes:
        return y, (hi, lo)
    return y if want_f32 else (hi, lo)


def window_attn(qkv, bias_table, B, H, W, C, heads, ws, shift):
    _chk_cuda(qkv, bias_table)
    rows = qkv.shape[0]
    hi = torch.empty((rows, C), dtype=torch.bfloat16, device=qkv.device)
    lo = torch.empty((rows, C), dtype=torch.bfloat16, device=qkv.device)
    _lib.check(_lib.lib().rba_k_window_attn(_p(qkv), _p(bias_table), B, H, W, C, heads, ws, shift, _p(hi), _p(lo), _stream()))
    return hi, lo


def window_attn_planes(qkv, bias_table, B, H, W, C, heads, ws, shift):
    """Tensor-core window attention; qkv = (hi, lo) planes [rows, 3C]."""
    q_hi, q_lo = qkv
    _chk_cuda(q_hi, q_lo, bias_table)
    rows = q_hi.shape[0]
    hi = torch.empty((rows, C), dtype=torch.bfloat16, device=q_hi.device)
    lo = torch.empty((rows, C), dtype=torch.bfloat16, device=q_hi.device)
    _lib.check(_lib.lib().rba_k_window_attn_planes(_p(q_hi), _p(q_lo), _p(bias_table), B, H, W, C, heads, ws, shift, _p(hi), _p(lo), _stream()))
    return hi, lo


def qkv_to_tiles(planes, heads):
    """Row-major q|k|v planes [rows, 3C] -> the tiled layout rba_gemm_args.qkv_tile_heads writes (torch permute; test aid)."""
    out = []
    for t in planes:
        rows = t.shape[0]
        out.append(t.view(rows // 144, 144, 3, heads, 4, 8).permute(0, 2, 3, 4, 1, 5).contiguous().view(rows, -1))
    return tuple(out)


def window_attn_tc(qkv, bias_table, B, H, W, C, heads, ws, shift, tiled=False):
    """tcgen05 + TMA window attention; qkv = (hi, lo) planes [rows, 3C] (tiled=True: in the tiled layout of qkv_to_tiles /
    the QKV GEMM's qkv_tile_heads mode); bias_table is the reference's relative_position_bias_table [(2ws-1)^2, heads]
    (prepared per call here; the engine prepares it once at load)."""
    q_hi, q_lo = qkv
    _chk_cuda(q_hi, q_lo, bias_table)
    rows = q_hi.shape[0]
    L = _lib.lib()
    prep = torch.empty(int(L.rba_k_window_attn_bias_floats(heads)), dtype=torch.float32, device=q_hi.device)
    _lib.check(L.rba_k_window_attn_prepare_bias(_p(bias_table), heads, _p(prep), _stream()))
    hi = torch.empty((rows, C), dtype=torch.bfloat16, device=q_hi.device)
    lo = torch.empty((rows, C), dtype=torch.bfloat16, device=q_hi.device)
    fn = L.rba_k_window_attn_tc_tiled if tiled else L.rba_k_window_attn_tc
    _lib.check(fn(_p(q_hi), _p(q_lo), _p(prep), B, H, W, C, heads, ws, shift, _p(hi), _p(lo), _stream()))
    return hi, lo


def mha(q, k, v, mask, heads):
    """q (B,Lq,E), k/v (B,Lk,E) fp32 projected; mask (B,Lq,Lk) uint8 (1 = blocked) or None -> planes (B*Lq, E)."""
    _chk_cuda(q, k, v, mask)
    B, Lq, E = q.shape
    Lk = k.shape[1]
    hi = torch.empty((B * Lq, E), dtype=torch.bfloat16, device=q.device)
    lo = torch.empty((B * Lq, E), dtype=torch.bfloat16, device=q.device)
    ws = torch.empty(max(1, int(_lib.lib().rba_k_mha_workspace_floats(B, Lq, Lk, heads))), dtype=torch.float32, device=q.device)
    _lib.check(_lib.lib().rba_k_mha(_p(q), _p(k), _p(v), _p(mask), B, Lq, Lk, E, heads, _p(hi), _p(lo), _p(ws), _stream()))
    return hi, lo


def groupnorm(x, gamma, beta, groups=32, eps=1e-5, prev=None, relu=False, want_f32=True, want_planes=False):
    """x fp32 (B,H,W,C) NHWC; y = GN(x) [+ bilinear_up(prev (B,hp,wp,C))] [relu]."""
    _chk_cuda(x, gamma, beta, prev)
    B, H, W, C = x.shape
    hp, wp = (prev.shape[1], prev.shape[2]) if prev is not None else (0, 0)
    n = int(_lib.lib().rba_k_groupnorm_ws(B, H, W, C, groups))
    ws = torch.empty(n, dtype=torch.float64, device=x.device)
    y = torch.empty_like(x) if want_f32 else None
    hi = lo = None
    if want_planes:
        hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().rba_k_groupnorm(_p(x), _p(gamma), _p(beta), B, H, W, C, groups, eps, _p(prev), hp, wp, int(relu),
                                         _p(y), _p(hi), _p(lo), _p(ws), _stream()))
    if want_f32 and want_planes:
        return y, (hi, lo)
    return y if want_f32 else (hi, lo)


def patch_embed(images, Hp, Wp, mean, std, conv_w, conv_b, gamma, beta):
    """images (B,3,H,W) uint8 or float32 -> tokens (B, (Hp/4)*(Wp/4), C)."""
    _chk_cuda(images, conv_w, conv_b, gamma, beta)
    B, _, H, W = images.shape
    C = conv_w.shape[0]
    dt = _lib.RBA_IMG_U8 if images.dtype == torch.uint8 else _lib.RBA_IMG_F32
    if images.dtype not in (torch.uint8, torch.float32):
        raise RbaError("patch_embed: images must be uint8 or float32")
    tok = torch.empty((B, (Hp // 4) * (Wp // 4), C), dtype=torch.float32, device=images.device)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    _lib.check(_lib.lib().rba_k_patch_embed(_p(images), dt, B, H, W, Hp, Wp, m, s, _p(conv_w), _p(conv_b), _p(gamma), _p(beta),
                                           C, _p(tok), _stream()))
    return tok


def attn_mask(masks, target_hw):
    """masks (B,Q,h,w) fp32 -> (B,Q,th*tw) uint8, 1 = blocked (mask2former_transformer_decoder.py:483-486,:433)."""
    _chk_cuda(masks)
    B, Q, h, w = masks.shape
    th, tw = target_hw
    out = torch.empty((B, Q, th * tw), dtype=torch.uint8, device=masks.device)
    _lib.check(_lib.lib().rba_k_attn_mask(_p(masks), B, Q, h, w, th, tw, _p(out), _stream()))
    return out

"""CPU oracle for the RbA hot path (Mask2Former forward + Rejected-by-All score).

TEST INFRASTRUCTURE — NOT PRODUCT.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import this file, and only
as the checker / the CPU baseline.  The product path (rba_b200/) never routes
through it and fails loudly when its CUDA library is missing.

It is a functional fp32 PyTorch-CPU restatement of the reference's algorithm: plain
functions over a reference-layout `state_dict` (same key names as the reference's
`MaskFormer.state_dict()`), no nn.Module of the reference, no detectron2.  Each
function cites the reference file:line it follows (paths relative to /root/reference).

PARITY PIN: the reference has no golden vectors for this path except the MSDeformAttn
shapes/seed in mask2former/modeling/pixel_decoder/ops/test.py:24-63.  This oracle is
therefore pinned against *outputs of the reference itself run in the build container*
(`oracle/ref_loader.py` imports the unmodified reference modules; `oracle/make_golden.py`
writes tests/golden/*.pt; `tests/test_oracle_golden.py` re-checks the oracle against
those fixtures on every run, and `tests/test_oracle_vs_reference.py` against the live
reference whenever /root/reference is present).

Tolerance: the north star states 1e-3 max-abs in fp32; oracle-vs-reference agrees to
fp32 round-off (<= 2e-5 on every tap, see tests).
"""
import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------


def config_from_yaml_dict(y):
    """Pulls the hot-path knobs out of a dumped ckpt YAML (ckpts/*/config.yaml).
    Keys follow mask2former/config.py:74-90 (SWIN), :40-64 (MASK_FORMER), :67,170-172."""
    M = y["MODEL"]
    mf, sh, sw = M["MASK_FORMER"], M["SEM_SEG_HEAD"], M["SWIN"]
    assert M["BACKBONE"]["NAME"] in ("D2SwinTransformer", "build_resnet_backbone")
    assert sh["PIXEL_DECODER_NAME"] == "MSDeformAttnPixelDecoder"
    assert mf["TRANSFORMER_DECODER_NAME"] == "MultiScaleMaskedTransformerDecoder"
    assert not mf["PRE_NORM"] and not sw["APE"]
    return SimpleNamespace(
        backbone="resnet" if M["BACKBONE"]["NAME"] == "build_resnet_backbone" else "swin",
        resnet_depth=int(M.get("RESNETS", {}).get("DEPTH", 50)),
        embed_dim=sw["EMBED_DIM"], depths=list(sw["DEPTHS"]), num_heads=list(sw["NUM_HEADS"]),
        window_size=sw["WINDOW_SIZE"], mlp_ratio=sw["MLP_RATIO"], patch_size=sw["PATCH_SIZE"],
        conv_dim=sh["CONVS_DIM"], mask_dim=sh["MASK_DIM"], num_classes=sh["NUM_CLASSES"],
        in_features=list(sh["IN_FEATURES"]),
        transformer_in_features=list(sh["DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES"]),
        common_stride=sh["COMMON_STRIDE"], enc_layers=sh["TRANSFORMER_ENC_LAYERS"],
        enc_heads=mf["NHEADS"], enc_points=4, enc_ffn=1024,          # msdeformattn.py:315, ms_deform_attn.py:35
        hidden_dim=mf["HIDDEN_DIM"], nheads=mf["NHEADS"], dim_feedforward=mf["DIM_FEEDFORWARD"],
        dec_layers=mf["DEC_LAYERS"] - 1,                              # mask2former_transformer_decoder.py:387-388
        num_queries=mf["NUM_OBJECT_QUERIES"],
        size_divisibility=mf["SIZE_DIVISIBILITY"],
        pixel_mean=list(M["PIXEL_MEAN"]), pixel_std=list(M["PIXEL_STD"]),
    )


# --------------------------------------------------------------------------------------
# Swin backbone  (mask2former/modeling/backbone/swin.py)
# --------------------------------------------------------------------------------------


def window_partition(x, ws):
    """swin.py:44-55"""
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def window_reverse(windows, ws, H, W):
    """swin.py:58-71"""
    B = windows.shape[0] // ((H // ws) * (W // ws))
    x = windows.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def relative_position_index(ws):
    """swin.py:110-120 — (ws*ws, ws*ws) index into the (2ws-1)^2 bias table."""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def shift_attn_mask(H, W, ws, shift, device=None):
    """swin.py:413-440 — (nW, ws*ws, ws*ws) of {0, -100}."""
    Hp = int(math.ceil(H / ws)) * ws
    Wp = int(math.ceil(W / ws)) * ws
    img_mask = torch.zeros((1, Hp, Wp, 1), device=device)
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img_mask[:, h, w, :] = cnt
            cnt += 1
    mw = window_partition(img_mask, ws).view(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, -100.0).masked_fill(am == 0, 0.0)


def window_attention(sd, p, x, mask, num_heads, ws):
    """swin.py:131-171  x: (nW*B, ws*ws, C)"""
    B_, N, C = x.shape
    qkv = F.linear(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"])
    qkv = qkv.reshape(B_, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * ((C // num_heads) ** -0.5)
    attn = q @ k.transpose(-2, -1)
    idx = sd.get(p + "relative_position_index", None)
    if idx is None:
        idx = relative_position_index(ws).to(x.device)
    bias = sd[p + "relative_position_bias_table"][idx.view(-1)].view(N, N, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, num_heads, N, N)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(x, sd[p + "proj.weight"], sd[p + "proj.bias"])


def swin_block(sd, p, x, H, W, num_heads, ws, shift, mask_matrix):
    """swin.py:235-295"""
    B, L, C = x.shape
    shortcut = x
    x = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    x = x.view(B, H, W, C)
    pad_r = (ws - W % ws) % ws
    pad_b = (ws - H % ws) % ws
    x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))
    _, Hp, Wp, _ = x.shape
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
        am = mask_matrix
    else:
        am = None
    xw = window_partition(x, ws).view(-1, ws * ws, C)
    aw = window_attention(sd, p + "attn.", xw, am, num_heads, ws)
    x = window_reverse(aw.view(-1, ws, ws, C), ws, Hp, Wp)
    if shift > 0:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    if pad_r > 0 or pad_b > 0:
        x = x[:, :H, :W, :].contiguous()
    x = shortcut + x.view(B, H * W, C)
    y = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    y = F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    y = F.gelu(y)  # nn.GELU() exact erf, swin.py:25,37
    y = F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + y


def patch_merging(sd, p, x, H, W):
    """swin.py:311-337"""
    B, L, C = x.shape
    x = x.view(B, H, W, C)
    if H % 2 == 1 or W % 2 == 1:
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, 0::2, 0::2, :], x[:, 1::2, 0::2, :], x[:, 0::2, 1::2, :], x[:, 1::2, 1::2, :]], -1)
    x = x.view(B, -1, 4 * C)
    x = F.layer_norm(x, (4 * C,), sd[p + "norm.weight"], sd[p + "norm.bias"])
    return F.linear(x, sd[p + "reduction.weight"])


def swin_forward(sd, cfg, x, p="backbone."):
    """swin.py:479-495 (PatchEmbed) + :651-678.  x: normalised (B,3,H,W) fp32.
    Returns {"res2".."res5"} NCHW."""
    ps = cfg.patch_size
    _, _, H, W = x.shape
    if W % ps != 0:
        x = F.pad(x, (0, ps - W % ps))
    if H % ps != 0:
        x = F.pad(x, (0, 0, 0, ps - H % ps))
    x = F.conv2d(x, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=ps)
    Wh, Ww = x.size(2), x.size(3)
    x = x.flatten(2).transpose(1, 2)
    C = cfg.embed_dim
    x = F.layer_norm(x, (C,), sd[p + "patch_embed.norm.weight"], sd[p + "patch_embed.norm.bias"])
    outs = {}
    ws = cfg.window_size
    for i, depth in enumerate(cfg.depths):
        Ci = C * 2 ** i
        mask = shift_attn_mask(Wh, Ww, ws, ws // 2, device=x.device)
        for j in range(depth):
            x = swin_block(sd, f"{p}layers.{i}.blocks.{j}.", x, Wh, Ww, cfg.num_heads[i], ws,
                           0 if j % 2 == 0 else ws // 2, mask)
        xo = F.layer_norm(x, (Ci,), sd[f"{p}norm{i}.weight"], sd[f"{p}norm{i}.bias"])
        outs[f"res{i + 2}"] = xo.view(-1, Wh, Ww, Ci).permute(0, 3, 1, 2).contiguous()
        if i < len(cfg.depths) - 1:
            x = patch_merging(sd, f"{p}layers.{i}.downsample.", x, Wh, Ww)
            Wh, Ww = (Wh + 1) // 2, (Ww + 1) // 2
    return outs


# --------------------------------------------------------------------------------------
# ResNet backbone  (detectron2 `build_resnet_backbone`, selected by configs/cityscapes/semantic-segmentation/
# Base-Cityscapes-SemanticSegmentation.yaml:4,8-15; call site maskformer_model.py:109)
# --------------------------------------------------------------------------------------
# PARITY UNPINNED: detectron2 is neither vendored by the reference nor pinned (INSTALL.md:13-19).  This restates the published
# architecture -- BasicStem (7x7/2 conv, norm, ReLU, 3x3/2 max-pool), bottleneck blocks with the stride in the 3x3 conv
# (RESNETS.STRIDE_IN_1X1: False), eval-mode batch norm -- over detectron2's state_dict names, which are the names
# tools/convert-torchvision-to-d2.py:33-44 maps torchvision's to.  tests/test_resnet_oracle.py pins it on torchvision's
# resnet50 / resnet101 under that mapping.

RESNET_BLOCKS = {50: [3, 4, 6, 3], 101: [3, 4, 23, 3]}


def _conv_bn(sd, p, x, stride=1, padding=0, eps=1e-5):
    x = F.conv2d(x, sd[p + "weight"], None, stride=stride, padding=padding)
    return F.batch_norm(x, sd[p + "norm.running_mean"], sd[p + "norm.running_var"], sd[p + "norm.weight"], sd[p + "norm.bias"],
                        training=False, eps=eps)


def resnet_forward(sd, cfg, x, p="backbone."):
    """x: normalised (B,3,H,W) fp32.  Returns {"res2".."res5"} NCHW with 256 / 512 / 1024 / 2048 channels."""
    x = F.relu(_conv_bn(sd, p + "stem.conv1.", x, stride=2, padding=3))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = {}
    for i, n in enumerate(RESNET_BLOCKS[getattr(cfg, "resnet_depth", 50)]):
        for j in range(n):
            q = f"{p}res{i + 2}.{j}."
            stride = 2 if (j == 0 and i > 0) else 1
            out = F.relu(_conv_bn(sd, q + "conv1.", x))
            out = F.relu(_conv_bn(sd, q + "conv2.", out, stride=stride, padding=1))
            out = _conv_bn(sd, q + "conv3.", out)
            sc = _conv_bn(sd, q + "shortcut.", x, stride=stride) if (q + "shortcut.weight") in sd else x
            x = F.relu(out + sc)
        outs[f"res{i + 2}"] = x
    return outs


# --------------------------------------------------------------------------------------
# Pixel decoder  (mask2former/modeling/pixel_decoder/msdeformattn.py, ops/)
# --------------------------------------------------------------------------------------


def position_embedding_sine(B, H, W, num_pos_feats, temperature=10000, scale=2 * math.pi, device=None):
    """transformer_decoder/position_encoding.py:29-52 with mask=None, normalize=True."""
    not_mask = torch.ones((B, H, W), dtype=torch.bool, device=device)
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    eps = 1e-6
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32, device=device)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def msda_bilinear_gather(value, spatial_shapes, sampling_locations, attention_weights):
    """Explicit restatement of the reference CUDA forward kernel's arithmetic
    (ops/src/cuda/ms_deform_im2col_cuda.cuh:38-89 bilinear, :242-304 main loop):
    h_im = loc_y*H - 0.5, w_im = loc_x*W - 0.5; a sample contributes only if
    h_im > -1 && w_im > -1 && h_im < H && w_im < W; 4 taps, each zero outside the map.
    Equivalent to the reference's own CPU statement `ms_deform_attn_core_pytorch`
    (ops/functions/ms_deform_attn_func.py:52-72: grid_sample bilinear/zeros/align_corners=False).
    value (N,S,M,D), sampling_locations (N,Lq,M,L,P,2) in [0,1] (x,y), weights (N,Lq,M,L,P)
    -> (N,Lq,M*D)."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    out = torch.zeros(N, Lq, M, D, dtype=value.dtype)
    start = 0
    for lid, (H, W) in enumerate([(int(h), int(w)) for h, w in spatial_shapes]):
        v = value[:, start:start + H * W].reshape(N, H, W, M, D)
        start += H * W
        loc = sampling_locations[:, :, :, lid]  # N,Lq,M,P,2
        w_im = loc[..., 0] * W - 0.5
        h_im = loc[..., 1] * H - 0.5
        valid = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
        h_low = torch.floor(h_im)
        w_low = torch.floor(w_im)
        lh, lw = h_im - h_low, w_im - w_low
        hh, hw = 1 - lh, 1 - lw
        h_low, w_low = h_low.long(), w_low.long()
        aw = attention_weights[:, :, :, lid]  # N,Lq,M,P
        n_idx = torch.arange(N).view(N, 1, 1, 1).expand(N, Lq, M, P)
        m_idx = torch.arange(M).view(1, 1, M, 1).expand(N, Lq, M, P)
        for dy, dx, wt in ((0, 0, hh * hw), (0, 1, hh * lw), (1, 0, lh * hw), (1, 1, lh * lw)):
            yy, xx = h_low + dy, w_low + dx
            ok = valid & (yy >= 0) & (yy <= H - 1) & (xx >= 0) & (xx <= W - 1)
            g = v[n_idx, yy.clamp(0, H - 1), xx.clamp(0, W - 1), m_idx]  # N,Lq,M,P,D
            out += (g * (wt * aw * ok.to(value.dtype)).unsqueeze(-1)).sum(3)
    return out.view(N, Lq, M * D)


def msda_bilinear_backward(value, spatial_shapes, sampling_locations, attention_weights, grad_output):
    """Explicit restatement of the reference CUDA backward's arithmetic
    (ops/src/cuda/ms_deform_im2col_cuda.cuh:92-163 `ms_deform_attn_col2im_bilinear`, driven by the kernels at :306-925):
      top_grad_value = grad_out * attn_weight;  grad_value[tap] += w_tap * top_grad_value  (atomicAdd, :127-155)
      grad_attn_weight = sum_d grad_out * val                                              (:158)
      grad_sampling_loc.x = W * sum_d grad_w_weight * top_grad_value, .y = H * sum_d grad_h_weight * top_grad_value
                                                                                           (:159-160)
    with grad_w_weight = -hh v1 + hh v2 - lh v3 + lh v4 and grad_h_weight = -hw v1 - lw v2 + hw v3 + lw v4, taps outside
    the map reading 0 and samples outside (-1,H)x(-1,W) skipped.  Returns (grad_value, grad_sampling_loc, grad_attn_weight)."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    go = grad_output.view(N, Lq, M, 1, D)
    g_value = torch.zeros_like(value)
    g_loc = torch.zeros_like(sampling_locations)
    g_aw = torch.zeros_like(attention_weights)
    start = 0
    for lid, (H, W) in enumerate([(int(h), int(w)) for h, w in spatial_shapes]):
        v = value[:, start:start + H * W].reshape(N, H, W, M, D)
        gv = torch.zeros_like(v)
        loc = sampling_locations[:, :, :, lid]
        w_im = loc[..., 0] * W - 0.5
        h_im = loc[..., 1] * H - 0.5
        valid = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
        h_low, w_low = torch.floor(h_im), torch.floor(w_im)
        lh, lw = h_im - h_low, w_im - w_low
        hh, hw = 1 - lh, 1 - lw
        h_low, w_low = h_low.long(), w_low.long()
        aw = attention_weights[:, :, :, lid]
        n_idx = torch.arange(N).view(N, 1, 1, 1).expand(N, Lq, M, P)
        m_idx = torch.arange(M).view(1, 1, M, 1).expand(N, Lq, M, P)
        top = go * aw.unsqueeze(-1)                                         # N,Lq,M,P,D
        taps = []
        for dy, dx, wt in ((0, 0, hh * hw), (0, 1, hh * lw), (1, 0, lh * hw), (1, 1, lh * lw)):
            yy, xx = h_low + dy, w_low + dx
            ok = (valid & (yy >= 0) & (yy <= H - 1) & (xx >= 0) & (xx <= W - 1)).to(value.dtype).unsqueeze(-1)
            yc, xc = yy.clamp(0, H - 1), xx.clamp(0, W - 1)
            taps.append(v[n_idx, yc, xc, m_idx] * ok)
            gv.index_put_((n_idx, yc, xc, m_idx), top * wt.unsqueeze(-1) * ok, accumulate=True)
        v1, v2, v3, v4 = taps
        e = lambda t: t.unsqueeze(-1)  # noqa: E731
        val = e(hh * hw) * v1 + e(hh * lw) * v2 + e(lh * hw) * v3 + e(lh * lw) * v4
        gw_w = -e(hh) * v1 + e(hh) * v2 - e(lh) * v3 + e(lh) * v4
        gh_w = -e(hw) * v1 - e(lw) * v2 + e(hw) * v3 + e(lw) * v4
        g_aw[:, :, :, lid] = (go * val).sum(-1)
        g_loc[:, :, :, lid, :, 0] = W * (gw_w * top).sum(-1)
        g_loc[:, :, :, lid, :, 1] = H * (gh_w * top).sum(-1)
        g_value[:, start:start + H * W] = gv.reshape(N, H * W, M, D)
        start += H * W
    return g_value, g_loc, g_aw


def msda_core_grid_sample(value, spatial_shapes, sampling_locations, attention_weights):
    """ops/functions/ms_deform_attn_func.py:52-72 (the reference's CPU statement); used for
    full-size runs because it is much faster than msda_bilinear_gather."""
    N_, S_, M_, D_ = value.shape
    _, Lq_, _, L_, P_, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in spatial_shapes]
    value_list = value.split([h * w for h, w in shapes], dim=1)
    grids = 2 * sampling_locations - 1
    sv = []
    for lid, (H_, W_) in enumerate(shapes):
        v_l = value_list[lid].flatten(2).transpose(1, 2).reshape(N_ * M_, D_, H_, W_)
        g_l = grids[:, :, :, lid].transpose(1, 2).flatten(0, 1)
        sv.append(F.grid_sample(v_l, g_l, mode="bilinear", padding_mode="zeros", align_corners=False))
    aw = attention_weights.transpose(1, 2).reshape(N_ * M_, 1, Lq_, L_ * P_)
    out = (torch.stack(sv, dim=-2).flatten(-2) * aw).sum(-1).view(N_, M_ * D_, Lq_)
    return out.transpose(1, 2).contiguous()


def msdeform_attn(sd, p, query, reference_points, src, spatial_shapes, n_heads, n_levels, n_points, explicit=False):
    """ops/modules/ms_deform_attn.py:82-125 (reference_points last dim == 2, no padding mask)."""
    N, Lq, C = query.shape
    _, Lin, _ = src.shape
    value = F.linear(src, sd[p + "value_proj.weight"], sd[p + "value_proj.bias"]).view(N, Lin, n_heads, C // n_heads)
    off = F.linear(query, sd[p + "sampling_offsets.weight"], sd[p + "sampling_offsets.bias"])
    off = off.view(N, Lq, n_heads, n_levels, n_points, 2)
    aw = F.linear(query, sd[p + "attention_weights.weight"], sd[p + "attention_weights.bias"])
    aw = F.softmax(aw.view(N, Lq, n_heads, n_levels * n_points), -1).view(N, Lq, n_heads, n_levels, n_points)
    normalizer = torch.tensor([[w, h] for h, w in spatial_shapes], dtype=torch.float32, device=query.device)
    loc = reference_points[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
    core = msda_bilinear_gather if explicit else msda_core_grid_sample
    out = core(value, spatial_shapes, loc, aw)
    return F.linear(out, sd[p + "output_proj.weight"], sd[p + "output_proj.bias"])


def encoder_reference_points(spatial_shapes, B, device=None):
    """msdeformattn.py:150-162 with valid_ratios == 1 (masks are all-False, :71)."""
    refs = []
    for H_, W_ in spatial_shapes:
        ry, rx = torch.meshgrid(torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device),
                                torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device), indexing="ij")
        ry = ry.reshape(-1)[None] / (torch.ones(B, 1, device=device) * H_)
        rx = rx.reshape(-1)[None] / (torch.ones(B, 1, device=device) * W_)
        refs.append(torch.stack((rx, ry), -1))
    ref = torch.cat(refs, 1)  # B, S, 2
    L = len(spatial_shapes)
    return ref[:, :, None] * torch.ones(B, 1, L, 2, device=device)


def pixel_decoder_forward(sd, cfg, feats, p="sem_seg_head.pixel_decoder.", explicit_msda=False, taps=None):
    """msdeformattn.py:323-367 (+ :70-98 encoder-only transformer, :131-140 layer).
    feats: {"res2".."res5"} NCHW.  Returns (mask_features, multi_scale_features)."""
    names = sorted(cfg.in_features)                      # res2..res5 by stride (msdeformattn.py:209-212)
    strides = {n: 2 ** (int(n[3:])) for n in names}       # res2->4 ... res5->32
    tin = sorted(cfg.transformer_in_features)
    D = cfg.conv_dim
    srcs, pos = [], []
    for idx, f in enumerate(tin[::-1]):                   # low-res first, :328-331
        x = feats[f].float()
        y = F.conv2d(x, sd[f"{p}input_proj.{idx}.0.weight"], sd[f"{p}input_proj.{idx}.0.bias"])
        y = F.group_norm(y, 32, sd[f"{p}input_proj.{idx}.1.weight"], sd[f"{p}input_proj.{idx}.1.bias"])
        srcs.append(y)
        pos.append(position_embedding_sine(x.shape[0], x.shape[2], x.shape[3], D // 2, device=x.device))
    B = srcs[0].shape[0]
    spatial_shapes = [(s.shape[2], s.shape[3]) for s in srcs]
    src_flat = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    lvl_pos = torch.cat([pe.flatten(2).transpose(1, 2) + sd[p + "transformer.level_embed"][l].view(1, 1, -1)
                         for l, pe in enumerate(pos)], 1)
    ref = encoder_reference_points(spatial_shapes, B, device=src_flat.device)
    out = src_flat
    L = len(spatial_shapes)
    if taps is not None:
        taps["enc_in"] = src_flat
    for i in range(cfg.enc_layers):
        q = f"{p}transformer.encoder.layers.{i}."
        s2 = msdeform_attn(sd, q + "self_attn.", out + lvl_pos, ref, out, spatial_shapes,
                           cfg.enc_heads, L, cfg.enc_points, explicit=explicit_msda)
        out = F.layer_norm(out + s2, (D,), sd[q + "norm1.weight"], sd[q + "norm1.bias"])
        s2 = F.linear(F.relu(F.linear(out, sd[q + "linear1.weight"], sd[q + "linear1.bias"])),
                      sd[q + "linear2.weight"], sd[q + "linear2.bias"])
        out = F.layer_norm(out + s2, (D,), sd[q + "norm2.weight"], sd[q + "norm2.bias"])
        if taps is not None:
            taps[f"enc_l{i}"] = out
    sizes = [h * w for h, w in spatial_shapes]
    outs = [z.transpose(1, 2).reshape(B, -1, h, w) for z, (h, w) in zip(out.split(sizes, dim=1), spatial_shapes)]
    # extra FPN levels :267-268, :352-360
    min_stride = min(strides[f] for f in tin)
    num_fpn = int(math.log2(min_stride) - math.log2(cfg.common_stride))
    for idx, f in enumerate(names[:num_fpn][::-1]):
        k = num_fpn - idx                                 # adapter_k / layer_k, :293-294,300-301
        x = feats[f].float()
        cur = F.conv2d(x, sd[f"{p}adapter_{k}.weight"])
        cur = F.group_norm(cur, 32, sd[f"{p}adapter_{k}.norm.weight"], sd[f"{p}adapter_{k}.norm.bias"])
        y = cur + F.interpolate(outs[-1], size=cur.shape[-2:], mode="bilinear", align_corners=False)
        y = F.conv2d(y, sd[f"{p}layer_{k}.weight"], padding=1)
        y = F.relu(F.group_norm(y, 32, sd[f"{p}layer_{k}.norm.weight"], sd[f"{p}layer_{k}.norm.bias"]))
        if taps is not None:
            taps[f"fpn_{f}"] = y
        outs.append(y)
    multi_scale = outs[:len(tin)]                         # :362-365
    mask_features = F.conv2d(outs[-1], sd[p + "mask_features.weight"], sd[p + "mask_features.bias"])
    return mask_features, multi_scale


# --------------------------------------------------------------------------------------
# Transformer decoder  (transformer_decoder/mask2former_transformer_decoder.py)
# --------------------------------------------------------------------------------------


def multihead_attention(sd, p, query, key, value, nheads, attn_mask=None):
    """torch.nn.MultiheadAttention forward (batch_first=False, separate q/k/v inputs) as used at
    mask2former_transformer_decoder.py:52-53,110-113: in_proj split in thirds, q scaled by
    head_dim**-0.5, boolean attn_mask True = not allowed (-inf), softmax, out_proj.
    query (Lq,B,E), key/value (Lk,B,E); attn_mask (B*nheads, Lq, Lk) bool."""
    Lq, B, E = query.shape
    Lk = key.shape[0]
    hd = E // nheads
    w, b = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    q = F.linear(query, w[:E], b[:E])
    k = F.linear(key, w[E:2 * E], b[E:2 * E])
    v = F.linear(value, w[2 * E:], b[2 * E:])
    q = q.view(Lq, B * nheads, hd).transpose(0, 1) * (hd ** -0.5)
    k = k.view(Lk, B * nheads, hd).transpose(0, 1)
    v = v.view(Lk, B * nheads, hd).transpose(0, 1)
    attn = q @ k.transpose(1, 2)
    if attn_mask is not None:
        attn = attn.masked_fill(attn_mask, float("-inf"))
    attn = attn.softmax(-1)
    o = (attn @ v).transpose(0, 1).reshape(Lq, B, E)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def prediction_heads(sd, p, output, mask_features, target_size, nheads, taps=None):
    """mask2former_transformer_decoder.py:472-489"""
    D = output.shape[-1]
    dec = F.layer_norm(output, (D,), sd[p + "decoder_norm.weight"], sd[p + "decoder_norm.bias"]).transpose(0, 1)
    cls = F.linear(dec, sd[p + "class_embed.weight"], sd[p + "class_embed.bias"])
    me = dec
    for i in range(3):
        me = F.linear(me, sd[f"{p}mask_embed.layers.{i}.weight"], sd[f"{p}mask_embed.layers.{i}.bias"])
        if i < 2:
            me = F.relu(me)
    masks = torch.einsum("bqc,bchw->bqhw", me, mask_features)
    am = F.interpolate(masks, size=target_size, mode="bilinear", align_corners=False)
    if taps is not None:   # distance of the closest boolean decision (sigmoid(x) < 0.5 <=> x < 0) from its threshold
        taps["am_margin"] = min(taps.get("am_margin", float("inf")), float(am.abs().min()))
        taps.setdefault("head_masks", []).append(masks)            # every head whose decisions feed a cross-attention
        taps.setdefault("head_sizes", []).append(tuple(target_size))
    am = (am.sigmoid().flatten(2).unsqueeze(1).repeat(1, nheads, 1, 1).flatten(0, 1) < 0.5).bool()
    return cls, masks, am


def transformer_decoder_forward(sd, cfg, multi_scale, mask_features, p="sem_seg_head.predictor.", taps=None):
    """mask2former_transformer_decoder.py:398-470 (post-norm, dropout 0)."""
    nl = len(multi_scale)
    D = cfg.hidden_dim
    src, pos, sizes = [], [], []
    for i in range(nl):
        x = multi_scale[i]
        sizes.append(x.shape[-2:])
        pos.append(position_embedding_sine(x.shape[0], x.shape[2], x.shape[3], D // 2, device=x.device).flatten(2).permute(2, 0, 1))
        if (p + f"input_proj.{i}.weight") in sd:          # :353-358 (empty Sequential when in_channels == hidden_dim)
            x = F.conv2d(x, sd[p + f"input_proj.{i}.weight"], sd[p + f"input_proj.{i}.bias"])
        src.append((x.flatten(2) + sd[p + "level_embed.weight"][i][None, :, None]).permute(2, 0, 1))
    bs = src[0].shape[1]
    qe = sd[p + "query_embed.weight"].unsqueeze(1).repeat(1, bs, 1)
    out = sd[p + "query_feat.weight"].unsqueeze(1).repeat(1, bs, 1)
    cls, masks, am = prediction_heads(sd, p, out, mask_features, sizes[0], cfg.nheads, taps)
    if taps is not None:
        taps["head0_logits"], taps["head0_masks"] = cls, masks
    for i in range(cfg.dec_layers):
        lvl = i % nl
        am = am.clone()
        am[torch.where(am.sum(-1) == am.shape[-1])] = False       # :433
        q = f"{p}transformer_cross_attention_layers.{i}."
        t2 = multihead_attention(sd, q + "multihead_attn.", out + qe, src[lvl] + pos[lvl], src[lvl], cfg.nheads, am)
        out = F.layer_norm(out + t2, (D,), sd[q + "norm.weight"], sd[q + "norm.bias"])
        if taps is not None:
            taps[f"dec{i}_cross"] = out
        q = f"{p}transformer_self_attention_layers.{i}."
        t2 = multihead_attention(sd, q + "self_attn.", out + qe, out + qe, out, cfg.nheads)
        out = F.layer_norm(out + t2, (D,), sd[q + "norm.weight"], sd[q + "norm.bias"])
        q = f"{p}transformer_ffn_layers.{i}."
        t2 = F.linear(F.relu(F.linear(out, sd[q + "linear1.weight"], sd[q + "linear1.bias"])),
                      sd[q + "linear2.weight"], sd[q + "linear2.bias"])
        out = F.layer_norm(out + t2, (D,), sd[q + "norm.weight"], sd[q + "norm.bias"])
        if taps is not None:
            taps[f"dec{i}_out"] = out
        cls, masks, am = prediction_heads(sd, p, out, mask_features, sizes[(i + 1) % nl], cfg.nheads,
                                          taps if i + 1 < cfg.dec_layers else None)   # the last mask is never used
    return cls, masks


# --------------------------------------------------------------------------------------
# Meta-arch + score  (mask2former/maskformer_model.py, evaluate_ood.py)
# --------------------------------------------------------------------------------------


def semantic_inference(mask_cls, mask_pred):
    """maskformer_model.py:381-386"""
    mask_cls = F.softmax(mask_cls, dim=-1)[..., :-1]
    return torch.einsum("qc,qhw->chw", mask_cls, mask_pred.sigmoid())


def rba_from_sem_seg(sem_seg):
    """evaluate_ood.py:143-150 (get_RbA): -tanh(logits).sum(0)"""
    return -sem_seg.tanh().sum(dim=0)


def score_from_head_outputs(pred_logits, pred_masks, padded_hw, image_hw):
    """maskformer_model.py:294-299 (x4 bilinear up to padded size), :381-386, sem_seg_postprocess crop
    (:330-333; resize is the identity because evaluate_ood.py passes no height/width), evaluate_ood.py:148-150.
    pred_logits (Q,K+1), pred_masks (Q,h,w) for ONE image -> (sem_seg (K,H,W), rba (H,W))."""
    up = F.interpolate(pred_masks[None], size=padded_hw, mode="bilinear", align_corners=False)[0]
    sem = semantic_inference(pred_logits, up)[:, :image_hw[0], :image_hw[1]]
    return sem, rba_from_sem_seg(sem)


def ood_pred_head(sd, mask_features, image_size, p="sem_seg_head.predictor.ood_pred."):
    """DenseHybrid head: BNReluConv(hidden_dim, 2, k=1, bias=True) on mask_features
    (mask2former_transformer_decoder.py:216-230,365-366,467-468: BatchNorm2d in eval mode -> ReLU -> 1x1 conv), then
    F.interpolate(..., size=images.image_sizes[0], mode='bilinear', align_corners=True) (maskformer_model.py:303-305).
    Returns (B,2,H,W)."""
    x = F.batch_norm(mask_features, sd[p + "norm.running_mean"], sd[p + "norm.running_var"], sd[p + "norm.weight"],
                     sd[p + "norm.bias"], training=False, eps=1e-5)
    x = F.conv2d(F.relu(x), sd[p + "conv.weight"], sd[p + "conv.bias"])
    return F.interpolate(x, size=tuple(image_size), mode="bilinear", align_corners=True)


def densehybrid_from_outputs(sem_seg, ood_pred):
    """evaluate_ood.py:161-173 get_densehybrid_score for ONE image: sem_seg (K,H,W), ood_pred (2,H,W) -> (H,W)."""
    p2 = F.softmax(ood_pred, dim=0)[1]
    return -torch.logsumexp(sem_seg, dim=0) + (p2 + 1e-9).log()


def outlier_loss(pred_masks, pred_logits, outlier_masks, target="nls", score_norm="tanh", t_in=-1.0, t_out=-0.1):
    """SetCriterion.outlier_loss, squared-hinge branch (mask2former/modeling/criterion.py:435-487):
    class softmax without void x mask sigmoid contracted over queries (:448-453), score = -sum_c f(.) or -logsumexp
    (:455-466), bilinear align_corners=True resize to the label size (:474-475), squared hinge on the inlier / outlier
    pixels with the 0.5 only when outliers exist (:480-487).  Differentiable torch statement (autograd = the reference's
    backward)."""
    ood, ind = outlier_masks == 1, outlier_masks == 0
    logits = torch.einsum("bqc,bqhw->bchw", F.softmax(pred_logits, dim=-1)[..., :-1], pred_masks.sigmoid())
    if target == "nls":
        f = {"tanh": torch.tanh, "sigmoid": torch.sigmoid}.get(score_norm, lambda t: t)
        score = -f(logits).sum(dim=1)
    elif target == "energy":
        score = -torch.logsumexp(logits, dim=1)
    else:
        raise ValueError(target)
    score = F.interpolate(score.unsqueeze(1), size=outlier_masks.shape[-2:], mode="bilinear", align_corners=True).squeeze(1)
    loss = F.relu(score[ind] - t_in).pow(2).mean()
    if ood.sum() > 0:
        loss = 0.5 * (loss + F.relu(t_out - score[ood]).pow(2).mean())
    return loss


def preprocess(images, cfg):
    """maskformer_model.py:255-257: (x - mean)/std per image, zero-pad bottom/right to a multiple of
    SIZE_DIVISIBILITY (detectron2 ImageList.from_tensors, pad_value 0)."""
    dev = images[0].device
    mean = torch.tensor(cfg.pixel_mean, device=dev).view(-1, 1, 1)
    std = torch.tensor(cfg.pixel_std, device=dev).view(-1, 1, 1)
    ims = [(x.float() - mean) / std for x in images]
    sizes = [(x.shape[-2], x.shape[-1]) for x in ims]
    s = cfg.size_divisibility
    Hm = max(h for h, _ in sizes)
    Wm = max(w for _, w in sizes)
    if s > 1:
        Hm, Wm = (Hm + s - 1) // s * s, (Wm + s - 1) // s * s
    out = torch.zeros(len(ims), 3, Hm, Wm, device=dev)
    for i, x in enumerate(ims):
        out[i, :, :x.shape[-2], :x.shape[-1]] = x
    return out, sizes


@torch.no_grad()
def forward(sd, cfg, images, explicit_msda=False, want_taps=False):
    """MaskFormer.forward eval branch (maskformer_model.py:227-356) + get_RbA.
    images: list of (3,H,W) uint8/float tensors.  Returns dict with per-image lists
    `sem_seg`, `rba`, plus batched `pred_logits`, `pred_masks` and (optionally) taps."""
    taps = {} if want_taps else None
    x, sizes = preprocess(images, cfg)
    feats = resnet_forward(sd, cfg, x) if getattr(cfg, "backbone", "swin") == "resnet" else swin_forward(sd, cfg, x)
    mask_features, multi_scale = pixel_decoder_forward(sd, cfg, feats, explicit_msda=explicit_msda, taps=taps)
    cls, masks = transformer_decoder_forward(sd, cfg, multi_scale, mask_features, taps=taps)
    sem_seg, rba = [], []
    for b in range(len(images)):
        s, r = score_from_head_outputs(cls[b], masks[b], x.shape[-2:], sizes[b])
        sem_seg.append(s)
        rba.append(r)
    res = {"pred_logits": cls, "pred_masks": masks, "sem_seg": sem_seg, "rba": rba}
    if "sem_seg_head.predictor.ood_pred.conv.weight" in sd:
        res["ood_pred"] = ood_pred_head(sd, mask_features, sizes[0])
        res["densehybrid"] = [densehybrid_from_outputs(sem_seg[b], res["ood_pred"][b]) for b in range(len(images))]
    if want_taps:
        taps.update(feats)
        taps["mask_features"] = mask_features
        taps["multi_scale0"] = multi_scale[0]
        res["taps"] = taps
    return res


# --------------------------------------------------------------------------------------
# synthetic weights
# --------------------------------------------------------------------------------------


def perturb_state_dict(sd, seed=1234, scale=0.02):
    """The reference's random init leaves many tensors at 0 / 1 (biases, LayerNorm/GroupNorm affine,
    MSDeformAttn offset/attention weights: ms_deform_attn.py:66-80), which would hide indexing
    bugs.  Adds seeded N(0, scale) noise to every floating tensor so that both implementations
    (which load the SAME state_dict) are exercised on non-degenerate weights."""
    g = torch.Generator().manual_seed(seed)
    out = type(sd)()
    if hasattr(sd, "_metadata"):      # keeps module versions so the reference's legacy-key upgrade
        out._metadata = sd._metadata  # (mask_former_head.py:31-53) does not fire on reload
    for k, v in sd.items():
        if v.is_floating_point():
            out[k] = (v + scale * torch.randn(v.shape, generator=g)).contiguous()
        else:
            out[k] = v.clone()
    return out

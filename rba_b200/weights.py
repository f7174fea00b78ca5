"""Reference-layout parameter inventory and random initialisation.

Key names and shapes are exactly those of the reference's `MaskFormer.state_dict()` for the built
architectures (Swin backbone swin.py:559-614, MSDeformAttnPixelDecoder msdeformattn.py:221-301,
MultiScaleMaskedTransformerDecoder mask2former_transformer_decoder.py:302-366), so reference checkpoints
load unchanged and `state_dict()` round-trips through DetectionCheckpointer."""
import math
from collections import OrderedDict

import torch


def relative_position_index(ws):
    """swin.py:110-120"""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def param_specs(mc):
    """OrderedDict name -> (shape, init kind) in the reference's registration order."""
    S = OrderedDict()
    C0, ws, D = mc.embed_dim, mc.window_size, mc.conv_dim
    FC = mc.feature_channels
    if mc.backbone == "resnet":
        _resnet_specs(S, mc)
    else:
        _swin_specs(S, mc)
    _head_specs(S, mc, FC)
    return S


RESNET_BLOCKS = {50: [3, 4, 6, 3], 101: [3, 4, 23, 3]}


def _resnet_specs(S, mc):
    """detectron2 ResNet state_dict names (= tools/convert-torchvision-to-d2.py:33-44 applied to torchvision's), batch norm
    with running statistics (RESNETS.NORM "SyncBN" at training time; an affine map in eval)."""
    def conv_bn(p, cout, cin, k):
        S[p + "weight"] = ((cout, cin, k, k), "msra")
        S[p + "norm.weight"] = ((cout,), "ones")
        S[p + "norm.bias"] = ((cout,), "zeros")
        S[p + "norm.running_mean"] = ((cout,), "bn_mean")
        S[p + "norm.running_var"] = ((cout,), "bn_var")
        S[p + "norm.num_batches_tracked"] = ((), "counter")
    conv_bn("backbone.stem.conv1.", 64, 3, 7)
    cin, width = 64, 64
    for i, n in enumerate(RESNET_BLOCKS[mc.resnet_depth]):
        cout = width * 4
        for j in range(n):
            p = f"backbone.res{i + 2}.{j}."
            if cin != cout:
                conv_bn(p + "shortcut.", cout, cin, 1)
            conv_bn(p + "conv1.", width, cin, 1)
            conv_bn(p + "conv2.", width, width, 3)
            conv_bn(p + "conv3.", cout, width, 1)
            cin = cout
        width *= 2


def _swin_specs(S, mc):
    C0, ws = mc.embed_dim, mc.window_size
    S["backbone.patch_embed.proj.weight"] = ((C0, 3, 4, 4), "conv")
    S["backbone.patch_embed.proj.bias"] = ((C0,), "conv_bias:48")
    S["backbone.patch_embed.norm.weight"] = ((C0,), "ones")
    S["backbone.patch_embed.norm.bias"] = ((C0,), "zeros")
    for i, depth in enumerate(mc.depths):
        C = C0 << i
        for j in range(depth):
            p = f"backbone.layers.{i}.blocks.{j}."
            S[p + "norm1.weight"] = ((C,), "ones")
            S[p + "norm1.bias"] = ((C,), "zeros")
            S[p + "attn.relative_position_bias_table"] = (((2 * ws - 1) ** 2, mc.num_heads[i]), "trunc02")
            S[p + "attn.relative_position_index"] = ((ws * ws, ws * ws), "rel_index")
            S[p + "attn.qkv.weight"] = ((3 * C, C), "linear")
            S[p + "attn.qkv.bias"] = ((3 * C,), f"linear_bias:{C}")
            S[p + "attn.proj.weight"] = ((C, C), "linear")
            S[p + "attn.proj.bias"] = ((C,), f"linear_bias:{C}")
            S[p + "norm2.weight"] = ((C,), "ones")
            S[p + "norm2.bias"] = ((C,), "zeros")
            S[p + "mlp.fc1.weight"] = ((4 * C, C), "linear")
            S[p + "mlp.fc1.bias"] = ((4 * C,), f"linear_bias:{C}")
            S[p + "mlp.fc2.weight"] = ((C, 4 * C), "linear")
            S[p + "mlp.fc2.bias"] = ((C,), f"linear_bias:{4 * C}")
        if i < 3:
            p = f"backbone.layers.{i}.downsample."
            S[p + "reduction.weight"] = ((2 * C, 4 * C), "linear")
            S[p + "norm.weight"] = ((4 * C,), "ones")
            S[p + "norm.bias"] = ((4 * C,), "zeros")
    for i in range(4):
        S[f"backbone.norm{i}.weight"] = ((C0 << i,), "ones")
        S[f"backbone.norm{i}.bias"] = ((C0 << i,), "zeros")


def _head_specs(S, mc, FC):
    D = mc.conv_dim
    pd = "sem_seg_head.pixel_decoder."
    L = mc.num_enc_levels
    for idx in range(L):
        Cin = FC[3 - idx]
        S[f"{pd}input_proj.{idx}.0.weight"] = ((D, Cin, 1, 1), "xavier")
        S[f"{pd}input_proj.{idx}.0.bias"] = ((D,), "zeros")
        S[f"{pd}input_proj.{idx}.1.weight"] = ((D,), "ones")
        S[f"{pd}input_proj.{idx}.1.bias"] = ((D,), "zeros")
    S[pd + "transformer.level_embed"] = ((L, D), "normal")
    M, P = mc.enc_heads, mc.enc_points
    for i in range(mc.enc_layers):
        p = f"{pd}transformer.encoder.layers.{i}."
        S[p + "self_attn.sampling_offsets.weight"] = ((M * L * P * 2, D), "zeros")
        S[p + "self_attn.sampling_offsets.bias"] = ((M * L * P * 2,), f"msda_grid:{M}:{L}:{P}")
        S[p + "self_attn.attention_weights.weight"] = ((M * L * P, D), "zeros")
        S[p + "self_attn.attention_weights.bias"] = ((M * L * P,), "zeros")
        S[p + "self_attn.value_proj.weight"] = ((D, D), "xavier")
        S[p + "self_attn.value_proj.bias"] = ((D,), "zeros")
        S[p + "self_attn.output_proj.weight"] = ((D, D), "xavier")
        S[p + "self_attn.output_proj.bias"] = ((D,), "zeros")
        S[p + "norm1.weight"] = ((D,), "ones")
        S[p + "norm1.bias"] = ((D,), "zeros")
        S[p + "linear1.weight"] = ((mc.enc_ffn, D), "xavier")
        S[p + "linear1.bias"] = ((mc.enc_ffn,), f"linear_bias:{D}")
        S[p + "linear2.weight"] = ((D, mc.enc_ffn), "xavier")
        S[p + "linear2.bias"] = ((D,), f"linear_bias:{mc.enc_ffn}")
        S[p + "norm2.weight"] = ((D,), "ones")
        S[p + "norm2.bias"] = ((D,), "zeros")
    S[pd + "mask_features.weight"] = ((mc.mask_dim, D, 1, 1), "c2_xavier")
    S[pd + "mask_features.bias"] = ((mc.mask_dim,), "zeros")
    num_fpn = 3 if L == 1 else 1
    for k in range(1, num_fpn + 1):
        Cin = FC[k - 1]
        S[f"{pd}adapter_{k}.weight"] = ((D, Cin, 1, 1), "c2_xavier")
        S[f"{pd}adapter_{k}.norm.weight"] = ((D,), "ones")
        S[f"{pd}adapter_{k}.norm.bias"] = ((D,), "zeros")
        S[f"{pd}layer_{k}.weight"] = ((D, D, 3, 3), "c2_xavier")
        S[f"{pd}layer_{k}.norm.weight"] = ((D,), "ones")
        S[f"{pd}layer_{k}.norm.bias"] = ((D,), "zeros")
    pr = "sem_seg_head.predictor."
    for kind, attn in (("self", "self_attn"), ("cross", "multihead_attn")):
        for i in range(mc.dec_layers):
            p = f"{pr}transformer_{kind}_attention_layers.{i}."
            S[p + attn + ".in_proj_weight"] = ((3 * D, D), "xavier")
            S[p + attn + ".in_proj_bias"] = ((3 * D,), "zeros")
            S[p + attn + ".out_proj.weight"] = ((D, D), "xavier")
            S[p + attn + ".out_proj.bias"] = ((D,), "zeros")
            S[p + "norm.weight"] = ((D,), "ones")
            S[p + "norm.bias"] = ((D,), "zeros")
    for i in range(mc.dec_layers):
        p = f"{pr}transformer_ffn_layers.{i}."
        S[p + "linear1.weight"] = ((mc.dim_feedforward, D), "xavier")
        S[p + "linear1.bias"] = ((mc.dim_feedforward,), f"linear_bias:{D}")
        S[p + "linear2.weight"] = ((D, mc.dim_feedforward), "xavier")
        S[p + "linear2.bias"] = ((D,), f"linear_bias:{mc.dim_feedforward}")
        S[p + "norm.weight"] = ((D,), "ones")
        S[p + "norm.bias"] = ((D,), "zeros")
    S[pr + "decoder_norm.weight"] = ((D,), "ones")
    S[pr + "decoder_norm.bias"] = ((D,), "zeros")
    S[pr + "query_feat.weight"] = ((mc.num_queries, D), "normal")
    S[pr + "query_embed.weight"] = ((mc.num_queries, D), "normal")
    S[pr + "level_embed.weight"] = ((L, D), "normal")
    S[pr + "class_embed.weight"] = ((mc.num_classes + 1, D), "linear")
    S[pr + "class_embed.bias"] = ((mc.num_classes + 1,), f"linear_bias:{D}")
    for i in range(3):
        S[f"{pr}mask_embed.layers.{i}.weight"] = ((mc.mask_dim if i == 2 else D, D), "linear")
        S[f"{pr}mask_embed.layers.{i}.bias"] = ((mc.mask_dim if i == 2 else D,), f"linear_bias:{D}")
    if getattr(mc, "ood_prediction", False):   # BNReluConv(hidden_dim, 2, k=1, bias=True), mask2former_transformer_decoder.py:216-230,365-366
        S[pr + "ood_pred.norm.weight"] = ((D,), "ones")
        S[pr + "ood_pred.norm.bias"] = ((D,), "zeros")
        S[pr + "ood_pred.norm.running_mean"] = ((D,), "zeros")
        S[pr + "ood_pred.norm.running_var"] = ((D,), "bn_var")
        S[pr + "ood_pred.norm.num_batches_tracked"] = ((), "counter")
        S[pr + "ood_pred.conv.weight"] = ((2, D, 1, 1), "conv")
        S[pr + "ood_pred.conv.bias"] = ((2,), f"conv_bias:{D}")
    S["criterion.empty_weight"] = ((mc.num_classes + 1,), "ones")  # training buffer kept for key parity


def _fans(shape):
    rf = 1
    for s in shape[2:]:
        rf *= s
    return shape[1] * rf, shape[0] * rf


def init_state_dict(mc, seed=0, perturb=0.0):
    """Random init with the reference's distributions (nn.Linear/Conv defaults, trunc_normal 0.02 for the bias
    tables, xavier_uniform for the transformer, MSDeformAttn._reset_parameters ms_deform_attn.py:66-80).
    `perturb` > 0 adds N(0, perturb) to every float tensor so that zero/one-initialised tensors (biases,
    norm affines, sampling-offset weights) are exercised too."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    for name, (shape, kind) in param_specs(mc).items():
        if kind == "ones":
            t = torch.ones(shape)
        elif kind == "zeros":
            t = torch.zeros(shape)
        elif kind == "rel_index":
            sd[name] = relative_position_index(mc.window_size)
            continue
        elif kind == "counter":
            sd[name] = torch.zeros((), dtype=torch.int64)
            continue
        elif kind == "bn_var":                                   # positive: 1 + |N(0, 10 perturb)|
            t = torch.ones(shape) + (10.0 * perturb * torch.randn(shape, generator=g)).abs()
            sd[name] = t.float().contiguous()
            continue
        elif kind == "bn_mean":                                  # running mean: 0 (+ perturbation below)
            t = torch.zeros(shape)
        elif kind == "msra":                                     # c2_msra_fill: kaiming_normal_(mode="fan_out", relu)
            _, fan_out = _fans(shape)
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out)
        elif kind == "trunc02":
            t = torch.nn.init.trunc_normal_(torch.empty(shape), std=0.02, generator=g)
        elif kind == "normal":
            t = torch.randn(shape, generator=g)
        elif kind in ("linear", "conv"):
            fan_in, _ = _fans(shape)
            t = uni(shape, 1.0 / math.sqrt(fan_in))            # kaiming_uniform(a=sqrt(5))
        elif kind.startswith("linear_bias:") or kind.startswith("conv_bias:"):
            t = uni(shape, 1.0 / math.sqrt(int(kind.split(":")[1])))
        elif kind == "xavier":
            fan_in, fan_out = _fans(shape)
            t = uni(shape, math.sqrt(6.0 / (fan_in + fan_out)))
        elif kind == "c2_xavier":
            fan_in, _ = _fans(shape)
            t = uni(shape, math.sqrt(3.0 / fan_in))            # kaiming_uniform(a=1)
        elif kind.startswith("msda_grid:"):
            M, L, P = (int(v) for v in kind.split(":")[1:])
            th = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
            grid = torch.stack([th.cos(), th.sin()], -1)
            grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(M, 1, 1, 2).repeat(1, L, P, 1)
            for i in range(P):
                grid[:, :, i, :] *= i + 1
            t = grid.reshape(-1)
        else:
            raise KeyError(kind)
        if perturb > 0:
            t = t + perturb * torch.randn(shape, generator=g)
        sd[name] = t.float().contiguous()
    return sd

"""GPU parity tests, kernel by kernel, through the C ABI (rba_b200.ops -> librba_b200.so), against the oracle /
plain fp32-or-better PyTorch CPU statements of the same reference op.  Tolerances are written per test;
the end-to-end bar is 1e-3 (north star), kernels are held to round-off."""
import math

import pytest
import torch
import torch.nn.functional as F

import rba_oracle as O
from conftest import load_golden
from rba_b200 import ops

pytestmark = pytest.mark.gpu


def planes(x, dev):
    """CPU fp32 -> (hi, lo) bf16 planes on the device + the exact fp32 value they represent."""
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    return (hi.to(dev).contiguous(), lo.to(dev).contiguous()), hi.float() + lo.float()


def unplanes(p):
    return (p[0].float() + p[1].float()).cpu()


def rel_err(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def test_split_planes(dev):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(37, 64, generator=g) * torch.logspace(-6, 4, 64)
    hi, lo = ops.split_planes(x.to(dev))
    back = unplanes((hi, lo))
    assert ((back - x).abs() <= x.abs() * 2.0 ** -16 + 1e-30).all()
    assert torch.equal(hi.cpu(), x.bfloat16())


@pytest.mark.parametrize("C", [32, 128, 192, 1024])
def test_layernorm_plain(dev, C):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(77, C, generator=g) * 3 + 1
    gm, bt = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ref = F.layer_norm(x, (C,), gm, bt)
    y, p = ops.layernorm(x.to(dev), gm.to(dev), bt.to(dev), mode=0, B=1, H=1, W=77, want_planes=True)
    assert (y.cpu() - ref).abs().max() < 2e-5
    assert (unplanes(p) - ref).abs().max() < 2e-4


def test_layernorm_narrow_rows_many(dev):
    """The 4-rows-per-warp kernel used for C <= 128 and >= 256 rows (stage-0 LayerNorms), row count not a multiple of 4."""
    for C, rows in ((128, 1001), (32, 258), (96, 4099)):
        g = torch.Generator().manual_seed(C)
        x = torch.randn(rows, C, generator=g) * 3 + 1
        gm, bt = torch.randn(C, generator=g), torch.randn(C, generator=g)
        ref = F.layer_norm(x, (C,), gm, bt)
        y, p = ops.layernorm(x.to(dev), gm.to(dev), bt.to(dev), mode=0, B=1, H=1, W=rows, want_planes=True)
        assert (y.cpu() - ref).abs().max() < 2e-5
        assert (unplanes(p) - ref).abs().max() < 2e-4


@pytest.mark.parametrize("H,W,shift", [(24, 36, 0), (30, 41, 6), (7, 50, 6), (12, 12, 0)])
def test_layernorm_window_gather(dev, H, W, shift):
    """swin.py:247-271: norm1, pad, roll(-shift), window_partition."""
    B, C, ws = 2, 64, 12
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, H * W, C, generator=g)
    gm, bt = torch.randn(C, generator=g), torch.randn(C, generator=g)
    xn = F.layer_norm(x, (C,), gm, bt).view(B, H, W, C)
    xn = F.pad(xn, (0, 0, 0, (ws - W % ws) % ws, 0, (ws - H % ws) % ws))
    if shift:
        xn = torch.roll(xn, shifts=(-shift, -shift), dims=(1, 2))
    ref = O.window_partition(xn, ws).reshape(-1, C)
    y = ops.layernorm(x.view(-1, C).to(dev), gm.to(dev), bt.to(dev), mode=1, B=B, H=H, W=W, ws=ws, shift=shift)
    assert y.shape == ref.shape
    assert (y.cpu() - ref).abs().max() < 2e-5


def test_layernorm_patch_merging(dev):
    """swin.py:327-334"""
    B, H, W, C = 2, 6, 10, 32
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, H * W, C, generator=g)
    gm, bt = torch.randn(4 * C, generator=g), torch.randn(4 * C, generator=g)
    xv = x.view(B, H, W, C)
    cat = torch.cat([xv[:, 0::2, 0::2], xv[:, 1::2, 0::2], xv[:, 0::2, 1::2], xv[:, 1::2, 1::2]], -1).view(-1, 4 * C)
    ref = F.layer_norm(cat, (4 * C,), gm, bt)
    y = ops.layernorm(x.view(-1, C).to(dev), gm.to(dev), bt.to(dev), mode=2, B=B, H=H, W=W)
    assert (y.cpu() - ref).abs().max() < 2e-5


@pytest.mark.parametrize("M,N,K", [(300, 20, 256), (129, 96, 128), (1000, 384, 64), (5, 1, 256), (257, 130, 2304)])
@pytest.mark.parametrize("act", [ops.RBA_ACT_NONE, ops.RBA_ACT_RELU, ops.RBA_ACT_GELU])
def test_gemm_ffma(dev, M, N, K, act):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    bias = torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g)
    ap, av = planes(a, dev)
    wp, wv = planes(w, dev)
    z = (av.double() @ wv.double().t() + bias.double()).float()
    z = {0: z, 1: F.relu(z), 2: F.gelu(z)}[act] + res
    c, cp = ops.gemm(ap, wp, bias=bias.to(dev), act=act, residual=res.to(dev), out_planes=True)
    assert (c.cpu() - z).abs().max() < 2e-5 * max(1.0, float(z.abs().max()))
    assert (unplanes(cp) - z).abs().max() < 1e-4 * max(1.0, float(z.abs().max()))


def test_gemm_batched_row_bias(dev):
    """mask einsum form: per-image A (Q,K) x per-image W (HW,K), bias per row (mask2former_transformer_decoder.py:479)."""
    Bn, Q, HW, K = 3, 100, 200, 64
    g = torch.Generator().manual_seed(5)
    a, w, b = torch.randn(Bn, Q, K, generator=g), torch.randn(Bn, HW, K, generator=g), torch.randn(Bn, Q, generator=g)
    ap, av = planes(a, dev)
    wp, wv = planes(w, dev)
    ref = (torch.einsum("bqk,bpk->bqp", av.double(), wv.double()) + b.double()[:, :, None]).float()
    c = ops.gemm(ap, wp, bias=b.to(dev), bias_per_row=True)
    assert c.shape == ref.shape
    assert (c.cpu() - ref).abs().max() < 5e-5


@pytest.mark.parametrize("H,W,shift", [(24, 36, 0), (30, 41, 6)])
def test_gemm_swin_scatter(dev, H, W, shift):
    """proj + window_reverse + roll back + crop + residual (swin.py:169,277-292)."""
    B, C, ws = 2, 64, 12
    nWh, nWw = -(-H // ws), -(-W // ws)
    rows = B * nWh * nWw * ws * ws
    g = torch.Generator().manual_seed(6)
    a = torch.randn(rows, C, generator=g)
    w = torch.randn(C, C, generator=g) / 8
    bias = torch.randn(C, generator=g)
    shortcut = torch.randn(B * H * W, C, generator=g)
    ap, av = planes(a, dev)
    wp, wv = planes(w, dev)
    y = (av.double() @ wv.double().t() + bias.double()).float()
    y = O.window_reverse(y.view(-1, ws, ws, C), ws, nWh * ws, nWw * ws)
    if shift:
        y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
    ref = shortcut + y[:, :H, :W, :].reshape(B * H * W, C)
    c = ops.gemm(ap, wp, bias=bias.to(dev), residual=shortcut.to(dev), swin=(B, H, W, ws, shift))
    assert (c.cpu() - ref).abs().max() < 5e-5


def test_conv3x3(dev):
    """msdeformattn.py:281-290 output_conv (bias-free), NHWC planes in, NHWC fp32 out."""
    B, H, W, Cin, Cout = 2, 9, 13, 32, 64
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, H, W, Cin, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin)
    xp, xv = planes(x, dev)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()       # k = (ky*3+kx)*Cin + ci
    wp, wkv = planes(wk, dev)
    wv = wkv.view(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
    ref = F.conv2d(xv.permute(0, 3, 1, 2).double(), wv.double(), padding=1).permute(0, 2, 3, 1).float()
    y = ops.conv3x3(xp, wp)
    assert (y.cpu() - ref).abs().max() < 5e-5


@pytest.mark.parametrize("H,W,shift", [(24, 24, 0), (30, 41, 6), (12, 12, 6)])
def test_window_attention(dev, H, W, shift):
    """swin.py:145-168 on given qkv."""
    B, heads, ws = 2, 2, 12
    C = heads * 32
    nWh, nWw = -(-H // ws), -(-W // ws)
    nW = nWh * nWw
    g = torch.Generator().manual_seed(8)
    qkv = torch.randn(B * nW, ws * ws, 3 * C, generator=g)
    table = torch.randn((2 * ws - 1) ** 2, heads, generator=g)
    q, k, v = qkv.reshape(B * nW, ws * ws, 3, heads, 32).permute(2, 0, 3, 1, 4)
    attn = (q * 32 ** -0.5) @ k.transpose(-2, -1)
    idx = O.relative_position_index(ws)
    attn = attn + table[idx.view(-1)].view(ws * ws, ws * ws, -1).permute(2, 0, 1).unsqueeze(0)
    if shift:
        mask = O.shift_attn_mask(H, W, ws, shift)
        attn = (attn.view(B, nW, heads, ws * ws, ws * ws) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, ws * ws, ws * ws)
    ref = (attn.softmax(-1) @ v).transpose(1, 2).reshape(B * nW * ws * ws, C)
    out = ops.window_attn(qkv.view(-1, 3 * C).to(dev), table.to(dev), B, H, W, C, heads, ws, shift)
    assert (unplanes(out) - ref).abs().max() < 1e-4


@pytest.mark.parametrize("Lk,masked,Lq", [(100, False, 100), (77, True, 100), (2048, True, 100), (5000, True, 100), (40000, True, 7), (600, False, 130)])
def test_mha(dev, Lk, masked, Lq):
    """nn.MultiheadAttention core as used by the decoder (mask2former_transformer_decoder.py:52-53,110-113): one and many key
    splits, ragged last tile, unaligned mask rows, more queries than one CTA holds."""
    B, E, heads = 2, 256, 8
    g = torch.Generator().manual_seed(9)
    q, k, v = torch.randn(B, Lq, E, generator=g), torch.randn(B, Lk, E, generator=g), torch.randn(B, Lk, E, generator=g)
    mask = None
    if masked:
        mask = torch.rand(B, Lq, Lk, generator=g) < 0.6
        mask[0, 3] = True
        mask[0, 3, 5] = False                   # a single open key
    qh = q.view(B, Lq, heads, 32).transpose(1, 2) * 32 ** -0.5
    kh = k.view(B, Lk, heads, 32).transpose(1, 2)
    vh = v.view(B, Lk, heads, 32).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    if masked:
        s = s.masked_fill(mask[:, None], float("-inf"))
    ref = (s.softmax(-1) @ vh).transpose(1, 2).reshape(B * Lq, E)
    out = ops.mha(q.to(dev), k.to(dev), v.to(dev), mask.to(torch.uint8).to(dev) if masked else None, heads)
    assert (unplanes(out) - ref).abs().max() < 1e-4


@pytest.mark.parametrize("with_prev,relu", [(False, False), (True, False), (False, True)])
def test_groupnorm_fused(dev, with_prev, relu):
    """GroupNorm(32) (+ bilinear x2 up of the previous FPN level) (+ ReLU), msdeformattn.py:356-360."""
    B, H, W, C = 2, 10, 14, 256
    g = torch.Generator().manual_seed(10)
    x = torch.randn(B, H, W, C, generator=g) * 2 + 0.5
    gm, bt = torch.randn(C, generator=g), torch.randn(C, generator=g)
    prev = torch.randn(B, H // 2, W // 2, C, generator=g) if with_prev else None
    ref = F.group_norm(x.permute(0, 3, 1, 2), 32, gm, bt)
    if with_prev:
        ref = ref + F.interpolate(prev.permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=False)
    if relu:
        ref = F.relu(ref)
    ref = ref.permute(0, 2, 3, 1)
    y, p = ops.groupnorm(x.to(dev), gm.to(dev), bt.to(dev), prev=prev.to(dev) if with_prev else None, relu=relu,
                         want_planes=True)
    assert (y.cpu() - ref).abs().max() < 3e-5
    assert (unplanes(p) - ref).abs().max() < 2e-4


@pytest.mark.parametrize("Wp,C", [(64, 32), (56, 128), (96, 192)])   # 56: the last group of 8 tokens is partial; 192: two channel groups
@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32])
def test_patch_embed(dev, dtype, Wp, C):
    """maskformer_model.py:255-257 + swin.py:479-495"""
    B, H, W = 2, 30, 45
    Hp = 32
    g = torch.Generator().manual_seed(11)
    img = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=g).to(dtype)
    cw, cb = torch.randn(C, 3, 4, 4, generator=g) / 7, torch.randn(C, generator=g)
    gm, bt = torch.randn(C, generator=g), torch.randn(C, generator=g)
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    x = (img.float() - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    x = F.pad(x, (0, Wp - W, 0, Hp - H))
    ref = F.layer_norm(F.conv2d(x, cw, cb, stride=4).flatten(2).transpose(1, 2), (C,), gm, bt)
    tok = ops.patch_embed(img.to(dev), Hp, Wp, mean, std, cw.to(dev), cb.to(dev), gm.to(dev), bt.to(dev))
    assert (tok.cpu() - ref).abs().max() < 3e-5


@pytest.mark.parametrize("h,w,th,tw", [(32, 64, 4, 8), (16, 24, 4, 6), (16, 24, 8, 12), (16, 24, 2, 3)])
def test_attn_mask(dev, h, w, th, tw):
    B, Q = 2, 10
    g = torch.Generator().manual_seed(12)
    m = torch.randn(B, Q, h, w, generator=g)
    m[0, 1] = m[0, 1].abs() + 0.1               # all-open row
    m[1, 2] = -m[1, 2].abs() - 0.1              # all-blocked row -> reset to open (:433)
    am = F.interpolate(m, size=(th, tw), mode="bilinear", align_corners=False)
    ref = (am.sigmoid().flatten(2) < 0.5)
    ref[torch.where(ref.sum(-1) == ref.shape[-1])] = False
    out = ops.attn_mask(m.to(dev), (th, tw)).cpu().bool()
    # values whose interpolated logit is within float round-off of 0 may legitimately flip
    near0 = am.flatten(2).abs() < 1e-6
    assert ((out != ref) & ~near0).sum() == 0
    assert not out[1, 2].any() and not out[0, 1].any()


def test_msda_reference_test_shapes(dev):
    """Same shapes / seed / tolerance as the reference's own ops/test.py:24-63 check_forward_equal_with_pytorch_float."""
    fix = load_golden("msda.pt")
    for f in (fix, fix["big"]):
        shapes = f["shapes"]
        lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
        out = ops.ms_deform_attn_forward(f["value"].to(dev), shapes.to(dev), lsi.to(dev), f["loc"].to(dev), f["aw"].to(dev), 2)
        assert torch.allclose(out.cpu(), f["out"], rtol=1e-2, atol=1e-3)      # reference tolerance (ops/test.py:59)
        assert (out.cpu() - f["out"]).abs().max() < 1e-5                      # ours
    with pytest.raises(ops.RbaError):                                        # batch % im2col_step (ms_deform_attn_cuda.cu:55-57)
        f = fix["big"]
        v3 = torch.cat([f["value"], f["value"][:1]]).to(dev)
        l3 = torch.cat([f["loc"], f["loc"][:1]]).to(dev)
        a3 = torch.cat([f["aw"], f["aw"][:1]]).to(dev)
        ops.ms_deform_attn_forward(v3, f["shapes"], lsi, l3, a3, 2)


def _msda_grad_inputs(c):
    """Same recipe as oracle/make_golden_msda_grad.py::make_inputs (only the reference's gradients are stored)."""
    shapes = torch.as_tensor(c["shapes"], dtype=torch.long)
    g = torch.Generator().manual_seed(c["seed"])
    S = int((shapes[:, 0] * shapes[:, 1]).sum())
    lo, hi = c.get("loc_range", (0.0, 1.0))
    value = torch.rand(c["N"], S, c["M"], c["D"], generator=g) * c.get("value_scale", 0.01)
    loc = torch.rand(c["N"], c["Lq"], c["M"], c["L"], c["P"], 2, generator=g) * (hi - lo) + lo
    aw = torch.rand(c["N"], c["Lq"], c["M"], c["L"], c["P"], generator=g) + 1e-5
    aw = aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    go = torch.randn(c["N"], c["Lq"], c["M"] * c["D"], generator=g)
    return shapes, value, loc, aw, go


def test_msda_backward_matches_reference_gradients(dev):
    """ms_deform_attn_backward against float64 autograd gradients of the reference's own ms_deform_attn_core_pytorch
    (tests/golden/msda_grad.pt) at the shapes / channel counts of the reference's check_gradient_numerical
    (ops/test.py:66-88), directly and through the MSDeformAttnFunction autograd wrapper."""
    fix = load_golden("msda_grad.pt")
    for name, f in fix.items():
        shapes, value, loc, aw, go = _msda_grad_inputs(f["case"])
        assert abs(float(value.double().sum() + loc.double().sum() + go.double().sum()) - f["in_checksum"]) < 1e-6 * max(1.0, abs(f["in_checksum"]))
        lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
        step = 2
        gv, gl, ga = ops.ms_deform_attn_backward(value.to(dev), shapes, lsi, loc.to(dev), aw.to(dev), go.to(dev), step)
        for got, want, nm in ((gv, f["grad_value"], "grad_value"), (gl, f["grad_loc"], "grad_loc"), (ga, f["grad_aw"], "grad_aw")):
            err = (got.cpu() - want).abs().max().item()
            assert err <= 2e-5 * max(1.0, want.abs().max().item()), (name, nm, err)
        # shapes as DEVICE int64 tensors (what the reference's caller passes): the sync-free *_dev entry points
        gv2, gl2, ga2 = ops.ms_deform_attn_backward(value.to(dev), shapes.to(dev), lsi.to(dev), loc.to(dev), aw.to(dev), go.to(dev), step)
        assert torch.equal(gl2, gl) and torch.equal(ga2, ga) and (gv2 - gv).abs().max() <= 1e-6 * max(1.0, gv.abs().max().item())
        o_host = ops.ms_deform_attn_forward(value.to(dev), shapes, lsi, loc.to(dev), aw.to(dev), step)
        o_dev = ops.ms_deform_attn_forward(value.to(dev), shapes.to(dev), lsi.to(dev), loc.to(dev), aw.to(dev), step)
        assert torch.equal(o_host, o_dev)
        v, l, a = (t.to(dev).requires_grad_(True) for t in (value, loc, aw))
        out = ops.MSDeformAttnFunction.apply(v, shapes, lsi, l, a, step)
        assert (out.detach().cpu() - f["out"]).abs().max() < 1e-5 * max(1.0, f["out"].abs().max().item())
        out.backward(go.to(dev))
        assert torch.equal(l.grad, gl) and torch.equal(a.grad, ga)
        assert (v.grad - gv).abs().max() <= 1e-6 * max(1.0, gv.abs().max().item())     # atomics: order may differ
    with pytest.raises(ops.RbaError):
        ops.ms_deform_attn_backward(value.to(dev), shapes, lsi, loc.to(dev), aw.to(dev), go[:, :1].contiguous().to(dev), 2)


def _outlier_inputs(c):
    """Same recipe as oracle/make_golden_outlier_loss.py::make_inputs (only loss + gradients are stored)."""
    g = torch.Generator().manual_seed(c["seed"])
    masks = torch.randn(c["B"], c["Q"], c["h"], c["w"], generator=g) * 0.99 - 0.54
    logits = torch.randn(c["B"], c["Q"], c["K"] + 1, generator=g)
    r = torch.rand(c["B"], c["H"], c["W"], generator=g)
    labels = torch.full((c["B"], c["H"], c["W"]), 255, dtype=torch.int64)
    labels[r < 0.6] = 0
    labels[r > 1.0 - c["p_ood"]] = 1
    return masks, logits, labels


def test_outlier_loss_matches_reference_criterion(dev):
    """ops.outlier_loss (fused forward + backward) against the loss value and float64 autograd gradients of the
    reference's own SetCriterion.outlier_loss (tests/golden/outlier_loss.pt): every shipped configuration, int64 and
    uint8 labels, ignored pixels, no-outlier batches, non-integer resize ratios."""
    fix = load_golden("outlier_loss.pt")
    for name, f in fix.items():
        c = f["case"]
        masks, logits, labels = _outlier_inputs(c)
        assert abs(float(masks.double().sum() + logits.double().sum() + labels.double().sum()) - f["in_checksum"]) < 1e-6 * abs(f["in_checksum"])
        for lab in (labels, labels.to(torch.uint8)):
            m, l = masks.to(dev).requires_grad_(True), logits.to(dev).requires_grad_(True)
            loss = ops.outlier_loss(m, l, lab.to(dev), outlier_loss_target=c["target"], score_norm=c["norm"],
                                    inlier_upper_threshold=c["t_in"], outlier_lower_threshold=c["t_out"])
            assert abs(loss.item() - f["loss"]) <= 2e-5 * max(1.0, abs(f["loss"])), (name, loss.item(), f["loss"])
            (3.0 * loss).backward()                                   # upstream gradient is applied
            for got, want, nm in ((m.grad, f["d_masks"], "d_masks"), (l.grad, f["d_logits"], "d_logits")):
                err = (got.cpu() / 3.0 - want).abs().max().item()
                assert err <= 2e-5 * max(want.abs().max().item(), 1e-3), (name, nm, err, want.abs().max().item())
    # no gradient requested: loss only
    with torch.no_grad():
        loss2 = ops.outlier_loss(masks.to(dev), logits.to(dev), labels.to(dev), outlier_loss_target=c["target"], score_norm=c["norm"],
                                 inlier_upper_threshold=c["t_in"], outlier_lower_threshold=c["t_out"])
    assert abs(loss2.item() - f["loss"]) <= 2e-5 * max(1.0, abs(f["loss"]))
    # no in-distribution pixel at all: mean of an empty set is NaN, as in the reference
    nan_loss = ops.outlier_loss(masks.to(dev), logits.to(dev), torch.full_like(labels, 255).to(dev))
    assert torch.isnan(nan_loss)
    for bad in (dict(outlier_loss_func="mse"), dict(outlier_loss_target="softmax_entropy"), dict(score_norm="softplus")):
        with pytest.raises(ops.RbaError):
            ops.outlier_loss(masks.to(dev), logits.to(dev), labels.to(dev), **bad)


def test_score_golden_and_edges(dev):
    fix = load_golden("score.pt")
    for nm, f in fix.items():
        rba, sem = ops.score_fused(f["masks"].to(dev), f["logits"].to(dev), (f["H"], f["W"]), want_sem_seg=True)
        assert (sem.cpu() - f["sem_seg"]).abs().max() < 2e-5, nm
        assert (rba.cpu() - f["rba"]).abs().max() < 2e-5, nm
        rba2 = ops.score_fused(f["masks"].to(dev), f["logits"].to(dev), (f["H"], f["W"]))
        assert torch.equal(rba2, rba)
    # empty batch
    e = ops.score_fused(torch.empty(0, 100, 4, 4, device=dev), torch.empty(0, 100, 20, device=dev), (16, 16))
    assert e.shape == (0, 16, 16)
    with pytest.raises(ops.RbaError):
        ops.score_fused(torch.zeros(1, 100, 4, 4, device=dev), torch.zeros(1, 100, 20, device=dev), (17, 16))


def test_score_full_size_properties(dev):
    """BASELINE size (8 x 100 x 256 x 512 -> 8 x 1024 x 2048): size-independent properties + oracle on crops."""
    B, Q, K, h, w = 2, 100, 19, 256, 512
    g = torch.Generator().manual_seed(13)
    masks = (torch.randn(B, Q, h, w, generator=g) * 0.99 - 0.54)
    logits = torch.randn(B, Q, K + 1, generator=g)
    rba, sem = ops.score_fused(masks.to(dev), logits.to(dev), (4 * h, 4 * w), want_sem_seg=True)
    rba, sem = rba.cpu(), sem.cpu()
    assert rba.shape == (B, 4 * h, 4 * w) and torch.isfinite(rba).all()
    assert (rba <= 0).all() and (rba >= -K).all()
    assert (rba - (-sem.tanh().sum(1))).abs().max() < 1e-5            # rba is consistent with its own sem_seg
    p = logits.softmax(-1)[..., :-1]
    assert (sem.sum(1) <= p.sum(-1).sum(-1)[:, None, None] + 1e-3).all()  # sigmoid <= 1
    # oracle on random 32x32 output crops (bilinear x4 is local: a crop needs a 1-pixel low-res halo)
    for (b, y0, x0) in [(0, 0, 0), (1, 4 * h - 32, 4 * w - 32), (0, 512, 1000), (1, 100, 4)]:
        ly0, lx0 = max(y0 // 4 - 1, 0), max(x0 // 4 - 1, 0)
        ly1, lx1 = min((y0 + 32) // 4 + 1, h), min((x0 + 32) // 4 + 1, w)
        up = F.interpolate(masks[b:b + 1], size=(4 * h, 4 * w), mode="bilinear", align_corners=False)[0, :, y0:y0 + 32, x0:x0 + 32]
        ref = O.semantic_inference(logits[b], up)
        assert (sem[b, :, y0:y0 + 32, x0:x0 + 32] - ref).abs().max() < 2e-5
        assert ly1 > ly0 and lx1 > lx0
    # constant mask logits: sem_seg[c] = sum_q p[q,c] * sigmoid(m_q) everywhere
    mc = torch.linspace(-3, 3, Q).view(1, Q, 1, 1).expand(1, Q, 8, 8).contiguous()
    r, s = ops.score_fused(mc.to(dev), logits[:1].to(dev), (32, 32), want_sem_seg=True)
    expect = (p[0] * torch.sigmoid(torch.linspace(-3, 3, Q))[:, None]).sum(0)
    assert (s.cpu()[0] - expect[:, None, None]).abs().max() < 1e-5


def _einsum_score_reference(E, Fm, bias, logits, H, W):
    """fp64 einsum (mask2former_transformer_decoder.py:479) on the exact values the planes hold, then the oracle."""
    masks = torch.einsum("bqc,bhwc->bqhw", E.double(), Fm.double())
    if bias is not None:
        masks = masks + bias.double()[:, :, None, None]
    masks = masks.float()
    h, w = Fm.shape[1:3]
    sems, rbas = [], []
    for b in range(E.shape[0]):
        sem, r = O.score_from_head_outputs(logits[b], masks[b], (4 * h, 4 * w), (H, W))
        sems.append(sem)
        rbas.append(r)
    return torch.stack(sems), torch.stack(rbas), masks


@pytest.mark.parametrize("B,Q,K,D,h,w,crop", [(2, 100, 19, 256, 20, 37, (0, 0)), (1, 100, 19, 256, 8, 16, (3, 5)),
                                               (1, 37, 3, 64, 5, 3, (0, 1)), (3, 104, 13, 128, 15, 31, (1, 0)),
                                               (1, 100, 19, 256, 64, 128, (0, 0))])
def test_einsum_score_fused(dev, B, Q, K, D, h, w, crop):
    """Fused mask einsum + x4 upsample + semantic_inference + RbA (score_fused.cu) against the oracle on the same
    inputs: several tiles per image, image borders, cropped outputs, Q not a multiple of 8/16, batch > 1."""
    g = torch.Generator().manual_seed(B * 1000 + Q + h + w)
    E = torch.randn(B, Q, D, generator=g) / math.sqrt(D) * 2.0
    Fm = torch.randn(B, h, w, D, generator=g)
    bias = torch.randn(B, Q, generator=g) * 0.5 - 0.5
    logits = torch.randn(B, Q, K + 1, generator=g)
    H, W = 4 * h - crop[0], 4 * w - crop[1]
    (e_hi, e_lo), Ex = planes(E.view(B * Q, D), dev)
    (f_hi, f_lo), Fx = planes(Fm.view(-1, D), dev)
    e_pl = (e_hi.view(B, Q, D), e_lo.view(B, Q, D))
    f_pl = (f_hi.view(B, h, w, D), f_lo.view(B, h, w, D))
    sem_ref, rba_ref, masks = _einsum_score_reference(Ex.view(B, Q, D), Fx.view(B, h, w, D), bias, logits, H, W)
    rba, sem = ops.einsum_score_fused(e_pl, f_pl, logits.to(dev), (H, W), bias=bias.to(dev), want_sem_seg=True)
    assert (sem.cpu() - sem_ref).abs().max() < 5e-5, float((sem.cpu() - sem_ref).abs().max())
    assert (rba.cpu() - rba_ref).abs().max() < 5e-5
    for variant in (1, 2, 3):                   # RbA-only launch: mma.sync cells / tcgen05 score phase / runs in registers
        try:
            ops.set_fused_score_variant(variant)
            rba2 = ops.einsum_score_fused(e_pl, f_pl, logits.to(dev), (H, W), bias=bias.to(dev))
        finally:
            ops.set_fused_score_variant(0)
        assert (rba2 - rba).abs().max() < 2e-5, variant     # same value as the sem_seg launch, other rounding
        assert (rba2.cpu() - rba_ref).abs().max() < 5e-5, variant
    # same answer as the two-kernel path (GEMM -> pred_masks -> rba_score_fused)
    rba3 = ops.score_fused(masks.to(dev), logits.to(dev), (H, W))
    assert (rba3 - rba).abs().max() < 5e-5
    # no bias
    sem_ref0, rba_ref0, _ = _einsum_score_reference(Ex.view(B, Q, D), Fx.view(B, h, w, D), None, logits, H, W)
    rba0 = ops.einsum_score_fused(e_pl, f_pl, logits.to(dev), (H, W))
    assert (rba0.cpu() - rba_ref0).abs().max() < 5e-5


@pytest.mark.parametrize("variant", [3, 2, 1])
@pytest.mark.parametrize("scale,bias_shift", [(12.0, 0.0), (30.0, 0.0), (60.0, -20.0), (400.0, 50.0)])
def test_einsum_score_fused_large_logits(dev, variant, scale, bias_shift):
    """Mask logits far outside the comfortable range (|x| up to several hundred, sharp sign changes between neighbouring
    low-res pixels): the score kernels may clamp only INTERPOLATED logits, never the taps (a clamped tap shifts every
    output pixel it is interpolated with), and the product form 2^(x0 + j d) = 2^x0 (2^d)^j of the second-generation
    kernel must hand over to the exact path beyond |u| = 60.  Reference: fp64 interpolate -> sigmoid -> class sums -> tanh."""
    B, Q, K, D, h, w = 2, 100, 19, 64, 17, 33
    g = torch.Generator().manual_seed(int(scale) + variant)
    E = torch.randn(B, Q, D, generator=g) / math.sqrt(D) * scale
    Fm = torch.randn(B, h, w, D, generator=g)
    bias = torch.randn(B, Q, generator=g) + bias_shift
    logits = torch.randn(B, Q, K + 1, generator=g) * 2.0
    H, W = 4 * h, 4 * w - 2
    (e_hi, e_lo), Ex = planes(E.view(B * Q, D), dev)
    (f_hi, f_lo), Fx = planes(Fm.view(-1, D), dev)
    e_pl = (e_hi.view(B, Q, D), e_lo.view(B, Q, D))
    f_pl = (f_hi.view(B, h, w, D), f_lo.view(B, h, w, D))
    masks = torch.einsum("bqc,bhwc->bqhw", Ex.view(B, Q, D).double(), Fx.view(B, h, w, D).double()) + bias.double()[:, :, None, None]
    assert masks.abs().max() > 2.5 * scale
    up = F.interpolate(masks.float().double(), size=(4 * h, 4 * w), mode="bilinear", align_corners=False)[:, :, :H, :W]
    sem = torch.einsum("bqc,bqhw->bchw", logits.double().softmax(-1)[..., :-1], up.sigmoid())
    ref = -sem.tanh().sum(1)
    try:
        ops.set_fused_score_variant(variant)
        rba = ops.einsum_score_fused(e_pl, f_pl, logits.to(dev), (H, W), bias=bias.to(dev))
    finally:
        ops.set_fused_score_variant(0)
    # fp32 einsum of logits of magnitude `scale * 3` carries ~1e-7 relative error into the exponent
    tol = 5e-5 if scale <= 30 else 2e-4
    err = (rba.cpu().double() - ref).abs().max().item()
    assert err < tol, err


@pytest.mark.parametrize("variant", [3, 2, 1])
def test_einsum_score_fused_many_tiles_per_cta(dev, variant):
    """More tiles than SMs (each persistent CTA walks 3-4 tiles and crosses image boundaries): the deferred epilogue of a tile's
    last block, the per-image class-probability buffers and the alternating half block of the second-generation kernel.
    Reference: the same arithmetic in PyTorch fp32 on the GPU."""
    B, Q, K, D, h, w = 3, 100, 19, 64, 96, 160
    g = torch.Generator().manual_seed(11)
    E = torch.randn(B, Q, D, generator=g) / math.sqrt(D) * 3.0
    Fm = torch.randn(B, h, w, D, generator=g)
    bias = torch.randn(B, Q, generator=g) - 1.0
    logits = torch.randn(B, Q, K + 1, generator=g) * 2.0
    H, W = 4 * h - 3, 4 * w
    (e_hi, e_lo), Ex = planes(E.view(B * Q, D), dev)
    (f_hi, f_lo), Fx = planes(Fm.view(-1, D), dev)
    e_pl = (e_hi.view(B, Q, D), e_lo.view(B, Q, D))
    f_pl = (f_hi.view(B, h, w, D), f_lo.view(B, h, w, D))
    masks = (torch.einsum("bqc,bhwc->bqhw", Ex.view(B, Q, D).double().to(dev), Fx.view(B, h, w, D).double().to(dev))
             + bias.double().to(dev)[:, :, None, None]).float()
    ref = torch.empty(B, H, W, device=dev)
    for b in range(B):
        up = F.interpolate(masks[b:b + 1], size=(4 * h, 4 * w), mode="bilinear", align_corners=False)[0, :, :H, :W]
        sem = torch.einsum("qc,qhw->chw", logits[b].to(dev).softmax(-1)[:, :-1], up.sigmoid())
        ref[b] = -sem.tanh().sum(0)
    try:
        ops.set_fused_score_variant(variant)
        rba = ops.einsum_score_fused(e_pl, f_pl, logits.to(dev), (H, W), bias=bias.to(dev))
    finally:
        ops.set_fused_score_variant(0)
    err = (rba - ref).abs().max().item()
    assert err < 5e-5, err


@pytest.mark.parametrize("Q,K,void", [(96, 19, False), (91, 19, False), (85, 7, True), (16, 19, False), (5, 3, False), (104, 22, True),
                                      (100, 19, True)])
def test_einsum_score_fused_v3_query_counts(dev, Q, K, void):
    """Third-generation kernel (score_fused3.cu): every shape of its query loop -- whole k16 steps only (96, 16), a remainder
    above 8 that takes a full step (91), the k8 tail step with two (85, 104) and with one query per lane (100, 5; 5 has no full
    step at all) -- and semantic_inference_with_void (K + 1 class columns, up to the 24 the three class tiles hold).
    Reference: fp64 interpolate -> sigmoid -> class sums -> tanh on the values the planes hold."""
    B, D, h, w = 2, 64, 13, 21
    g = torch.Generator().manual_seed(Q * 31 + K)
    E = torch.randn(B, Q, D, generator=g) / math.sqrt(D) * 3.0
    Fm = torch.randn(B, h, w, D, generator=g)
    bias = torch.randn(B, Q, generator=g) - 0.5
    logits = torch.randn(B, Q, K + 1, generator=g) * 2.0
    H, W = 4 * h - 1, 4 * w - 3
    (e_hi, e_lo), Ex = planes(E.view(B * Q, D), dev)
    (f_hi, f_lo), Fx = planes(Fm.view(-1, D), dev)
    e_pl = (e_hi.view(B, Q, D), e_lo.view(B, Q, D))
    f_pl = (f_hi.view(B, h, w, D), f_lo.view(B, h, w, D))
    masks = torch.einsum("bqc,bhwc->bqhw", Ex.view(B, Q, D).double(), Fx.view(B, h, w, D).double()) + bias.double()[:, :, None, None]
    up = F.interpolate(masks.float().double(), size=(4 * h, 4 * w), mode="bilinear", align_corners=False)[:, :, :H, :W]
    probs = logits.double().softmax(-1)
    sem = torch.einsum("bqc,bqhw->bchw", probs if void else probs[..., :-1], up.sigmoid())
    ref = -sem.tanh().sum(1)
    out = {}
    for variant in (3, 1):
        try:
            ops.set_fused_score_variant(variant)
            out[variant] = ops.einsum_score_fused(e_pl, f_pl, logits.to(dev), (H, W), bias=bias.to(dev), include_void=void)
        finally:
            ops.set_fused_score_variant(0)
        err = (out[variant].cpu().double() - ref).abs().max().item()
        assert err < 5e-5, (variant, err)
    assert (out[3] - out[1]).abs().max() < 2e-5


def test_einsum_score_fused_energy_and_void(dev):
    """Other score heads on the same kernel: get_energy (evaluate_ood.py:152-159) and
    semantic_inference_with_void (maskformer_model.py:388-392)."""
    B, Q, K, D, h, w = 2, 100, 19, 256, 9, 20
    g = torch.Generator().manual_seed(77)
    E = torch.randn(B, Q, D, generator=g) / math.sqrt(D) * 2.0
    Fm = torch.randn(B, h, w, D, generator=g)
    bias = torch.randn(B, Q, generator=g) * 0.5 - 0.5
    logits = torch.randn(B, Q, K + 1, generator=g)
    H, W = 4 * h - 1, 4 * w
    (e_hi, e_lo), Ex = planes(E.view(B * Q, D), dev)
    (f_hi, f_lo), Fx = planes(Fm.view(-1, D), dev)
    e_pl = (e_hi.view(B, Q, D), e_lo.view(B, Q, D))
    f_pl = (f_hi.view(B, h, w, D), f_lo.view(B, h, w, D))
    masks = (torch.einsum("bqc,bhwc->bqhw", Ex.view(B, Q, D).double(), Fx.view(B, h, w, D).double())
             + bias.double()[:, :, None, None]).float()
    up = F.interpolate(masks, size=(4 * h, 4 * w), mode="bilinear", align_corners=False)[:, :, :H, :W]
    sem = torch.einsum("bqc,bqhw->bchw", logits.softmax(-1)[..., :-1], up.sigmoid())
    sem_void = torch.einsum("bqc,bqhw->bchw", logits.softmax(-1), up.sigmoid())
    en, s1 = ops.einsum_score_fused(e_pl, f_pl, logits.to(dev), (H, W), bias=bias.to(dev), want_sem_seg=True, score_func="energy")
    assert (s1.cpu() - sem).abs().max() < 5e-5
    assert (en.cpu() - (-torch.logsumexp(sem, dim=1))).abs().max() < 5e-5
    rv, s2 = ops.einsum_score_fused(e_pl, f_pl, logits.to(dev), (H, W), bias=bias.to(dev), want_sem_seg=True, include_void=True)
    assert s2.shape == (B, K + 1, H, W)
    assert (s2.cpu() - sem_void).abs().max() < 5e-5
    assert (rv.cpu() - (-sem_void.tanh().sum(1))).abs().max() < 5e-5
    ev = ops.einsum_score_fused(e_pl, f_pl, logits.to(dev), (H, W), bias=bias.to(dev), score_func="pebal", include_void=True)
    assert (ev.cpu() - (-torch.logsumexp(sem_void, dim=1))).abs().max() < 5e-5


def test_einsum_score_fused_limits(dev):
    z = lambda *s: torch.zeros(*s, dtype=torch.bfloat16, device=dev)  # noqa: E731
    with pytest.raises(ops.RbaError):      # Q > 104
        ops.einsum_score_fused((z(1, 105, 64), z(1, 105, 64)), (z(1, 4, 4, 64), z(1, 4, 4, 64)), torch.zeros(1, 105, 4, device=dev), (16, 16))
    with pytest.raises(ops.RbaError):      # D not a multiple of 64
        ops.einsum_score_fused((z(1, 10, 32), z(1, 10, 32)), (z(1, 4, 4, 32), z(1, 4, 4, 32)), torch.zeros(1, 10, 4, device=dev), (16, 16))
    with pytest.raises(ops.RbaError):      # output larger than 4x
        ops.einsum_score_fused((z(1, 10, 64), z(1, 10, 64)), (z(1, 4, 4, 64), z(1, 4, 4, 64)), torch.zeros(1, 10, 4, device=dev), (17, 16))
    e = ops.einsum_score_fused((z(0, 10, 64), z(0, 10, 64)), (z(0, 4, 4, 64), z(0, 4, 4, 64)), torch.zeros(0, 10, 4, device=dev), (16, 16))
    assert e.shape == (0, 16, 16)


# ------------------------------------------------------------------------------------------------
# tcgen05 bf16x3 backend: same contract as the FFMA kernel; tolerance = bf16x3 truncation (~1e-5 relative / product)
# ------------------------------------------------------------------------------------------------
TC_TOL = 2e-4


@pytest.mark.parametrize("M,N,K", [(300, 20, 256), (129, 96, 128), (1000, 384, 64), (5, 1, 256), (257, 130, 2304),
                                   (128, 128, 32), (700, 512, 512), (300, 512, 1024), (129, 256, 2048),   # last two: 256-wide N tiles
                                   (1024, 256, 512), (2049, 768, 576), (40000, 1536, 512)])   # CTA pairs (M >= 512, N % 256 == 0, K >= 512), > 1 tile per pair
@pytest.mark.parametrize("act", [ops.RBA_ACT_NONE, ops.RBA_ACT_GELU])
def test_gemm_tc(dev, M, N, K, act):
    g = torch.Generator().manual_seed(M + N + K + 1)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    bias = torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g)
    ap, av = planes(a, dev)
    wp, wv = planes(w, dev)
    z = (av.double() @ wv.double().t() + bias.double()).float()
    z = {0: z, 2: F.gelu(z)}[act] + res
    c, cp = ops.gemm(ap, wp, bias=bias.to(dev), act=act, residual=res.to(dev), out_planes=True, backend=ops.RBA_GEMM_TC)
    torch.cuda.synchronize()
    assert (c.cpu() - z).abs().max() < TC_TOL * max(1.0, float(z.abs().max()))
    assert (unplanes(cp) - z).abs().max() < 2 * TC_TOL * max(1.0, float(z.abs().max()))
    # and against the exact-arithmetic kernel on the same planes
    c2 = ops.gemm(ap, wp, bias=bias.to(dev), act=act, residual=res.to(dev), backend=ops.RBA_GEMM_FFMA)
    assert (c - c2).abs().max() < TC_TOL * max(1.0, float(z.abs().max()))


def test_gemm_tc_batched_and_swin(dev):
    Bn, Q, HW, K = 3, 100, 300, 256
    g = torch.Generator().manual_seed(21)
    a, w, b = torch.randn(Bn, Q, K, generator=g), torch.randn(Bn, HW, K, generator=g) / 16, torch.randn(Bn, Q, generator=g)
    ap, av = planes(a, dev)
    wp, wv = planes(w, dev)
    ref = (torch.einsum("bqk,bpk->bqp", av.double(), wv.double()) + b.double()[:, :, None]).float()
    c = ops.gemm(ap, wp, bias=b.to(dev), bias_per_row=True, backend=ops.RBA_GEMM_TC)
    assert (c.cpu() - ref).abs().max() < TC_TOL * max(1.0, float(ref.abs().max()))
    # swin scatter epilogue
    B, H, W, C, ws, shift = 2, 30, 41, 64, 12, 6
    nWh, nWw = -(-H // ws), -(-W // ws)
    rows = B * nWh * nWw * ws * ws
    a = torch.randn(rows, C, generator=g)
    w = torch.randn(C, C, generator=g) / 8
    bias = torch.randn(C, generator=g)
    shortcut = torch.randn(B * H * W, C, generator=g)
    ap, av = planes(a, dev)
    wp, wv = planes(w, dev)
    y = (av.double() @ wv.double().t() + bias.double()).float()
    y = torch.roll(O.window_reverse(y.view(-1, ws, ws, C), ws, nWh * ws, nWw * ws), shifts=(shift, shift), dims=(1, 2))
    ref = shortcut + y[:, :H, :W, :].reshape(B * H * W, C)
    c = ops.gemm(ap, wp, bias=bias.to(dev), residual=shortcut.to(dev), swin=(B, H, W, ws, shift), backend=ops.RBA_GEMM_TC)
    assert (c.cpu() - ref).abs().max() < TC_TOL * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("H,W,Cout", [(16, 32, 128), (9, 13, 128), (17, 20, 256)])    # Cout 256: the 256-wide N tile
def test_conv3x3_tc(dev, H, W, Cout):
    B, Cin = 2, 64
    g = torch.Generator().manual_seed(22)
    x = torch.randn(B, H, W, Cin, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin)
    xp, xv = planes(x, dev)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    wp, wkv = planes(wk, dev)
    wv = wkv.view(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
    ref = F.conv2d(xv.permute(0, 3, 1, 2).double(), wv.double(), padding=1).permute(0, 2, 3, 1).float()
    y = ops.conv3x3(xp, wp, backend=ops.RBA_GEMM_TC)
    assert (y.cpu() - ref).abs().max() < TC_TOL * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("kernel", ["tcgen05_tiled", "tcgen05", "mma_sync"])
@pytest.mark.parametrize("H,W,shift,B,heads", [(24, 24, 0, 2, 2), (30, 41, 6, 2, 2), (12, 12, 6, 2, 2), (12, 12, 0, 1, 1),
                                               (96, 180, 6, 3, 4), (40, 40, 6, 1, 16)])
def test_window_attention_tensor_core(dev, H, W, shift, B, heads, kernel):
    """swin.py:145-168 with q/k/v as split planes, both contractions on tensor cores (bf16x3): the tcgen05 + TMA kernel the
    engine runs (wattn_tc.cu: more items than SMs in the big case, so the persistent pipeline wraps its TMEM / stage rings)
    and the mma.sync kernel kept as its cross-check."""
    ws = 12
    C = heads * 32
    nWh, nWw = -(-H // ws), -(-W // ws)
    nW = nWh * nWw
    g = torch.Generator().manual_seed(18)
    qkv = torch.randn(B * nW, ws * ws, 3 * C, generator=g) * 1.5
    table = torch.randn((2 * ws - 1) ** 2, heads, generator=g)
    qp, qv = planes(qkv.view(-1, 3 * C), dev)
    qkv = qv.view(B * nW, ws * ws, 3 * C)
    q, k, v = qkv.reshape(B * nW, ws * ws, 3, heads, 32).permute(2, 0, 3, 1, 4)
    attn = (q * 32 ** -0.5) @ k.transpose(-2, -1)
    idx = O.relative_position_index(ws)
    attn = attn + table[idx.view(-1)].view(ws * ws, ws * ws, -1).permute(2, 0, 1).unsqueeze(0)
    if shift:
        mask = O.shift_attn_mask(H, W, ws, shift)
        attn = (attn.view(B, nW, heads, ws * ws, ws * ws) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, ws * ws, ws * ws)
    ref = (attn.softmax(-1) @ v).transpose(1, 2).reshape(B * nW * ws * ws, C)
    if kernel == "tcgen05_tiled":      # q|k|v in the tiled layout the engine's QKV GEMM writes: one bulk copy per operand tile
        out = ops.window_attn_tc(ops.qkv_to_tiles(qp, heads), table.to(dev), B, H, W, C, heads, ws, shift, tiled=True)
    else:
        fn = ops.window_attn_tc if kernel == "tcgen05" else ops.window_attn_planes
        out = fn(qp, table.to(dev), B, H, W, C, heads, ws, shift)
    err = (unplanes(out) - ref).abs()
    assert err.max() < 3e-4, (float(err.max()), int(err.argmax()) // C, int(err.argmax()) % C)


# ---------------------------------------------------------------------------------------------------------------------
# ResNet backbone pieces (detectron2 build_resnet_backbone; resnet.cu)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32])
def test_resnet_stem_conv_and_maxpool(dev, dtype):
    """normalise + zero-pad + 7x7/2 conv + folded BN + ReLU, then the 3x3/2 max-pool, against torch."""
    import ctypes
    from rba_b200 import _lib
    g = torch.Generator().manual_seed(31)
    B, H, W, Hp, Wp = 2, 70, 100, 96, 128
    img = torch.randint(0, 256, (B, 3, H, W), generator=g).to(dtype)
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.1
    bias = torch.randn(64, generator=g) * 0.1
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    xn = (img.float() - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    xn = F.pad(xn, (0, Wp - W, 0, Hp - H))
    ref = F.relu(F.conv2d(xn, w, bias, stride=2, padding=3))
    ref_pool = F.max_pool2d(ref, 3, 2, 1)
    out = torch.empty(B, Hp // 2, Wp // 2, 64, device=dev)
    L = _lib.lib()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    imd, wd, bd = img.to(dev).contiguous(), w.reshape(64, -1).to(dev).contiguous(), bias.to(dev)
    _lib.check(L.rba_k_stem_conv(p(imd), 0 if dtype == torch.uint8 else 1, B, H, W, Hp, Wp, (ctypes.c_float * 3)(*mean),
                                 (ctypes.c_float * 3)(*std), p(wd), p(bd), p(out), st))
    assert (out.cpu().permute(0, 3, 1, 2) - ref).abs().max() < 2e-4
    y = torch.empty(B, Hp // 4, Wp // 4, 64, device=dev)
    hi = torch.empty(B, Hp // 4, Wp // 4, 64, dtype=torch.bfloat16, device=dev)
    lo = torch.empty_like(hi)
    _lib.check(L.rba_k_maxpool3x3s2(p(out), B, Hp // 2, Wp // 2, 64, p(y), p(hi), p(lo), st))
    assert (y.cpu().permute(0, 3, 1, 2) - ref_pool).abs().max() < 2e-4
    assert (hi.float() + lo.float() - y).abs().max() < 1e-4


@pytest.mark.parametrize("stride,relu,with_bias", [(1, 1, True), (2, 1, True), (2, 0, False), (1, 1, False)])
def test_resnet_bias_act_sub(dev, stride, relu, with_bias):
    import ctypes
    from rba_b200 import _lib
    g = torch.Generator().manual_seed(32)
    B, H, W, C = 2, 12, 20, 64
    x = torch.randn(B, H, W, C, generator=g)
    bias = torch.randn(C, generator=g) if with_bias else None
    ref = x[:, ::stride, ::stride] + (bias if with_bias else 0.0)
    if relu:
        ref = ref.relu()
    xd = x.to(dev)
    y = torch.empty(B, H // stride, W // stride, C, device=dev)
    hi = torch.empty(y.shape, dtype=torch.bfloat16, device=dev)
    lo = torch.empty_like(hi)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
    _lib.check(_lib.lib().rba_k_bias_act_sub(p(xd), p(bias.to(dev) if with_bias else None), B, H, W, C, stride, relu, p(y), p(hi),
                                             p(lo), st))
    assert torch.equal(y.cpu(), ref)
    assert (hi.float() + lo.float() - y).abs().max() < 1e-4
    if stride == 1:                                              # in place
        _lib.check(_lib.lib().rba_k_bias_act_sub(p(xd), p(bias.to(dev) if with_bias else None), B, H, W, C, 1, relu, p(xd), None,
                                                 None, st))
        assert torch.equal(xd.cpu(), ref)


def test_gemm_tc_writes_tiled_qkv_planes(dev):
    """rba_gemm_args.qkv_tile_heads: the QKV GEMM's planes in the (window, part, head) tiled layout equal a permutation of its
    plain row-major planes (bitwise)."""
    g = torch.Generator().manual_seed(41)
    heads, nwin = 4, 5
    C, M = 32 * heads, 144 * nwin
    a = torch.randn(M, C, generator=g)
    w = torch.randn(3 * C, C, generator=g) / math.sqrt(C)
    bias = torch.randn(3 * C, generator=g)
    ap, _ = planes(a, dev)
    wp, _ = planes(w, dev)
    plain = ops.gemm(ap, wp, bias=bias.to(dev), out_planes=True, out_f32=False, backend=ops.RBA_GEMM_TC)
    tiled = ops.gemm(ap, wp, bias=bias.to(dev), out_planes=True, out_f32=False, backend=ops.RBA_GEMM_TC, qkv_tile_heads=heads)
    want = ops.qkv_to_tiles(plain, heads)
    assert torch.equal(tiled[0].view(-1), want[0].view(-1)) and torch.equal(tiled[1].view(-1), want[1].view(-1))
    with pytest.raises(ops.RbaError):
        ops.gemm(ap, wp, out_planes=True, out_f32=False, backend=ops.RBA_GEMM_FFMA, qkv_tile_heads=heads)

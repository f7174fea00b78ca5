"""LIVE parity on the GPU box against the UNMODIFIED reference running on the same B200 (baseline/_ref, installed by
tools/make_baseline_ref.py and shipped with the snapshot; nothing here reads /root/reference).

  * the reference's own MultiScaleDeformableAttention CUDA extension (kernels untouched, rebuilt for sm_100a) against
    rba_msda_forward through the reference's own Python wrapper `MSDeformAttnFunction`
    (ops/functions/ms_deform_attn_func.py:32-49) bound to this repo's shim module;
  * the reference model's forward on cuda (TF32 off, so it is the fp32 statement) at the metric's shape, 1 x 1024 x 2048
    Swin-B 1dl, against the engine: every output within 1e-3 when the cross-attention reads the reference's own boolean
    decisions, and the number of near-threshold decisions that differ when free-running.
Skipped when baseline/_ref is absent."""
import glob
import os
import sys

import pytest
import torch

import rba_b200
from conftest import ROOT, TOL
from rba_b200 import ops, weights

pytestmark = pytest.mark.gpu

BREF = os.path.join(ROOT, "baseline", "_ref")
HAVE_REF = os.path.isdir(os.path.join(BREF, "mask2former"))
HAVE_EXT = bool(glob.glob(os.path.join(BREF, "MultiScaleDeformableAttention*.so")))


def _ref_loader(native):
    os.environ["RBA_REFERENCE_ROOT"] = BREF
    import ref_loader
    ref_loader.REF_ROOT = BREF
    if native and "mask2former.maskformer_model" not in sys.modules and HAVE_EXT:
        ref_loader.use_native_msda()
    return ref_loader


def _load_ext():
    import importlib.util
    so = glob.glob(os.path.join(BREF, "MultiScaleDeformableAttention*.so"))[0]
    spec = importlib.util.spec_from_file_location("MultiScaleDeformableAttention", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not HAVE_EXT, reason="baseline/_ref has no reference MSDA extension")
@pytest.mark.parametrize("levels", [1, 3])
def test_msda_forward_matches_reference_cuda_extension(dev, levels):
    """Same six arguments, same result as the reference FFI (ops/src/vision.cpp:18-21) at the model's shapes:
    1 level (32x64, Swin-B 1dl at 1024x2048) and 3 levels (res5, res4, res3)."""
    ext = _load_ext()
    shapes = [(32, 64)] if levels == 1 else [(32, 64), (64, 128), (128, 256)]
    B, M, D, P = 2, 8, 32, 4
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    g = torch.Generator(device=dev).manual_seed(5)
    value = torch.randn(B, S, M, D, device=dev, generator=g)
    loc = torch.rand(B, S, M, L, P, 2, device=dev, generator=g) * 1.2 - 0.1          # some samples fall outside
    aw = torch.softmax(torch.randn(B, S, M, L * P, device=dev, generator=g), -1).view(B, S, M, L, P)
    ss = torch.as_tensor(shapes, dtype=torch.long, device=dev)
    lsi = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
    ref = ext.ms_deform_attn_forward(value, ss, lsi, loc, aw, 128)
    got = ops.ms_deform_attn_forward(value, ss, lsi, loc, aw, 128)
    assert got.shape == ref.shape
    assert (got - ref).abs().max() < 2e-5
    # the reference's own autograd wrapper bound to this repo's module (the drop-in of INTEGRATION.md §1)
    from rba_b200.compat import MultiScaleDeformableAttention as shim
    out = shim.ms_deform_attn_forward(value, ss, lsi, loc, aw, 128)
    assert (out - ref).abs().max() < 2e-5


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref absent (run tools/make_baseline_ref.py in the build container)")
def test_live_reference_forward_parity_1024x2048(dev):
    rl = _ref_loader(native=True)
    mc = rba_b200.config.swin_b_1dl()
    sd = weights.init_state_dict(mc, seed=13, perturb=0.02)
    cfg = rl.load_cfg("swin_b_1dl")
    model = rl.build_reference_model(cfg, seed=0)
    rl.load_state_dict_into(model, sd)
    model.to(dev).eval()
    am_list = rl.spy_attention_decisions(model)
    caps = {}
    model.sem_seg_head.register_forward_hook(lambda m, i, o: caps.__setitem__("head", o))
    g = torch.Generator().manual_seed(41)
    img = torch.randint(0, 256, (3, 1024, 2048), dtype=torch.uint8, generator=g).to(dev)
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            sem_ref = model([{"image": img}])[0]["sem_seg"]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf
    rba_ref = -sem_ref.tanh().sum(0)
    ref_dec = am_list[0].clone()
    ref_dec[ref_dec.all(-1)] = False                                 # mask2former_transformer_decoder.py:433
    ref_dec = ref_dec.to(torch.uint8).contiguous()
    logits_ref, masks_ref = caps["head"]["pred_logits"], caps["head"]["pred_masks"]
    del model
    torch.cuda.empty_cache()

    e = rba_b200.Engine(mc, dev.index).load_state_dict(sd)
    e.set_gemm_backend("tc")
    dump = torch.zeros_like(ref_dec)
    e.debug_attn_mask(0, dump=dump)
    free = e.forward(img[None].contiguous(), rba=True, logits=True, masks=True)
    torch.cuda.synchronize()
    flips = int((dump != ref_dec).sum())
    free_err = {"pred_logits": float((free["pred_logits"] - logits_ref).abs().max()),
                "pred_masks": float((free["pred_masks"] - masks_ref).abs().max()),
                "rba": float((free["rba"][0] - rba_ref).abs().max())}
    print("free-running vs reference-on-GPU:", free_err, "decisions differing:", flips, "of", ref_dec.numel())
    assert flips <= ref_dec.numel() // 1000                           # only near-threshold decisions may differ
    e.debug_attn_mask(0, force=ref_dec)
    out = e.forward(img[None].contiguous(), rba=True, sem_seg=True, logits=True, masks=True)
    torch.cuda.synchronize()
    err = {"pred_logits": float((out["pred_logits"] - logits_ref).abs().max()),
           "pred_masks": float((out["pred_masks"] - masks_ref).abs().max()),
           "sem_seg": float((out["sem_seg"][0] - sem_ref).abs().max()),
           "rba": float((out["rba"][0] - rba_ref).abs().max())}
    print("on the reference's decisions:", err)
    import json
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/live_reference_parity.json", "w") as f:
        json.dump({"free": free_err, "flips": flips, "decisions": ref_dec.numel(), "forced": err}, f, indent=1)
    for k, v in err.items():
        assert v < TOL, (k, v)
    if flips == 0:
        for k, v in free_err.items():
            assert v < TOL, ("free-running", k, v)

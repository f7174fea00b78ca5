"""GPU parity of the whole hot path against (a) the committed golden outputs of the unmodified reference and
(b) the oracle's stage-boundary tensors, through the reference-facing interface (rba_b200.MaskFormer) and the
engine C ABI.  Bar: 1e-3 max-abs (north star), on pred_logits, pred_masks, pre-tanh sem_seg and rba."""
import pytest
import torch

import rba_oracle as O
from conftest import TOL, load_golden
from golden_cases import CASES, case_images, case_model_config, state_checksum
import rba_b200
from rba_b200 import weights

pytestmark = pytest.mark.gpu


def _engine(mc, sd, dev, taps=False):
    e = rba_b200.Engine(mc, dev.index).load_state_dict(sd)
    if taps:
        e.set_option("taps", 1)
    return e


@pytest.mark.parametrize("name", list(CASES))
def test_forward_matches_reference_golden(dev, name):
    fix = load_golden(f"model_{name}.pt")
    case = fix["case"]
    mc = case_model_config(case)
    sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
    assert abs(state_checksum(sd) - fix["state_checksum"]) <= 1e-6 * fix["state_checksum"]
    imgs = torch.stack(case_images(case)).to(dev)
    e = _engine(mc, sd, dev)
    out = e.forward(imgs, rba=True, sem_seg=True, logits=True, masks=True)
    torch.cuda.synchronize()
    errs = {
        "pred_logits": (out["pred_logits"].cpu() - fix["pred_logits"]).abs().max().item(),
        "pred_masks": (out["pred_masks"].cpu() - fix["pred_masks"]).abs().max().item(),
        "sem_seg": (out["sem_seg"].cpu()[:, :, ::4, ::4] - fix["sem_seg_s4"]).abs().max().item(),
        "rba": (out["rba"].cpu() - fix["rba"]).abs().max().item(),
    }
    print(name, errs)
    for k, v in errs.items():
        assert v < TOL, (k, v)


def test_stage_taps_match_oracle(dev):
    """Every stage boundary: res2..res5, encoder output, FPN levels, decoder state."""
    case = CASES["tiny_1dl"]
    mc = case_model_config(case)
    sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
    imgs = case_images(case)
    ref = O.forward(sd, mc, imgs, want_taps=True)
    e = _engine(mc, sd, dev, taps=True)
    out = e.forward(torch.stack(imgs).to(dev), rba=True, logits=True, masks=True)
    B = len(imgs)
    t = ref["taps"]
    for i, k in enumerate(["res2", "res3", "res4", "res5"]):
        r = t[k].permute(0, 2, 3, 1).reshape(B, -1, t[k].shape[1])
        g = e.tap(k).cpu().view_as(r)
        assert (g - r).abs().max() < 2e-4, k
    enc = t[f"enc_l{mc.enc_layers - 1}"]
    assert (e.tap("enc_out").cpu().view_as(enc) - enc).abs().max() < 5e-4
    for k in ["res4", "res3", "res2"]:
        r = t[f"fpn_{k}"].permute(0, 2, 3, 1)
        assert (e.tap(f"fpn_{k}").cpu().view_as(r) - r).abs().max() < 5e-4, k
    dec = t[f"dec{mc.dec_layers - 1}_out"].transpose(0, 1)
    assert (e.tap("dec_out").cpu().view_as(dec) - dec).abs().max() < TOL
    assert (out["pred_masks"].cpu() - ref["pred_masks"]).abs().max() < TOL


def test_maskformer_module_interface(dev):
    """model([{'image': ...}]) -> [{'sem_seg': (K,H,W)}] + get_RbA arithmetic (evaluate_ood.py:108-150)."""
    case = CASES["tiny_1dl"]
    mc = case_model_config(case)
    sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
    fix = load_golden("model_tiny_1dl.pt")
    model = rba_b200.MaskFormer(mc)
    missing = model.load_state_dict(sd)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert list(model.state_dict().keys()) == list(sd.keys())
    model.to(dev)
    model.eval()
    x = case_images(case)[0]
    with torch.no_grad():
        out = model([{"image": x.to(dev)}])
    logits = out[0]["sem_seg"]
    assert logits.shape == (mc.num_classes, x.shape[1], x.shape[2]) and logits.device.type == "cuda"
    rba = -logits.tanh().sum(dim=0)                                   # evaluate_ood.py:148-150, unchanged
    assert (rba.cpu() - fix["rba"][0]).abs().max() < TOL
    fused = model.rba([{"image": x.to(dev)}])[0]
    assert (fused - rba).abs().max() < 1e-5
    with pytest.raises(rba_b200.RbaError):
        model([{"image": x.to(dev)}], return_ood_pred=True)
    with pytest.raises(rba_b200.RbaError):
        model.train()


def test_batch_invariance_and_graph_replay(dev):
    """Images are independent units: a batch equals per-image runs bitwise; CUDA-graph replay equals eager."""
    case = CASES["tiny_3lvl"]
    mc = case_model_config(case)
    sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
    imgs = torch.stack(case_images(case)).to(dev)
    e = _engine(mc, sd, dev)
    both = e.forward(imgs, rba=True)["rba"].clone()
    one = [e.forward(imgs[i:i + 1].contiguous(), rba=True)["rba"].clone() for i in range(imgs.shape[0])]
    assert torch.equal(both, torch.cat(one))
    static_in, static_out, g = e.graphed(imgs, rba=True)
    static_in.copy_(imgs)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out["rba"], both)
    static_in.copy_(imgs.flip(0))
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out["rba"], both.flip(0))


def test_errors_are_loud(dev):
    mc = rba_b200.config.tiny_test()
    sd = weights.init_state_dict(mc, seed=0)
    e = rba_b200.Engine(mc, dev.index)
    with pytest.raises(rba_b200.RbaError):
        e.forward(torch.zeros(1, 3, 64, 64, dtype=torch.uint8, device=dev))      # not loaded
    bad = dict(sd)
    bad.pop("backbone.norm2.weight")
    with pytest.raises(rba_b200.RbaError, match="backbone.norm2.weight"):
        rba_b200.Engine(mc, dev.index).load_state_dict(bad)
    e.load_state_dict(sd)
    with pytest.raises(rba_b200.RbaError):
        e.forward(torch.zeros(1, 3, 64, 64, dtype=torch.float16, device=dev))
    with pytest.raises(rba_b200.RbaError):
        e.forward(torch.zeros(1, 3, 64, 64, dtype=torch.uint8))                   # CPU tensor


@pytest.mark.parametrize("name", list(CASES))
def test_forward_tc_backend_matches_reference_golden(dev, name):
    """Same bar with every GEMM / 3x3 conv on tcgen05 tensor cores (bf16x3 split precision)."""
    fix = load_golden(f"model_{name}.pt")
    case = fix["case"]
    mc = case_model_config(case)
    sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
    imgs = torch.stack(case_images(case)).to(dev)
    e = _engine(mc, sd, dev)
    e.set_gemm_backend("tc")
    out = e.forward(imgs, rba=True, sem_seg=True, logits=True, masks=True)
    torch.cuda.synchronize()
    errs = {
        "pred_logits": (out["pred_logits"].cpu() - fix["pred_logits"]).abs().max().item(),
        "pred_masks": (out["pred_masks"].cpu() - fix["pred_masks"]).abs().max().item(),
        "sem_seg": (out["sem_seg"].cpu()[:, :, ::4, ::4] - fix["sem_seg_s4"]).abs().max().item(),
        "rba": (out["rba"].cpu() - fix["rba"]).abs().max().item(),
    }
    print(name, "tc", errs)
    for k, v in errs.items():
        assert v < TOL, (k, v)


@pytest.mark.parametrize("name", list(CASES))
def test_forward_fused_score_matches_reference_golden(dev, name):
    """The product path of model.rba() / model(): last mask einsum fused into the score kernel (pred_masks never
    materialised), against the same golden outputs of the unmodified reference; and against the two-kernel path."""
    fix = load_golden(f"model_{name}.pt")
    case = fix["case"]
    mc = case_model_config(case)
    sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
    imgs = torch.stack(case_images(case)).to(dev)
    e = _engine(mc, sd, dev)
    n0 = rba_b200.launch_count()
    out = e.forward(imgs, rba=True, sem_seg=True, logits=True)          # no pred_masks -> fused kernel
    n_fused = rba_b200.launch_count() - n0
    torch.cuda.synchronize()
    errs = {
        "pred_logits": (out["pred_logits"].cpu() - fix["pred_logits"]).abs().max().item(),
        "sem_seg": (out["sem_seg"].cpu()[:, :, ::4, ::4] - fix["sem_seg_s4"]).abs().max().item(),
        "rba": (out["rba"].cpu() - fix["rba"]).abs().max().item(),
    }
    print(name, "fused", errs)
    for k, v in errs.items():
        assert v < TOL, (k, v)
    e.set_option("fused_score", 0)
    n0 = rba_b200.launch_count()
    two = e.forward(imgs, rba=True, sem_seg=True)
    n_two = rba_b200.launch_count() - n0
    assert n_fused == n_two - 1, (n_fused, n_two)                        # one GEMM launch less
    assert (two["rba"] - out["rba"]).abs().max() < 1e-4
    assert (two["sem_seg"] - out["sem_seg"]).abs().max() < 1e-4


def test_compat_plugin_flow_on_gpu(dev, tmp_path):
    """evaluate_ood.get_model / get_RbA flow (evaluate_ood.py:108-150) through the drop-in plumbing, on the GPU box
    (no reference checkout here): get_cfg -> merge_from_file -> build_model -> DetectionCheckpointer -> model(...)."""
    import yaml
    from rba_b200 import compat
    compat.plug_in()
    from detectron2.checkpoint import DetectionCheckpointer
    from detectron2.config import get_cfg
    from detectron2.modeling import build_model
    case = CASES["tiny_1dl"]
    mc = case_model_config(case)
    sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
    fix = load_golden("model_tiny_1dl.pt")
    y = {"MODEL": {
        "META_ARCHITECTURE": "MaskFormer", "DEVICE": "cuda", "WEIGHTS": "",
        "PIXEL_MEAN": list(mc.pixel_mean), "PIXEL_STD": list(mc.pixel_std),
        "BACKBONE": {"NAME": "D2SwinTransformer"},
        "SWIN": {"EMBED_DIM": mc.embed_dim, "DEPTHS": list(mc.depths), "NUM_HEADS": list(mc.num_heads), "WINDOW_SIZE": 12,
                 "MLP_RATIO": 4.0, "PATCH_SIZE": 4, "APE": False, "QKV_BIAS": True, "PATCH_NORM": True},
        "SEM_SEG_HEAD": {"NAME": "MaskFormerHead", "PIXEL_DECODER_NAME": "MSDeformAttnPixelDecoder", "NORM": "GN",
                         "CONVS_DIM": 256, "MASK_DIM": 256, "NUM_CLASSES": mc.num_classes,
                         "IN_FEATURES": ["res2", "res3", "res4", "res5"],
                         "DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES": list(mc.transformer_in_features),
                         "COMMON_STRIDE": 4, "TRANSFORMER_ENC_LAYERS": mc.enc_layers},
        "MASK_FORMER": {"TRANSFORMER_DECODER_NAME": "MultiScaleMaskedTransformerDecoder", "PRE_NORM": False, "NHEADS": 8,
                        "HIDDEN_DIM": 256, "DIM_FEEDFORWARD": 2048, "DEC_LAYERS": mc.dec_layers + 1,
                        "NUM_OBJECT_QUERIES": mc.num_queries, "SIZE_DIVISIBILITY": 32}}}
    cfg_path, ckpt_path = str(tmp_path / "config.yaml"), str(tmp_path / "model_final.pth")
    with open(cfg_path, "w") as f:
        yaml.safe_dump(y, f)
    torch.save({"model": sd}, ckpt_path)
    cfg = get_cfg()
    cfg.merge_from_file(cfg_path)
    cfg.merge_from_list(["OUTPUT_DIR", str(tmp_path / "out")])
    cfg.freeze()
    model = build_model(cfg)
    assert isinstance(model, rba_b200.MaskFormer)
    DetectionCheckpointer(model, save_dir=cfg.OUTPUT_DIR).resume_or_load(ckpt_path, resume=False)
    model.to(dev)
    model.eval()
    x = case_images(case)[0]
    with torch.no_grad():
        out = model([{"image": x.to(dev)}])
    rba = -out[0]["sem_seg"].tanh().sum(dim=0)
    assert (rba.cpu() - fix["rba"][0]).abs().max() < TOL


def test_model_energy_score_and_include_void(dev):
    """model.score(..., "pebal") == get_energy on the model's own sem_seg (evaluate_ood.py:152-159);
    model(..., include_void=True) returns K+1 planes whose first K equal the default call (maskformer_model.py:381-392)."""
    case = CASES["tiny_1dl"]
    mc = case_model_config(case)
    sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
    model = rba_b200.MaskFormer(mc)
    model.load_state_dict(sd)
    model.to(dev).eval()
    x = case_images(case)[0].to(dev)
    sem = model([{"image": x}])[0]["sem_seg"]
    energy = model.score([{"image": x}], "pebal")[0]
    assert (energy - (-torch.logsumexp(sem, dim=0))).abs().max() < 1e-5
    semv = model([{"image": x}], include_void=True)[0]["sem_seg"]
    assert semv.shape[0] == mc.num_classes + 1
    assert (semv[:-1] - sem).abs().max() < 1e-6
    ref = O.forward(sd, mc, [x.cpu()])
    cls = ref["pred_logits"][0].softmax(-1)
    up = torch.nn.functional.interpolate(ref["pred_masks"], scale_factor=4, mode="bilinear", align_corners=False)[0]
    void = torch.einsum("q,qhw->hw", cls[:, -1], up.sigmoid())[: x.shape[1], : x.shape[2]]
    assert (semv[-1].cpu() - void).abs().max() < TOL
    again = model.rba([{"image": x}])[0]                       # options reset to the default score
    assert (again - (-sem.tanh().sum(0))).abs().max() < 1e-5


def test_score_stream_overlapped_pipeline(dev):
    """rba_b200.ScoreStream (pinned H2D / forward / D2H on three streams): results, in order and bitwise, equal the
    synchronous forward of the same batches; misuse fails loudly."""
    case = CASES["tiny_1dl"]
    mc = case_model_config(case)
    sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
    e = _engine(mc, sd, dev)
    g = torch.Generator().manual_seed(11)
    B, H, W = 2, 64, 96
    batches = [torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=g).pin_memory() for _ in range(5)]
    want = [e.forward(b.to(dev), rba=True)["rba"].cpu() for b in batches]
    for use_graph in (True, False):
        stream = rba_b200.ScoreStream(e, B, H, W, use_graph=use_graph)
        got = [r.clone() for r in stream.run(batches)]
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert torch.equal(a, b)
        assert stream.h2d_bytes_per_step == B * 3 * H * W and stream.d2h_bytes_per_step == B * H * W * 4
    with pytest.raises(rba_b200.RbaError):
        stream.submit(torch.zeros((B, 3, H, W), dtype=torch.uint8))          # not pinned
    with pytest.raises(rba_b200.RbaError):
        stream.submit(torch.zeros((B, 3, H + 32, W), dtype=torch.uint8).pin_memory())
    with pytest.raises(rba_b200.RbaError):
        stream.collect()


def test_densehybrid_head_matches_reference_golden(dev):
    """DenseHybrid: `model(x, return_ood_pred=True)` (maskformer_model.py:303-305,350-351) and the fused
    `--score_func dense_hybrid` score (evaluate_ood.py:161-173) against the unmodified reference with
    MODEL.MASK_FORMER.DENSE_HYBRID_LOSS: True (tests/golden/model_tiny_ood.pt)."""
    fix = load_golden("model_tiny_ood.pt")
    case = fix["case"]
    mc = case_model_config(case)
    assert mc.ood_prediction
    sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
    assert abs(state_checksum(sd) - fix["state_checksum"]) <= 1e-6 * fix["state_checksum"]
    imgs = case_images(case)
    model = rba_b200.MaskFormer(mc)
    model.load_state_dict(sd)
    model.to(dev).eval()
    res, ood_pred = model([{"image": im.to(dev)} for im in imgs], return_ood_pred=True)
    assert tuple(ood_pred.shape) == tuple(fix["ood_pred"].shape)
    assert (ood_pred.cpu() - fix["ood_pred"]).abs().max() < TOL
    sem = torch.stack([r["sem_seg"] for r in res])
    # the reference caller's own arithmetic on our outputs (evaluate_ood.py:165-172)
    p2 = torch.softmax(ood_pred, dim=1)[:, 1]
    dh_ref_style = -torch.logsumexp(sem, dim=1) + (p2 + 1e-9).log()
    assert (dh_ref_style.cpu() - fix["densehybrid"]).abs().max() < TOL
    # fused: energy from the einsum+score kernel, head added in place
    dh = model.score([{"image": im.to(dev)} for im in imgs], "dense_hybrid")
    assert (dh.cpu() - fix["densehybrid"]).abs().max() < TOL
    en = model.score([{"image": im.to(dev)} for im in imgs], "energy")
    assert (en.cpu() - fix["energy"]).abs().max() < TOL
    # a model without the head refuses loudly
    plain = rba_b200.MaskFormer(case_model_config(CASES["tiny_1dl"]))
    plain.to(dev)
    with pytest.raises(rba_b200.RbaError):
        plain([{"image": imgs[0].to(dev)}], return_ood_pred=True)
    with pytest.raises(rba_b200.RbaError):
        plain.score([{"image": imgs[0].to(dev)}], "dense_hybrid")


def test_full_size_properties_swin_b_1024x2048(dev):
    """BASELINE.json's full size (Swin-B 1dl, 1024 x 2048), where the CPU oracle would take minutes per image: size-independent
    properties instead.  (1) the fused score equals the reference caller's own arithmetic on the materialised sem_seg
    (evaluate_ood.py:148-150); (2) the two routes to the score (mask einsum fused into the score kernel vs stored pred_masks
    -> Variant-B score kernel) agree; (3) images are independent units: permuting the batch permutes the result bitwise;
    (4) the tensor-core and the fp32 CUDA-core backends agree within the parity bar; (5) every score is finite and inside the
    range of -sum_c tanh (-K, 0]."""
    mc = rba_b200.config.swin_b_1dl()
    sd = weights.init_state_dict(mc, seed=0, perturb=0.02)
    e = _engine(mc, sd, dev)
    e.set_gemm_backend("tc")
    g = torch.Generator().manual_seed(21)
    imgs = torch.randint(0, 256, (2, 3, 1024, 2048), dtype=torch.uint8, generator=g).to(dev)
    out = e.forward(imgs, rba=True, sem_seg=True)
    rba, sem = out["rba"], out["sem_seg"]
    assert torch.isfinite(rba).all() and float(rba.max()) <= 1e-6 and float(rba.min()) > -mc.num_classes - 1e-3
    assert (rba - (-sem.tanh().sum(1))).abs().max() < 5e-5                      # (1)
    del sem, out
    out2 = e.forward(imgs, rba=True, masks=True)                                 # pred_masks requested -> un-fused route
    assert (out2["rba"] - rba).abs().max() < 2e-4                                # (2)
    del out2
    rba_only = e.forward(imgs, rba=True)["rba"]                                  # the score-only launch (pre-scaled probabilities)
    assert (rba_only - rba).abs().max() < 1e-5
    flipped = e.forward(imgs.flip(0).contiguous(), rba=True)["rba"]
    assert torch.equal(flipped.flip(0), rba_only)                                # (3)
    del flipped, rba_only
    e.set_gemm_backend("ffma")
    ref = e.forward(imgs[:1].contiguous(), rba=True, logits=True)
    e.set_gemm_backend("tc")
    tc = e.forward(imgs[:1].contiguous(), rba=True, logits=True)
    assert (tc["pred_logits"] - ref["pred_logits"]).abs().max() < TOL           # (4)
    assert (tc["rba"] - ref["rba"]).abs().max() < TOL


# ---------------------------------------------------------------------------------------------------------------------
# Parity at the metric's shape against the UNMODIFIED reference (tests/golden/model_full_*.pt, oracle/make_golden_fullsize.py)
# ---------------------------------------------------------------------------------------------------------------------
def _unpack_ref_decisions(fix, hd):
    """The reference's boolean decisions of prediction head `hd` (pre-reset, bit-packed) -> uint8 (B,Q,S) with the
    all-blocked-row reset of mask2former_transformer_decoder.py:433 applied (that is what the cross-attention reads)."""
    import numpy as np
    shp = fix["attn_mask_shapes"][hd]
    n = int(np.prod(shp))
    bits = np.unpackbits(fix["attn_masks"][hd].numpy())[:n].reshape(shp)
    am = torch.from_numpy(bits.astype("bool"))
    am = am.clone()
    am[am.all(-1)] = False
    return am.to(torch.uint8)


def _full_errs(out, fix):
    s = fix["sub"]
    return {
        "pred_logits": (out["pred_logits"].cpu() - fix["pred_logits"]).abs().max().item(),
        "pred_masks": (out["pred_masks"].cpu()[:, :, ::s, ::s] - fix["pred_masks_sub"]).abs().max().item(),
        "sem_seg": (out["sem_seg"].cpu()[:, :, ::s, ::s] - fix["sem_seg_sub"]).abs().max().item(),
        "rba": (out["rba"].cpu()[:, ::s, ::s] - fix["rba_sub"]).abs().max().item(),
    }


FULL_RESULTS = {}


@pytest.mark.parametrize("name", ["swin_b_1dl_1024x2048", "swin_l_1dl_256x512", "swin_b_3lvl_256x512", "r50_1dl_512x1024",
                                  "r50_3lvl_192x320"])
@pytest.mark.parametrize("backend", ["tc", "ffma"])
def test_full_size_matches_reference_golden(dev, name, backend):
    """North star: "outputs match the reference's own forward on identical random-init weights and synthetic 1024x2048
    inputs within 1e-3 fp32".  The fixture is the unmodified reference run at the metric's shape.  With 100 x 2048 boolean
    attention-mask decisions per image (mask2former_transformer_decoder.py:483-486) some sit within ~1e-6 of their
    threshold (fixture `am_margin`), where ANY implementation whose mask logits differ by round-off may decide
    differently, and a flipped decision changes the outputs by more than round-off.  So parity is checked in two parts:
      (a) arithmetic: stage tensors before any decision (res2..res5 against the oracle's, which equals the reference to
          round-off -- fixture `oracle_vs_reference`; for the r50_* cases the backbone under the reference's own pixel decoder /
          transformer decoder is the detectron2 stand-in of oracle/ref_shims, BASELINE.json configs[0], backbone parity
          unpinned), and ALL outputs with the cross-attention reading the
          reference's own decisions: < 1e-3 max-abs;
      (b) decisions: every decision where this engine differs from the reference is within 1e-3 of its threshold (in
          the fixture's near list); the free-running outputs are reported, and must also meet 1e-3 when no decision flipped.
    """
    if backend == "ffma" and name == "swin_b_1dl_1024x2048":
        pytest.skip("the fp32 CUDA-core backend is the cross-check at the smaller cases; 1024x2048 runs on tcgen05")
    fix = load_golden(f"model_full_{name}.pt")
    case = fix["case"]
    mc = case_model_config(case)
    sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
    assert abs(state_checksum(sd) - fix["state_checksum"]) <= 1e-6 * fix["state_checksum"]
    assert max(fix["oracle_vs_reference"].values()) < 2e-5          # the oracle restates the reference to fp32 round-off
    imgs = torch.stack(case_images(case)).to(dev)
    B = imgs.shape[0]
    e = _engine(mc, sd, dev, taps=True)
    e.set_gemm_backend(backend)
    L = mc.dec_layers
    ref_dec = [_unpack_ref_decisions(fix, hd).to(dev) for hd in range(L)]
    dumps = [torch.zeros_like(r) for r in ref_dec]
    for hd in range(L):
        e.debug_attn_mask(hd, dump=dumps[hd])
    free = e.forward(imgs, rba=True, sem_seg=True, logits=True, masks=True)
    torch.cuda.synchronize()
    # (a1) backbone stage tensors, before any decision
    t = fix["taps"]
    stage_err = {}
    for k, (cs, ss) in {"res2": (8, 8), "res3": (8, 4), "res4": (8, 2), "res5": (8, 1)}.items():
        r = t[f"{k}_sub"]
        Cc = r.shape[1] * cs if cs > 1 else r.shape[1]
        g = e.tap(k).cpu().view(B, -1, Cc)
        hw = g.shape[1]
        # token-major (B, H*W, C) -> NCHW
        Hs = {"res2": 4, "res3": 8, "res4": 16, "res5": 32}[k]
        Hh, Ww = e.padded_hw(*imgs.shape[-2:])
        g = g.view(B, Hh // Hs, Ww // Hs, Cc).permute(0, 3, 1, 2)[:, ::cs, ::ss, ::ss]
        # relative to the stage's magnitude: Swin stages are LayerNorm outputs (O(1)); un-normalised ResNet activations of a
        # random-init network grow to O(100) by res4 / res5
        stage_err[k] = (g - r).abs().max().item() / max(1.0, r.abs().max().item())
        assert hw == (Hh // Hs) * (Ww // Hs)
    print(name, backend, "stage max-abs vs reference (relative to max(1, |ref|max))", stage_err)
    for k, v in stage_err.items():
        assert v < 5e-4, (k, v)
    # (b) decisions
    near = {(hd, b, q, p) for hd, b, q, p, _ in fix["am_near"]}
    rows_near = {(hd, b, q) for hd, b, q, p in near}
    flips, bad = [], []
    for hd in range(L):
        d = (dumps[hd] != ref_dec[hd]).nonzero().cpu().tolist()
        here = [(hd, b, q, p) for b, q, p in d]
        # Only the FIRST head in which anything differs is judged against the reference's near-threshold list: once a
        # legitimate near-threshold flip has happened, the decoder state of the later heads is no longer the reference's and
        # their decisions are compared with thresholds that have moved.  (A whole row may differ through the all-blocked
        # reset when one decision in it flipped: only the rows' own flips count.)
        if not flips:
            bad = [f for f in here if f not in near and (f[0], f[1], f[2]) not in rows_near]
        flips += here
    n_dec = sum(int(r.numel()) for r in ref_dec)
    free_errs = _full_errs(free, fix)
    print(name, backend, f"decisions differing from the reference: {len(flips)} of {n_dec} (near-threshold list: {len(near)}, "
          f"closest margin {fix['am_margin']:.2e}); free-running max-abs {free_errs}")
    assert not bad, f"decisions differ away from the threshold: {bad[:5]}"
    if not flips:
        for k, v in free_errs.items():
            assert v < TOL, ("free-running", k, v)
    # (a2) arithmetic parity of everything downstream, on the reference's decisions
    for hd in range(L):
        e.debug_attn_mask(hd, force=ref_dec[hd])
    forced = e.forward(imgs, rba=True, sem_seg=True, logits=True, masks=True)
    torch.cuda.synchronize()
    forced_errs = _full_errs(forced, fix)
    print(name, backend, "max-abs on the reference's decisions", forced_errs)
    for k, v in forced_errs.items():
        assert v < TOL, ("forced", k, v)
    # the fused route (pred_masks never materialised) on the same decisions
    fused = e.forward(imgs, rba=True, sem_seg=True)
    s = fix["sub"]
    assert (fused["rba"].cpu()[:, ::s, ::s] - fix["rba_sub"]).abs().max() < TOL
    assert (fused["sem_seg"].cpu()[:, :, ::s, ::s] - fix["sem_seg_sub"]).abs().max() < TOL
    # fraction of score pixels within the bar when free-running (what a user sees)
    frac = ((free["rba"].cpu()[:, ::s, ::s] - fix["rba_sub"]).abs() < TOL).float().mean().item()
    FULL_RESULTS[(name, backend)] = dict(stage=stage_err, flips=len(flips), decisions=n_dec, free=free_errs, forced=forced_errs,
                                         free_rba_frac_within_tol=frac)
    print(name, backend, f"free-running: {100 * frac:.3f}% of score pixels within {TOL}")
    import json
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/full_size_parity.json", "w") as f:
        json.dump({f"{k[0]}/{k[1]}": v for k, v in FULL_RESULTS.items()}, f, indent=1)


def test_arena_growth_keeps_captured_graphs_valid(dev):
    """growing the arena must not leave a captured graph replaying into freed memory.  A graph captured at a
    small shape stays replayable (its arena is retired, not freed) after a larger forward re-allocated the arena, the
    generation counter tells holders to re-capture, and ScoreStream does."""
    case = CASES["tiny_1dl"]
    mc = case_model_config(case)
    sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
    e = _engine(mc, sd, dev)
    g = torch.Generator().manual_seed(3)
    small = torch.randint(0, 256, (1, 3, 64, 96), dtype=torch.uint8, generator=g).to(dev)
    big = torch.randint(0, 256, (2, 3, 128, 192), dtype=torch.uint8, generator=g).to(dev)
    want = e.forward(small, rba=True)["rba"].clone()
    static_in, static_out, graph = e.graphed(small, rba=True)
    gen0 = e.arena_generation()
    stream = rba_b200.ScoreStream(e, 1, 64, 96)
    e.forward(big, rba=True)                                    # outgrows the arena
    assert e.arena_generation() > gen0
    filler = torch.full((64 << 20,), 1.0, device=dev)           # would land in the freed arena if it had been freed
    static_in.copy_(small)
    graph.replay()                                              # the OLD graph: still valid
    torch.cuda.synchronize()
    assert torch.equal(static_out["rba"], want)
    got = list(stream.run([small.cpu().pin_memory()]))          # ScoreStream re-captures by itself
    assert torch.equal(got[0], want.cpu())
    s2, o2, g2 = e.graphed(small, rba=True)                     # the engine's cache re-captures in the new arena
    assert g2 is not graph
    s2.copy_(small)
    g2.replay()
    torch.cuda.synchronize()
    assert torch.equal(o2["rba"], want)
    del filler
    e.release_retired()


def test_panoptic_on_model_matches_host_inference(dev):
    """MODEL.MASK_FORMER.TEST.PANOPTIC_ON: `model(x)[0]["panoptic_seg"]` (maskformer_model.py:336-340) = the reference's
    panoptic_inference restated in rba_b200/panoptic.py (pinned on the reference in tests/test_panoptic.py) applied to this
    model's own post-processed head outputs; the OoD segments come from the fused kernel's RbA score."""
    from rba_b200.panoptic import panoptic_inference
    case = CASES["tiny_1dl"]
    mc = case_model_config(case)
    sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
    y = {"MODEL": {"META_ARCHITECTURE": "MaskFormer", "PIXEL_MEAN": list(mc.pixel_mean), "PIXEL_STD": list(mc.pixel_std),
                   "BACKBONE": {"NAME": "D2SwinTransformer"},
                   "SWIN": {"EMBED_DIM": mc.embed_dim, "DEPTHS": list(mc.depths), "NUM_HEADS": list(mc.num_heads), "WINDOW_SIZE": 12,
                            "MLP_RATIO": 4.0, "PATCH_SIZE": 4, "APE": False, "QKV_BIAS": True, "PATCH_NORM": True},
                   "SEM_SEG_HEAD": {"PIXEL_DECODER_NAME": "MSDeformAttnPixelDecoder", "NORM": "GN", "CONVS_DIM": 256, "MASK_DIM": 256,
                                    "NUM_CLASSES": mc.num_classes, "IN_FEATURES": ["res2", "res3", "res4", "res5"],
                                    "DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES": list(mc.transformer_in_features),
                                    "COMMON_STRIDE": 4, "TRANSFORMER_ENC_LAYERS": mc.enc_layers},
                   "MASK_FORMER": {"TRANSFORMER_DECODER_NAME": "MultiScaleMaskedTransformerDecoder", "PRE_NORM": False, "NHEADS": 8,
                                   "HIDDEN_DIM": 256, "DIM_FEEDFORWARD": 2048, "DEC_LAYERS": mc.dec_layers + 1,
                                   "NUM_OBJECT_QUERIES": mc.num_queries, "SIZE_DIVISIBILITY": 32, "OPEN_PANOPTIC": True,
                                   "TEST": {"SEMANTIC_ON": True, "PANOPTIC_ON": True, "INSTANCE_ON": False,
                                            "OBJECT_MASK_THRESHOLD": 0.0, "OVERLAP_THRESHOLD": 0.3}}},
         "DATASETS": {"TRAIN": ["cityscapes_fine_sem_seg_train"]}}
    model = rba_b200.MaskFormer(y)
    assert model.panoptic_on and model.open_panoptic
    model.thing_ids = [11, 12, 13, 14, 15, 16, 17, 18]
    model.load_state_dict(sd)
    model.to(dev).eval()
    x = case_images(case)[0].to(dev)
    H, W = x.shape[-2:]
    out = model([{"image": x}], panoptic_ood_threshold=-15.0, panoptic_pixel_min=4, return_panoptic_ood=True)[0]
    pan, info, ood = out["panoptic_seg"]
    assert pan.shape == (H, W) and pan.dtype == torch.int32 and out["sem_seg"].shape == (mc.num_classes, H, W)
    # the same from the plain model's head outputs
    plain = rba_b200.MaskFormer(mc)
    plain.load_state_dict(sd)
    plain.to(dev).eval()
    res, cls, up = plain([{"image": x}], return_separately=True)
    want = panoptic_inference(cls, up[:, :H, :W], mc.num_classes, 0.0, 0.3, model.thing_ids, True, -15.0, 4, True)
    assert (ood - want[2]).abs().max() < 2e-4                       # fused-kernel RbA vs the reference arithmetic on the masks
    agree = (pan == want[0]).float().mean().item()
    assert agree > 0.999 and len(info) == len(want[1]), (agree, len(info), len(want[1]))
    assert (out["sem_seg"] - res[0]["sem_seg"]).abs().max() < 1e-5


def test_mixed_size_batch_is_padded_like_image_list(dev):
    """Images of different sizes in one batch: detectron2's ImageList.from_tensors pads them (zeros after normalisation) to the
    largest one (maskformer_model.py:255-257) and sem_seg_postprocess crops each result to its image (:318-322).  Oracle: the
    restatement of exactly that (oracle.preprocess / forward)."""
    case = CASES["tiny_1dl"]
    mc = case_model_config(case)
    sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
    g = torch.Generator().manual_seed(5)
    ims = [torch.randint(0, 256, (3, 64, 96), generator=g, dtype=torch.uint8),
           torch.randint(0, 256, (3, 48, 80), generator=g, dtype=torch.uint8),
           torch.randint(0, 256, (3, 64, 70), generator=g, dtype=torch.uint8)]
    ref = O.forward(sd, mc, ims)
    model = rba_b200.MaskFormer(mc)
    model.load_state_dict(sd)
    model.to(dev).eval()
    out = model([{"image": im.to(dev)} for im in ims])
    scores = model.rba([{"image": im.to(dev)} for im in ims])
    for b, im in enumerate(ims):
        assert out[b]["sem_seg"].shape[-2:] == im.shape[-2:] == scores[b].shape
        assert (out[b]["sem_seg"].cpu() - ref["sem_seg"][b]).abs().max() < 1e-3
        assert (scores[b].cpu() - ref["rba"][b]).abs().max() < 1e-3

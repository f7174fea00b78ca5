"""Host-side mirror of the reference's plugin interface for the hot path.

`MaskFormer` here is what `META_ARCH_REGISTRY.get("MaskFormer")` resolves to when rba_b200 is plugged in
(INTEGRATION.md): same constructor contract (`MaskFormer(cfg)` / `from_config`), same `state_dict` keys as the
reference model (mask2former/maskformer_model.py:23-24), same call contract
`model([{"image": CHW tensor, ...}]) -> [{"sem_seg": (K,H,W) tensor}]` (maskformer_model.py:227-356), same
error behaviour for the things it does not do (training, panoptic/instance inference raise).  All arithmetic
runs in librba_b200.so; torch only carries buffers."""
from collections import OrderedDict

import torch
from torch import nn
from torch.nn import functional as F

from . import weights
from ._lib import RbaError
from .config import ModelConfig, model_config_from_cfg
from .engine import Engine


def _opt(node, key, default):
    try:
        return node[key] if isinstance(node, dict) else getattr(node, key)
    except (KeyError, AttributeError):
        return default


def _thing_ids(cfg):
    """MetadataCatalog.get(cfg.DATASETS.TRAIN[0]).thing_dataset_id_to_contiguous_id.values() (maskformer_model.py:204,424)."""
    try:
        from detectron2.data import MetadataCatalog
        ds = cfg["DATASETS"]["TRAIN"][0] if isinstance(cfg, dict) else cfg.DATASETS.TRAIN[0]
        return sorted(int(v) for v in MetadataCatalog.get(ds).thing_dataset_id_to_contiguous_id.values())
    except Exception:
        return []


class MaskFormer(nn.Module):
    """B200-native drop-in for the reference meta-architecture (eval branch: semantic + panoptic / open-panoptic)."""

    def __init__(self, cfg, seed=0):
        super().__init__()
        self.mc = cfg.validate() if isinstance(cfg, ModelConfig) else model_config_from_cfg(cfg)
        self.num_queries = self.mc.num_queries
        self.size_divisibility = self.mc.size_divisibility
        self.semantic_on, self.panoptic_on, self.instance_on = True, False, False
        self.sem_seg_postprocess_before_inference = False
        # panoptic / open-panoptic inference (maskformer_model.py:202-220,336-340,394-481): host-side post-processing of the
        # head outputs, rba_b200/panoptic.py
        self.object_mask_threshold, self.overlap_threshold, self.open_panoptic = 0.8, 0.8, True
        self.thing_ids = []
        if not isinstance(cfg, ModelConfig):
            M = cfg["MODEL"] if isinstance(cfg, dict) else cfg.MODEL
            mf = M["MASK_FORMER"] if isinstance(M, dict) else M.MASK_FORMER
            test = _opt(mf, "TEST", None)
            if test is not None:
                self.panoptic_on = bool(_opt(test, "PANOPTIC_ON", False))
                self.semantic_on = bool(_opt(test, "SEMANTIC_ON", True))
                self.object_mask_threshold = float(_opt(test, "OBJECT_MASK_THRESHOLD", 0.8))
                self.overlap_threshold = float(_opt(test, "OVERLAP_THRESHOLD", 0.8))
                if bool(_opt(test, "INSTANCE_ON", False)):
                    raise RbaError("MODEL.MASK_FORMER.TEST.INSTANCE_ON is not built")
            self.open_panoptic = bool(_opt(mf, "OPEN_PANOPTIC", True))
            if self.panoptic_on:
                self.thing_ids = _thing_ids(cfg)
        self.sem_seg_postprocess_before_inference = self.panoptic_on
        # host master copy of the reference-layout state_dict (what DetectionCheckpointer reads and writes)
        self._sd = weights.init_state_dict(self.mc, seed=seed)
        self._device = torch.device("cpu")
        self._engine = None
        self.training = False

    @classmethod
    def from_config(cls, cfg):
        return {"cfg": cfg}

    # ---- nn.Module surface used by the reference's callers (evaluate_ood.py:108-124) ----
    @property
    def device(self):
        return self._device

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        out = destination if destination is not None else OrderedDict()
        for k, v in self._sd.items():
            out[prefix + k] = v
        return out

    def load_state_dict(self, state_dict, strict=True, assign=False):
        sd = OrderedDict(state_dict)
        # legacy key upgrades of the reference (mask_former_head.py:31-53, mask2former_transformer_decoder.py:237-258)
        for k in list(sd.keys()):
            nk = k
            if k.startswith("sem_seg_head.") and not k.startswith("sem_seg_head.predictor") \
                    and not k.startswith("sem_seg_head.pixel_decoder."):
                nk = k.replace("sem_seg_head.", "sem_seg_head.pixel_decoder.", 1)
            nk = nk.replace("static_query", "query_feat")
            if nk != k:
                sd[nk] = sd.pop(k)
        missing = [k for k in self._sd if k not in sd]
        unexpected = [k for k in sd if k not in self._sd]
        errors = []
        for k, v in sd.items():
            if k in self._sd:
                if tuple(v.shape) != tuple(self._sd[k].shape):
                    errors.append(f"size mismatch for {k}: {tuple(v.shape)} vs {tuple(self._sd[k].shape)}")
                else:
                    self._sd[k] = v.detach().to("cpu", self._sd[k].dtype).contiguous().clone()
        if errors or (strict and (missing or unexpected)):
            raise RuntimeError("Error(s) in loading state_dict for MaskFormer: " + "; ".join(
                errors + ([f"missing {missing}"] if strict and missing else []) + ([f"unexpected {unexpected}"] if strict and unexpected else [])))
        self._engine = None  # weights changed: rebuild lazily
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def to(self, device=None, *args, **kwargs):
        if device is not None:
            device = torch.device(device)
            if device.type == "cuda" and device.index is None:
                device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
            if device != self._device:
                self._device = device
                self._engine = None
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device if device is not None else 0))

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        if mode:
            raise RbaError("rba_b200.MaskFormer is inference-only (training is out of scope of the hot path)")
        return self.eval()

    # ---- engine ----
    def engine(self):
        if self._engine is None:
            if self._device.type != "cuda":
                raise RbaError("rba_b200.MaskFormer runs on CUDA only: call .to('cuda') (there is no CPU fallback)")
            self._engine = Engine(self.mc, self._device).load_state_dict(self._sd)
        return self._engine

    def _batch(self, batched_inputs):
        """(B,3,Hmax,Wmax) batch + the per-image (h, w).  Images of different sizes are padded bottom / right to the largest one
        like detectron2's ImageList.from_tensors does (maskformer_model.py:255-257: zeros AFTER normalisation, i.e. the pixel
        mean before it); the engine then pads the batch to SIZE_DIVISIBILITY the same way."""
        ims = [x["image"] for x in batched_inputs]
        sizes = [(int(im.shape[-2]), int(im.shape[-1])) for im in ims]
        if len(set(sizes)) == 1:
            dt = {im.dtype for im in ims}
            tgt = torch.uint8 if dt == {torch.uint8} else torch.float32
            ims = [im.to(self._device, tgt, non_blocking=True) for im in ims]
            return torch.stack(ims).contiguous(), sizes
        Hm, Wm = max(h for h, _ in sizes), max(w for _, w in sizes)
        mean = torch.tensor(self.mc.pixel_mean, dtype=torch.float32, device=self._device).view(3, 1, 1)
        batch = mean.expand(3, Hm, Wm).repeat(len(ims), 1, 1, 1).contiguous()
        for b, im in enumerate(ims):
            batch[b, :, :sizes[b][0], :sizes[b][1]] = im.to(self._device, torch.float32, non_blocking=True)
        return batch, sizes

    @torch.no_grad()
    def forward(self, batched_inputs, include_void=False, return_separately=False, return_aux=False,
                return_ood_pred=False, panoptic_ood_threshold=-0.3, panoptic_pixel_min=200, return_panoptic_ood=False):
        """maskformer_model.py:227-356, eval branch (SEMANTIC_ON, and PANOPTIC_ON with the open-panoptic OoD segments)."""
        if return_aux:
            raise RbaError("return_aux (auxiliary decoder outputs) is not built")
        if self.panoptic_on:
            return self._forward_panoptic(batched_inputs, include_void, panoptic_ood_threshold, panoptic_pixel_min, return_panoptic_ood)
        if return_ood_pred and not self.mc.ood_prediction:
            raise RbaError("return_ood_pred needs the DenseHybrid head (MODEL.MASK_FORMER.DENSE_HYBRID_LOSS: True)")
        images, sizes = self._batch(batched_inputs)
        B, _, H, W = images.shape
        eng = self.engine()
        eng.set_score("rba", include_void=include_void)     # semantic_inference_with_void (maskformer_model.py:388-392)
        out = eng.forward(images, rba=False, sem_seg=True, logits=return_separately, masks=return_separately,
                          ood_pred=return_ood_pred)
        results = []
        for b, inp in enumerate(batched_inputs):
            hi, wi = sizes[b]
            r = out["sem_seg"][b][:, :hi, :wi]               # sem_seg_postprocess: crop to the image ...
            h, w = inp.get("height", hi), inp.get("width", wi)
            if (h, w) != (hi, wi):  # ... and resize (detectron2): bilinear, align_corners=False
                r = F.interpolate(r[None], size=(h, w), mode="bilinear", align_corners=False)[0]
            results.append({"sem_seg": r})
        if return_separately:
            Hp, Wp = self.engine().padded_hw(H, W)
            up = F.interpolate(out["pred_masks"][-1:], size=(Hp, Wp), mode="bilinear", align_corners=False)[0]
            return results, out["pred_logits"][-1], up
        if return_ood_pred:                                  # maskformer_model.py:303-305,350-351
            return results, out["ood_pred"]
        return results

    def _forward_panoptic(self, batched_inputs, include_void, ood_threshold, pixel_min, return_panoptic_ood):
        """PANOPTIC_ON: sem_seg_postprocess runs BEFORE inference (maskformer_model.py:206-209,318-322), then semantic and panoptic
        inference on the post-processed masks (:325-340).  The engine supplies sem_seg, the RbA score (= the open-panoptic
        branch's ood_mask, :456-458) and the head outputs in one forward."""
        from .panoptic import panoptic_inference
        images, sizes = self._batch(batched_inputs)
        if len(set(sizes)) != 1:
            raise RbaError("rba_b200.MaskFormer: the panoptic branch takes images of one size per batch")
        B, _, H, W = images.shape
        eng = self.engine()
        eng.set_score("rba", include_void=include_void)
        out = eng.forward(images, rba=True, sem_seg=self.semantic_on, logits=True, masks=True)
        Hp, Wp = eng.padded_hw(H, W)
        results = []
        for b, inp in enumerate(batched_inputs):
            h, w = inp.get("height", H), inp.get("width", W)
            up = F.interpolate(out["pred_masks"][b:b + 1], size=(Hp, Wp), mode="bilinear", align_corners=False)[:, :, :H, :W]
            rba = out["rba"][b]
            r = {}
            if (h, w) != (H, W):                 # sem_seg_postprocess: crop (above) + bilinear resize to the requested size
                up = F.interpolate(up, size=(h, w), mode="bilinear", align_corners=False)
                rba = None                       # the score of the resized masks is recomputed from them
            if self.semantic_on:
                sem = out["sem_seg"][b]
                if (h, w) != (H, W):
                    sem = F.interpolate(sem[None], size=(h, w), mode="bilinear", align_corners=False)[0]
                r["sem_seg"] = sem
            r["panoptic_seg"] = panoptic_inference(out["pred_logits"][b], up[0], self.mc.num_classes, self.object_mask_threshold,
                                                   self.overlap_threshold, self.thing_ids, self.open_panoptic, ood_threshold,
                                                   pixel_min, return_panoptic_ood, ood_mask=rba)
            results.append(r)
        return results

    @torch.no_grad()
    def rba(self, batched_inputs):
        """Fused path: evaluate_ood.get_RbA (evaluate_ood.py:143-150) without materialising sem_seg.
        Returns a (B,H,W) tensor of anomaly scores."""
        return self.score(batched_inputs, "rba")

    @torch.no_grad()
    def score(self, batched_inputs, score_func="rba"):
        """Fused anomaly score of evaluate_ood.py --score_func: "rba" (get_RbA, :143-150), "pebal"/"energy"
        (get_energy, :152-159: -logsumexp over the class planes) or "dense_hybrid" (get_densehybrid_score, :161-173:
        energy + log p(outlier) from the ood_pred head).  sem_seg is never materialised.  Returns (B,H,W), or a list of per-image
        maps when the images of the batch differ in size."""
        images, sizes = self._batch(batched_inputs)
        eng = self.engine()
        eng.set_score(score_func, include_void=False)
        score = eng.forward(images, rba=True)["rba"]
        if len(set(sizes)) == 1:
            return score
        return [score[b, :h, :w] for b, (h, w) in enumerate(sizes)]      # mixed sizes: one cropped map per image


def build_model(cfg):
    """detectron2.modeling.build_model equivalent for this meta-arch: construct + move to cfg.MODEL.DEVICE."""
    name = cfg["MODEL"]["META_ARCHITECTURE"] if isinstance(cfg, dict) else cfg.MODEL.META_ARCHITECTURE
    if name != "MaskFormer":
        raise KeyError(f"No object named '{name}' found in 'META_ARCH' registry!")
    m = MaskFormer(cfg)
    dev = cfg["MODEL"]["DEVICE"] if isinstance(cfg, dict) else cfg.MODEL.DEVICE
    return m.to(torch.device(dev))

"""Python handle on the C++/CUDA engine (rba_model_* / rba_forward in include/rba_b200.h)."""
import ctypes

import torch

from . import _lib
from ._lib import RBA_GEMM_FFMA, RBA_GEMM_TC, RbaError


class Engine:
    """Owns one `rba_model` on one CUDA device.  torch is used for device buffers and the stream only."""

    def __init__(self, model_config, device=0):
        if not torch.cuda.is_available():
            raise RbaError("rba_b200.Engine needs a CUDA device: the product path has no CPU fallback")
        self.mc = model_config.validate()
        self.device = torch.device("cuda", device if isinstance(device, int) else (device.index or 0))
        self._h = ctypes.c_void_p()
        cfg = self.mc.to_ctypes()
        _lib.check(_lib.lib().rba_model_create(ctypes.byref(cfg), self.device.index, ctypes.byref(self._h)))
        self._finalized = False
        self._graphs = {}
        self._opts = {}

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().rba_model_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights ----
    def load_state_dict(self, sd):
        """sd: reference-layout state_dict (name -> tensor).  Integer buffers (relative_position_index) and
        training-only tensors are skipped; missing tensors are reported by finalize()."""
        L = _lib.lib()
        for k, v in sd.items():
            if not torch.is_tensor(v) or not v.is_floating_point():
                continue
            t = v.detach().to("cpu", torch.float32).contiguous()
            if t.numel() == 0:
                continue
            shape = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
            _lib.check(L.rba_model_load_tensor(self._h, k.encode(), ctypes.c_void_p(t.data_ptr()), shape, t.dim()))
        _lib.check(L.rba_model_finalize(self._h))
        self._finalized = True
        return self

    def set_option(self, name, value):
        _lib.check(_lib.lib().rba_model_set_option(self._h, name.encode(), int(value)))
        self._opts[name] = int(value)
        self._graphs.clear()

    def set_score(self, func="rba", include_void=False):
        """Per-pixel reduction of the score output: "rba" (evaluate_ood.get_RbA), "energy"/"pebal" (get_energy) or
        "densehybrid" (get_densehybrid_score; needs the ood_pred head's weights);
        include_void keeps the void column (semantic_inference_with_void): sem_seg gets K+1 planes."""
        from .ops import SCORE_FUNCS
        if func not in SCORE_FUNCS:
            raise RbaError(f"unknown score function {func!r} (one of {sorted(SCORE_FUNCS)})")
        if self._opts.get("score_func", 0) != SCORE_FUNCS[func]:
            self.set_option("score_func", SCORE_FUNCS[func])
        if self._opts.get("include_void", 0) != int(bool(include_void)):
            self.set_option("include_void", int(bool(include_void)))

    def set_gemm_backend(self, name):
        """'tc': tcgen05 GEMMs + tcgen05 window attention; 'ffma': exact fp32 CUDA-core kernels everywhere."""
        self.set_option("gemm_backend", {"ffma": RBA_GEMM_FFMA, "tc": RBA_GEMM_TC}[name])
        self.set_option("attn_backend", 2 if name == "tc" else 0)

    def reserve(self, B, H, W):
        _lib.check(_lib.lib().rba_model_reserve(self._h, B, H, W))

    def arena_generation(self):
        """Increments whenever the workspace arena was re-allocated (a larger shape was seen): graphs captured before
        that must be re-captured.  `graphed()` checks it; holders of a graph (ScoreStream) compare it themselves."""
        return int(_lib.lib().rba_model_arena_generation(self._h))

    def release_retired(self):
        """Frees outgrown arenas that earlier graph captures ran in.  Call only after dropping every graph captured
        before the last growth (this engine's own cache is dropped here)."""
        self._graphs.clear()
        _lib.check(_lib.lib().rba_model_release_retired(self._h))

    def debug_attn_mask(self, head, force=None, dump=None):
        """Parity aid (rba_model_debug_attn_mask): uint8 CUDA tensors (B,Q,S_l); keeps them alive on the engine."""
        self._dbg = getattr(self, "_dbg", {})
        self._dbg[head] = (force, dump)
        _lib.check(_lib.lib().rba_model_debug_attn_mask(
            self._h, head, ctypes.c_void_p(force.data_ptr()) if force is not None else None,
            ctypes.c_void_p(dump.data_ptr()) if dump is not None else None))
        self._graphs.clear()

    # ---- forward ----
    def padded_hw(self, H, W):
        s = self.mc.size_divisibility
        return (H + s - 1) // s * s, (W + s - 1) // s * s

    def alloc_outputs(self, B, H, W, rba=True, sem_seg=False, logits=False, masks=False, ood_pred=False):
        Hp, Wp = self.padded_hw(H, W)
        dev, K, Q = self.device, self.mc.num_classes, self.mc.num_queries
        out = {}
        if rba:
            out["rba"] = torch.empty((B, H, W), dtype=torch.float32, device=dev)
        if sem_seg:
            out["sem_seg"] = torch.empty((B, K + self._opts.get("include_void", 0), H, W), dtype=torch.float32, device=dev)
        if logits:
            out["pred_logits"] = torch.empty((B, Q, K + 1), dtype=torch.float32, device=dev)
        if masks:
            out["pred_masks"] = torch.empty((B, Q, Hp // 4, Wp // 4), dtype=torch.float32, device=dev)
        if ood_pred:
            out["ood_pred"] = torch.empty((B, 2, H, W), dtype=torch.float32, device=dev)
        return out

    def forward_into(self, images, out):
        """images: (B,3,H,W) uint8 or float32 CUDA tensor (RAW pixel values, like the reference's input dicts).
        Launches on torch's current stream; `out` from alloc_outputs()."""
        if not self._finalized:
            raise RbaError("Engine.forward: load_state_dict() first")
        if not images.is_cuda or images.device != self.device or not images.is_contiguous():
            raise RbaError("Engine.forward: images must be a contiguous CUDA tensor on the engine's device")
        if images.dtype == torch.uint8:
            dt = _lib.RBA_IMG_U8
        elif images.dtype == torch.float32:
            dt = _lib.RBA_IMG_F32
        else:
            raise RbaError(f"Engine.forward: unsupported image dtype {images.dtype}")
        B, C, H, W = images.shape
        if C != 3:
            raise RbaError("Engine.forward: images must be (B,3,H,W)")
        p = lambda k: out[k].data_ptr() if k in out else None  # noqa: E731
        st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        o = _lib.RbaOutputs(p("rba"), p("sem_seg"), p("pred_logits"), p("pred_masks"), p("ood_pred"))
        _lib.check(_lib.lib().rba_forward_ex(self._h, ctypes.c_void_p(images.data_ptr()), dt, B, H, W, ctypes.byref(o), st))
        return out

    def forward(self, images, rba=True, sem_seg=False, logits=False, masks=False, ood_pred=False):
        B, _, H, W = images.shape
        return self.forward_into(images, self.alloc_outputs(B, H, W, rba, sem_seg, logits, masks, ood_pred))

    def tap(self, name):
        """Stage-boundary tensor of the last forward (needs set_option('taps', 1) before that forward)."""
        n = ctypes.c_int64()
        L = _lib.lib()
        _lib.check(L.rba_model_get_tap(self._h, name.encode(), None, 0, ctypes.byref(n), None))
        t = torch.empty(n.value, dtype=torch.float32, device=self.device)
        st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(L.rba_model_get_tap(self._h, name.encode(), ctypes.c_void_p(t.data_ptr()), n.value, ctypes.byref(n), st))
        return t

    # ---- CUDA graph replay of a fixed-shape forward ----
    def graphed(self, images, rba=True, sem_seg=False, logits=False, masks=False):
        """Captures rba_forward for this (shape, dtype, outputs) once and returns (static_images, static_out, replay)."""
        key = (tuple(images.shape), images.dtype, rba, sem_seg, logits, masks)
        B, _, H, W = images.shape
        self.reserve(B, H, W)
        if key in self._graphs and self._graphs[key][3] != self.arena_generation():
            del self._graphs[key]                       # captured in an arena that has since been outgrown
        if key not in self._graphs:
            static_in = torch.empty_like(images)
            static_in.copy_(images)
            out = self.alloc_outputs(B, H, W, rba, sem_seg, logits, masks)
            s = torch.cuda.Stream(self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                self.forward_into(static_in, out)       # warm-up outside capture (position tables, func attributes)
            torch.cuda.current_stream(self.device).wait_stream(s)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.forward_into(static_in, out)
            self._graphs[key] = (static_in, out, g, self.arena_generation())
        return self._graphs[key][:3]

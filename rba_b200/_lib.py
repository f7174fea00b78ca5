"""ctypes binding of include/rba_b200.h.  Loads the in-tree librba_b200.so and FAILS LOUDLY if it is
missing or a call errors: there is no CPU / PyTorch fallback anywhere in this package."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RBA_B200_LIB") or os.path.join(_HERE, "lib", "librba_b200.so")   # env: A/B builds (tools/)

RBA_IMG_U8, RBA_IMG_F32 = 0, 1
RBA_ACT_NONE, RBA_ACT_RELU, RBA_ACT_GELU = 0, 1, 2
RBA_GEMM_FFMA, RBA_GEMM_TC = 0, 1
RBA_SCORE_RBA, RBA_SCORE_ENERGY, RBA_SCORE_DENSEHYBRID = 0, 1, 2


class RbaError(RuntimeError):
    pass


class RbaConfig(Structure):
    _fields_ = [
        ("embed_dim", c_int32), ("depths", c_int32 * 4), ("num_heads", c_int32 * 4), ("window_size", c_int32),
        ("conv_dim", c_int32), ("mask_dim", c_int32), ("num_classes", c_int32), ("num_queries", c_int32),
        ("nheads", c_int32), ("dim_feedforward", c_int32), ("dec_layers", c_int32), ("enc_layers", c_int32),
        ("enc_points", c_int32), ("enc_ffn", c_int32), ("num_enc_levels", c_int32), ("size_divisibility", c_int32),
        ("pixel_mean", c_float * 3), ("pixel_std", c_float * 3),
        ("backbone_type", c_int32), ("resnet_depth", c_int32),
    ]


class RbaOutputs(Structure):
    _fields_ = [("rba", c_void_p), ("sem_seg", c_void_p), ("pred_logits", c_void_p), ("pred_masks", c_void_p),
                ("ood_pred", c_void_p)]


class RbaGemmArgs(Structure):
    _fields_ = [
        ("a_hi", c_void_p), ("a_lo", c_void_p), ("lda", c_int64),
        ("w_hi", c_void_p), ("w_lo", c_void_p), ("ldw", c_int64),
        ("M", c_int32), ("N", c_int32), ("K", c_int32), ("batch", c_int32),
        ("a_bstride", c_int64), ("w_bstride", c_int64),
        ("bias", c_void_p), ("bias_per_row", c_int32), ("bias_bstride", c_int64),
        ("act", c_int32), ("residual", c_void_p),
        ("c", c_void_p), ("ldc", c_int64), ("c_bstride", c_int64),
        ("c_hi", c_void_p), ("c_lo", c_void_p), ("ldcp", c_int64), ("cp_bstride", c_int64),
        ("swin_map", c_int32), ("sw_H", c_int32), ("sw_W", c_int32), ("sw_ws", c_int32), ("sw_shift", c_int32),
        ("backend", c_int32), ("qkv_tile_heads", c_int32),
    ]


# symbol -> (restype, argtypes); every symbol declared in include/rba_b200.h appears here
PROTOTYPES = {
    "rba_last_error": (c_char_p, []),
    "rba_version": (c_int, []),
    "rba_launch_count": (c_int64, []),
    "rba_model_create": (c_int, [POINTER(RbaConfig), c_int, POINTER(c_void_p)]),
    "rba_model_destroy": (None, [c_void_p]),
    "rba_model_load_tensor": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int]),
    "rba_model_finalize": (c_int, [c_void_p]),
    "rba_model_set_option": (c_int, [c_void_p, c_char_p, c_int]),
    "rba_model_reserve": (c_int, [c_void_p, c_int, c_int, c_int]),
    "rba_model_arena_generation": (c_int64, [c_void_p]),
    "rba_model_release_retired": (c_int, [c_void_p]),
    "rba_model_debug_attn_mask": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "rba_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rba_forward_ex": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, POINTER(RbaOutputs), c_void_p]),
    "rba_model_get_tap": (c_int, [c_void_p, c_char_p, c_void_p, c_int64, POINTER(c_int64), c_void_p]),
    "rba_score_fused": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rba_einsum_score_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                       c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rba_k_set_fused_score_variant": (c_int, [c_int]),
    "rba_jpeg_available": (c_int, []),
    "rba_jpeg_info": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "rba_jpeg_decode": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "rba_ood_hist_bytes": (c_int64, []),
    "rba_ood_workspace_bytes": (c_int64, []),
    "rba_ood_hist_update": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "rba_ood_hist_finalize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "rba_msda_forward": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64), c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "rba_msda_backward": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64), c_void_p, c_void_p, c_void_p, c_int, c_int,
                                  c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rba_msda_forward_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                     c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "rba_msda_backward_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                      c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rba_k_ood_pred_resize": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rba_outlier_loss_workspace_floats": (c_int64, [c_int, c_int, c_int, c_int, c_int]),
    "rba_outlier_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rba_k_split": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    "rba_k_gemm": (c_int, [POINTER(RbaGemmArgs), c_void_p]),
    "rba_k_conv3x3": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "rba_k_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "rba_k_window_attn": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rba_k_window_attn_planes": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rba_k_window_attn_bias_floats": (c_int64, [c_int]),
    "rba_k_window_attn_prepare_bias": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "rba_k_window_attn_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rba_k_window_attn_tc_tiled": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rba_k_mha_workspace_floats": (c_int64, [c_int, c_int, c_int, c_int]),
    "rba_k_mha": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                          c_void_p]),
    "rba_k_groupnorm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_int,
                                c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rba_k_groupnorm_ws": (c_int64, [c_int, c_int, c_int, c_int, c_int]),
    "rba_k_patch_embed": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float),
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "rba_k_stem_conv": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float),
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "rba_k_maxpool3x3s2": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rba_k_bias_act_sub": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rba_k_attn_mask": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
}

_lib = None


def lib():
    """The loaded library (ctypes.CDLL) with prototypes applied."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RbaError(
                f"{LIB_PATH} is missing: build it with `python -m rba_b200.build` (or __graft_entry__.build()). "
                "rba_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RbaError(lib().rba_last_error().decode("utf-8", "replace") + f" (rba status {rc})")


def launch_count():
    return int(lib().rba_launch_count())

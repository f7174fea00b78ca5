"""Drop-in for the reference's compiled extension module `MultiScaleDeformableAttention`
(mask2former/modeling/pixel_decoder/ops/src/vision.cpp:18-21).  Put this directory on sys.path (or copy this
file next to the reference) and the reference's unmodified Python wrapper
ops/functions/ms_deform_attn_func.py:21,36 (`import MultiScaleDeformableAttention as MSDA`) binds to the
B200 kernel.  Same two functions, same argument order and results (forward and backward)."""
from rba_b200.ops import ms_deform_attn_backward, ms_deform_attn_forward  # noqa: F401

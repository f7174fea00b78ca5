"""Generates tests/golden/msda_grad.pt: gradients of the reference's own CPU statement of MSDeformAttn
(`ms_deform_attn_core_pytorch`, ops/functions/ms_deform_attn_func.py:52-72, imported UNMODIFIED from /root/reference)
by autograd in float64 -- what the reference's `check_gradient_numerical` (ops/test.py:66-88) compares its CUDA
backward against -- at the reference test's shapes / seed and its channel list (30, 32, 64, 71, 1025; the larger ones
are omitted to keep the fixture small), plus an out-of-range-heavy case at the model's head layout.
TEST INFRASTRUCTURE; run manually:   python oracle/make_golden_msda_grad.py
"""
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
warnings.filterwarnings("ignore")
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def make_inputs(seed, N, M, D, Lq, L, P, shapes, loc_range=(0.0, 1.0), value_scale=0.01):
    """Deterministic CPU inputs (same recipe in tests/test_kernels_gpu.py: only the gradients are stored)."""
    g = torch.Generator().manual_seed(seed)
    S = int((shapes[:, 0] * shapes[:, 1]).sum())
    value = torch.rand(N, S, M, D, generator=g) * value_scale
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * (loc_range[1] - loc_range[0]) + loc_range[0]
    aw = torch.rand(N, Lq, M, L, P, generator=g) + 1e-5
    aw = aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    go = torch.randn(N, Lq, M * D, generator=g)
    return value, loc, aw, go


CASES = {
    # ops/test.py:24-28 shapes, channels of ops/test.py:82-88
    **{f"ref_c{c}": dict(seed=3 + c, N=1, M=2, D=c, Lq=2, L=2, P=2, shapes=[(6, 4), (3, 2)]) for c in (30, 32, 64, 71, 1025)},
    # the model's head layout, 3 levels, a third of the samples outside the maps
    "big": dict(seed=101, N=2, M=8, D=32, Lq=37, L=3, P=4, shapes=[(5, 7), (10, 14), (3, 4)], loc_range=(-0.3, 1.3), value_scale=1.0),
}


def main():
    core = ref_loader.msda_core_pytorch()
    fix = {}
    for name, c in CASES.items():
        shapes = torch.as_tensor(c["shapes"], dtype=torch.long)
        value, loc, aw, go = make_inputs(c["seed"], c["N"], c["M"], c["D"], c["Lq"], c["L"], c["P"], shapes,
                                         c.get("loc_range", (0.0, 1.0)), c.get("value_scale", 0.01))
        v, l, a = (t.double().requires_grad_(True) for t in (value, loc, aw))
        out = core(v, shapes, l, a)
        out.backward(go.double())
        fix[name] = {"case": c, "out": out.detach().float(), "grad_value": v.grad.float(), "grad_loc": l.grad.float(),
                     "grad_aw": a.grad.float(), "in_checksum": float(value.double().sum() + loc.double().sum() + go.double().sum())}
        print(name, tuple(out.shape), float(v.grad.abs().max()), float(l.grad.abs().max()), float(a.grad.abs().max()))
    torch.save(fix, os.path.join(OUT, "msda_grad.pt"))


if __name__ == "__main__":
    main()

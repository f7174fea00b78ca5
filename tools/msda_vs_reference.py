"""rba_msda_forward against the reference's own CUDA kernel (baseline/_ref/MultiScaleDeformableAttention*.so, kernels
untouched, rebuilt for sm_100a) on the model's shapes: us per call (CUDA events, 50 calls after 10 warm-ups) and max-abs
difference.  Also the engine's fused variant's context: the drop-in op is what a user of the reference's FFI gets."""
import glob
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rba_b200 import ops  # noqa: E402


def load_ext():
    so = glob.glob(os.path.join(ROOT, "baseline", "_ref", "MultiScaleDeformableAttention*.so"))[0]
    spec = importlib.util.spec_from_file_location("MultiScaleDeformableAttention", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def timeit(fn, n=50, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    ext = load_ext()
    dev = torch.device("cuda", 0)
    out = []
    for name, shapes, B in [("1 level 32x64 (Swin-B 1dl @1024x2048), B=8", [(32, 64)], 8),
                            ("3 levels res5/res4/res3 @1024x2048, B=8", [(32, 64), (64, 128), (128, 256)], 8),
                            ("3 levels @1024x2048, B=1", [(32, 64), (64, 128), (128, 256)], 1)]:
        M, D, P = 8, 32, 4
        L = len(shapes)
        S = sum(h * w for h, w in shapes)
        g = torch.Generator(device=dev).manual_seed(5)
        value = torch.randn(B, S, M, D, device=dev, generator=g)
        loc = torch.rand(B, S, M, L, P, 2, device=dev, generator=g)
        aw = torch.softmax(torch.randn(B, S, M, L * P, device=dev, generator=g), -1).view(B, S, M, L, P)
        ss = torch.as_tensor(shapes, dtype=torch.long, device=dev)
        lsi = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
        r = ext.ms_deform_attn_forward(value, ss, lsi, loc, aw, 128)
        o = ops.ms_deform_attn_forward(value, ss, lsi, loc, aw, 128)
        t_ref = timeit(lambda: ext.ms_deform_attn_forward(value, ss, lsi, loc, aw, 128))
        t_our = timeit(lambda: ops.ms_deform_attn_forward(value, ss, lsi, loc, aw, 128))
        out.append({"case": name, "reference_ext_us": t_ref, "rba_msda_forward_us": t_our, "speedup": t_ref / t_our,
                    "max_abs_diff": float((r - o).abs().max())})
        print(out[-1])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "msda_vs_reference.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

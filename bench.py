#!/usr/bin/env python
"""bench.py — images/s of the RbA hot path (Mask2Former Swin-B 1dl forward + Rejected-by-All score) at
1024x2048, BASELINE.json configs[1] ("Swin-B 1dl, batch 8x1024x2048 synthetic, 1xB200").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--height H] [--width W]

A "step" = one pass of the hot path over one batch of B synthetic uint8 images per GPU: patch-embed ... decoder ...
fused score -> (B,H,W) anomaly-score maps (what evaluate_ood.get_RbA returns).  Random-init weights of the
ckpts/swin_b_1dl architecture (no network for checkpoints), synthetic uint8 images.

  value  images/s with inputs resident in HBM (CUDA-graph replay of rba_forward; N>1: one process per GPU, images
         sharded, + ONE NCCL all-gather of the score maps per step, overlapped with the next step's forward on a side
         stream), device-timed with CUDA events, max over ranks.
  e2e    the same metric through the public streaming call (rba_b200.ScoreStream) with HOST buffers: every step's
         pinned H2D of its uint8 batch, its forward and the D2H of its score maps are inside the timed region; the
         three legs of consecutive steps overlap on three CUDA streams.
  roofline     the fused mask-einsum + upsample + sigmoid + contraction + tanh score kernel (the kernel BASELINE's metric
               names; score_fused3.cu), timed alone with CUDA events on its launch stream on inputs > L2 (1.07 GB at B=8).
  cpu_baseline the oracle port (oracle/rba_oracle.py, PyTorch CPU fp32, all host threads) on a bounded sample.

`--impl reference` times that CPU port of the reference's path as the reference arm (rank 0 only).
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec at 1024x2048 Swin-B 1dl (Mask2Former forward + RbA score)"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step")
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--model", default="swin_b_1dl", choices=["swin_b_1dl", "swin_l_1dl", "swin_b_full", "r50_1dl", "r50_full", "tiny"])
    ap.add_argument("--backend", default=os.environ.get("RBA_GEMM_BACKEND", "auto"), choices=["auto", "ffma", "tc"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-images", type=int, default=1)
    return ap.parse_args()


def model_config(name):
    from rba_b200 import config
    return {"swin_b_1dl": config.swin_b_1dl, "swin_l_1dl": config.swin_l_1dl, "swin_b_full": config.swin_b_full,
            "r50_1dl": config.r50_1dl, "r50_full": config.r50_full, "tiny": config.tiny_test}[name]()


# ------------------------------------------------------------------------------------------------
# clocks (pynvml; nvidia-smi CSV fallback)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU port (oracle) timing — cpu_baseline leg and the reference arm
# ------------------------------------------------------------------------------------------------
def cpu_port_images_per_s(mc, H, W, n_images, warmup, steps, budget_s=240.0):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rba_oracle as O
    from rba_b200 import weights
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = weights.init_state_dict(mc, seed=0)
    g = torch.Generator().manual_seed(1)
    imgs = [torch.randint(0, 256, (3, H, W), dtype=torch.uint8, generator=g) for _ in range(n_images)]
    t_start = time.perf_counter()
    for _ in range(warmup):
        O.forward(sd, mc, imgs)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        out = O.forward(sd, mc, imgs)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    assert torch.isfinite(out["rba"][0]).all()
    total = sum(times)
    return n_images * len(times) / total, len(times), total / len(times), threads


# ------------------------------------------------------------------------------------------------
# the UNMODIFIED reference modules (baseline/_ref, installed by tools/make_baseline_ref.py; travels to the GPU box)
# ------------------------------------------------------------------------------------------------
REF_CKPT = {"swin_b_1dl": "swin_b_1dl", "swin_l_1dl": "swin_l_1dl", "swin_b_full": "swin_b_1dl", "r50_1dl": "swin_b_1dl",
            "r50_full": "swin_b_1dl"}


def reference_available(model_name):
    return model_name in REF_CKPT and os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "mask2former"))


def build_reference(model_name, mc, device, native_msda):
    """MaskFormer(cfg) of the reference itself (its own swin.py / msdeformattn.py / decoder / maskformer_model.py under
    the detectron2 / fvcore / timm stand-ins of oracle/ref_shims), with the SAME seeded weights this repo's arm uses."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.setdefault("RBA_REFERENCE_ROOT", os.path.join(ROOT, "baseline", "_ref"))
    import ref_loader
    from rba_b200 import weights
    if native_msda:
        ref_loader.use_native_msda()
    over = {}
    if model_name == "swin_b_full":
        over = {"MODEL.SEM_SEG_HEAD.DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES": ["res3", "res4", "res5"],
                "MODEL.MASK_FORMER.DEC_LAYERS": mc.dec_layers + 1}
    if model_name.startswith("r50"):     # the reference's modules around the detectron2 ResNet stand-in (oracle/ref_shims)
        over.update(ref_loader.r50_overrides(dec_layers=mc.dec_layers, levels=mc.num_enc_levels))
    cfg = ref_loader.load_cfg(REF_CKPT[model_name], over)
    model = ref_loader.build_reference_model(cfg, seed=0)
    ref_loader.load_state_dict_into(model, weights.init_state_dict(mc, seed=0))
    return model.to(device).eval()


def reference_cpu_images_per_s(model_name, mc, H, W, n_images, warmup, steps, budget_s=240.0):
    """The reference's own CPU path: model([{"image": x}, ...]) + get_RbA (evaluate_ood.py:143-150), all host threads."""
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = build_reference(model_name, mc, torch.device("cpu"), native_msda=False)
    g = torch.Generator().manual_seed(1)
    imgs = [torch.randint(0, 256, (3, H, W), dtype=torch.uint8, generator=g) for _ in range(n_images)]

    def one():
        with torch.no_grad():
            out = model([{"image": x} for x in imgs])
            return [-o["sem_seg"].tanh().sum(dim=0) for o in out]

    t_start = time.perf_counter()
    for _ in range(warmup):
        one()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        r = one()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    assert torch.isfinite(r[0]).all()
    total = sum(times)
    return n_images * len(times) / total, len(times), total / len(times), threads


def reference_gpu_baseline(model_name, mc, B, H, W, dev, warmup=2, steps=5, ours_rba=None, ours_img=None):
    """SURVEY §8d "Reference GPU path (the >=4x denominator)": the reference modules on the same B200, fp32, default
    torch flags (matmul TF32 off, cuDNN conv TF32 on: what a user gets), eval + no_grad, the batch as a list of B dicts,
    MSDeformAttn through the reference's own CUDA extension; CUDA-event timing.  `value` has the images resident on the
    device; `e2e` copies each image from pinned host memory and brings every score map back (evaluate_ood.py:143-150)."""
    model = build_reference(model_name, mc, dev, native_msda=True)
    g = torch.Generator().manual_seed(1)
    host = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=g).pin_memory()
    imgs = host.to(dev)

    def fwd(batch):
        with torch.no_grad():
            out = model([{"image": x} for x in batch])
            return [-o["sem_seg"].tanh().sum(dim=0) for o in out]

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for _ in range(warmup):
        fwd(imgs)
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    ms = timed(lambda: fwd(imgs), steps)

    def e2e_step():
        r = fwd([host[b].to(dev, non_blocking=True) for b in range(B)])
        return [x.cpu() for x in r]

    e2e_step()
    ms_e2e = timed(e2e_step, max(2, steps // 2))
    clk = clocks.stop()
    res = {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch": B, "steps": steps, "warmup": warmup,
           "e2e": {"value": B / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * 3 * H * W, "d2h_bytes_per_step": B * H * W * 4},
           "kind": "reference", "clocks": clk, "torch": torch.__version__,
           "flags": {"matmul_allow_tf32": torch.backends.cuda.matmul.allow_tf32, "cudnn_allow_tf32": torch.backends.cudnn.allow_tf32},
           "how": "unmodified reference modules (baseline/_ref) under oracle/ref_shims on cuda, MSDeformAttn through the reference's "
                  "own CUDA extension rebuilt for sm_100a, batch as a list of B dicts, eval + no_grad, CUDA events"}
    if ours_rba is not None and ours_img is not None:
        # live cross-check on this box: the reference with TF32 off against this repo's scores of the same image
        tf = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        r = fwd([ours_img])[0]
        torch.backends.cudnn.allow_tf32 = tf
        d = (r - ours_rba).abs()
        res["live_check"] = {"rba_max_abs_vs_ours": float(d.max()), "rba_frac_within_1e-3": float((d < 1e-3).float().mean()),
                             "note": "free-running (near-threshold attention-mask decisions may differ; tests/test_reference_gpu.py "
                                     "separates them)"}
    del model
    torch.cuda.empty_cache()
    return res


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mc = model_config(args.model)
    warm = min(args.warmup, 1)
    kind = "reference" if reference_available(args.model) else "port"
    if kind == "reference":
        ips, steps, sec_per_step, threads = reference_cpu_images_per_s(args.model, mc, args.height, args.width,
                                                                       args.cpu_sample_images, warm, args.steps)
    else:
        ips, steps, sec_per_step, threads = cpu_port_images_per_s(mc, args.height, args.width, args.cpu_sample_images, warm, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} Mask2Former forward + RbA score, {args.height}x{args.width}, "
                               f"{args.cpu_sample_images} image/step (bounded sample of the 8-image batch)"},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{steps} step(s) x {args.cpu_sample_images} image(s) at {args.height}x{args.width}; "
                                   + ("the reference's own modules (baseline/_ref: mask2former/*.py unmodified, under the detectron2 / "
                                      "fvcore / timm stand-ins of oracle/ref_shims; MSDeformAttn through its own PyTorch statement "
                                      "ms_deform_attn_core_pytorch, as on any CPU run of the reference)" if kind == "reference" else
                                      "oracle/rba_oracle.py = PyTorch-CPU fp32 restatement of the reference modules "
                                      "(baseline/_ref is absent)")},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import rba_b200
    from rba_b200 import ops, weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner to stdout (NCCL_DEBUG=VERSION/WARN in this image): stdout must carry only the ONE
        # JSON line, so file descriptor 1 points at stderr while the communicator is created and warmed up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    mc = model_config(args.model)
    B, H, W = args.batch, args.height, args.width
    sd = weights.init_state_dict(mc, seed=0)
    eng = rba_b200.Engine(mc, local).load_state_dict(sd)
    backend = args.backend
    if backend == "auto":
        backend = os.environ.get("RBA_BENCH_BACKEND", "tc")
    eng.set_gemm_backend(backend)
    del sd

    g = torch.Generator().manual_seed(1 + rank)
    host_imgs = [torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=g).pin_memory() for _ in range(2)]
    dev_imgs = [h.to(dev) for h in host_imgs]
    host_out = torch.empty((B, H, W), dtype=torch.float32).pin_memory()
    from rba_b200.parallel import OverlappedGather, PeerGather
    # the one exchange of the path: score maps to rank 0 (what evaluate_ood consumes).  Default: copy-engine peer writes into
    # rank 0's symmetric-memory buffer (no collective kernel on the SMs); RBA_GATHER=nccl: one NCCL all-gather per step
    og, collective = None, None
    if world > 1:
        if os.environ.get("RBA_GATHER", "peer") == "peer":
            try:
                og = PeerGather((B, H, W), dev)
                collective = "copy-engine peer writes of the score maps into rank 0's symmetric-memory buffer (NVLink), side stream"
            except Exception as ex:                      # report the transport that actually ran
                print(f"[bench] symmetric-memory gather unavailable ({type(ex).__name__}: {ex}); using NCCL all-gather", file=sys.stderr)
        if og is None:
            og = OverlappedGather((B, H, W), dev)
            collective = "one NCCL all-gather of the score maps per step, side stream"

    # one eager forward: warms position tables / function attributes and counts this library's launches per step
    n0 = rba_b200.launch_count()
    out = eng.forward(dev_imgs[0], rba=True)
    torch.cuda.synchronize()
    launches_per_step = rba_b200.launch_count() - n0
    assert torch.isfinite(out["rba"]).all(), "non-finite scores"

    use_graph = not args.no_graph
    if use_graph:
        try:
            static_in, static_out, graph = eng.graphed(dev_imgs[0], rba=True)
        except Exception as e:  # report, then fall back to eager launches (same kernels)
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({e}); timing eager launches", file=sys.stderr)
            use_graph = False
    if not use_graph:
        static_in = torch.empty_like(dev_imgs[0])
        static_out = eng.alloc_outputs(B, H, W, rba=True)

    def step_device(i):
        static_in.copy_(dev_imgs[i & 1], non_blocking=True)      # alternate inputs (device-resident)
        if use_graph:
            graph.replay()
        else:
            eng.forward_into(static_in, static_out)
        if world > 1:
            og.submit(static_out["rba"])

    # e2e: the public streaming call (rba_b200.ScoreStream): pinned H2D of batch i+1, forward of batch i and D2H of the
    # scores of batch i-1 run on three streams; every step copies its own inputs in and its own results out
    post = (lambda out: og.submit(out["rba"])) if world > 1 else None
    stream = rba_b200.ScoreStream(eng, B, H, W, use_graph=use_graph, post_forward=post)

    def timed_e2e(warmup, steps):
        for i in range(warmup):
            stream.step(host_imgs[i & 1])
        stream.drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        chk = 0.0
        for i in range(steps):
            r = stream.step(host_imgs[i & 1])                    # returns the host scores of the previous step
            if r is not None:
                chk += float(r[0, 0, 0])                         # the caller reads the scores every step
        for r in stream.drain():                                 # blocks until the last D2H has landed
            chk += float(r[0, 0, 0])
        if og is not None:
            og.wait()
        e1.record()
        torch.cuda.synchronize()
        assert chk == chk, "non-finite scores on the host"
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def timed(fn, warmup, steps):
        for i in range(warmup):
            fn(i)
        if og is not None:
            og.wait()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        if og is not None:
            og.wait()                                            # the last all-gather is inside the timed region
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            per = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(per, ms)                              # every rank's own device time (diagnostic, see `per_rank`)
            timed.last_per_rank = [float(t.item()) for t in per]
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- roofline of the fused mask-einsum + RbA score kernel (score_fused3.cu), timed ALONE (before the long timed loops
    # put the GPU at its power cap) with CUDA events on the stream it is launched
    # on; its inputs (feature planes 134 MB/img) exceed L2 at every batch size ----
    Q, K, D = mc.num_queries, mc.num_classes, mc.conv_dim
    Hp, Wp = eng.padded_hw(H, W)
    h4, w4 = Hp // 4, Wp // 4
    gk = torch.Generator(device=dev).manual_seed(2)
    # synthetic operands with the statistics of the real ones (mask logits ~ N(-0.54, 0.99^2))
    feat = torch.randn(B * h4 * w4, D, device=dev, generator=gk)
    emb = torch.randn(B * Q, D, device=dev, generator=gk) * (0.99 / D ** 0.5)
    f_pl = tuple(t.view(B, h4, w4, D) for t in ops.split_planes(feat))
    e_pl = tuple(t.view(B, Q, D) for t in ops.split_planes(emb))
    del feat, emb
    kbias = torch.full((B, Q), -0.54, device=dev)
    klogits = torch.randn(B, Q, K + 1, device=dev, generator=gk)
    for _ in range(3):
        ops.einsum_score_fused(e_pl, f_pl, klogits, (H, W), bias=kbias)
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.einsum_score_fused(e_pl, f_pl, klogits, (H, W), bias=kbias)
    e1.record()
    torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / reps
    # SURVEY §8d "Variant A": feature planes + mask embeddings + class logits + bias in, score map out
    alg_bytes = B * (4 * D * h4 * w4 + 4 * Q * D + 4 * Q * (K + 1) + 4 * Q + 4 * H * W)
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    # DRAM traffic of this kernel from the committed `ncu --set full` capture (dram__bytes_read.sum + dram__bytes_write.sum,
    # per image; profiles/r2t_fused_score_traffic.json), scaled to this launch's batch
    traffic, issue = None, None
    tp = os.path.join(ROOT, "profiles", "r2t_fused_score_traffic.json")
    other_bounds = None
    if os.path.exists(tp):
        cap = json.load(open(tp))
        traffic = cap["dram_bytes_per_image"] * B
        if "warp_instructions_per_image" in cap:
            # the resources that actually bind this kernel (three co-equal floors, from the committed ncu capture, at
            # sm_max_mhz): warp-instruction issue (4 schedulers x 1 / clk / SM), the XU (MUFU) pipe and the mma.sync pipe
            sm_clock_hz = 1e6 * float(peaks.get("sm_max_mhz", 1965.0))
            floor_ms = cap["warp_instructions_per_image"] * B / (4.0 * 148 * sm_clock_hz) * 1e3
            issue = {"warp_instructions_per_launch": cap["warp_instructions_per_image"] * B, "floor_ms": floor_ms,
                     "frac": floor_ms / k_ms, "source": "smsp__inst_executed.sum of the committed ncu capture, at sm_max_mhz"}
            other_bounds = {}
            for name, key in (("mufu", "xu_busy_cycles_per_sm_per_image"), ("tensor_mma_sync", "tensor_busy_cycles_per_sm_per_image")):
                if key in cap:
                    f_ms = cap[key] * B / sm_clock_hz * 1e3
                    other_bounds[name] = {"floor_ms": f_ms, "frac": f_ms / k_ms}
            other_bounds["issue"] = {"floor_ms": floor_ms, "frac": floor_ms / k_ms}
    roofline = {"kernel": "rba_einsum_score3_kernel (tcgen05 mask einsum -> x4 bilinear in the thread -> run-form sigmoids -> (Q,K) "
                          "contraction on f16 hi/lo mma.sync -> tanh -> class sum; one HBM pass)", "bound": "hbm",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst, kernel timed alone)" if peaks else "fallback 6.65 TB/s",
                "traffic": traffic, "ms_per_launch": k_ms, "algorithmic_bytes_per_launch": alg_bytes, "issue_bound": issue,
                "binding_bounds": other_bounds,
                "note": "fp32 semantics make this kernel issue/MUFU-bound, not HBM-bound (SURVEY §0.5): per output pixel "
                        "Q sigmoids of individually interpolated logits (1 MUFU + ~9 other instructions each) + 2*K*Q contraction FLOP vs 68 B"}

    del f_pl, e_pl, kbias, klogits
    torch.cuda.empty_cache()

    clocks = ClockSampler(local)
    clocks.start()
    ms_total = timed(step_device, args.warmup, args.steps)
    clk = clocks.stop()
    per_rank = None
    if world > 1:
        # diagnostic (not the metric): each rank's own device time per step with the exchange, and for the forward alone, plus
        # its median SM clock -- tells board-power skew between the GPUs (the step is the MAX over ranks) from communication
        with_x = [t / args.steps for t in timed.last_per_rank]

        def step_local(i):
            static_in.copy_(dev_imgs[i & 1], non_blocking=True)
            if use_graph:
                graph.replay()
            else:
                eng.forward_into(static_in, static_out)
        og_saved, og = og, None
        timed(step_local, 1, args.steps)
        og = og_saved
        alone = [t / args.steps for t in timed.last_per_rank]
        mhz = torch.tensor([float(clk.get("sm_mhz") or 0)], device=dev)
        mhzs = [torch.zeros_like(mhz) for _ in range(world)]
        dist.all_gather(mhzs, mhz)
        per_rank = {"ms_per_step": with_x, "ms_per_step_forward_only": alone, "sm_mhz": [float(t.item()) for t in mhzs]}
    ms_e2e = timed_e2e(max(2, args.warmup // 2), args.steps)
    value = world * B * args.steps / (ms_total * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    gpu_baseline = None
    if world == 1 and not args.no_gpu_baseline and reference_available(args.model):
        try:
            ours_img = dev_imgs[0][0].contiguous()
            ours_rba = eng.forward(ours_img[None], rba=True)["rba"][0].clone()
            gpu_baseline = reference_gpu_baseline(args.model, mc, B, H, W, dev, ours_rba=ours_rba, ours_img=ours_img)
            gpu_baseline["speedup_value"] = value / gpu_baseline["value"]
            gpu_baseline["speedup_e2e"] = e2e / gpu_baseline["e2e"]["value"]
        except Exception as ex:          # the baseline leg must never take the product's line down
            gpu_baseline = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        if reference_available(args.model):
            ips, steps, sps, threads = reference_cpu_images_per_s(args.model, mc, H, W, args.cpu_sample_images, 0, 1)
            cpu_baseline = {"value": ips, "unit": UNIT, "cores": threads, "kind": "reference",
                            "sample": f"{steps} step x {args.cpu_sample_images} image at {H}x{W} ({sps:.1f} s); the reference's own "
                                      "modules from baseline/_ref on the host cores"}
        else:
            ips, steps, sps, threads = cpu_port_images_per_s(mc, H, W, args.cpu_sample_images, 0, 1)
            cpu_baseline = {"value": ips, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": f"{steps} step x {args.cpu_sample_images} image at {H}x{W} ({sps:.1f} s); oracle/rba_oracle.py, "
                                      "PyTorch-CPU fp32 restatement of the reference modules"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (GEMM operands as bf16 hi+lo split planes, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": f"{args.model} Mask2Former forward + RbA score, batch {B}x{H}x{W} uint8 per GPU "
                               + ("(BASELINE.json configs[1])" if args.model == "swin_b_1dl" and B == 8 else ""),
                   "gemm_backend": backend, "cuda_graph": use_graph,
                   "l2": "activations are several GB per step (>> 126 MB L2); two input batches alternate",
                   "parallelism": f"dp{world}: images sharded, weights replicated" + (f"; {collective}" if world > 1 else "")},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": B * 3 * H * W, "d2h_bytes_per_step": B * H * W * 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches_per_step * args.steps),
        "gpu_launches_per_step": int(launches_per_step),
        "clocks": clk, "roofline": roofline,
    }
    if per_rank:
        line["per_rank"] = per_rank
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    if gpu_baseline:
        line["gpu_baseline"] = gpu_baseline
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

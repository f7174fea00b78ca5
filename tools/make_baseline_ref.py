"""Installs the UNMODIFIED reference into baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box),
so the reference's own GPU path can be timed on the B200 next to this repo's (bench.py `gpu_baseline` leg, SURVEY §8d
"Reference GPU path (the >=4x denominator)") and its evaluate_ood.py can run unchanged there (BASELINE config 5).

    python tools/make_baseline_ref.py            # run in the build container, where /root/reference exists

What it does:
  1. copies mask2former/, datasets/, ckpts/*/config.yaml, evaluate_ood.py, support.py, train_net.py verbatim
     (`pip install /root/reference` is not possible: the reference has no setup.py / pyproject and its dependencies
     detectron2 / fvcore / timm are not in the wheelhouse);
  2. builds the reference's MultiScaleDeformableAttention CUDA extension for sm_100a from a scratch copy of ops/ with
     the ONE patch torch >= 2.x needs (`value.type()` -> `value.scalar_type()` in AT_DISPATCH_FLOATING_TYPES,
     ops/src/cuda/ms_deform_attn_cuda.cu:69,139) and drops the .so into baseline/_ref/.  The kernels are untouched:
     this is "the reference GPU kernel to beat" of SURVEY §2.2.
Nothing under baseline/_ref/ is product code and nothing in rba_b200/ imports it.
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

REF = os.environ.get("RBA_REFERENCE_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")


def _writable(path):
    for r, ds, fs in os.walk(path):          # the reference checkout is read-only; the copy must stay replaceable
        for n in ds + fs:
            os.chmod(os.path.join(r, n), 0o755 if n in ds else 0o644)


def copy_tree():
    os.makedirs(DST, exist_ok=True)
    _writable(DST)
    for d in ("mask2former", "datasets"):
        shutil.rmtree(os.path.join(DST, d), ignore_errors=True)
        shutil.copytree(os.path.join(REF, d), os.path.join(DST, d),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "build", "*.egg-info"))
    for f in ("evaluate_ood.py", "support.py", "train_net.py"):
        shutil.copy2(os.path.join(REF, f), os.path.join(DST, f))
    for cfg in glob.glob(os.path.join(REF, "ckpts", "*", "config.yaml")):
        out = os.path.join(DST, "ckpts", os.path.basename(os.path.dirname(cfg)))
        os.makedirs(out, exist_ok=True)
        shutil.copy2(cfg, out)
    cfgs = os.path.join(REF, "configs")
    if os.path.isdir(cfgs):
        shutil.rmtree(os.path.join(DST, "configs"), ignore_errors=True)
        shutil.copytree(cfgs, os.path.join(DST, "configs"))


def build_msda_ext():
    src = os.path.join(REF, "mask2former", "modeling", "pixel_decoder", "ops")
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "ops")
        shutil.copytree(src, work)
        cu = os.path.join(work, "src", "cuda", "ms_deform_attn_cuda.cu")
        s = open(cu).read()
        n = s.count("AT_DISPATCH_FLOATING_TYPES(value.type()")
        assert n == 2, f"expected two AT_DISPATCH sites to patch, found {n}"
        s = s.replace("AT_DISPATCH_FLOATING_TYPES(value.type()", "AT_DISPATCH_FLOATING_TYPES(value.scalar_type()")
        s = s.replace(".type().is_cuda()", ".is_cuda()")           # DeprecatedTypeProperties::is_cuda is gone in torch 2.x
        open(cu, "w").write(s)
        env = dict(os.environ, FORCE_CUDA="1", TORCH_CUDA_ARCH_LIST="10.0a", MAX_JOBS="8")
        out = os.path.join(tmp, "out")
        r = subprocess.run([sys.executable, "setup.py", "build_ext", "--build-lib", out], cwd=work, env=env,
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-4000:] + r.stderr[-4000:])
            raise SystemExit("building the reference MSDA extension failed")
        sos = glob.glob(os.path.join(out, "MultiScaleDeformableAttention*.so"))
        assert len(sos) == 1, sos
        shutil.copy2(sos[0], DST)
        return os.path.join(DST, os.path.basename(sos[0]))


if __name__ == "__main__":
    assert os.path.isdir(os.path.join(REF, "mask2former")), f"{REF} is not the reference checkout"
    copy_tree()
    _writable(DST)
    so = build_msda_ext()
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write("verbatim copy of NazirNayal8/RbA (mask2former/, datasets/, configs/, ckpts/*/config.yaml, evaluate_ood.py, "
                "support.py, train_net.py) made by tools/make_baseline_ref.py; MultiScaleDeformableAttention*.so built from "
                "ops/ with value.type() -> value.scalar_type() (the AT_DISPATCH_FLOATING_TYPES sites) for sm_100a.\n")
    print("baseline/_ref ready:", so)

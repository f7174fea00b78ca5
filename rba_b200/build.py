"""In-tree build of librba_b200.so (sm_100a only).  nvcc cross-compiles without a GPU.

    python -m rba_b200.build            # incremental
    python -m rba_b200.build --force    # rebuild everything
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(OUT_DIR, "librba_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "rba_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, force, hm):
    obj = os.path.join(OUT_DIR, src[:-3] + ".o")
    sp = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(sp), hm):
        return src, 0, "(cached)"
    cmd = [NVCC] + ARCH + FLAGS + ["-c", sp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build_library(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    hm = _headers_mtime()
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force, hm), srcs))
    failed = False
    for src, rc, log in results:
        if rc != 0:
            failed = True
            sys.stderr.write(f"--- {src} FAILED ---\n{log}\n")
        elif verbose:
            sys.stderr.write(f"--- {src} ---\n{log}\n")
    if failed:
        raise RuntimeError("nvcc failed")
    objs = [os.path.join(OUT_DIR, s[:-3] + ".o") for s in srcs]
    if (force or not os.path.exists(LIB_PATH)
            or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs)):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB_PATH] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)

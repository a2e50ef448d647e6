#!/usr/bin/env python
"""Build the UNMODIFIED reference CUDA extensions for sm_100a into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  The reference's five pybind11 extension modules
(raymarching, hashencoder, gridencoder, freqencoder, shencoder) are compiled from
the sources *where they lie* under /root/reference (never copied into this repo)
and linked into oracle/_ref/_<name>.so (git-ignored, but shipped to the GPU box).  The reference's Python files
(model, renderer, wrappers, configs) are mirrored next to them into oracle/_ref/py/ (git-ignored as well) so
that the reference's own NeRFNetwork / NeRFRenderer can run on the GPU box (oracle/ref_model.py).
`tests/` (-m gpu) import these modules as the ground-truth checker for our own
kernels; nothing in the product path (envidr_b200/) may import them.

We do not run the reference's own build system (its setup.py pins -std=c++14,
which torch 2.11 headers reject); this is the short recipe instead:
    nvcc -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a <pkg>/src/<pkg>.cu
    g++  -std=c++17 <pkg>/src/bindings.cpp
    g++  -shared ... -ltorch -ltorch_python -lc10 -lc10_cuda -lcudart
Each .cu takes 4-6 minutes (torch headers under nvcc); the five are built in parallel.
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

REF = os.environ.get("ENVIDR_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
PKGS = ["raymarching", "hashencoder", "gridencoder", "freqencoder", "shencoder"]


def _flags():
    from torch.utils import cpp_extension as ce
    import torch
    inc = []
    for p in ce.include_paths("cuda"):
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    return inc, libdir


def build_one(pkg, inc, libdir, force=False):
    so = os.path.join(OUT, f"_{pkg}.so")
    src_cu = os.path.join(REF, pkg, "src", f"{pkg}.cu")
    src_cpp = os.path.join(REF, pkg, "src", "bindings.cpp")
    if not os.path.exists(src_cu):
        return pkg, "skipped (reference sources not present)"
    if os.path.exists(so) and not force and os.path.getmtime(so) > os.path.getmtime(src_cu):
        return pkg, "up to date"
    tmp = os.path.join(OUT, "build")
    os.makedirs(tmp, exist_ok=True)
    defs = [f"-DTORCH_EXTENSION_NAME=_{pkg}", "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=1"]
    o_cu = os.path.join(tmp, f"{pkg}.cu.o")
    o_cpp = os.path.join(tmp, f"{pkg}.bindings.o")
    nvcc = ["nvcc", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
            "-gencode", "arch=compute_100a,code=sm_100a",
            "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
            "-U__CUDA_NO_HALF2_OPERATORS__", "-Xcompiler", "-fPIC"]
    if pkg == "freqencoder":           # reference: freqencoder/setup.py:10
        nvcc.append("-use_fast_math")
    subprocess.check_call(nvcc + defs + inc + ["-c", src_cu, "-o", o_cu])
    subprocess.check_call(["g++", "-O3", "-std=c++17", "-fPIC"] + defs + inc + ["-c", src_cpp, "-o", o_cpp])
    subprocess.check_call(["g++", "-shared", o_cu, o_cpp, "-o", so, f"-L{libdir}",
                           "-L/usr/local/cuda/lib64", "-lc10", "-lc10_cuda", "-ltorch_cpu",
                           "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
                           f"-Wl,-rpath,{libdir}"])
    return pkg, "built"


def copy_py_tree():
    """The reference's Python files (its model / renderer / wrappers / configs) -> oracle/_ref/py/ (git-ignored, shipped to the
    GPU box): /root/reference does not exist there, and tests/test_gpu_refmodel.py + bench.py's gpu_reference run the reference's
    OWN NeRFNetwork / NeRFRenderer.render / run_cuda on the B200 (oracle/ref_model.py).  Nothing is edited; data, figures,
    notebooks and checkpoints are not copied."""
    import shutil
    dst_root = os.path.join(OUT, "py")
    n = 0
    for sub in ("nerf", "raymarching", "hashencoder", "gridencoder", "freqencoder", "shencoder", "ide_encoder", "configs"):
        for dirpath, dirnames, filenames in os.walk(os.path.join(REF, sub)):
            dirnames[:] = [d for d in dirnames if d not in ("src", "__pycache__", "build")]
            for fn in filenames:
                if not fn.endswith((".py", ".ini", ".txt")):
                    continue
                src = os.path.join(dirpath, fn)
                dst = os.path.join(dst_root, os.path.relpath(src, REF))
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
                    shutil.copy2(src, dst)
                    n += 1
    for fn in ("encoding.py", "activation.py", "loss.py"):
        src, dst = os.path.join(REF, fn), os.path.join(dst_root, fn)
        if os.path.exists(src) and (not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src)):
            shutil.copy2(src, dst)
            n += 1
    print(f"[oracle/_ref] py/: {n} file(s) refreshed")


def main(argv):
    os.makedirs(OUT, exist_ok=True)
    if not os.path.isdir(REF):
        print(f"[oracle/_ref] {REF} not present; using prebuilt files in {OUT} if any")
        return 0
    copy_py_tree()
    inc, libdir = _flags()
    pkgs = [p for p in argv if p in PKGS] or PKGS
    with ThreadPoolExecutor(max_workers=5) as ex:
        for pkg, status in ex.map(lambda p: build_one(p, inc, libdir, "--force" in argv), pkgs):
            print(f"[oracle/_ref] _{pkg}.so: {status}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))

"""Build libenvidr_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m envidr_b200.build [--force]

No torch headers are involved (each TU compiles in seconds); nvcc cross-compiles without a GPU.
The shared library lands in envidr_b200/_lib/ (git-ignored, shipped to the GPU box by gpurun).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libenvidr_b200.so")
SOURCES = ["error.cu", "raymarching.cu", "gridenc.cu", "direnc.cu", "field.cu", "render.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _sources():
    return [s for s in SOURCES + EXTRA_SOURCES if os.path.exists(os.path.join(CSRC, s))]


EXTRA_SOURCES = ["field_tc.cu", "tc_probe.cu", "geom_tc.cu", "shade_tc.cu", "linear_tc.cu", "density.cu", "optim.cu", "epilogue.cu", "neus.cu", "neus_field.cu", "neus_geom_tc.cu", "env_train_tc.cu"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "envidr_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []

    def compile_one(src):
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        cmd = ["nvcc"] + NVCC_FLAGS + (["-DENVIDR_TC_WATCHDOG"] if os.environ.get("ENVIDR_TC_WATCHDOG") == "1" else []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = ["nvcc", "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Build ``libb200xai.so`` from ``csrc/*.cu`` with nvcc for sm_100a (cross-compiles without a GPU).

The library is built IN-TREE (next to this file) so that it travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
# developer overrides for A/B builds of kernel variants: B200X_LIB_OUT=<path of the .so>  B200X_NVCC_EXTRA="-DNAME=value ..."
LIB = Path(os.environ["B200X_LIB_OUT"]) if os.environ.get("B200X_LIB_OUT") else HERE / "libb200xai.so"
EXTRA = os.environ.get("B200X_NVCC_EXTRA", "").split()
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libb200xai.so")
    return exe


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [HERE.parent / "include" / "b200xai.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    obj_dir = HERE / ("build" if not EXTRA else "build_" + "_".join(t.strip("-D").replace("=", "") for t in EXTRA))
    obj_dir.mkdir(exist_ok=True)

    def compile_one(src: Path):
        obj = obj_dir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *EXTRA, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (obj_dir / (src.stem + ".ptxas.log")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stderr[-4000:]}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs), "-cudart", "shared",
           # the CUDA runtime is resolved at load time (the process normally has torch's libcudart.so.12 mapped already);
           # RUNPATH covers a bare `ctypes.CDLL` without LD_LIBRARY_PATH
           "-Xlinker", "-rpath=/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/cuda_runtime/lib:/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))

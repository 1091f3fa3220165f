"""Build libvipe_ba.so (the C-ABI CUDA library) in-tree for sm_100a with plain nvcc.

No torch headers are involved (the library takes raw pointers), so a full rebuild is seconds.
Output: vipe_b200/lib/libvipe_ba.so  (git-ignored, shipped to the GPU box by gpurun).
"""

from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib"
SO = LIB / "libvipe_ba.so"
SOURCES = ["ba_kernels.cu", "ba_linearize2.cu", "ba_lin3.cu", "ba_lin4.cu", "chol.cu", "ba_api.cu", "geom_ops.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    return os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")


def _deps() -> list[Path]:
    deps = [p for p in CSRC.glob("*") if p.suffix in (".cu", ".cuh", ".h")]
    deps.append(PKG.parent / "include" / "vipe_ba.h")
    return deps


def needs_build() -> bool:
    if not SO.is_file():
        return True
    t = SO.stat().st_mtime
    return any(p.stat().st_mtime > t for p in _deps())


def build(force: bool = False, verbose: bool = False, ptxas_verbose: bool = False) -> Path:
    if not force and not needs_build():
        return SO
    LIB.mkdir(exist_ok=True)
    srcs = [s for s in SOURCES if (CSRC / s).is_file()]
    flags = ["-O3", "-std=c++17", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC", f"-I{PKG.parent / 'include'}"]
    if ptxas_verbose:
        flags += ["-Xptxas", "-v"]
    flags += [f"-D{d}" for d in os.environ.get("VIPE_BA_DEFINES", "").split() if d]  # developer switches, e.g. VBA_LIN3_TRACE

    def compile_one(src: str):
        obj = LIB / (Path(src).stem + ".o")
        cmd = [_nvcc(), "-c", str(CSRC / src), "-o", str(obj), *flags]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose or ptxas_verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [_nvcc(), "-shared", "-o", str(SO), *map(str, objs), *ARCH, "-cudart", "static"]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    for o in objs:
        o.unlink(missing_ok=True)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_verbose="-v" in sys.argv))

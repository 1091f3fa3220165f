"""TEST INFRASTRUCTURE ONLY.  Build the reference's own `slam_ext` (CUDA) into oracle/_ref/.

The reference sources are compiled WHERE THEY LIE under /root/reference (nothing is copied into this
repo): csrc/slam_ext/geom_kernels.cu and csrc/slam_ext/slam.cpp, unmodified, with the reference's own
flags (`-O3 -DWITH_CUDA --use_fast_math`, vipe/ext/specs.py:34-39) plus the sm_100a gencode.  The only
thing that is not the reference is Eigen, which this image lacks: `-I oracle/eigen_stub` supplies a
dense-backed stand-in for the few Eigen types the host code touches (see eigen_stub/eigen3/Eigen/Sparse).

Output: oracle/_ref/vipe_ref_ext.so (git-ignored, travels to the GPU box with gpurun).
Only runs where /root/reference exists (the build container); on the GPU box the prebuilt file is used.

Usage: python oracle/build_ref.py [--force]
"""

from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path(os.environ.get("VIPE_REFERENCE_ROOT", "/root/reference"))
OUT = HERE / "_ref"
NAME = "vipe_ref_ext"


def ref_available() -> bool:
    return (REF / "csrc/slam_ext/geom_kernels.cu").is_file()


def so_path() -> Path:
    return OUT / f"{NAME}.so"


def build(force: bool = False, verbose: bool = True) -> Path | None:
    if not ref_available():
        return so_path() if so_path().is_file() else None
    srcs = [REF / "csrc/slam_ext/geom_kernels.cu", REF / "csrc/slam_ext/slam.cpp", HERE / "ref_module.cpp",
            HERE / "eigen_stub/eigen3/Eigen/Sparse"]
    out = so_path()
    if out.is_file() and not force:
        newest = max(p.stat().st_mtime for p in srcs)
        if out.stat().st_mtime >= newest:
            return out
    OUT.mkdir(exist_ok=True)

    import torch
    from torch.utils import cpp_extension as ce

    incs = [str(HERE / "eigen_stub")] + ce.include_paths() + [sysconfig.get_paths()["include"]]
    inc_flags = [f"-I{p}" for p in incs]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    common = ["-O3", "-DWITH_CUDA", "-std=c++17", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
              f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H"]
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")

    def run(cmd):
        if verbose:
            print(" ".join(map(str, cmd)), flush=True)
        subprocess.check_call(list(map(str, cmd)))

    objs = []
    o = OUT / "geom_kernels.o"
    run([nvcc, "-c", srcs[0], "-o", o, *common, "--use_fast_math", "--expt-relaxed-constexpr",
         "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC,-fopenmp", *inc_flags])
    objs.append(o)
    for s in srcs[1:3]:
        o = OUT / (Path(s).stem + ".o")
        run(["g++", "-c", s, "-o", o, *common, "-fPIC", "-fopenmp", *inc_flags])
        objs.append(o)
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    run(["g++", "-shared", "-o", out, *objs, f"-L{tlib}", f"-Wl,-rpath,{tlib}",
         "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
         "-L/usr/local/cuda/lib64", "-lcudart", "-fopenmp"])
    for o in objs:
        Path(o).unlink(missing_ok=True)
    return out


def load():
    """Import the prebuilt reference module (never builds; returns None when absent)."""
    p = so_path()
    if not p.is_file():
        return None
    import importlib.machinery
    import importlib.util

    import torch  # noqa: F401  (libtorch must be loaded first)

    loader = importlib.machinery.ExtensionFileLoader(NAME, str(p))
    spec = importlib.util.spec_from_loader(NAME, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print("reference build:", p)
